#!/usr/bin/env python
"""diagnostics: cycle counters inside k_chain for one batch (tools only, not part of the bench)"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import ssq_testlib as T, bench
gl = int(sys.argv[1]) if len(sys.argv) > 1 else bench.GENOME_LEN
s = T.SSQ()
fa, g = bench.ensure_reference(os.path.join(ROOT, "data_cache"), gl, lambda f: s.index_build(f))
idx = s.index_load(fa)
seq, off = bench.fast_pairs(g, 1000000, 150, 1000)
h = C.c_void_p()
s.ck(s.lib.ssq_batch_create(idx, s.opts, C.c_int(len(off) - 1), seq.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), C.byref(h)), "create")
for _ in range(2):
    s.ck(s.lib.ssq_batch_run(h), "run")
c = [int(s.lib.ssq_batch_counter(h, i)) for i in range(23)]
ms = [s.lib.ssq_batch_stage_ms(h, i) for i in range(5)]
thr = 148 * 12 * 128
print("genome", gl, "stage ms", [round(x, 1) for x in ms], "seeds", c[7])
print("k_chain cycles per thread (avg): add_seed %.1fM finish %.1fM fetch/init %.1fM ; kernel %.1fM cycles" % (c[12] / thr / 1e6, c[13] / thr / 1e6, c[14] / thr / 1e6, ms[2] * 1.965e3 / 1e3))
print("slowest single read: %.2fM cycles (%.2f ms) with n_seeds=%d n_chains=%d" % (c[15] / 1e6, c[15] / 1.965e6, c[16] >> 32, c[16] & 0xffffffff))
print("per seed: add %.0f cycles" % (c[12] / max(1, c[7])))
print("tiers: light %.1f ms, heavy %.1f ms, heavy reads %d (of which %d overflowed to the giant tier), cut %d seeds" % (c[17] / 1e3, c[18] / 1e3, c[19], c[21], c[20]))
print("giant tier alone: %.1f ms" % (c[22] / 1e3))
