#!/usr/bin/env python
"""throughput of the streaming duplicate set (ssq_dupset_mark_dev: two stable radix-sort passes + adjacent-equal mark inside the chunk,
binary search against the sorted set of earlier chunks, one merge to absorb the new signatures): chunks of 1 M pair signatures on device
pointers, the set growing to tens of millions.  Prints one line per decade of set size and a JSON summary (25 algorithmic bytes per pair:
16 B signature + 8 B ordinal in, 1 B out — SURVEY.md §8d).  Checks the first chunks against the oracle."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from speedseq_b200 import capi

n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 64
chunk = 1_000_000
s = capi.SSQ()
L = s.lib
h = C.c_void_p()
s.ck(L.ssq_dupset_create(0, C.byref(h)), "create")
L.ssq_dupset_mark_dev.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
L.ssq_dupset_size.restype = C.c_uint64
L.ssq_dupset_size.argtypes = [C.c_void_p]
rng = np.random.default_rng(1)
genome = 3_100_000_000
st = torch.cuda.Stream()
rows, all_sig, all_dup = [], [], []
for c in range(n_chunks):
    p1 = rng.integers(1, genome, chunk).astype(np.uint64)
    p2 = p1 + rng.integers(100, 900, chunk).astype(np.uint64)
    ndup = chunk // 10  # 10 % copies of signatures of this and earlier chunks
    src = rng.integers(0, chunk, ndup)
    dst = rng.choice(chunk, ndup, replace=False)
    if all_sig and c % 2:
        o1, o2 = all_sig[int(rng.integers(0, len(all_sig)))]
        p1[dst], p2[dst] = o1[src], o2[src]
    else:
        p1[dst], p2[dst] = p1[src], p2[src]
    if c < 4:
        all_sig.append((p1.copy(), p2.copy()))
    k1 = torch.from_numpy(((p1 << np.uint64(1)) | rng.integers(0, 2, chunk).astype(np.uint64) * 0).view(np.int64)).cuda()
    k2 = torch.from_numpy((p2 << np.uint64(1)).view(np.int64)).cuda()
    va = torch.ones(chunk, dtype=torch.uint8, device="cuda")
    du = torch.zeros(chunk, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        s.ck(L.ssq_dupset_mark_dev(h, chunk, k1.data_ptr(), k2.data_ptr(), va.data_ptr(), du.data_ptr(), st.cuda_stream), "mark")
        e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rows.append((int(L.ssq_dupset_size(h)), ms))
    if c < 4:
        all_dup.append(du.cpu().numpy())
# oracle check on the first four chunks as one stream
import ssq_testlib as T
o = T.Oracle()
sig = np.zeros(4 * chunk, T.DUPSIG_DT)
sig["pos1"] = np.concatenate([a for a, _ in all_sig]); sig["pos2"] = np.concatenate([b for _, b in all_sig]); sig["valid"] = 1
ok = bool(np.array_equal(o.dupmark(sig), np.concatenate(all_dup)))
for sz, ms in rows[:: max(1, n_chunks // 8)]:
    print("set %9d signatures: chunk of 1 M pairs in %6.3f ms -> %6.1f M pairs/s, %5.1f GB/s algorithmic" % (sz, ms, chunk / ms / 1e3, 25 * chunk / ms / 1e6))
first, last = np.mean([m for _, m in rows[1:5]]), np.mean([m for _, m in rows[-4:]])
print(json.dumps({"chunks": n_chunks, "chunk_pairs": chunk, "identical_to_oracle_first_4M": ok, "ms_per_chunk_small_set": first, "ms_per_chunk_final_set": last, "final_set": rows[-1][0],
                  "M_pairs_per_s_final": chunk / last / 1e3, "algorithmic_GBps_final": 25 * chunk / last / 1e6}))
sys.exit(0 if ok else 1)
