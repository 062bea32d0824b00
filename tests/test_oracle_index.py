"""Oracle vs the reference's own golden index files (the only reference-owned known-answer data for this path)."""
import hashlib
import json
import os

import numpy as np

import ssq_testlib as T


def test_index_matches_reference_goldens(ex_index):
    gold = json.load(open(os.path.join(T.GOLDEN, "ex_index.sha256.json")))
    for ext, g in gold.items():
        data = open(ex_index + "." + ext, "rb").read()
        assert len(data) == g["size"], ext
        assert hashlib.sha256(data).hexdigest() == g["sha256"], ext


def test_index_loads_and_sa_is_consistent(oracle, ex_index):
    idx = oracle.load(ex_index)
    l_pac, n, primary = oracle.info(idx, 0), oracle.info(idx, 1), oracle.info(idx, 2)
    assert n == 2 * l_pac == 643270 and primary == 405526
    # SA look-up of every 997th row must be a permutation-consistent position in [0, n]
    rows = np.arange(1, n + 1, 997, dtype=np.uint64)
    pos = oracle.sa_batch(idx, rows)
    assert pos.max() <= n and len(np.unique(pos)) == len(pos)


def test_sais_against_naive(oracle):
    import ctypes as C
    rng = np.random.default_rng(3)
    for n in (1, 2, 17, 300, 5000):
        s = np.concatenate([rng.integers(1, 5, n), [0]]).astype(np.int32)
        sa = np.zeros(n + 1, np.int32)
        oracle.lib.ssqo_sais(s.ctypes.data_as(C.c_void_p), sa.ctypes.data_as(C.c_void_p), C.c_int32(n + 1), C.c_int32(5))
        b = s.tobytes()
        naive = sorted(range(n + 1), key=lambda i: s[i:].tolist())
        assert sa.tolist() == naive


def test_multi_contig_index_roundtrip(oracle, syn_index):
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    assert oracle.info(idx, 0) == len(g) and oracle.info(idx, 3) == 3
