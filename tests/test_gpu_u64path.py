"""The GRCh37-class code path on the GPU: indexes with >= 2^32 BWT rows cannot be re-blocked into 32-byte u32 rank sectors, so they
run the 64-bit seeding machine k_smem_m<u64,5> on the 64-byte on-disk rank blocks (ScalarFm::occ4 u64), the u64 SA sample and the
u64 interval arithmetic.  No such index can be built on this box, so the same instantiations are forced on small indexes:
SSQ_NO_BWT32=1 (index load keeps only the on-disk blocks), SSQ_SMEM_M64=1 (64-bit machine even when bwt32 exists), SSQ_SA_DENSE=0
(walk to the on-disk every-32nd-row sample, u64 entries), SSQ_SA_U64=1 (u64 entries in the load-time densified sample).  Every kernel-level entry point and the region pipeline must stay
bit-identical to the oracle."""
import numpy as np
import pytest

import ssq_testlib as T

pytestmark = pytest.mark.gpu

MODES = [
    pytest.param({"SSQ_NO_BWT32": "1"}, id="no_bwt32"),
    pytest.param({"SSQ_NO_BWT32": "1", "SSQ_SA_DENSE": "0"}, id="no_bwt32+ondisk_sa"),
    pytest.param({"SSQ_SMEM_M64": "1"}, id="m64_on_bwt32"),
    pytest.param({"SSQ_NO_BWT32": "1", "SSQ_SA_DENSE": "4", "SSQ_SA_U64": "1"}, id="no_bwt32+dense4_u64"),
]


def _load(ssq, monkeypatch, prefix, env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    return ssq.index_load(prefix)


@pytest.mark.parametrize("env", MODES)
def test_u64_path_kernel_entries(ssq, oracle, ex_index, ex_reads, monkeypatch, env):
    idx = oracle.load(ex_index)
    h = _load(ssq, monkeypatch, ex_index, env)
    try:
        if "SSQ_NO_BWT32" in env:
            assert int(ssq.lib.ssq_index_info(h, 7)) == 64  # bytes per rank query: the on-disk block
        n = oracle.info(idx, 1)
        rows = np.concatenate([np.arange(0, 2000, dtype=np.uint64), np.random.default_rng(1).integers(0, n + 1, 30000).astype(np.uint64),
                               np.array([oracle.info(idx, 2), n, n - 1], np.uint64)])
        assert np.array_equal(ssq.sa_lookup_batch(h, rows), oracle.sa_batch(idx, rows))
        seq, off = T.encode_reads(ex_reads[1])
        a, ao = oracle.smem_batch(idx, seq, off)
        b, bo = ssq.smem_batch(h, seq, off)
        assert np.array_equal(ao, bo) and np.array_equal(a, b)
        a, ao = oracle.align_batch(idx, seq, off)
        b, bo = ssq.align_batch(h, seq, off)
        assert np.array_equal(ao, bo) and np.array_equal(a, b)
    finally:
        ssq.index_free(h)


@pytest.mark.parametrize("env", MODES[:2])
@pytest.mark.parametrize("rl,seed", [(75, 1), (150, 2), (250, 3)])
def test_u64_path_synthetic(ssq, oracle, syn_index, monkeypatch, env, rl, seed):
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    h = _load(ssq, monkeypatch, fa, env)
    try:
        names, seqs, quals = T.simulate_pairs(g, bounds, 1200, rl, seed, err=0.01, indel=0.002, n_frac=0.003)
        seqs = seqs + ["", "A", "N" * 30, "AC" * 70, "A" * 150, "ACGT" * 37, "T" * 100]  # edge cases + interval-rich reads (overflow pass)
        seq, off = T.encode_reads(seqs)
        a, ao = oracle.smem_batch(idx, seq, off)
        b, bo = ssq.smem_batch(h, seq, off)
        assert np.array_equal(ao, bo) and np.array_equal(a, b)
        a = oracle.chain_batch(idx, seq, off)
        b = ssq.chain_batch(h, seq, off)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        a, ao = oracle.align_batch(idx, seq, off)
        b, bo = ssq.align_batch(h, seq, off)
        assert np.array_equal(ao, bo) and np.array_equal(a, b)
    finally:
        ssq.index_free(h)
