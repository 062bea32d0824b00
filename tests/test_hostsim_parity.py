"""The kernels' device routines (speedseq_b200/csrc/ssq_dev.cuh) compiled for the HOST by the test-only harness
tests/hostsim and compared with the oracle.  This validates the arithmetic the GPU executes on a box without a GPU; the
GPU parity proper is tests/test_gpu_parity.py.  The harness is not part of the product."""
import numpy as np
import pytest

import ssq_testlib as T


def _cmp_all(oracle, hostsim, idx, seqs):
    seq, off = T.encode_reads(seqs)
    a, ao = oracle.smem_batch(idx, seq, off)
    b, bo = hostsim.smem_batch(idx, seq, off)
    assert np.array_equal(ao, bo) and np.array_equal(a, b)
    a = oracle.chain_batch(idx, seq, off)
    b = hostsim.chain_batch(idx, seq, off)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    a, ao = oracle.align_batch(idx, seq, off)
    b, bo = hostsim.align_batch(idx, seq, off)
    assert np.array_equal(ao, bo) and np.array_equal(a, b)
    return len(a)


def test_example_reads(oracle, hostsim, ex_index, ex_reads):
    idx = oracle.load(ex_index)
    assert _cmp_all(oracle, hostsim, idx, ex_reads[1][:1500]) > 1000


def test_synthetic_reads_with_repeats_indels_and_Ns(oracle, hostsim, syn_index):
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    for rl, seed in ((75, 1), (150, 2), (250, 3)):
        names, seqs, quals = T.simulate_pairs(g, bounds, 300, rl, seed, err=0.01, indel=0.002, n_frac=0.003)
        _cmp_all(oracle, hostsim, idx, seqs)


@pytest.mark.parametrize("env", ["HOSTSIM_M64", "HOSTSIM_STRAIGHT", "HOSTSIM_SPLIT", "HOSTSIM_SPLIT_LEAN", "HOSTSIM_SPLIT_LEAN,HOSTSIM_KMER=7", "HOSTSIM_KMER=8"])
def test_seeding_formulations_agree(oracle, hostsim, syn_index, monkeypatch, env):
    """default = the 32-bit-row state machine (what the GPU runs when the index has < 2^32 rows); also the 64-bit machine and the
    straight-line smem1()/seed_strategy1() form, and the phase-split form (forward walks, backward sweeps and the greedy
    pass as separate stages connected by call records — what the GPU's split kernels do)"""
    for kv in env.split(","):  # the last one: the split form with the short strings answered from the k-mer jump-start table
        monkeypatch.setenv(*(kv.split("=") if "=" in kv else (kv, "1")))
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    names, seqs, quals = T.simulate_pairs(g, bounds, 400, 150, 21, err=0.015, indel=0.003, n_frac=0.004)
    seq, off = T.encode_reads(seqs + ["", "N" * 40, "AC" * 70, "T" * 120])
    a, ao = oracle.smem_batch(idx, seq, off)
    b, bo = hostsim.smem_batch(idx, seq, off)
    assert np.array_equal(ao, bo) and np.array_equal(a, b)


def test_edge_cases(oracle, hostsim, syn_index):
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    acgt = "ACGT"
    ref = "".join(acgt[x] for x in g[5000:5400])
    seqs = ["", "A", "N" * 30, "ACGT" * 4, ref[:18], ref[:19], ref[:20], ref[:150], "N" + ref[1:150], ref[:75] + "N" + ref[76:150],
            "AC" * 60,  # microsatellite, highly repetitive
            ref[:60] + ref[200:290],  # split read
            "".join(acgt[x] for x in g[int(bounds[1]) - 70:int(bounds[1]) + 80]),  # spans a contig boundary
            "T" * 100]
    _cmp_all(oracle, hostsim, idx, seqs)


def test_sw_extend_random_tasks(oracle, hostsim):
    rng = np.random.default_rng(5)
    tasks, q, t = T.extension_tasks(rng, 4000, qmax=255)
    a = oracle.sw_extend_batch(tasks, q, t)
    b = hostsim.sw_extend_batch(tasks, q, t)
    assert np.array_equal(a, b)
    assert (a["score"] >= tasks["h0"]).all()


def test_chain_tree_order_equals_ordered_array(hostsim):
    """ChainBuilder keeps chains in a search tree; its look-up / insert must reproduce the ordered-array semantics of the oracle
    (first chain with an equal position, else the predecessor; insert right after it) — exercised with many tied positions"""
    assert hostsim.lib.hostsim_chain_selftest(3000, 12345) == 0
