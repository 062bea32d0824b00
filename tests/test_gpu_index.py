"""GPU index builder (`bwa index` replacement) against the reference's own golden index files and against the oracle."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

import ssq_testlib as T

pytestmark = pytest.mark.gpu


def test_gpu_index_matches_reference_goldens(ssq, tmp_path):
    fa = str(tmp_path / "ex.fa")
    open(fa, "wb").write(gzip.open(os.path.join(T.GOLDEN, "ex_ref.fa.gz")).read())
    ssq.index_build(fa)
    gold = json.load(open(os.path.join(T.GOLDEN, "ex_index.sha256.json")))
    for ext, g in gold.items():
        data = open(fa + "." + ext, "rb").read()
        assert len(data) == g["size"], ext
        assert hashlib.sha256(data).hexdigest() == g["sha256"], ext


def test_gpu_index_gz_input_and_prefix(ssq, tmp_path):
    gz = os.path.join(T.GOLDEN, "ex_ref.fa.gz")
    prefix = str(tmp_path / "pfx")
    ssq.index_build(gz, prefix)
    gold = json.load(open(os.path.join(T.GOLDEN, "ex_index.sha256.json")))
    assert hashlib.sha256(open(prefix + ".bwt", "rb").read()).hexdigest() == gold["bwt"]["sha256"]


@pytest.mark.parametrize("n,nc,seed", [(1000, 1, 1), (4097, 3, 2), (250000, 5, 3), (1 << 20, 2, 4)])
def test_gpu_index_equals_oracle_on_synthetic_genomes(ssq, oracle, tmp_path, n, nc, seed):
    """multi-contig genomes with N runs (hole table + lrand48 replacement), long exact repeats and l_pac % 4 in {0,1,2,3}"""
    g, bounds = T.synth_genome(n, seed, n_contigs=nc)
    a, b = str(tmp_path / "a.fa"), str(tmp_path / "b.fa")
    T.write_fasta(a, g, bounds)
    # plant ambiguity codes
    txt = open(a).read().split("\n")
    rng = np.random.default_rng(seed)
    for k in rng.integers(1, len(txt) - 1, 6):
        if txt[k] and not txt[k].startswith(">"):
            txt[k] = txt[k][:5] + "NNNNnnRY" + txt[k][13:]
    open(a, "w").write("\n".join(txt))
    open(b, "w").write("\n".join(txt))
    oracle.index_build(a)
    ssq.index_build(b)
    for ext in ("amb", "ann", "pac", "bwt", "sa"):
        assert open(a + "." + ext, "rb").read() == open(b + "." + ext, "rb").read(), ext


def test_gpu_index_then_align(ssq, oracle, tmp_path):
    g, bounds = T.synth_genome(300000, 12, n_contigs=2)
    fa = str(tmp_path / "x.fa")
    T.write_fasta(fa, g, bounds)
    ssq.index_build(fa)
    h = ssq.index_load(fa)
    oidx = oracle.load(fa)  # the oracle loads the GPU-built files
    names, seqs, quals = T.simulate_pairs(g, bounds, 500, 150, 3)
    seq, off = T.encode_reads(seqs)
    a, ao = oracle.align_batch(oidx, seq, off)
    b, bo = ssq.align_batch(h, seq, off)
    assert np.array_equal(ao, bo) and np.array_equal(a, b)
    ssq.index_free(h)
