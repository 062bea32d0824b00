"""The HBM-resident `bwa mem | samblaster` pipeline (speedseq_b200/csrc/ssq_pipe.cu) on the CPU: tests/hostsim runs the very
same per-thread bodies the kernels run (ssq_dev3.cuh: sort/dedup/patch, insert-size votes, mate rescue, primary marking, pairing,
MAPQ, CIGAR/NM/MD, samblaster's signature / discordant / splitter tests, SAM text) in plain loops, with the same host-side
reductions (ssq_pipe_host.h).  Output must equal the oracle's `bwa mem` text, and the oracle's `bwa mem | samblaster` streams,
byte for byte.  GPU parity proper: tests/test_gpu_pipe.py."""
import subprocess

import numpy as np
import pytest

import ssq_testlib as T


def stress_reads(g, bounds, n_pairs, rl, seed, **kw):
    """simulated pairs plus what simulated pairs never contain: unmappable mates (orphans, mate rescue), chimeric reads
    (supplementary lines, SA tags, splitters), exact duplicate pairs, junk pairs (both ends unmapped), improper pairs"""
    names, seqs, quals = T.simulate_pairs(g, bounds, n_pairs, rl, seed, **kw)
    rng = np.random.default_rng(seed)
    rnd = lambda: "".join("ACGT"[x] for x in rng.integers(0, 4, rl))
    for k in range(0, len(seqs), 40):
        seqs[k + 1] = rnd()
    for k in range(10, len(seqs), 50):
        seqs[k] = seqs[k][: rl // 2] + seqs[(k + 200) % len(seqs)][rl // 2:]
    for k in range(4, len(seqs) - 2, 120):  # junk pair
        seqs[k], seqs[k + 1] = rnd(), rnd()
    for k in range(20, len(seqs) - 2, 90):  # improper pair: mate taken from another fragment
        seqs[k + 1] = seqs[(k + 301) % len(seqs) | 1]
    nd = n_pairs // 10
    for k in range(nd):  # duplicates of earlier pairs under new names (some with the ends swapped)
        j = int(rng.integers(0, n_pairs))
        a, b = (2 * j, 2 * j + 1) if rng.random() < 0.7 else (2 * j + 1, 2 * j)
        names += ["dup%d" % k] * 2
        seqs += [seqs[a], seqs[b]]
        quals += [quals[a], quals[b]]
    return names, seqs, quals


def oracle_streams(oracle, idx, fa_header, names, seqs, quals, n_processed, rg, sb_args, tmp_path):
    body = oracle.mem_pe(idx, names, seqs, quals, n_processed, 4, rg)
    spl, disc = str(tmp_path / "o.spl"), str(tmp_path / "o.disc")
    out = subprocess.run([T.ORACLE_BIN, "samblaster"] + sb_args + ["--splitterFile", spl, "--discordantFile", disc], input=(fa_header + body).encode(), check=True,
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rec = lambda t: "".join(l for l in t.splitlines(True) if not l.startswith("@"))
    return body, rec(out), rec(open(spl).read()), rec(open(disc).read())


def sq_header(oracle, idx, fa):
    return "".join("@SQ\tSN:%s\tLN:%d\n" % (l.split()[1], int(m.split()[1])) for l, m in zip(*[iter(open(fa + ".ann").read().splitlines()[1:])] * 2))


def test_plain_bwa_mem_example_reads(oracle, hostsim, ex_index, ex_reads):
    idx = oracle.load(ex_index)
    names, seqs, quals = ex_reads
    assert hostsim.mem_pe(idx, names, seqs, quals, 0, b"NA12878") == oracle.mem_pe(idx, names, seqs, quals, 0, 4, b"NA12878")


@pytest.mark.parametrize("rl,seed,kw", [(75, 1, {}), (150, 2, dict(err=0.02, indel=0.004, n_frac=0.004)), (250, 3, dict(err=0.01, indel=0.003)), (101, 4, dict(ins_mean=250, ins_sd=80))])
def test_plain_bwa_mem_synthetic_stress(oracle, hostsim, syn_index, rl, seed, kw):
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    names, seqs, quals = stress_reads(g, bounds, 500, rl, seed, **kw)
    assert hostsim.mem_pe(idx, names, seqs, quals, 1000, b"rg1") == oracle.mem_pe(idx, names, seqs, quals, 1000, 4, b"rg1")


@pytest.mark.parametrize("env", [{"HOSTSIM_RESCUE_SPEC": "0"}, {"HOSTSIM_RESCUE_DROP": "1"}, {"HOSTSIM_RESCUE_DROP": "2"}, {"HOSTSIM_RESCUE_DROP": "5"}])
def test_mate_rescue_computed_ahead_and_in_the_replay(oracle, hostsim, syn_index, monkeypatch, env):
    """the rescue alignments are computed ahead as tasks and looked up by the sequential replay (ssq_dev2.cuh: RTask / RCache); a
    replay step that finds no result computes it itself.  Same records whether everything is computed in the replay (SPEC=0), nothing
    is found ahead (DROP=1) or every 2nd / 5th result is withheld"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    names, seqs, quals = stress_reads(g, bounds, 500, 150, 2, err=0.02, indel=0.004, n_frac=0.004)
    assert hostsim.mem_pe(idx, names, seqs, quals, 1000, b"rg1") == oracle.mem_pe(idx, names, seqs, quals, 1000, 4, b"rg1")


def test_plain_bwa_mem_single_end(oracle, hostsim, syn_index):
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    names, seqs, quals = stress_reads(g, bounds, 300, 150, 9)
    names = ["s%d" % i for i in range(len(names))]
    lib = oracle.lib
    # the oracle's single-end form: ssqo_api_mem_pe pairs adjacent reads, so go through the CLI with a FASTQ (no -p)
    import os, tempfile
    with tempfile.TemporaryDirectory() as d:
        fq = os.path.join(d, "se.fq")
        with open(fq, "w") as f:
            for n, s, q in zip(names, seqs, quals):
                f.write("@%s\n%s\n+\n%s\n" % (n, s, q))
        ref = subprocess.run([T.ORACLE_BIN, "mem", "-t", "2", fa, fq], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    ref = "".join(l for l in ref.splitlines(True) if not l.startswith("@"))
    assert hostsim.mem_pe(idx, names, seqs, quals, 0, b"", paired=0) == ref


@pytest.mark.parametrize("rl,seed,sb,args", [
    (150, 5, (1, 1, 2, 20, 0), ["--excludeDups", "--addMateTags", "--maxSplitCount", "2", "--minNonOverlap", "20"]),   # speedseq's invocation
    (101, 6, (0, 1, 3, 10, 0), ["--addMateTags", "--maxSplitCount", "3", "--minNonOverlap", "10"]),
    (250, 7, (0, 0, 2, 20, 1), ["--removeDups", "--maxSplitCount", "2", "--minNonOverlap", "20"]),
])
def test_fused_pipeline_equals_bwa_pipe_samblaster(oracle, hostsim, syn_index, tmp_path, rl, seed, sb, args):
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    names, seqs, quals = stress_reads(g, bounds, 600, rl, seed, err=0.01, indel=0.002)
    _, o_main, o_spl, o_disc = oracle_streams(oracle, idx, sq_header(oracle, idx, fa), names, seqs, quals, 0, b"rgX", args, tmp_path)
    h_main, h_spl, h_disc = hostsim.pipe(idx, names, seqs, quals, 0, b"rgX", 1, sb)
    assert h_main == o_main
    assert h_spl == o_spl
    assert h_disc == o_disc
    flags = [int(l.split("\t")[1]) for l in h_main.splitlines()]
    assert (sum(1 for f in flags if f & 0x400) > 50 or sb[4]) and h_spl.count("\n") > 10 and h_disc.count("\n") > 20  # --removeDups drops them instead
    assert "\t77\t" in h_main and "\t77\t" not in h_disc


def test_fused_pipeline_batches_share_the_dup_set(oracle, hostsim, syn_index, tmp_path):
    """three consecutive batches through one run: per-batch insert-size statistics, global read ordinals, duplicates whose first
    occurrence lies in an earlier batch"""
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    names, seqs, quals = stress_reads(g, bounds, 900, 150, 11, err=0.01)
    cuts = [0, 700, 1300, len(names)]
    hdr = sq_header(oracle, idx, fa)
    body = "".join(oracle.mem_pe(idx, names[a:b], seqs[a:b], quals[a:b], a, 4, b"r") for a, b in zip(cuts, cuts[1:]))
    spl, disc = str(tmp_path / "o.spl"), str(tmp_path / "o.disc")
    out = subprocess.run([T.ORACLE_BIN, "samblaster", "--excludeDups", "--addMateTags", "--splitterFile", spl, "--discordantFile", disc], input=(hdr + body).encode(), check=True,
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rec = lambda t: "".join(l for l in t.splitlines(True) if not l.startswith("@"))
    got = ["", "", ""]
    for k, (a, b) in enumerate(zip(cuts, cuts[1:])):
        r = hostsim.pipe(idx, names[a:b], seqs[a:b], quals[a:b], a, b"r", 1, (1, 1, 2, 20, 0), reset=1 if k == 0 else 0)
        for i in range(3):
            got[i] += r[i]
    assert got[0] == rec(out) and got[1] == rec(open(spl).read()) and got[2] == rec(open(disc).read())


def test_fused_pipeline_single_end_reads(oracle, hostsim, syn_index, tmp_path):
    """single-end input through the fused stage: every read is its own block (samblaster's lone-record rule: signature from the one mapped
    record, duplicates by 5' position and strand, no mate tags, never discordant or splitter)"""
    import os
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    names, seqs, quals = stress_reads(g, bounds, 400, 150, 17)
    names = ["s%d" % i for i in range(len(names))]
    fq = str(tmp_path / "se.fq")
    with open(fq, "w") as f:
        for n, s, q in zip(names, seqs, quals):
            f.write("@%s\n%s\n+\n%s\n" % (n, s, q))
    sam = subprocess.run([T.ORACLE_BIN, "mem", "-t", "2", fa, fq], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    spl, disc = str(tmp_path / "o.spl"), str(tmp_path / "o.disc")
    out = subprocess.run([T.ORACLE_BIN, "samblaster", "--excludeDups", "--addMateTags", "--splitterFile", spl, "--discordantFile", disc], input=sam, check=True,
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rec = lambda t: "".join(l for l in t.splitlines(True) if not l.startswith("@"))
    h_main, h_spl, h_disc = hostsim.pipe(idx, names, seqs, quals, 0, b"", 0, (1, 1, 2, 20, 0))
    assert h_main == rec(out) and h_spl == rec(open(spl).read()) and h_disc == rec(open(disc).read())
    assert sum(1 for l in h_main.splitlines() if int(l.split("\t")[1]) & 0x400) > 20 and h_disc == ""
