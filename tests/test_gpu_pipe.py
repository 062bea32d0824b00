"""GPU parity of the HBM-resident `bwa mem | samblaster` pipeline (ssq_aligner_* in include/ssq.h, speedseq_b200/csrc/ssq_pipe.cu)
through the C-ABI: the plain records must equal the oracle's `bwa mem` text, the fused streams must equal the oracle's
`bwa mem | samblaster` main / splitter / discordant streams, byte for byte — including across several batches of one run
(per-batch insert-size statistics, global read ordinals, duplicates whose first occurrence lies in an earlier batch), with -I,
-C comments, single-end input and FASTA input (no qualities)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import ssq_testlib as T
from test_hostsim_pipe import stress_reads, sq_header, oracle_streams

pytestmark = pytest.mark.gpu
SB_SPEEDSEQ = dict(exclude_dups=1, add_mate_tags=1, max_split_count=2, min_non_overlap=20)
SB_ARGS = ["--excludeDups", "--addMateTags", "--maxSplitCount", "2", "--minNonOverlap", "20"]


@pytest.fixture(scope="module")
def gpu_syn(ssq, syn_index):
    h = ssq.index_load(syn_index[0])
    yield h
    ssq.index_free(h)


def _run(ssq, h, names, seqs, quals, n_processed=0, rg=b"", paired=1, sb=None, comments=None, pes=None, al=None):
    own = al is None
    if own:
        al = ssq.aligner_create(h, sb, rg)
    rd, keep = T.pack_reads(names, seqs, quals, comments, paired, n_processed)
    txt, out = ssq.aligner_run(al, rd, pes)
    if own:
        ssq.aligner_free(al)
    return [t.decode() for t in txt], out


def test_plain_records_example_reads(ssq, oracle, ex_index, ex_reads):
    idx = oracle.load(ex_index)
    h = ssq.index_load(ex_index)
    names, seqs, quals = ex_reads
    txt, out = _run(ssq, h, names, seqs, quals, 0, b"NA12878")
    assert txt[0] == oracle.mem_pe(idx, names, seqs, quals, 0, 8, b"NA12878")
    ro = out["read_off"]
    assert ro[0] == 0 and ro[-1] == len(txt[0]) and (np.diff(ro.astype(np.int64)) > 0).all()
    ssq.index_free(h)


@pytest.mark.parametrize("rl,seed,kw", [(75, 1, {}), (150, 2, dict(err=0.02, indel=0.004, n_frac=0.004)), (250, 3, dict(err=0.01, indel=0.003)), (101, 4, dict(ins_mean=250, ins_sd=80))])
def test_plain_records_synthetic_stress(ssq, oracle, syn_index, gpu_syn, rl, seed, kw):
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    names, seqs, quals = stress_reads(g, bounds, 1500, rl, seed, **kw)
    txt, _ = _run(ssq, gpu_syn, names, seqs, quals, 1000, b"rg1")
    assert txt[0] == oracle.mem_pe(idx, names, seqs, quals, 1000, 8, b"rg1")


def test_single_end_fasta_and_comments(ssq, oracle, syn_index, gpu_syn):
    fa, g, bounds = syn_index
    names, seqs, quals = stress_reads(g, bounds, 400, 150, 9)
    names = ["s%d" % i for i in range(len(names))]
    cm = ["BC:Z:%d" % (i % 7) if i % 3 else "" for i in range(len(names))]
    with tempfile.TemporaryDirectory() as d:
        fq, fasta = os.path.join(d, "se.fq"), os.path.join(d, "se.fa")
        with open(fq, "w") as f:
            for n, s, q, c in zip(names, seqs, quals, cm):
                f.write("@%s%s\n%s\n+\n%s\n" % (n, " " + c if c else "", s, q))
        with open(fasta, "w") as f:
            for n, s in zip(names, seqs):
                f.write(">%s\n%s\n" % (n, s))
        rec = lambda a: "".join(l for l in subprocess.run([T.ORACLE_BIN, "mem", "-t", "2"] + a, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode().splitlines(True) if not l.startswith("@"))
        assert _run(ssq, gpu_syn, names, seqs, quals, paired=0, comments=cm)[0][0] == rec(["-C", fa, fq])
        assert _run(ssq, gpu_syn, names, seqs, None, paired=0)[0][0] == rec([fa, fasta])


def test_insert_size_override(ssq, oracle, syn_index, gpu_syn):
    """-I 300,30 (bin/speedseq:286): the statistics of every batch are replaced; wide windows exercise the rescue scratch sizing"""
    fa, g, bounds = syn_index
    names, seqs, quals = stress_reads(g, bounds, 500, 101, 12, ins_mean=300, ins_sd=30)
    with tempfile.TemporaryDirectory() as d:
        fq = os.path.join(d, "i.fq")
        T.write_fastq(fq, names, seqs, quals)
        for spec, pes1 in (("300,30", (180, 420, 0, 300.0, 30.0)), ("2000,400,6000,100", (100, 6000, 0, 2000.0, 400.0))):
            ref = subprocess.run([T.ORACLE_BIN, "mem", "-t", "2", "-p", "-I", spec, fa, fq], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
            ref = "".join(l for l in ref.splitlines(True) if not l.startswith("@"))
            pes = [(0, 0, 1, 0, 0), pes1, (0, 0, 1, 0, 0), (0, 0, 1, 0, 0)]
            assert _run(ssq, gpu_syn, names, seqs, quals, pes=pes)[0][0] == ref, spec


@pytest.mark.parametrize("rl,seed,sb,args", [
    (150, 5, SB_SPEEDSEQ, SB_ARGS),
    (101, 6, dict(add_mate_tags=1, max_split_count=3, min_non_overlap=10), ["--addMateTags", "--maxSplitCount", "3", "--minNonOverlap", "10"]),
    (250, 7, dict(remove_dups=1), ["--removeDups", "--maxSplitCount", "2", "--minNonOverlap", "20"]),
])
def test_fused_streams_equal_bwa_pipe_samblaster(ssq, oracle, syn_index, gpu_syn, tmp_path, rl, seed, sb, args):
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    names, seqs, quals = stress_reads(g, bounds, 1500, rl, seed, err=0.01, indel=0.002)
    _, o_main, o_spl, o_disc = oracle_streams(oracle, idx, sq_header(oracle, idx, fa), names, seqs, quals, 0, b"rgX", args, tmp_path)
    txt, out = _run(ssq, gpu_syn, names, seqs, quals, 0, b"rgX", 1, sb)
    assert txt[0] == o_main
    assert txt[1] == o_spl
    assert txt[2] == o_disc
    n_dup_blocks = len(set(l.split("\t")[0] for l in o_main.splitlines() if int(l.split("\t")[1]) & 0x400))
    if not sb.get("remove_dups"):
        assert out["n_dup"] == n_dup_blocks > 50
    assert out["n_ids"] == len(names) // 2 and txt[1].count("\n") > 10 and txt[2].count("\n") > 20


def test_fused_streams_across_batches(ssq, oracle, syn_index, gpu_syn, tmp_path):
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    names, seqs, quals = stress_reads(g, bounds, 2000, 150, 11, err=0.01)
    cuts = [0, 1500, 2900, len(names)]
    hdr = sq_header(oracle, idx, fa)
    body = "".join(oracle.mem_pe(idx, names[a:b], seqs[a:b], quals[a:b], a, 8, b"r") for a, b in zip(cuts, cuts[1:]))
    spl, disc = str(tmp_path / "o.spl"), str(tmp_path / "o.disc")
    out = subprocess.run([T.ORACLE_BIN, "samblaster"] + SB_ARGS + ["--splitterFile", spl, "--discordantFile", disc], input=(hdr + body).encode(), check=True,
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    rec = lambda t: "".join(l for l in t.splitlines(True) if not l.startswith("@"))
    al = ssq.aligner_create(gpu_syn, SB_SPEEDSEQ, b"r")
    got = ["", "", ""]
    for a, b in zip(cuts, cuts[1:]):
        txt, _ = _run(ssq, gpu_syn, names[a:b], seqs[a:b], quals[a:b], a, al=al)
        for i in range(3):
            got[i] += txt[i]
    # an empty batch in the middle of a run is legal
    txt, o = _run(ssq, gpu_syn, [], [], [], len(names), al=al)
    assert txt == ["", "", ""] and o["n_ids"] == 0
    ssq.aligner_free(al)
    assert got[0] == rec(out) and got[1] == rec(open(spl).read()) and got[2] == rec(open(disc).read())


def test_bench_scale_sample_matches_oracle(ssq, oracle, syn_index, gpu_syn, tmp_path):
    """a 60 k-read batch (every kernel at real occupancy, pools and slabs past their first sizes): the first 6 k reads compared
    exactly with the oracle (same batch statistics supplied to both through -I-style overrides), the rest through invariants"""
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    names, seqs, quals = stress_reads(g, bounds, 30000, 150, 21)
    pes = [(0, 0, 1, 0, 0), (230, 570, 0, 400.0, 40.0), (0, 0, 1, 0, 0), (0, 0, 1, 0, 0)]
    txt, out = _run(ssq, gpu_syn, names, seqs, quals, 0, b"b", 1, SB_SPEEDSEQ, pes=pes)
    m = 6000
    sub, _ = _run(ssq, gpu_syn, names[:m], seqs[:m], quals[:m], 0, b"b", 1, None, pes=pes)
    with tempfile.TemporaryDirectory() as d:
        fq = os.path.join(d, "s.fq")
        T.write_fastq(fq, names[:m], seqs[:m], quals[:m])
        ref = subprocess.run([T.ORACLE_BIN, "mem", "-t", "8", "-p", "-I", "400,40,570,230", "-R", r"@RG\tID:b", fa, fq], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
    ref = "".join(l for l in ref.splitlines(True) if not l.startswith("@"))
    assert sub[0] == ref
    # the big batch: its first m reads' records are the same records (plus dup flag / mate tags), every read is present once, in order
    lines = txt[0].splitlines()
    qn = [l.split("\t", 1)[0] for l in lines]
    order = [k for i, k in enumerate(qn) if i == 0 or qn[i - 1] != k]
    assert order == [n for i, n in enumerate(names) if i % 2 == 0]
    strip = lambda l: "\t".join(f for f in l.split("\t") if not f.startswith(("MC:Z:", "MQ:i:")))
    unflag = lambda l: "\t".join([l.split("\t")[0], str(int(l.split("\t")[1]) & ~0x400)] + l.split("\t")[2:])
    big = [unflag(strip(l)) for l in lines[: ref.count("\n")]]
    assert big == ref.splitlines()


def test_invariants_on_200k_reads(ssq, syn_index, gpu_syn):
    """bwa-free invariants of the GPU's records at a scale the oracle is too slow for (200 k reads, three batches of one run): every
    alignment re-scored from its CIGAR agrees with AS, NM and MD agree with the reference, the simulated origin is recovered, flags of a
    pair are consistent, every read appears exactly once as a primary record, in input order, and duplicates planted under new names carry 0x400"""
    import re
    fa, g, bounds = syn_index
    ref = "".join("ACGT"[b] for b in g)
    names, seqs, quals = T.simulate_pairs(g, bounds, 90000, 150, 77)
    rng = np.random.default_rng(5)
    nd = 10000
    for k in range(nd):  # planted duplicates of earlier pairs
        j = int(rng.integers(0, 90000))
        names += ["dup%d" % k] * 2; seqs += [seqs[2 * j], seqs[2 * j + 1]]; quals += [quals[2 * j], quals[2 * j + 1]]
    al = ssq.aligner_create(gpu_syn, SB_SPEEDSEQ, b"inv")
    cuts = [0, 70000, 140000, len(names)]
    lines = []
    for a, b in zip(cuts, cuts[1:]):
        rd, keep = T.pack_reads(names[a:b], seqs[a:b], quals[a:b], None, 1, a)
        txt, info = ssq.aligner_run(al, rd)
        lines += txt[0].decode().splitlines()
    ssq.aligner_free(al)
    ctg = {"ctg%d" % (i + 1): int(bounds[i]) for i in range(len(bounds) - 1)}
    prim = [l for l in lines if not int(l.split("\t", 2)[1]) & 0x900]
    assert len(prim) == len(names) and [l.split("\t", 1)[0] for l in prim] == names
    n_checked = n_hit = n_dupflag = 0
    for l in lines:
        f = l.split("\t")
        flag = int(f[1])
        if f[0].startswith("dup") and flag & 0x400:
            n_dupflag += 1
        if flag & 4 or f[5] == "*" or f[9] == "*":
            continue
        tags = {t[:2]: t[5:] for t in f[11:]}
        pos = ctg[f[2]] + int(f[3]) - 1
        q = r = sc = nm = run = 0
        md = []
        for ln, op in re.findall(r"(\d+)([MIDSH])", f[5]):
            ln = int(ln)
            if op == "M":
                qs, rs = f[9][q:q + ln], ref[pos + r:pos + r + ln]
                if qs == rs and "N" not in qs:
                    sc += ln; run += ln
                else:
                    for a_, b_ in zip(qs, rs):
                        if a_ == b_ and a_ != "N":
                            sc += 1; run += 1
                        else:
                            sc -= 1 if "N" in (a_, b_) else 4
                            nm += 1; md.append(str(run) + b_); run = 0
                q += ln; r += ln
            elif op == "I":
                sc -= 6 + ln; nm += ln; q += ln
            elif op == "D":
                sc -= 6 + ln; nm += ln; md.append(str(run) + "^" + ref[pos + r:pos + r + ln]); run = 0; r += ln
            elif op == "S":
                q += ln
        md.append(str(run))
        assert q == len(f[9]) and nm == int(tags["NM"]) and "".join(md) == tags["MD"], l
        assert 0 <= int(tags["AS"]) - sc <= 8 and 0 <= int(f[4]) <= 60, l
        if flag & 1:
            assert bool(flag & 0x40) != bool(flag & 0x80)
        n_checked += 1
        if not flag & 0x900 and f[0].startswith("r"):
            _, c, p = f[0].split("_")
            n_hit += abs(pos - (int(bounds[int(c)]) + int(p))) < 700
    assert n_checked > 195000 and n_hit > 0.97 * 180000
    assert n_dupflag >= 2 * nd * 0.95  # (a planted copy of a pair that did not align at all cannot be marked)
