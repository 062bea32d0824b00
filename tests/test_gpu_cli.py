"""The product's `bwa mem` path (ssq_mem_batch_sam + the CLI shims speedseq_b200/bin/{bwa,samblaster}) against the oracle:
SAM records, dup flags, MC/MQ tags, discordant and splitter streams must be identical byte for byte (everything except
the @PG header lines, which carry program paths)."""
import ctypes as C
import gzip
import os
import subprocess

import numpy as np
import pytest

import ssq_testlib as T

pytestmark = pytest.mark.gpu
BWA = os.path.join(T.ROOT, "speedseq_b200", "bin", "bwa")
SAMBLASTER = os.path.join(T.ROOT, "speedseq_b200", "bin", "samblaster")
RG = r"@RG\tID:NA12878\tSM:NA12878\tLB:lib1"


def _strip_pg(b):
    return b"".join(l for l in b.splitlines(True) if not l.startswith(b"@PG"))


def _mem_batch(ssq, idx, names, seqs, quals, n_processed, paired, rg):
    n = len(names)
    arr = lambda xs: (C.c_char_p * n)(*[x.encode() for x in xs])
    out, ln = C.c_char_p(), C.c_size_t(0)
    rc = ssq.lib.ssq_mem_batch_sam(idx, ssq.opts, C.c_int(n), arr(names), arr(seqs), arr(quals), None, C.c_int64(n_processed), C.c_int(paired), None, rg, C.c_int(0),
                                   C.byref(out), C.byref(ln), None)
    ssq.ck(rc, "ssq_mem_batch_sam")
    s = C.string_at(out, ln.value).decode()
    ssq.lib.ssq_free(out)
    return s


def test_mem_batch_sam_example_reads(ssq, oracle, ex_index, ex_reads):
    idx = oracle.load(ex_index)
    h = ssq.index_load(ex_index)
    names, seqs, quals = ex_reads
    a = oracle.mem_pe(idx, names, seqs, quals, 0, 8, b"NA12878")
    b = _mem_batch(ssq, h, names, seqs, quals, 0, 1, b"NA12878")
    assert a == b
    ssq.index_free(h)


@pytest.mark.parametrize("rl,seed,kw", [(75, 1, {}), (150, 2, dict(err=0.02, indel=0.004, n_frac=0.004)), (250, 3, dict(err=0.01, indel=0.003)), (101, 4, dict(ins_mean=250, ins_sd=80))])
def test_mem_batch_sam_synthetic_stress(ssq, oracle, syn_index, rl, seed, kw):
    fa, g, bounds = syn_index
    idx = oracle.load(fa)
    h = ssq.index_load(fa)
    names, seqs, quals = T.simulate_pairs(g, bounds, 1200, rl, seed, **kw)
    rng = np.random.default_rng(seed)
    for k in range(0, len(seqs), 40):   # unmappable mates: orphans / mate rescue
        seqs[k + 1] = "".join("ACGT"[x] for x in rng.integers(0, 4, rl))
    for k in range(10, len(seqs), 50):  # chimeric reads: supplementary lines, SA tags, splitters
        seqs[k] = seqs[k][: rl // 2] + seqs[(k + 200) % len(seqs)][rl // 2:]
    a = oracle.mem_pe(idx, names, seqs, quals, 1000, 8, b"rg1")
    b = _mem_batch(ssq, h, names, seqs, quals, 1000, 1, b"rg1")
    assert a == b
    ssq.index_free(h)


def test_cli_pipeline_matches_oracle_cli(ssq, tmp_path):
    """config 1 in miniature: bwa index + bwa mem -p | samblaster with FIFOs, product shims vs oracle CLI"""
    assert os.path.exists(BWA) and os.path.exists(SAMBLASTER), "run `make -C speedseq_b200/csrc cli`"
    fq = os.path.join(T.GOLDEN, "ex_reads_2k.fq.gz")
    outs = {}
    for tag, bwa, sb in (("oracle", [T.ORACLE_BIN], [T.ORACLE_BIN, "samblaster"]), ("b200", [BWA], [SAMBLASTER])):
        d = tmp_path / tag
        d.mkdir()
        fa = str(d / "ex.fa")
        open(fa, "wb").write(gzip.open(os.path.join(T.GOLDEN, "ex_ref.fa.gz")).read())
        subprocess.check_call(bwa + ["index", fa], stderr=subprocess.DEVNULL)
        sam = subprocess.run(bwa + ["mem", "-t", "2", "-p", "-R", RG, fa, fq], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
        spl, disc = str(d / "spl.sam"), str(d / "disc.sam")
        out = subprocess.run(sb + ["--excludeDups", "--addMateTags", "--maxSplitCount", "2", "--minNonOverlap", "20", "--splitterFile", spl, "--discordantFile", disc],
                             input=sam, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
        outs[tag] = (sam, out, open(spl, "rb").read(), open(disc, "rb").read(), [open(fa + "." + e, "rb").read() for e in ("amb", "ann", "pac", "bwt", "sa")])
    for i, what in enumerate(("bwa mem SAM", "samblaster SAM", "splitters", "discordants")):
        assert _strip_pg(outs["oracle"][i]) == _strip_pg(outs["b200"][i]), what
    assert outs["oracle"][4] == outs["b200"][4]
    assert outs["b200"][1].count(b"\n") > 4000


def test_samblaster_shim_dups_across_chunks(ssq, oracle, ex_index, ex_reads, tmp_path):
    """duplicates whose first occurrence lies many records earlier; also exercises ssq_dupset_mark across several calls"""
    idx = oracle.load(ex_index)
    names, seqs, quals = ex_reads
    names = names + ["dup_" + n for n in names[:1000]]
    seqs = seqs + seqs[:1000]
    quals = quals + quals[:1000]
    body = oracle.mem_pe(idx, names, seqs, quals, 0, 8, b"")
    sam = ("@SQ\tSN:20_slice\tLN:321635\n" + body).encode()
    res = {}
    for tag, cmd in (("oracle", [T.ORACLE_BIN, "samblaster"]), ("b200", [SAMBLASTER])):
        res[tag] = subprocess.run(cmd + ["--addMateTags"], input=sam, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    assert _strip_pg(res["oracle"]) == _strip_pg(res["b200"])
    assert res["b200"].count(b"\t1187\t") + res["b200"].count(b"\t1171\t") > 100  # flagged duplicates exist


def test_dupset_streaming_equals_one_shot(ssq, oracle):
    rng = np.random.default_rng(3)
    n = 60000
    sig = np.zeros(n, T.DUPSIG_DT)
    sig["pos1"] = rng.integers(0, 3000, n); sig["pos2"] = sig["pos1"] + rng.integers(0, 40, n)
    sig["strand1"] = rng.integers(0, 2, n); sig["strand2"] = rng.integers(0, 2, n); sig["valid"] = rng.random(n) > 0.03
    ref = oracle.dupmark(sig)
    h = C.c_void_p()
    ssq.ck(ssq.lib.ssq_dupset_create(0, C.byref(h)), "ssq_dupset_create")
    got = np.zeros(n, np.uint8)
    for lo in range(0, n, 7001):
        part = sig[lo:lo + 7001].copy()
        d = np.zeros(len(part), np.uint8)
        ssq.ck(ssq.lib.ssq_dupset_mark(h, C.c_uint64(len(part)), part.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p)), "ssq_dupset_mark")
        got[lo:lo + 7001] = d
    ssq.lib.ssq_dupset_size.restype = C.c_uint64
    assert np.array_equal(got, ref)
    assert int(ssq.lib.ssq_dupset_size(h)) == int(((ref == 0) & (sig["valid"] == 1)).sum())
    ssq.lib.ssq_dupset_free(h)
