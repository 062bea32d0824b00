"""Text logic of the `samblaster` shim (parsing, MC/MQ, discordant / splitter predicates, chunked dup-set calls) fuzzed against
the oracle's samblaster on the CPU: the shim is built against a test-only stand-in for ssq_dupset_* (tests/stubs/dupset_stub.c)
so that no GPU is needed.  Inputs contain what simulated reads never do: pairs with both ends unmapped (flags 77/141), orphans,
lone records, supplementary lines, equal 5' positions on different contigs."""
import os
import subprocess

import numpy as np
import pytest

import ssq_testlib as T


@pytest.fixture(scope="module")
def shim(tmp_path_factory, oracle):
    d = tmp_path_factory.mktemp("shim")
    exe = str(d / "samblaster_stub")
    subprocess.check_call(["gcc", "-O1", "-w", "-I" + os.path.join(T.ROOT, "include"), "-I" + os.path.join(T.ROOT, "speedseq_b200", "cli"), "-o", exe,
                           os.path.join(T.ROOT, "speedseq_b200", "cli", "samblaster_main.c"), os.path.join(T.ROOT, "tests", "stubs", "dupset_stub.c")])
    return exe


def _fuzz_sam(seed, n_blocks):
    rng = np.random.default_rng(seed)
    ctg = [("c1", 5000), ("c2", 3000), ("c3", 800)]
    out = ["@HD\tVN:1.3\tSO:unsorted"] + ["@SQ\tSN:%s\tLN:%d" % c for c in ctg] + ["@PG\tID:bwa\tPN:bwa"]

    def cigar(l):
        k = rng.integers(0, 6)
        if k == 0:
            return "%dM" % l
        if k == 1:
            s = int(rng.integers(1, 40)); return "%dS%dM" % (s, l - s)
        if k == 2:
            s = int(rng.integers(1, 40)); return "%dM%dS" % (l - s, s)
        if k == 3:
            return "%dM2D%dM" % (l // 2, l - l // 2)
        if k == 4:
            return "%dM3I%dM" % (l // 2, l - l // 2 - 3)
        s = int(rng.integers(1, 30)); return "%dH%dM%dS" % (s, l - 2 * s, s)

    for b in range(n_blocks):
        name = "q%d" % b
        kind = rng.random()
        l = 100
        seq, qual = "A" * l, "I" * l

        def rec(flag, c, pos, mq, cg, mc, mpos, tl, tags=""):
            return "\t".join([name, str(flag), c, str(pos), str(mq), cg, mc, str(mpos), str(tl), seq, qual]) + tags

        if kind < 0.12:  # both ends unmapped
            out += [rec(77, "*", 0, 0, "*", "*", 0, 0), rec(141, "*", 0, 0, "*", "*", 0, 0)]
        elif kind < 0.24:  # orphan: one end mapped, the mate placed at its coordinates
            c, ln = ctg[rng.integers(0, 3)]; p = int(rng.integers(1, 60)); rev = int(rng.integers(0, 2))
            a = rec(0x49 | (0x10 if rev else 0), c, p, 37, cigar(l), "=", p, 0)
            u = rec(0x85 | (0x20 if rev else 0), c, p, 0, "*", "=", p, 0)
            out += [u.replace("\t%d\t" % (0x85 | (0x20 if rev else 0)), "\t%d\t" % ((0x85 | (0x20 if rev else 0)) ^ 0xC0), 1), a.replace("\t%d\t" % (0x49 | (0x10 if rev else 0)), "\t%d\t" % ((0x49 | (0x10 if rev else 0)) ^ 0xC0), 1)] if rng.random() < 0.5 else [a, u]
        elif kind < 0.30:  # lone single-end record
            c, ln = ctg[rng.integers(0, 3)]
            out += [rec(int(rng.choice([0, 16, 4])), c, int(rng.integers(1, 50)), 20, cigar(l), "*", 0, 0)]
        else:  # mapped pair, few distinct positions so that duplicates are frequent; sometimes improper / split
            c1, _ = ctg[rng.integers(0, 3)]; c2 = c1 if rng.random() < 0.8 else ctg[rng.integers(0, 3)][0]
            p1, p2 = int(rng.integers(1, 12)), int(rng.integers(1, 12) + (200 if rng.random() < 0.7 else 0))
            r1, r2 = int(rng.integers(0, 2)), int(rng.integers(0, 2))
            proper = 2 if (c1 == c2 and rng.random() < 0.6) else 0
            f1 = 0x41 | proper | (0x10 if r1 else 0) | (0x20 if r2 else 0)
            f2 = 0x81 | proper | (0x10 if r2 else 0) | (0x20 if r1 else 0)
            cg1, cg2 = cigar(l), cigar(l)
            lines = [rec(f1, c1, p1, 60, cg1, "=" if c1 == c2 else c2, p2, 0, "\tNM:i:0"), rec(f2, c2, p2, int(rng.integers(0, 61)), cg2, "=" if c1 == c2 else c1, p1, 0, "\tNM:i:1\tXS:i:0")]
            if rng.random() < 0.3:  # supplementary line of read 1 (split read): clipped on opposite sides
                s = int(rng.integers(25, 70))
                lines[0] = rec(f1, c1, p1, 60, "%dM%dS" % (l - s, s), "=" if c1 == c2 else c2, p2, 0, "\tNM:i:0\tSA:Z:x")
                sc = c1 if rng.random() < 0.5 else "c2"
                lines.insert(1, rec(f1 | 0x800 | (0x10 if rng.random() < 0.3 else 0), sc, int(rng.integers(300, 700)), 30, "%dH%dM" % (l - s, s) if rng.random() < 0.7 else "%dH%dM%dH" % (l - s - 10, s, 10), "=" if sc == c2 else c2, p2, 0, "\tNM:i:0"))
            if rng.random() < 0.1:  # a secondary line: never part of the pair, never a splitter
                lines.append(rec(f2 | 0x100, c2, int(rng.integers(1, 900)), 0, "%dM" % l, "=", p1, 0))
            out += lines
    return ("\n".join(out) + "\n").encode()


def _strip_pg(b):
    return b"".join(x for x in b.splitlines(True) if not x.startswith(b"@PG\tID:SAMBLASTER"))


@pytest.mark.parametrize("seed,chunk,extra", [(1, None, []), (2, "7", ["--excludeDups"]), (3, "1", []), (4, "64", ["--excludeDups", "--maxSplitCount", "3", "--minNonOverlap", "10"])])
def test_shim_text_logic_matches_oracle(shim, oracle, tmp_path, seed, chunk, extra):
    sam = _fuzz_sam(seed, 1500)
    res = {}
    for tag, cmd in (("oracle", [T.ORACLE_BIN, "samblaster"]), ("shim", [shim])):
        spl, disc = str(tmp_path / (tag + ".spl")), str(tmp_path / (tag + ".disc"))
        env = dict(os.environ)
        if chunk:
            env["SSQ_SB_CHUNK"] = chunk
        out = subprocess.run(cmd + extra + ["--addMateTags", "--splitterFile", spl, "--discordantFile", disc], input=sam, check=True, stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, env=env).stdout
        res[tag] = (out, open(spl, "rb").read(), open(disc, "rb").read())
    for i, what in enumerate(("main SAM", "splitters", "discordants")):
        assert _strip_pg(res["oracle"][i]) == _strip_pg(res["shim"][i]), what
    assert b"\t77\t" in res["shim"][0] and b"\t77\t" not in res["shim"][2]          # unmapped pairs pass through, never discordant
    assert res["shim"][2].count(b"\n") > 50 and res["shim"][1].count(b"\n") > 20    # both side streams are exercised
    assert any((int(l.split(b"\t")[1]) & 0x400) for l in res["shim"][0].splitlines() if not l.startswith(b"@"))


def test_fused_stream_is_only_routed(shim, tmp_path):
    """fused mode (speedseq_b200/cli/ssq_fuse.h): behind the marker line the shim copies frames to stdout / --splitterFile /
    --discordantFile and never touches the GPU; a marker written under other options is refused"""
    import struct
    hdr = b"@SQ\tSN:c1\tLN:5000\n@PG\tID:bwa\tPN:bwa\n"
    marker = b"@CO\tssq-fused-v1\texcludeDups=1 addMateTags=1 removeDups=0 maxSplitCount=2 minNonOverlap=20 minIndelSize=50 maxUnmappedBases=50\n"
    frame = lambda k, p: b"SSQFRAME" + struct.pack("<QQ", k, len(p)) + p
    main1, main2, spl, disc = b"a\t99\tc1\t1\n" * 3, b"b\t1171\tc1\t9\n", b"a_1\t65\tc1\t1\n" * 2, b"d\t65\tc1\t1\nd\t129\tc1\t7\n"
    data = hdr + marker + frame(0, main1) + frame(1, spl) + frame(0, main2) + frame(2, disc)
    s, d = str(tmp_path / "s"), str(tmp_path / "d")
    args = ["--excludeDups", "--addMateTags", "--maxSplitCount", "2", "--minNonOverlap", "20", "--splitterFile", s, "--discordantFile", d]
    out = subprocess.run([shim] + args, input=data, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    body = lambda b: b"".join(l for l in b.splitlines(True) if not l.startswith(b"@PG\tID:SAMBLASTER"))
    assert body(out) == hdr + main1 + main2 and b"ssq-fused" not in out
    assert body(open(s, "rb").read()) == hdr + spl and body(open(d, "rb").read()) == hdr + disc
    assert out.count(b"@PG\tID:SAMBLASTER") == 1
    bad = subprocess.run([shim, "--addMateTags", "--splitterFile", s, "--discordantFile", d], input=data, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert bad.returncode != 0 and b"other options" in bad.stderr
