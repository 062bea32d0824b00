"""Checks on the oracle that need no bwa: alignments re-scored from their CIGAR, NM/MD consistency, SMEM maximality and
occurrence counts against brute force, samblaster rules against an independent Python restatement, self-golden hashes."""
import gzip
import hashlib
import json
import os
import re
import subprocess

import numpy as np

import ssq_testlib as T


def _ref_seq(fa):
    return "".join(l.strip() for l in open(fa) if not l.startswith(">")).upper()


def _sam(oracle, idx, reads, n=None):
    names, seqs, quals = reads
    n = n or len(names)
    return oracle.mem_pe(idx, names[:n], seqs[:n], quals[:n], 0, 4, b"NA12878").splitlines()


def test_sam_records_are_self_consistent(oracle, ex_index, ex_reads):
    idx = oracle.load(ex_index)
    ref = _ref_seq(ex_index)
    lines = _sam(oracle, idx, ex_reads, 2000)
    assert len(lines) >= 2000
    n_checked = 0
    for l in lines:
        f = l.split("\t")
        flag = int(f[1])
        if flag & 4 or f[5] == "*" or f[9] == "*":
            continue
        tags = {t[:2]: t[5:] for t in f[11:]}
        pos = int(f[3]) - 1
        q = r = sc = nm = 0
        md = []
        run = 0
        for ln, op in re.findall(r"(\d+)([MIDSH])", f[5]):
            ln = int(ln)
            if op == "M":
                for i in range(ln):
                    a, b = f[9][q + i], ref[pos + r + i]
                    if a == b and a != "N":
                        sc += 1; run += 1
                    else:
                        sc -= 1 if "N" in (a, b) else 4
                        nm += 1; md.append(str(run) + b); run = 0
                q += ln; r += ln
            elif op == "I":
                sc -= 6 + ln; nm += ln; q += ln
            elif op == "D":
                sc -= 6 + ln; nm += ln
                md.append(str(run) + "^" + ref[pos + r:pos + r + ln]); run = 0
                r += ln
            elif op == "S":
                q += ln
        md.append(str(run))
        assert q == len(f[9])
        assert nm == int(tags["NM"]), l
        assert "".join(md) == tags["MD"], l
        # AS is the best LOCAL extension score; when an end is extended to the read end despite a better clipped score the
        # CIGAR's score is lower by less than the clipping penalty (5) per end
        assert 0 <= int(tags["AS"]) - sc <= 8, l
        assert 0 <= int(f[4]) <= 60
        n_checked += 1
    assert n_checked > 1900


def test_smems_are_maximal_and_counts_match_bruteforce(oracle, ex_index, ex_reads):
    idx = oracle.load(ex_index)
    ref = _ref_seq(ex_index)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    both = ref + "".join(comp[c] for c in reversed(ref))
    seqs = ex_reads[1][:40]
    seq, off = T.encode_reads(seqs)
    iv, ioff = oracle.smem_batch(idx, seq, off)
    for r, s in enumerate(seqs):
        for v in iv[int(ioff[r]):int(ioff[r + 1])]:
            sub = s[int(v["qbeg"]):int(v["qend"])]
            assert len(sub) >= 19 and "N" not in sub
            occ = len(re.findall("(?=" + sub + ")", both))
            assert occ == int(v["s"]), (r, sub)


def test_extend_matches_plain_dp_when_nothing_is_trimmed(oracle):
    """with a huge h0 no cell ever dies, so band trimming/z-drop never act and ksw_extend2 must equal the plain recurrences
    H=max(M,E,F), E'=max(E-e,M-oe), F'=max(F-e,M-oe) evaluated over the full rectangle"""
    rng = np.random.default_rng(11)
    tasks, q, t = T.extension_tasks(rng, 150, qmax=40)
    tasks["h0"] = 3000; tasks["w"] = 1000; tasks["zdrop"] = 0; tasks["end_bonus"] = 5
    res = oracle.sw_extend_batch(tasks, q, t)
    for k, tk in enumerate(tasks):
        qs = q[int(tk["q_off"]):int(tk["q_off"]) + int(tk["qlen"])].astype(int)
        ts = t[int(tk["t_off"]):int(tk["t_off"]) + int(tk["tlen"])].astype(int)
        ql, tl, h0 = len(qs), len(ts), 3000
        Hprev = [h0] + [h0 - 7 - j for j in range(ql)]  # H(-1,-1), H(-1,0..ql-1)
        E = [0] * ql
        best, bi, bj, gs, gi = h0, -1, -1, -1, -1
        for i in range(tl):
            Hcur = [h0 - (6 + (i + 1))]  # H(i,-1)
            f = 0
            m, mj = 0, -1
            for j in range(ql):
                s = -1 if qs[j] > 3 or ts[i] > 3 else (1 if qs[j] == ts[i] else -4)
                M = Hprev[j] + s
                h = max(M, E[j], f)
                if h >= m:
                    m, mj = h, j
                E[j] = max(E[j] - 1, M - 7, 0)
                f = max(f - 1, M - 7, 0)
                Hcur.append(h)
            if Hcur[ql] >= gs:
                gs, gi = Hcur[ql], i
            if m > best:
                best, bi, bj = m, i, mj
            Hprev = Hcur
        r = res[k]
        assert (int(r["score"]), int(r["qle"]), int(r["tle"]), int(r["gscore"]), int(r["gtle"])) == (best, bj + 1, bi + 1, gs, gi + 1), k


def test_samblaster_against_python_restatement(oracle, ex_index, ex_reads, tmp_path):
    idx = oracle.load(ex_index)
    names, seqs, quals = ex_reads
    # duplicate the first 300 pairs under new names so that duplicates certainly exist
    names = names + ["dup_" + n for n in names[:600]]
    seqs = seqs + seqs[:600]
    quals = quals + quals[:600]
    body = oracle.mem_pe(idx, names, seqs, quals, 0, 4, b"")
    hdr = "@SQ\tSN:20_slice\tLN:321635\n"
    spl, disc = str(tmp_path / "spl.sam"), str(tmp_path / "disc.sam")
    out = subprocess.run([T.ORACLE_BIN, "samblaster", "--addMateTags", "--maxSplitCount", "2", "--minNonOverlap", "20", "--splitterFile", spl, "--discordantFile", disc],
                         input=(hdr + body).encode(), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout.decode()
    recs = [l.split("\t") for l in out.splitlines() if not l.startswith("@")]
    # independent restatement of the signature rule
    seen, expect_dup = set(), {}
    blocks = {}
    order = []
    for f in recs:
        if f[0] not in blocks:
            blocks[f[0]] = []; order.append(f[0])
        blocks[f[0]].append(f)

    def five_prime(f):
        flag = int(f[1]); ops = re.findall(r"(\d+)([MIDNSHP=X])", f[5])
        lead = 0; i = 0
        while i < len(ops) and ops[i][1] in "SH":
            lead += int(ops[i][0]); i += 1
        trail = 0; j = len(ops) - 1
        while j >= 0 and ops[j][1] in "SH":
            trail += int(ops[j][0]); j -= 1
        ral = sum(int(n) for n, o in ops if o in "MDN=X")
        return (f[2], int(f[3]) - lead, 0) if not flag & 16 else (f[2], int(f[3]) + ral + trail - 1, 1)
    for name in order:
        prim = [f for f in blocks[name] if not int(f[1]) & 0x900]
        r1 = [f for f in prim if int(f[1]) & 0x40][0]; r2 = [f for f in prim if int(f[1]) & 0x80][0]
        u1, u2 = int(r1[1]) & 4, int(r2[1]) & 4
        if u1 and u2:
            expect_dup[name] = False; continue
        if u1 or u2:
            m = r2 if u1 else r1
            key = ("orphan", five_prime(m))
        else:
            a, b = five_prime(r1), five_prime(r2)
            key = tuple(sorted([(a[1], a[0], a[2]), (b[1], b[0], b[2])]))
        expect_dup[name] = key in seen
        seen.add(key)
    n_dup = 0
    for name in order:
        flags = {bool(int(f[1]) & 0x400) for f in blocks[name]}
        assert flags == {expect_dup[name]}, name
        n_dup += expect_dup[name]
    assert n_dup >= 290
    # every paired line carries MC/MQ of its mate's primary line
    for name in order[:200]:
        prim = {bool(int(f[1]) & 0x40): f for f in blocks[name] if not int(f[1]) & 0x900}
        for f in blocks[name]:
            mate = prim[not bool(int(f[1]) & 0x40)]
            tags = {t[:2]: t[5:] for t in f[11:]}
            assert tags["MC"] == mate[5] and tags["MQ"] == mate[4]
    # discordant file: exactly the primary lines of pairs with both ends mapped and 0x2 clear
    dn = {l.split("\t")[0] for l in open(disc) if not l.startswith("@")}
    exp = set()
    for name in order:
        prim = [f for f in blocks[name] if not int(f[1]) & 0x900]
        if all(not int(f[1]) & 4 for f in prim) and not int(prim[0][1]) & 2:
            exp.add(name)
    assert dn == exp
    # splitter names carry _1/_2
    for l in open(spl):
        if not l.startswith("@"):
            assert re.search(r"_[12]$", l.split("\t")[0])


def test_self_golden_hashes(oracle, ex_index, tmp_path):
    """oracle output on the committed 2k-pair fixture is frozen (guards against accidental change; not a bwa known-answer)"""
    gold = json.load(open(os.path.join(T.GOLDEN, "ex_reads_2k.self.json")))
    fq = os.path.join(T.GOLDEN, "ex_reads_2k.fq.gz")
    sam = subprocess.run([T.ORACLE_BIN, "mem", "-t", "3", "-p", "-R", r"@RG\tID:NA12878\tSM:NA12878\tLB:lib1", ex_index, fq], check=True, stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL).stdout
    body = b"".join(l for l in sam.splitlines(True) if not l.startswith(b"@PG"))
    assert hashlib.sha256(body).hexdigest() == gold["bwa_mem_sam_sha256_without_PG"]
    spl, disc = str(tmp_path / "s.sam"), str(tmp_path / "d.sam")
    sb = subprocess.run([T.ORACLE_BIN, "samblaster", "--excludeDups", "--addMateTags", "--maxSplitCount", "2", "--minNonOverlap", "20", "--splitterFile", spl,
                         "--discordantFile", disc], input=sam, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    strip = lambda b: b"".join(l for l in b.splitlines(True) if not l.startswith(b"@PG"))
    assert hashlib.sha256(strip(sb)).hexdigest() == gold["samblaster_sam_sha256_without_PG"]
    assert hashlib.sha256(strip(open(spl, "rb").read())).hexdigest() == gold["splitters_sha256_without_PG"]
    assert hashlib.sha256(strip(open(disc, "rb").read())).hexdigest() == gold["discordants_sha256_without_PG"]
