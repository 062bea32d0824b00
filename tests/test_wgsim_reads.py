"""Reads simulated by the REFERENCE'S OWN simulator (samtools-1.3.1/misc/wgsim.c, compiled from the reference tree by
tests/golden/make_wgsim_reads.py; fixtures tests/golden/wgsim_r{1,2}.fq.gz) through the two-file mode of `bwa mem` (speedseq:468): the names
carry the true fragment coordinates, so mapping accuracy needs no bwa to be scored.  CPU: the oracle CLI.  GPU: the product's
`bwa` executable, which must also write exactly the oracle's bytes."""
import gzip
import os
import subprocess

import pytest

import ssq_testlib as T

BWA = os.path.join(T.ROOT, "speedseq_b200", "bin", "bwa")
R1, R2 = os.path.join(T.GOLDEN, "wgsim_r1.fq.gz"), os.path.join(T.GOLDEN, "wgsim_r2.fq.gz")


def _score(sam):
    """fraction of primary records of confidently placed reads (MAPQ >= 20) lying inside the fragment wgsim drew"""
    ok = n = 0
    for l in sam.decode().splitlines():
        if l.startswith("@"):
            continue
        f = l.split("\t")
        flag = int(f[1])
        if flag & 0x904 or int(f[4]) < 20:
            continue
        p = f[0].split("_")
        start, end = int(p[2]), int(p[3])  # 20_slice_<start>_<end>_...
        n += 1
        ok += start - 10 <= int(f[3]) <= end + 10
    return ok, n


def test_oracle_places_wgsim_reads(oracle, ex_index):
    sam = subprocess.run([T.ORACLE_BIN, "mem", "-t", "4", ex_index, R1, R2], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    ok, n = _score(sam)
    assert n > 2700 and ok >= 0.99 * n, (ok, n)


@pytest.mark.gpu
def test_product_places_wgsim_reads_and_matches_oracle(ssq, oracle, ex_index):
    rec = lambda b: b"".join(l for l in b.splitlines(True) if not l.startswith(b"@"))
    a = subprocess.run([T.ORACLE_BIN, "mem", "-t", "4", ex_index, R1, R2], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    b = subprocess.run([BWA, "mem", "-t", "4", ex_index, R1, R2], check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=300).stdout
    assert rec(a) == rec(b)
    ok, n = _score(b)
    assert n > 2700 and ok >= 0.99 * n, (ok, n)
