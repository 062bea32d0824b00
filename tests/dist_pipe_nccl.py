"""run under torchrun on N GPUs (tests/test_gpu_dist.py launches it when the box has >= 2): the fused `bwa mem | samblaster` pipeline with
batches dealt round-robin to the ranks and the duplicate stage going through libssq's NCCL exchange (csrc/ssq_dist.cu).  The three
streams, re-assembled in batch order, must equal the oracle's single-process `bwa mem | samblaster` over the whole input — i.e. the
multi-GPU run marks exactly the duplicates a single GPU (and the CPU) would."""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import ssq_testlib as T
from test_hostsim_pipe import stress_reads, sq_header

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s = T.SSQ()
L = s.lib
SB = dict(exclude_dups=1, add_mate_tags=1, max_split_count=2, min_non_overlap=20)
SB_ARGS = ["--excludeDups", "--addMateTags", "--maxSplitCount", "2", "--minNonOverlap", "20"]
wd = os.environ.get("SSQ_DIST_DIR") or tempfile.gettempdir()
fa = os.path.join(wd, "dist_ref.fa")
g, bounds = T.synth_genome(400000, 7, n_contigs=3)
if rank == 0:
    T.write_fasta(fa, g, bounds)
    s.index_build(fa, None, local)
dist.barrier()
h = s.index_load(fa, local)
# the communicator: rank 0 makes the id, torch broadcasts it
idb = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    buf = (C.c_uint8 * 128)()
    s.ck(L.ssq_comm_unique_id(buf), "ssq_comm_unique_id")
    idb = torch.tensor(list(buf), dtype=torch.uint8)
idb = idb.cuda()
dist.broadcast(idb, 0)
idbytes = bytes(idb.cpu().tolist())
comm = C.c_void_p()
s.ck(L.ssq_comm_create(idbytes, C.c_int(rank), C.c_int(world), C.c_int(local), C.byref(comm)), "ssq_comm_create")
al = s.aligner_create(h, SB, b"d")
s.ck(L.ssq_aligner_set_comm(al, comm), "ssq_aligner_set_comm")
# every rank generates the same reads and works on its share of the batches: round k -> batch k*world + rank
names, seqs, quals = stress_reads(g, bounds, 3000, 150, 5, err=0.01)
n_batches = 3 * world
per = (len(names) // n_batches) & ~1
cuts = [i * per for i in range(n_batches)] + [len(names)]
mine = []
for k in range(3):
    b = k * world + rank
    lo, hi = cuts[b], cuts[b + 1]
    if k == 1 and rank == world - 1:
        lo = hi  # one empty batch in the middle: the round must still complete
    rd, keep = T.pack_reads(names[lo:hi], seqs[lo:hi], quals[lo:hi], None, 1, lo)
    txt, info = s.aligner_run(al, rd)
    mine.append((b, txt))
L.ssq_comm_counter.restype = C.c_uint64
L.ssq_comm_counter.argtypes = [C.c_void_p, C.c_int]
traffic = (int(L.ssq_comm_counter(comm, 0)), int(L.ssq_comm_counter(comm, 1)), int(L.ssq_comm_counter(comm, 2)))
gathered = [None] * world
dist.all_gather_object(gathered, (mine, traffic))
ok = 1
if rank == 0:
    got = {b: t for m, _ in gathered for b, t in m}
    streams = [b"".join(got[b][i] for b in range(n_batches)) for i in range(3)]
    o = T.Oracle()
    oidx = o.load(fa)
    body = ""
    for b in range(n_batches):
        lo, hi = cuts[b], cuts[b + 1]
        if b == 1 * world + world - 1:
            continue  # the batch the last rank skipped
        body += o.mem_pe(oidx, names[lo:hi], seqs[lo:hi], quals[lo:hi], lo, 8, b"d")
    spl, disc = os.path.join(wd, "o.spl"), os.path.join(wd, "o.disc")
    out = subprocess.run([T.ORACLE_BIN, "samblaster"] + SB_ARGS + ["--splitterFile", spl, "--discordantFile", disc], input=(sq_header(o, oidx, fa) + body).encode(), check=True,
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout
    rec = lambda t: b"".join(l for l in t.splitlines(True) if not l.startswith(b"@"))
    same = [streams[0] == rec(out), streams[1] == rec(open(spl, "rb").read()), streams[2] == rec(open(disc, "rb").read())]
    n_dup = sum(1 for l in streams[0].splitlines() if int(l.split(b"\t")[1]) & 0x400)
    print("dist_pipe world=%d batches=%d identical_to_oracle=%s dup_lines=%d exchange(bytes out, back, rounds) per rank=%s" % (world, n_batches, same, n_dup, [t for _, t in gathered]))
    ok = int(all(same) and n_dup > 100)
okt = torch.tensor([ok], device="cuda")
dist.broadcast(okt, 0)
s.aligner_free(al)
L.ssq_comm_free(comm)
s.index_free(h)
dist.destroy_process_group()
sys.exit(0 if int(okt) else 1)
