"""world_size-2 gloo test (CPU) of the multi-GPU dup-mark exchange: routing by signature hash, ordering by global ordinal on
the owner, return trip.  The marking function is the oracle here; on GPUs it is libssq (tests/dist_dupmark_nccl.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ssq_testlib as T


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, T.ROOT)
    from speedseq_b200.dist import exchange_and_mark
    o = T.Oracle()
    rng = np.random.default_rng(5)  # same stream on every rank: the global signature list
    sig = np.zeros(n, T.DUPSIG_DT)
    sig["pos1"] = rng.integers(0, 800, n); sig["pos2"] = sig["pos1"] + rng.integers(0, 30, n)
    sig["strand1"] = rng.integers(0, 2, n); sig["strand2"] = rng.integers(0, 2, n); sig["valid"] = rng.random(n) > 0.05
    # batches of 1000 pairs dealt round-robin to the ranks (whole batches stay together, like bwa's read batches)
    mine = np.nonzero((np.arange(n) // 1000) % world == rank)[0]

    def mark(k1, k2, valid):
        s = np.zeros(k1.numel(), T.DUPSIG_DT)
        s["pos1"] = k1.numpy().view(np.uint64) >> 1; s["strand1"] = k1.numpy() & 1
        s["pos2"] = k2.numpy().view(np.uint64) >> 1; s["strand2"] = k2.numpy() & 1
        s["valid"] = valid.numpy()
        return torch.from_numpy(o.dupmark(s))
    k1 = torch.from_numpy(((sig["pos1"][mine] << 1) | sig["strand1"][mine]).astype(np.int64))
    k2 = torch.from_numpy(((sig["pos2"][mine] << 1) | sig["strand2"][mine]).astype(np.int64))
    got = exchange_and_mark(k1, k2, torch.from_numpy(sig["valid"][mine].copy()), torch.from_numpy(mine.astype(np.int64)), mark)
    ref = o.dupmark(sig)[mine]
    q.put((rank, bool(np.array_equal(got.numpy(), ref)), int(ref.sum())))
    dist.destroy_process_group()


def test_exchange_and_mark_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    ps = [ctx.Process(target=_worker, args=(r, 2, port, 20000, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    assert all(ok for _, ok, _ in res), res
    assert sum(nd for _, _, nd in res) > 1000
