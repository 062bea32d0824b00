"""run under torchrun on N GPUs: NCCL all-to-all + libssq device marking vs the oracle (not collected by pytest)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import ssq_testlib as T
from speedseq_b200.dist import exchange_and_mark, mark_cuda

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 4_000_000
rng = np.random.default_rng(7)
sig = np.zeros(n, T.DUPSIG_DT)
sig["pos1"] = rng.integers(0, 3_000_000, n); sig["pos2"] = sig["pos1"] + rng.integers(0, 600, n)
sig["strand1"] = rng.integers(0, 2, n); sig["strand2"] = rng.integers(0, 2, n); sig["valid"] = rng.random(n) > 0.02
sig["pos1"][::10] = sig["pos1"][1::10][: len(sig["pos1"][::10])]; sig["pos2"][::10] = sig["pos2"][1::10][: len(sig["pos2"][::10])]
sig["strand1"][::10] = sig["strand1"][1::10][: len(sig["strand1"][::10])]; sig["strand2"][::10] = sig["strand2"][1::10][: len(sig["strand2"][::10])]
mine = np.nonzero((np.arange(n) // 50000) % world == rank)[0]
s = T.SSQ()
dev = torch.device("cuda", local)
k1 = torch.from_numpy(((sig["pos1"][mine] << 1) | sig["strand1"][mine]).astype(np.int64)).to(dev)
k2 = torch.from_numpy(((sig["pos2"][mine] << 1) | sig["strand2"][mine]).astype(np.int64)).to(dev)
va = torch.from_numpy(sig["valid"][mine].copy()).to(dev)
od = torch.from_numpy(mine.astype(np.int64)).to(dev)
fn = mark_cuda(s.lib, local)
got = exchange_and_mark(k1, k2, va, od, fn)  # warm-up
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
got = exchange_and_mark(k1, k2, va, od, fn)
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ref = T.Oracle().dupmark(sig)[mine]
ok = torch.tensor([int(np.array_equal(got.cpu().numpy(), ref))], device=dev)
dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print("dist_dupmark world=%d pairs=%d identical_to_oracle=%d dups=%d  %.2f ms -> %.1f M pairs/s" % (world, n, int(ok), int(ref.sum()), float(ms), n / float(ms) / 1e3))
dist.destroy_process_group()
sys.exit(0 if int(ok) else 1)
