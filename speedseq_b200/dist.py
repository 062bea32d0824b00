"""Multi-GPU duplicate marking: the one real exchange on the `speedseq align` path.

Alignment shards by read batch with the index replicated, so it needs no collective.  "First pair seen with a signature is
kept" is global, though: every rank routes its pair signatures to an owner rank chosen by hash(signature) with one
all-to-all (over NCCL / NVLink when the tensors are CUDA tensors), the owner orders what it received by GLOBAL pair ordinal
and runs the sort-and-mark kernel of libssq on device pointers (ssq_dupmark_keys_dev), and a second all-to-all returns one
duplicate bit per pair to the rank that aligned it.  24 bytes per pair go out, 1 byte comes back.

`exchange_and_mark` is backend-agnostic plumbing (tested with gloo on CPU in tests/test_dist_gloo.py with the oracle as the
marking function); `mark_cuda` is the product's marking function and refuses to run without libssq and a GPU."""
import ctypes as C

import torch
import torch.distributed as dist

_MIX = 0x9E3779B97F4A7C15


def _owner(key1, key2, world):
    # torch has no uint64 arithmetic: work on the int64 bit patterns (wrap-around multiplication is what we want)
    h = key1 * torch.tensor(_MIX - (1 << 64), dtype=torch.int64, device=key1.device) ^ (key2 * 0x2545F491) ^ (key2 >> 29)
    h = h ^ (h >> 32)
    return (h & 0x7FFFFFFF) % world


def exchange_and_mark(key1, key2, valid, ordinal, mark_fn, group=None):
    """key1/key2/ordinal: int64 tensors (bit patterns of u64), valid: uint8; returns uint8 is_dup aligned with the inputs.
    mark_fn(key1, key2, valid) -> uint8 tensor marks later occurrences in ARRAY ORDER among what one owner received."""
    world = dist.get_world_size(group)
    dev = key1.device
    n = key1.numel()
    owner = _owner(key1, key2, world)
    order = torch.argsort(owner, stable=True)
    send_counts = torch.bincount(owner, minlength=world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc, rcv = send_counts.tolist(), recv_counts.tolist()
    payload = torch.stack([key1[order], key2[order], ordinal[order], valid[order].to(torch.int64)], dim=1).contiguous()
    got = torch.empty((sum(rcv), 4), dtype=torch.int64, device=dev)
    dist.all_to_all_single(got, payload, output_split_sizes=rcv, input_split_sizes=sc, group=group)
    # owner side: first-seen order is the global ordinal
    by_ord = torch.argsort(got[:, 2], stable=True)
    g = got[by_ord]
    d_sorted = mark_fn(g[:, 0].contiguous(), g[:, 1].contiguous(), g[:, 3].to(torch.uint8).contiguous())
    d_recv = torch.empty_like(d_sorted)
    d_recv[by_ord] = d_sorted
    back = torch.empty(n, dtype=torch.uint8, device=dev)
    dist.all_to_all_single(back, d_recv, output_split_sizes=sc, input_split_sizes=rcv, group=group)
    out = torch.empty(n, dtype=torch.uint8, device=dev)
    out[order] = back
    return out


def mark_cuda(lib, device):
    """marking function over CUDA tensors backed by libssq's device-pointer entry point"""
    def fn(k1, k2, valid):
        if not k1.is_cuda:
            raise RuntimeError("mark_cuda needs CUDA tensors: there is no CPU path in libssq")
        out = torch.empty(k1.numel(), dtype=torch.uint8, device=k1.device)
        st = torch.cuda.current_stream(k1.device).cuda_stream
        rc = lib.ssq_dupmark_keys_dev(C.c_int(device), C.c_uint64(k1.numel()), C.c_void_p(k1.data_ptr()), C.c_void_p(k2.data_ptr()), C.c_void_p(valid.data_ptr()),
                                      C.c_void_p(out.data_ptr()), C.c_void_p(st))
        if rc != 0:
            lib.ssq_last_error.restype = C.c_char_p
            raise RuntimeError("ssq_dupmark_keys_dev failed: %d %s" % (rc, lib.ssq_last_error().decode()))
        return out
    return fn
