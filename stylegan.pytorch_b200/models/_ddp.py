"""Data-parallel gradient exchange: one process per GPU; after each backward the gradients of one optimizer are packed
into flat fp32 buckets (one bucket per optimizer at the default size: ~92 MB for D, ~100 MB for G at 256x256) and
SUM-all-reduced over NCCL (NVLink 5 / NVSwitch); every `.grad` is then re-pointed into its bucket, so the multi-tensor
Adam kernel consumes the reduced values in place with the 1/world average folded into its gradient scale (and into the
clip norm).  The exchange is NOT overlapped with the backward pass: it starts when the backward has finished and the
optimizer waits for it (at 8 GPUs ~0.3 ms of NVLink time per optimizer against a ~20 ms step half; what the scaling
runs lose is mostly the max-over-ranks of the per-step time, which the CUDA-graph replay of the step removes).
Inside a captured step (models/_graphs.py) the all-reduce is a node of the graph.

The reference has no multi-GPU code at all (`# TODO` at models/GAN.py:509-510); semantics follow SURVEY.md 8(e):
identical replicas, per-rank local batch, gradients averaged, rank-0 broadcast of parameters and buffers at start,
Truncation.avg_latent averaged over the ranks after each update.  Works with the gloo backend on CPU tensors for the
host-logic tests."""
import os

import torch
import torch.distributed as dist


def bucketize(sizes, bucket_elems):
    """Greedy contiguous bucketing of tensor sizes -> list of lists of indices."""
    buckets, cur, acc = [], [], 0
    for i, n in enumerate(sizes):
        if cur and acc + n > bucket_elems:
            buckets.append(cur)
            cur, acc = [], 0
        cur.append(i)
        acc += n
    if cur:
        buckets.append(cur)
    return buckets


class DataParallelContext:
    def __init__(self, device, bucket_elems=64 * 1024 * 1024):
        self.device = device
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.bucket_elems = bucket_elems        # one NCCL call per optimizer: large messages ride NVLS / the ring at full rate

    @classmethod
    def from_env(cls, device):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return cls(device)
        return None

    def broadcast_module(self, module, src=0):
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src)

    def average_(self, t):
        """In-place mean of a small tensor over the ranks (Truncation.avg_latent after each update)."""
        if dist.get_backend() == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:                                   # gloo (CPU host-logic tests) has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.mul_(1.0 / self.world)
        return t

    def reduce_buckets(self, grads):
        """grads: list of tensors (same device).  Returns (flats, works, placement): flats are the flat buckets being
        SUM-reduced (asynchronously; wait on `works`), placement[i] = (bucket_index, offset)."""
        sizes = [g.numel() for g in grads]
        plan = bucketize(sizes, self.bucket_elems)
        flats, works, placement = [], [], [None] * len(grads)
        for b, idxs in enumerate(plan):
            off = 0
            for i in idxs:
                placement[i] = (b, off)
                off += sizes[i]
            flat = torch.cat([grads[i].reshape(-1) for i in idxs])
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))
            flats.append(flat)
        return flats, works, placement

    def allreduce_grads(self, optim):
        """SUM-reduce the gradients of `optim`'s parameters across ranks into flat buckets and re-point each `.grad`
        into its bucket.  Returns the scale (1/world) that turns the sums into averages; the fused Adam kernel applies
        it (to the clip norm as well), so no extra pass over the gradients is made."""
        params = [p for g in optim.param_groups for p in g["params"] if p.grad is not None]
        if not params:
            return 1.0
        grads = [p.grad for p in params]
        flats, works, placement = self.reduce_buckets(grads)
        for w in works:
            w.wait()
        for p, (b, off) in zip(params, placement):
            p.grad = flats[b][off:off + p.numel()].view_as(p)
        return 1.0 / self.world


def sharded_loader(dataset, batch_size, num_workers, rank, world, seed=0):
    """Per-rank loader for data-parallel training (rank r draws its own local batch): the reference's
    `data.get_data_loader` settings (`data/__init__.py:32-52`: shuffle, drop_last, pin_memory) with the shuffle moved
    into a DistributedSampler so that the ranks see disjoint, equally long shards of every epoch.  Call
    `loader.sampler.set_epoch(e)` before each epoch (StyleGAN.train does)."""
    from torch.utils.data import DataLoader
    from torch.utils.data.distributed import DistributedSampler
    sampler = DistributedSampler(dataset, num_replicas=world, rank=rank, shuffle=True, seed=seed, drop_last=True)
    return DataLoader(dataset, batch_size=batch_size, sampler=sampler, num_workers=num_workers, drop_last=True,
                      pin_memory=True)
