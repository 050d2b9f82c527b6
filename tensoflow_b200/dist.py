"""Ray-sharded data parallelism: one process per GPU, identical replicas, each
rank renders its slice of the ray batch, then ONE sum-allreduce of a flat fp32 gradient bucket
(VM planes + lines, MLP weights, variance, env cubemap) over NCCL / NVLink, plus a 2-float
(sum, count) allreduce for losses that are means over a rank-dependent number of samples.
The reference has no distributed code (train/trainer_inv.py:29,188 only skips `.cuda()`).
Works with any torch.distributed backend (tests use gloo on CPU)."""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_slice(n: int, rank: int, world: int) -> slice:
    """Contiguous, near-equal slice of n rays for this rank."""
    per = (n + world - 1) // world
    return slice(min(rank * per, n), min((rank + 1) * per, n))


def _flat_view(g: torch.Tensor) -> torch.Tensor:
    """1-D view of a gradient in its MEMORY order (channels-last factor grads stay views)."""
    if g.is_contiguous():
        return g.reshape(-1)
    if g.dim() == 4 and g.permute(0, 2, 3, 1).is_contiguous():
        return g.permute(0, 2, 3, 1).reshape(-1)
    return None


class FlatGradBucket:
    """One flat fp32 buffer holding every parameter gradient: a single collective per step."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.numel = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)

    def pack(self):
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self.flat[off:off + n].zero_()
            else:
                v = _flat_view(p.grad)
                self.flat[off:off + n].copy_(v if v is not None else p.grad.contiguous().reshape(-1))
            off += n

    def unpack(self):
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is not None:
                v = _flat_view(p.grad)
                if v is not None:
                    v.copy_(self.flat[off:off + n])
                else:
                    p.grad.copy_(self.flat[off:off + n].reshape(p.grad.shape))
            off += n

    def allreduce(self, average: bool = False, async_op: bool = False):
        """Sum the bucket over all ranks (NCCL picks NVLS / ring on the NVSwitch domain)."""
        self.pack()
        work = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
            if average and not async_op:
                self.flat.div_(dist.get_world_size())
        if not async_op:
            self.unpack()
        return work


def global_mean(local_sum: torch.Tensor, local_count: torch.Tensor) -> torch.Tensor:
    """Mean over ALL ranks' samples of a quantity whose per-rank sample count differs
    (loss_sparse / loss_hessian / eikonal: reference shapeRenderer.py:1152-1162)."""
    t = torch.stack([local_sum.reshape(()).float(), local_count.reshape(()).float()])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t[0] / t[1].clamp_min(1.0)


def gather_tiles(local: torch.Tensor) -> torch.Tensor:
    """Inference (config 5): concatenate per-rank image tiles [rays_r, C] in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n = torch.tensor([local.shape[0]], device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    m = int(max(s.item() for s in sizes))
    pad = torch.zeros(m, *local.shape[1:], device=local.device, dtype=local.dtype)
    pad[:local.shape[0]] = local
    outs = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    return torch.cat([o[:int(s.item())] for o, s in zip(outs, sizes)], 0)


def interleaved_ids(n: int, rank: int, world: int, device=None) -> torch.Tensor:
    """Pixel ids rank, rank+world, rank+2*world, ... : the inference split (config 5).  Object pixels cluster in the image, so
    contiguous tiles leave the ranks that own background rows idle; a strided split gives every rank the same mix."""
    return torch.arange(rank, max(n, rank), world, device=device)


def gather_interleaved(local: torch.Tensor, n: int) -> torch.Tensor:
    """All-gather per-rank results of `interleaved_ids` and put them back in pixel order -> [n, C] on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    full = gather_tiles(local)
    order = torch.cat([interleaved_ids(n, r, world, local.device) for r in range(world)])
    out = torch.empty_like(full)
    out[order] = full
    return out
