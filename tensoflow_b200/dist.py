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


def _dense(t: torch.Tensor) -> bool:
    """True when the tensor covers numel() consecutive elements in SOME dimension order (contiguous, channels-last, ...)."""
    expect = 1
    for size, stride in sorted(((s, st) for s, st in zip(t.shape, t.stride()) if s > 1), key=lambda x: x[1]):
        if stride != expect:
            return False
        expect *= size
    return True


def _aliases(a: torch.Tensor, b: torch.Tensor) -> bool:
    """Same elements at the same addresses (strides of size-1 dimensions are irrelevant)."""
    return a.data_ptr() == b.data_ptr() and a.shape == b.shape and \
        all(sa == sb for n, sa, sb in zip(a.shape, a.stride(), b.stride()) if n > 1)


class FlatGradBucket:
    """One flat fp32 buffer that IS the gradient storage of every parameter: `views[i]` has parameter i's shape and strides
    (channels-last factors included) and lives inside `flat`, so the collective runs in place and there is no pack / unpack
    copy on the common path:

        bucket.begin_step()      # flat.zero_(), p.grad = None, the kernels' gradient buffers are handed out from `flat`
        loss.backward()          # tensoflow_b200.ops backward functions scatter straight into the views; autograd adopts them
        bucket.allreduce(async_op=True)   # in-place sum over ranks on the backend's own stream (NCCL: NVLS / ring)
        ...                      # anything independent of the gradients overlaps the collective
        bucket.finish()          # wait, p.grad = view for EVERY parameter (also those this rank did not touch)

    A gradient that did not land in its view (a parameter fed by ordinary PyTorch ops, or by two backward calls in one
    step) is copied into it before the collective, so the result is the same either way."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        # every segment starts on a 256-byte boundary: the kernels scatter into the views with 16-byte vector reductions
        align = 64
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += (p.numel() + align - 1) // align * align
        self.numel = off
        dev = self.params[0].device
        self.flat = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.views: List[torch.Tensor] = []
        self._by_ptr = {}
        for i, (p, off) in enumerate(zip(self.params, offs)):
            n = p.numel()
            seg = self.flat[off:off + n]
            self.views.append(seg.as_strided(p.shape, p.stride()) if _dense(p) else seg.view(p.shape))
            self._by_ptr[p.data_ptr()] = i
        self._handed = set()
        self._work = None
        self._average = False

    # ---- gradient buffers for the kernels ------------------------------------------------------------
    def _alloc(self, like: torch.Tensor):
        """Zeroed gradient buffer for the parameter whose storage `like` aliases: its view, once per step."""
        i = self._by_ptr.get(like.data_ptr())
        if i is None or i in self._handed or like.numel() != self.params[i].numel():
            return None
        self._handed.add(i)
        return self.views[i]

    def begin_step(self):
        from . import ops
        self.flat.zero_()
        self._handed.clear()
        for p in self.params:
            p.grad = None
        ops.set_grad_allocator(self._alloc)

    # ---- the collective ----------------------------------------------------------------------------------
    def allreduce(self, average: bool = False, async_op: bool = False):
        """Sum (or average) every gradient over all ranks, in place in the flat buffer."""
        from . import ops
        ops.set_grad_allocator(None)
        for i, (p, v) in enumerate(zip(self.params, self.views)):
            g = p.grad
            if g is None:
                if i not in self._handed:
                    v.zero_()                      # this rank did not touch the parameter: it contributes zeros
            elif not _aliases(g, v):
                v.copy_(g)
        self._average = average
        self._work = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True)
        if not async_op:
            self.finish()
        return self._work

    def finish(self):
        """Wait for the collective and make the summed gradients the parameters' .grad (views, no copy)."""
        if self._work is not None:
            self._work.wait()
            self._work = None
            if self._average:
                self.flat.div_(dist.get_world_size())
        for p, v in zip(self.params, self.views):
            p.grad = v


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
