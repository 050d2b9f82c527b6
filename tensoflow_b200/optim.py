"""Optimizer step of the reference trainer (`Adam(grad_vars, betas=(0.9, 0.99))`, train/trainer_inv.py:112,124,212) on the
multi-tensor `tf_adam_step` kernel: one streaming pass over parameters, gradients and both moments (28 B/element)
instead of PyTorch's ~8 elementwise foreach passes.  Same `torch.optim.Optimizer` surface (param groups with their own
`lr`, `state_dict`), so the trainer's `param_group['lr'] *= lr_factor` schedule (:247-248) works unchanged.
No CPU fallback: parameters must live on a CUDA device."""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import _lib
from ._lib import check, stream_ptr


def _same_layout(a: torch.Tensor, b: torch.Tensor) -> bool:
    return a.shape == b.shape and a.stride() == b.stride()


def _dense(t: torch.Tensor) -> bool:
    """True when the tensor covers numel() consecutive elements in SOME dimension order (contiguous, channels-last, ...)."""
    expect = 1
    for size, stride in sorted(((s, st) for s, st in zip(t.shape, t.stride()) if s > 1), key=lambda x: x[1]):
        if stride != expect:
            return False
        expect *= size
    return True


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (no weight decay, no amsgrad: the reference uses neither)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.99), eps=1e-8):
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or eps < 0.0:
            raise ValueError("FusedAdam: betas in [0,1) and eps >= 0 required")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FusedAdam.step(closure): the reference trainer never passes one")
        lib = _lib.load()
        # every tensor of every group goes into ONE call (the ABI takes a learning rate per tensor); a call per distinct
        # (step count, betas, eps) only when parameters were added later (e.g. re-created by an upsampling)
        calls = {}
        keep = []                                           # gradient copies must outlive the launch
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdam needs fp32 CUDA parameters (tensoflow_b200 has no CPU fallback)")
                st = self.state[p]
                if not st:
                    if not _dense(p):
                        raise RuntimeError("FusedAdam needs dense parameters")
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    if not (_same_layout(st['exp_avg'], p) and _same_layout(st['exp_avg_sq'], p)):
                        raise RuntimeError("FusedAdam: optimizer state layout differs from the parameter's")
                g = p.grad
                if g.dtype != torch.float32 or not _same_layout(g, p):
                    g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)     # same memory order as p
                    keep.append(g)
                st['step'] += 1
                c = calls.setdefault((st['step'], float(b1), float(b2), float(group['eps'])), ([], [], [], [], [], []))
                c[0].append(p.data_ptr()); c[1].append(g.data_ptr()); c[2].append(st['exp_avg'].data_ptr())
                c[3].append(st['exp_avg_sq'].data_ptr()); c[4].append(p.numel()); c[5].append(float(group['lr']))
        stream = stream_ptr()
        for (k, b1, b2, eps), (ps, gs, ms, vs, numel, lr) in calls.items():
            n = len(ps)
            vp = C.c_void_p * n
            check(lib.tf_adam_step(n, vp(*ps), vp(*gs), vp(*ms), vp(*vs), (C.c_int64 * n)(*numel), (C.c_float * n)(*lr), b1, b2, eps,
                                   int(k), stream), "tf_adam_step")
        del keep
        if calls:
            # the kernel rewrote the parameters through raw pointers: autograd version counters did not move, so everything
            # derived from parameter values and cached on them (VM descriptors with their mip chains) must be dropped
            from . import ops
            ops.bump_param_epoch()
        return None
