"""Shape-stage per-ray work (reference network/shapeRenderer.py) on the sm_100a kernels.

`render_core` mirrors ShapeRenderer.render_core (reference shapeRenderer.py:1105-1277):
packed samples in, per-ray colour / acc / normal and the sample-level loss terms out.
Two kernels do the work: the fused TensoSDF stencil (field + FD normals + hessian) and
the NeuS alpha + compositing kernel; the few remaining elementwise ops are index
plumbing.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

from . import ops


def near_far_from_sphere(rays_o, dirs, radius=1.0):
    """reference shapeRenderer.py:676-684"""
    a = torch.sum(dirs ** 2, dim=-1, keepdim=True)
    b = 2.0 * torch.sum(rays_o * dirs, dim=-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return torch.clamp(mid - radius, min=1e-3), mid + radius


def compute_ball_radii(distance, radiis, cos):
    """reference shapeRenderer.py:966-970"""
    inverse_cos = 1.0 / cos
    tmp = (inverse_cos * inverse_cos - 1).sqrt() - radiis
    return distance * radiis * cos / (tmp * tmp + 1.0).sqrt()


def ray_offsets_from_indices(ray_indices: torch.Tensor, n_rays: int) -> torch.Tensor:
    """sorted int64 ray_indices [N] (nerfacc packing) -> CSR offsets int32 [n_rays+1]"""
    bounds = torch.arange(n_rays + 1, device=ray_indices.device, dtype=ray_indices.dtype)
    return torch.searchsorted(ray_indices.contiguous(), bounds).to(torch.int32)


def charbonnier(rgb_pr, rgb_gt, epsilon=0.001):
    """reference shapeRenderer.py:803-805"""
    return torch.sqrt(torch.sum((rgb_gt - rgb_pr) ** 2, dim=-1) + epsilon)


def render_core(field, variance: torch.Tensor, color_fn: Callable, rays_o, dirs, radiis, rays_cos, t_starts, t_ends,
                ray_indices, cos_anneal_ratio: float = 1.0, base_radii: Optional[float] = None, is_train: bool = True,
                train_variance: bool = True, white_bg: bool = True) -> Dict[str, torch.Tensor]:
    """reference shapeRenderer.py:1105-1206.  `field` is a tensoflow_b200.fields.TensoSDF,
    `color_fn(points, normals, view_dirs, feature_vectors) -> [N,3]` stands for the shading network."""
    n_rays = rays_o.shape[0]
    mid = (t_starts + t_ends) * 0.5
    dists = t_ends - t_starts
    viewdir = dirs[ray_indices]
    pts = rays_o[ray_indices] + viewdir * mid[:, None]
    if base_radii is None:
        base_radii = float(field.aabbSize[0] / 2 / field.gridSize[0])          # reference shapeRenderer.py:251
    ball = compute_ball_radii(mid[:, None], radiis[ray_indices], rays_cos[ray_indices])
    levels = torch.log2(ball / base_radii)
    sdf, feat, grads, hess = field.stencil(pts, levels)
    normals = F.normalize(grads, dim=-1)
    colors = color_fn(pts, normals, -viewdir, feat)
    offsets = ray_offsets_from_indices(ray_indices, n_rays)
    vals = torch.cat([colors, grads], -1)
    alpha, weights, acc, out = ops.NeusCompositeFunction.apply(sdf, grads, dists, dirs, offsets, variance,
                                                               float(cos_anneal_ratio), vals, train_variance)
    acc = acc[:, None]
    rgb = out[:, :3]
    if white_bg:
        rgb = rgb + (1 - acc)
    up = torch.tensor([0.0, 0.0, 1.0], device=rgb.device)
    normal = F.normalize(out[:, 3:6] * acc + (1.0 - acc) * up, dim=-1)
    res = {
        'ray_rgb': rgb, 'acc': acc, 'normal': normal,
        'gradient_error': (torch.linalg.norm(grads, ord=2, dim=-1) - 1.0) ** 2,
        'sample_num': pts.shape[0] / max(n_rays, 1),
        'sdf': sdf, 'alpha': alpha, 'weights': weights, 'gradients': grads, 'feat': feat, 'levels': levels,
        'points': pts,
    }
    if is_train and pts.shape[0] > 0:
        res['loss_sparse'] = torch.exp(-20.0 * sdf.abs()).mean()
        res['loss_hessian'] = hess.abs().mean()
        res['hessian'] = hess
        res['std'] = torch.mean(1 / torch.exp(variance * 10.0).clip(1e-6, 1e6))
    return res
