"""Hierarchical NeuS ray sampler of the shape stage on the `tf_sampler_*` kernels.

`hierarchical_sample` is `ShapeRenderer.sample_ray` of the reference (network/shapeRenderer.py:871-932 with upsample :820-849,
cat_z_vals :851-869 and utils/network_utils.py:117-147 sample_pdf, det=True): coarse stratified depths inside the box, then
`up_sample_steps` rounds of importance depths drawn from the NeuS section weights of the sorted list, each round followed by an
SDF-only field query of the new points (except the last), and finally the packed (t_starts, t_ends, ray_indices) of the
intervals whose mid point lies inside the box.  The reference runs ~100 sort / searchsorted / cumprod / gather tensor ops for
this; here it is 1 + (rounds + 1) + 2 kernels next to the field queries.  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr

_TABLES: dict = {}


def _table(kind: str, n: int, device) -> torch.Tensor:
    """torch.linspace tables exactly as the reference builds them: 'lin' = linspace(0, 1, n) (sample_ray),
    'u' = linspace(0.5/n, 1 - 0.5/n, n) (sample_pdf, det=True)."""
    key = (kind, n, str(device))
    t = _TABLES.get(key)
    if t is None:
        t = torch.linspace(0.0, 1.0, n, device=device) if kind == "lin" else torch.linspace(0. + 0.5 / n, 1. - 0.5 / n, steps=n, device=device)
        _TABLES[key] = t.contiguous()
    return t


def _c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


@torch.no_grad()
def hierarchical_sample(sdf_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], aabb: torch.Tensor, base_radii: float,
                        rays_o, dirs, near, far, radiis, rays_cos, n_samples: int, n_importance: int, up_sample_steps: int,
                        perturb: float, t_rand: Optional[torch.Tensor], variance: Optional[torch.Tensor], clip_sample_variance: bool):
    """sdf_fn(points [N,3], level [N]) -> [N] SDF values (no grad).  `variance` is the SingleVarianceNetwork parameter (device
    scalar, read by the kernel: no host sync) when clip_sample_variance, else the cap 64 * 2^i alone is used.
    Returns (t_starts [N], t_ends [N], ray_indices int64 [N], ray_offsets int32 [R+1])."""
    lib = _lib.load()
    R = int(rays_o.shape[0])
    dev = rays_o.device
    S = n_samples + (n_importance if n_importance > 0 else 0)
    ro, rd, nr, fr, rad, rc = _c(rays_o), _c(dirs), _c(near.reshape(-1)), _c(far.reshape(-1)), _c(radiis.reshape(-1)), _c(rays_cos.reshape(-1))
    ab_min, ab_max = ops._aabb_floats(aabb)
    aabb6 = (C.c_float * 6)(*ab_min, *ab_max)
    if R == 0:
        e = torch.zeros(0, device=dev)
        return e, e.clone(), torch.zeros(0, dtype=torch.long, device=dev), torch.zeros(1, dtype=torch.int32, device=dev)
    z = torch.empty(R, S, device=dev, dtype=torch.float32)
    sdf = torch.empty(R, S, device=dev, dtype=torch.float32)
    pts = torch.empty(R * n_samples, 3, device=dev, dtype=torch.float32)
    level = torch.empty(R * n_samples, device=dev, dtype=torch.float32)
    tr = None
    if perturb > 0:
        tr = _c(t_rand.reshape(-1)) if t_rand is not None else torch.rand(R, device=dev)
    check(lib.tf_sampler_init(ptr(ro), ptr(rd), ptr(nr), ptr(fr), ptr(rad), ptr(rc), ptr(_table("lin", n_samples, dev)), ptr(tr), aabb6,
                              float(base_radii), R, n_samples, S, ptr(z), ptr(pts), ptr(level), stream_ptr()), "tf_sampler_init")
    n = n_samples
    if n_importance > 0:
        m = n_importance // up_sample_steps
        sdf[:, :n_samples] = sdf_fn(pts, level).reshape(R, n_samples)
        u = _table("u", m, dev)
        var = _c(variance.reshape(1)) if (clip_sample_variance and variance is not None) else None
        new_z = new_sdf = None
        m_in = 0
        for i in range(up_sample_steps):
            out_z = torch.empty(R, m, device=dev, dtype=torch.float32)
            out_pts = torch.empty(R * m, 3, device=dev, dtype=torch.float32)
            out_lv = torch.empty(R * m, device=dev, dtype=torch.float32)
            check(lib.tf_sampler_upsample(ptr(ro), ptr(rd), ptr(rad), ptr(rc), ptr(z), ptr(sdf), ptr(new_z), ptr(new_sdf), m_in, ptr(u), m,
                                          ptr(var), float(64 * 2 ** i), float(base_radii), R, n, S, ptr(out_z), ptr(out_pts), ptr(out_lv),
                                          stream_ptr()), "tf_sampler_upsample")
            n += m_in
            new_z, m_in = out_z, m
            new_sdf = _c(sdf_fn(out_pts, out_lv)) if i + 1 < up_sample_steps else None        # cat_z_vals(last=True): no SDF
        check(lib.tf_sampler_upsample(ptr(ro), ptr(rd), ptr(rad), ptr(rc), ptr(z), ptr(sdf), ptr(new_z), None, m_in, None, 0, None, 0.0,
                                      float(base_radii), R, n, S, None, None, None, stream_ptr()), "tf_sampler_upsample (merge)")
        n += m_in
    counts = torch.empty(R, device=dev, dtype=torch.int32)
    check(lib.tf_sampler_finalize(ptr(ro), ptr(rd), ptr(z), aabb6, R, n, S, ptr(counts), None, None, None, None, stream_ptr()),
          "tf_sampler_finalize (count)")
    offsets = torch.zeros(R + 1, device=dev, dtype=torch.int64)
    torch.cumsum(counts, 0, out=offsets[1:])
    total = int(offsets[-1])                              # the one host sync of the sampler (the packed size)
    t_starts = torch.empty(total, device=dev, dtype=torch.float32)
    t_ends = torch.empty(total, device=dev, dtype=torch.float32)
    ray_indices = torch.empty(total, device=dev, dtype=torch.int64)
    if total > 0:
        check(lib.tf_sampler_finalize(ptr(ro), ptr(rd), ptr(z), aabb6, R, n, S, None, ptr(offsets), ptr(t_starts), ptr(t_ends),
                                      ptr(ray_indices), stream_ptr()), "tf_sampler_finalize (write)")
    return t_starts, t_ends, ray_indices, offsets.to(torch.int32)


@torch.no_grad()
def probe_sections(sdf_fn: Callable[[torch.Tensor], torch.Tensor], variance: torch.Tensor, origins, dirs, t0: Optional[torch.Tensor],
                   t1: torch.Tensor, sn0: int, sn1: int):
    """Two-stage NeuS probe of secondary / mesh-guided rays: sn0 uniform depths in [t0, t1] (t0 None: [0, t1]) -> weights ->
    sn1 importance depths (deterministic quantiles) -> weights again.  The body of the reference's get_intersection
    (utils/network_utils.py:172-202) and get_intersection_around_mesh (network/materialRenderer.py:281-313) on the
    `tf_probe_*` kernels.  sdf_fn(points [N,3]) -> [N].  Returns (z_mid, weights, mid_sdf), each [pn, sn1-1]."""
    lib = _lib.load()
    o, d = _c(origins), _c(dirs)
    pn, dev = int(o.shape[0]), o.device
    var = _c(variance.reshape(1))
    t1c = _c(t1.reshape(-1))
    t0c = None if t0 is None else _c(t0.reshape(-1))
    z = torch.empty(pn, sn0, device=dev, dtype=torch.float32)
    pts = torch.empty(pn * sn0, 3, device=dev, dtype=torch.float32)
    check(lib.tf_probe_init(ptr(o), ptr(d), ptr(t0c), ptr(t1c), ptr(_table("lin", sn0, dev)), pn, sn0, ptr(z), ptr(pts), stream_ptr()),
          "tf_probe_init")
    sdf = _c(sdf_fn(pts).reshape(pn, sn0))
    z_new = torch.empty(pn, sn1, device=dev, dtype=torch.float32)
    pts_new = torch.empty(pn * sn1, 3, device=dev, dtype=torch.float32)
    check(lib.tf_probe_weights(ptr(o), ptr(d), ptr(z), ptr(sdf), ptr(var), pn, sn0, ptr(_table("u", sn1, dev)), sn1, ptr(z_new), ptr(pts_new),
                               None, None, None, stream_ptr()), "tf_probe_weights (resample)")
    sdf_new = _c(sdf_fn(pts_new).reshape(pn, sn1))
    w = torch.empty(pn, sn1 - 1, device=dev, dtype=torch.float32)
    mid_sdf = torch.empty_like(w)
    z_mid = torch.empty_like(w)
    check(lib.tf_probe_weights(ptr(o), ptr(d), ptr(z_new), ptr(sdf_new), ptr(var), pn, sn1, None, 0, None, None, ptr(w), ptr(mid_sdf),
                               ptr(z_mid), stream_ptr()), "tf_probe_weights (weights)")
    return z_mid, w, mid_sdf
