"""Synthetic scenes of the BASELINE.json shapes: rays, packed samples and
perturbed fields.  Pure tensor generators with fixed seeds; usable on CPU or CUDA, shared
by tests/, bench.py and __graft_entry__.smoke()."""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def make_rays(n_rays: int, seed: int = 0, device="cpu", radius: float = 2.0, radii_jitter: bool = True):
    """Origins uniform on the sphere of `radius`, directions toward U[-0.3,0.3]^3 targets."""
    g = torch.Generator().manual_seed(seed)
    o = F.normalize(torch.randn(n_rays, 3, generator=g), dim=-1) * radius
    tgt = (torch.rand(n_rays, 3, generator=g) - 0.5) * 0.6
    d = F.normalize(tgt - o, dim=-1)
    if radii_jitter:   # raw mip levels ~ U[-0.6, 2.7] at t~2 for G=512 (config 2)
        radiis = 1.25e-3 * 2.0 ** (torch.rand(n_rays, 1, generator=g) * 3 - 1)
    else:
        radiis = torch.full((n_rays, 1), 1e-3)
    rays_cos = torch.ones(n_rays, 1)
    rgb = torch.rand(n_rays, 3, generator=g)
    return {k: v.to(device) for k, v in dict(rays_o=o, dirs=d, radiis=radiis, rays_cos=rays_cos, rgbs=rgb).items()}


def uniform_samples(rays_o, dirs, aabb, n_samples: int):
    """n_samples fixed-step intervals between the aabb entry and exit of every ray, packed
    (t_starts, t_ends, ray_indices) like nerfacc; rays that miss the box get no samples."""
    vec = torch.where(dirs == 0, torch.full_like(dirs, 1e-6), dirs)
    ra = (aabb[1] - rays_o) / vec
    rb = (aabb[0] - rays_o) / vec
    t_min = torch.minimum(ra, rb).amax(-1).clamp(min=1e-3)
    t_max = torch.maximum(ra, rb).amin(-1)
    hit = t_max > t_min
    step = (t_max - t_min) / n_samples
    k = torch.arange(n_samples, device=rays_o.device, dtype=rays_o.dtype)
    t0 = t_min[:, None] + step[:, None] * k[None, :]
    t1 = t0 + step[:, None]
    idx = torch.arange(rays_o.shape[0], device=rays_o.device)[:, None].expand(-1, n_samples)
    m = hit[:, None].expand(-1, n_samples)
    return t0[m].contiguous(), t1[m].contiguous(), idx[m].contiguous()


@torch.no_grad()
def perturb_field(field, seed: int = 1, noise: float = 1e-2):
    """N(0, noise) on planes/lines so mip levels and channels differ (works on the oracle
    TensoSDF and on tensoflow_b200.fields.TensoSDF alike)."""
    g = torch.Generator().manual_seed(seed)
    for plist in (field.sdf_plane, field.sdf_line):
        for p in plist:
            p.add_((torch.randn(p.shape, generator=g) * noise).to(p.device, p.dtype))


def simple_color_fn(points, normals, view_dirs, feat):
    """Stand-in for the shading network in field+raymarch-only runs: a smooth function of the
    first appearance features and the normal, so that d colour reaches the decoder."""
    return torch.sigmoid(feat[:, :3] + 0.5 * normals)


def copy_field_params(src, dst):
    """Copy planes/lines/MLP between two TensoSDF implementations with equal shapes."""
    with torch.no_grad():
        for a, b in zip(list(src.sdf_plane) + list(src.sdf_line), list(dst.sdf_plane) + list(dst.sdf_line)):
            b.copy_(a.to(b.device, b.dtype))
        for a, b in zip(src.sdf_mat.parameters(), dst.sdf_mat.parameters()):
            b.copy_(a.to(b.device, b.dtype))


def bumpy_sphere(nu: int, nv: int, r: float = 0.5, bump: float = 0.1):
    """Lat-long sphere of radius r with radial bumps: the synthetic mesh family of BASELINE config 3 (1000 x 500 quads there).
    Returns (vertices [nu*nv,3] fp32, triangles [2*nu*(nv-1),3] int32), both on the host."""
    u = torch.linspace(0, 2 * math.pi, nu + 1)[:-1]
    v = torch.linspace(0.05, math.pi - 0.05, nv)
    vv, uu = torch.meshgrid(v, u, indexing="ij")
    rad = r * (1 + bump * torch.sin(5 * uu) * torch.sin(4 * vv))
    verts = torch.stack([rad * torch.sin(vv) * torch.cos(uu), rad * torch.sin(vv) * torch.sin(uu), rad * torch.cos(vv)], -1).reshape(-1, 3)
    i = torch.arange(nv - 1)[:, None]
    j = torch.arange(nu)[None, :]
    a, b = i * nu + j, i * nu + (j + 1) % nu
    c, d = (i + 1) * nu + j, (i + 1) * nu + (j + 1) % nu
    tris = torch.cat([torch.stack([a, c, b], -1).reshape(-1, 3), torch.stack([b, c, d], -1).reshape(-1, 3)], 0)
    return verts.float(), tris.to(torch.int32)
