"""Pure-PyTorch restatement of the material-stage / lighting pieces of the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are relative to
/root/reference.

Parity status: cube-map lookups restate nvdiffrast's documented `boundary_mode=
'cube'` semantics ("parity unpinned": nvdiffrast is not vendored); the cube-face
convention itself is pinned twice by the reference (network/light_utils.py:24-31
and network/renderutils/c_src/cubemap.cu:32-60) and is checked in
tests/test_oracle_cube.py.  The prefilter maths restates cubemap.cu, which IS
in-tree source.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-6  # network/fields.py:18


# --------------------------------------------------------------------------
# cube map geometry
# --------------------------------------------------------------------------
def cube_to_dir(s: int, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """network/light_utils.py:24-31 (un-normalised)."""
    one = torch.ones_like(x)
    if s == 0:
        r = (one, -y, -x)
    elif s == 1:
        r = (-one, -y, x)
    elif s == 2:
        r = (x, one, y)
    elif s == 3:
        r = (x, -one, -y)
    elif s == 4:
        r = (x, -y, one)
    else:
        r = (-x, -y, -one)
    return torch.stack(r, -1)


def dir_to_face_xy(d: torch.Tensor):
    """Inverse of cube_to_dir: d [N,3] -> face [N] long, x,y in [-1,1] (x along W, y along H)."""
    ax = d.abs()
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    is_x = (ax[:, 0] >= ax[:, 1]) & (ax[:, 0] >= ax[:, 2])
    is_y = (~is_x) & (ax[:, 1] >= ax[:, 2])
    is_z = ~(is_x | is_y)
    face = torch.zeros(d.shape[0], dtype=torch.long, device=d.device)
    face = torch.where(is_x, torch.where(dx >= 0, 0, 1), face)
    face = torch.where(is_y, torch.where(dy >= 0, 2, 3), face)
    face = torch.where(is_z, torch.where(dz >= 0, 4, 5), face)
    m = torch.where(is_x, ax[:, 0], torch.where(is_y, ax[:, 1], ax[:, 2])).clamp_min(1e-30)
    x = torch.zeros_like(dx)
    y = torch.zeros_like(dx)
    x = torch.where(face == 0, -dz / m, x); y = torch.where(face == 0, -dy / m, y)
    x = torch.where(face == 1, dz / m, x);  y = torch.where(face == 1, -dy / m, y)
    x = torch.where(face == 2, dx / m, x);  y = torch.where(face == 2, dz / m, y)
    x = torch.where(face == 3, dx / m, x);  y = torch.where(face == 3, -dz / m, y)
    x = torch.where(face == 4, dx / m, x);  y = torch.where(face == 4, -dy / m, y)
    x = torch.where(face == 5, -dx / m, x); y = torch.where(face == 5, -dy / m, y)
    return face, x, y


def _fold_texel(face: torch.Tensor, ix: torch.Tensor, iy: torch.Tensor, R: int):
    """Map a texel (face, ix, iy) whose ix OR iy lies one step outside [0,R-1] to
    the texel across the cube edge (unfolded-cube neighbour).  Returns (face', ix', iy')."""
    fx = (2.0 * (ix.double() + 0.5) / R - 1.0)
    fy = (2.0 * (iy.double() + 0.5) / R - 1.0)
    out_face, out_x, out_y = face.clone(), ix.clone(), iy.clone()
    for s in range(6):
        sel = face == s
        if not bool(sel.any()):
            continue
        p = cube_to_dir(s, fx[sel], fy[sel])              # [n,3], major axis == +-1
        a = p.abs()
        major = int(s // 2)
        ex = (a - 1.0).clamp_min(0.0)                     # excess beyond the face square
        e = ex.sum(-1)                                     # exactly one axis exceeds
        over = e > 0
        q = p.clone()
        # pull the overshooting axis back to +-1 and push the old major axis in by e
        for k in range(3):
            if k == major:
                continue
            ok = ex[:, k] > 0
            q[ok, k] = torch.sign(p[ok, k])
            q[ok, major] = torch.sign(p[ok, major]) * (1.0 - e[ok])
        f2, x2, y2 = dir_to_face_xy(q)
        ix2 = torch.floor((x2 + 1.0) * 0.5 * R).long().clamp(0, R - 1)
        iy2 = torch.floor((y2 + 1.0) * 0.5 * R).long().clamp(0, R - 1)
        idx = torch.nonzero(sel)[:, 0]
        out_face[idx[over]] = f2[over]
        out_x[idx[over]] = ix2[over]
        out_y[idx[over]] = iy2[over]
    return out_face, out_x, out_y


def texture_cube(tex: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """dr.texture(tex[None], d, filter_mode='linear', boundary_mode='cube')
    tex [6,R,R,C], d [N,3] (need not be normalised) -> [N,C].
    Seamless bilinear: taps that fall off the face come from the adjacent face; a
    tap that falls off in both axes (cube corner) is dropped and the remaining
    weights renormalised.  (parity unpinned; call sites network/light.py:107,135)"""
    R = tex.shape[1]
    face, x, y = dir_to_face_xy(d)
    u = (x + 1.0) * 0.5 * R - 0.5
    v = (y + 1.0) * 0.5 * R - 0.5
    u0 = torch.floor(u); v0 = torch.floor(v)
    fu = u - u0; fv = v - v0
    u0 = u0.long(); v0 = v0.long()
    out = torch.zeros(d.shape[0], tex.shape[-1], dtype=tex.dtype, device=tex.device)
    wsum = torch.zeros(d.shape[0], 1, dtype=tex.dtype, device=tex.device)
    for du, dv in ((0, 0), (1, 0), (0, 1), (1, 1)):
        iu = u0 + du; iv = v0 + dv
        w = (fu if du else 1 - fu) * (fv if dv else 1 - fv)
        ou = (iu < 0) | (iu >= R)
        ov = (iv < 0) | (iv >= R)
        corner = ou & ov
        f2, iu2, iv2 = _fold_texel(face, iu, iv, R)
        iu2 = iu2.clamp(0, R - 1); iv2 = iv2.clamp(0, R - 1)
        w = torch.where(corner, torch.zeros_like(w), w).unsqueeze(-1)
        out = out + tex[f2, iv2, iu2] * w
        wsum = wsum + w
    return out / wsum


def texture_cube_mip(stack: Sequence[torch.Tensor], d: torch.Tensor, level: torch.Tensor) -> torch.Tensor:
    """linear-mipmap-linear over a user-supplied stack (network/light.py:111-118)."""
    n = len(stack)
    lv = level.reshape(-1).clamp(0.0, float(n - 1))
    l0 = torch.floor(lv)
    f = (lv - l0).unsqueeze(-1)
    l0 = l0.long()
    l1 = (l0 + 1).clamp(max=n - 1)
    out = torch.zeros(d.shape[0], stack[0].shape[-1], dtype=stack[0].dtype, device=d.device)
    for l in range(n):
        s = texture_cube(stack[l], d)
        w = (l0 == l).unsqueeze(-1) * (1 - f) + ((l1 == l) & (l0 != l)).unsqueeze(-1) * f
        out = out + s * w
    return out


# --------------------------------------------------------------------------
# cubemap prefilters: restates network/renderutils/c_src/cubemap.cu
# --------------------------------------------------------------------------
def _pixel_area(N: int, dtype, device):                                        # cubemap.cu:17-30
    if N <= 1:
        return torch.ones(1, 1, dtype=dtype, device=device)
    H = N // 2
    i = (torch.arange(N, device=device) - H).abs().to(dtype)
    dx = torch.atan((i + 1) / H) - torch.atan(i / H)
    return dx[None, :] * dx[:, None]      # [y,x]


def _texel_dirs(N: int, dtype, device):                                        # cubemap.cu:32-46
    c = 2.0 * ((torch.arange(N, device=device).to(dtype) + 0.5) / N) - 1.0
    fy, fx = torch.meshgrid(c, c, indexing="ij")
    dirs = torch.stack([F.normalize(cube_to_dir(s, fx, fy), dim=-1) for s in range(6)], 0)
    return dirs                                                                # [6,N,N,3]


def diffuse_cubemap(cubemap: torch.Tensor) -> torch.Tensor:
    """cubemap.cu:110-139: out(n) = sum_L clamp(n.L,0,0.999)*area(L)/3.141592 * c(L)."""
    N = cubemap.shape[1]
    dirs = _texel_dirs(N, cubemap.dtype, cubemap.device).reshape(-1, 3)
    area = _pixel_area(N, cubemap.dtype, cubemap.device).reshape(1, -1).repeat(6, 1).reshape(-1)
    w = (dirs @ dirs.T).clamp(0.0, 0.999) * area[None, :] / 3.141592
    return (w @ cubemap.reshape(-1, cubemap.shape[-1])).reshape(cubemap.shape)


_cutoff_cache = {}


def _ndf_cutoff(roughness: float, cutoff: float) -> float:                     # ops.py:427-438
    key = (roughness, cutoff)
    if key not in _cutoff_cache:
        a2 = roughness ** 4
        ct = np.cos(np.linspace(0, np.pi / 2.0, 1000000))
        c = np.clip(ct, 0.0, 1.0)
        dd = (c * a2 - c) * c + 1.0
        D = np.cumsum(a2 / (dd * dd * np.pi))
        _cutoff_cache[key] = float(ct[np.argmax(D >= D[-1] * cutoff)])
    return _cutoff_cache[key]


def specular_cubemap(cubemap: torch.Tensor, roughness: float, cutoff: float = 0.99, chunk: int = 2048) -> torch.Tensor:
    """cubemap.cu:246-298 + ops.py:446-458.  The AABB bounds of cubemap.cu:181-244
    only cull texels already rejected by the cone test, so the dense form is identical."""
    N = cubemap.shape[1]
    cos_cut = _ndf_cutoff(roughness, cutoff)
    dirs = _texel_dirs(N, cubemap.dtype, cubemap.device).reshape(-1, 3)
    area = _pixel_area(N, cubemap.dtype, cubemap.device).reshape(1, -1).repeat(6, 1).reshape(-1)
    a2 = (roughness * roughness) ** 2
    flat = cubemap.reshape(-1, cubemap.shape[-1])
    outs = []
    for i in range(0, dirs.shape[0], chunk):
        V = dirs[i:i + chunk]                                  # [m,3]
        LdV = V @ dirs.T                                       # [m,M]
        Hh = F.normalize(dirs[None, :, :] + V[:, None, :], dim=-1, eps=1e-20)
        VdH = (Hh * V[:, None, :]).sum(-1).clamp_min(0.0).clamp(0.0, 1.0)
        dd = (VdH * a2 - VdH) * VdH + 1.0
        ndf = a2 / (dd * dd * math.pi)
        w = LdV.clamp_min(0.0) * ndf * area[None, :] / 4.0
        w = torch.where(LdV >= cos_cut, w, torch.zeros_like(w))
        outs.append((w @ flat) / w.sum(-1, keepdim=True))
    return torch.cat(outs, 0).reshape(cubemap.shape)
