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


# --------------------------------------------------------------------------
# TensoFlow sampler (network/flow.py): prior, piecewise-quadratic coupling, VM feature
# --------------------------------------------------------------------------
from . import torch_oracle as _O  # noqa: E402
import torch.nn as nn  # noqa: E402


def sphere_prior_angles(num_samples: int, dtype=torch.float32) -> torch.Tensor:
    """SphereSampler.set_angle (network/flow.py:62-76): the last `num_samples` points of a
    Fibonacci sphere above 1 degree elevation, as (phi/2pi, theta/(pi/2)); built in float32
    like the reference, then cast."""
    ratio = (1 + 90) / 180
    num_points = int(num_samples // (1 - ratio))
    g = (np.sqrt(5) - 1.0) / 2.0
    phis, thetas = [], []
    for n in range(num_points - num_samples, num_points):
        z = 2.0 * n / num_points - 1.0
        phis.append(2 * np.pi * n * g % (2 * np.pi))
        thetas.append(np.arcsin(z))
    phi = torch.tensor(phis, dtype=torch.float32) / (2 * np.pi)
    theta = torch.tensor(thetas, dtype=torch.float32) / (0.5 * np.pi)
    return torch.stack([phi, theta], -1).to(dtype)


def sphere_prior_sample(pn: int, sn: int, phi_shift: Optional[torch.Tensor], dtype=torch.float32, device="cpu"):
    """SphereSampler.forward (flow.py:82-90).  phi_shift [pn,sn,1] stands for the
    torch.rand_like draw of training mode (None = eval)."""
    x = sphere_prior_angles(sn, dtype).to(device).expand(pn, sn, 2)
    if phi_shift is not None:
        x = torch.cat([(x[..., :1] + phi_shift) % 1, x[..., 1:]], -1)
    x = x.clamp(1e-6, 1 - 1e-6)
    logj = -torch.cos(x[..., 1:] * (0.5 * np.pi)).log()
    return x, logj


def _pwq_common(wv_tilde: torch.Tensor, clamp_w: bool):
    """wv_tilde [N,21] -> w [N,10], wsum [N,10], wsum_shift [N,11], v [N,11], vw [N,11]."""
    nv = int(np.ceil(wv_tilde.shape[-1] / 2))
    v_t, w_t = wv_tilde[:, :nv], wv_tilde[:, nv:]
    w = torch.exp(w_t)
    if clamp_w:
        w = w.clamp_min(1e-6)                                              # flow.py:343
    wsum = torch.cumsum(w, -1)
    wn = wsum[:, -1:]
    w = w / wn
    if clamp_w:
        w = w.clamp_min(1e-6)                                              # flow.py:346
    wsum = wsum / wn
    wsum_shift = torch.cat([torch.zeros_like(wsum[:, :1]), wsum], -1)
    v = torch.exp(v_t)
    v = (v / ((v[:, :-1] + v[:, 1:]) / 2 * w).sum(-1, keepdim=True)).clamp_min(1e-6)   # flow.py:166-168,350
    vw = torch.cat([torch.zeros_like(wsum[:, :1]), torch.cumsum((v[:, :-1] + v[:, 1:]) / 2 * w, -1)], -1)
    return w, wsum, wsum_shift, v, vw


def pwquad_forward(x: torch.Tensor, wv_tilde: torch.Tensor):
    """ElementWisePWQuadraticTransform.flow_inv (flow.py:332-413) for one coordinate:
    x [N] in (0,1) -> (out [N], logj [N]); used for density evaluation."""
    w, wsum, wsum_shift, v, vw = _pwq_common(wv_tilde, True)
    b = w.shape[-1]
    eps = torch.finfo(wsum.dtype).eps
    finder = torch.where(wsum > x[:, None], torch.zeros_like(wsum), torch.ones_like(wsum))
    mx = torch.argmax(torch.cat([torch.full_like(wsum[:, :1], eps), finder * wsum], -1), -1).clamp(0, b - 1)[:, None]
    wb, vb, vb1 = w.gather(-1, mx)[:, 0], v.gather(-1, mx)[:, 0], v.gather(-1, mx + 1)[:, 0]
    alphas = ((x - wsum_shift.gather(-1, mx)[:, 0]) / wb).clamp(0, 1)
    out = alphas ** 2 / 2 * ((vb1 - vb) * wb) + alphas * vb * wb + vw.gather(-1, mx)[:, 0]
    out = out.clamp(min=eps, max=1.0 - eps)
    logj = torch.log(torch.lerp(vb, vb1, alphas))
    return out, logj


def pwquad_inverse(y: torch.Tensor, wv_tilde: torch.Tensor):
    """ElementWisePWQuadraticTransform.flow (flow.py:415-525): y [N] -> (x [N], logj [N]);
    used for sampling."""
    w, wsum, wsum_shift, v, vw = _pwq_common(wv_tilde, False)
    b = w.shape[-1]
    eps = torch.finfo(vw.dtype).eps
    finder = torch.where(vw > y[:, None], torch.zeros_like(vw), torch.ones_like(vw))
    mx = torch.argmax(torch.cat([torch.full_like(vw[:, :1], eps), finder * (vw + 1)], -1), -1) - 1
    e = mx.clamp(0, b - 1)[:, None]
    wb, vb, vb1 = w.gather(-1, e)[:, 0], v.gather(-1, e)[:, 0], v.gather(-1, e + 1)[:, 0]
    a = (vb1 - vb) * wb
    bb = vb * wb
    c = vw.gather(-1, e)[:, 0] - y
    a = torch.where(a.abs() < eps, eps * torch.ones_like(a), a)
    d = (bb ** 2 - 2 * a * c).clamp_min(0)
    sol1 = (-bb - torch.sqrt(d)) / a
    sol2 = (-bb + torch.sqrt(d)) / a
    sol = torch.where((sol1 >= 0) & (sol1 < 1), sol1, sol2).clamp(min=eps, max=1.0 - eps)
    x = (wb * sol + wsum_shift.gather(-1, e)[:, 0]).clamp(min=eps, max=1.0 - eps)
    logj = -torch.log(torch.lerp(vb, vb1, sol))
    return x, logj


class FlowBlock(nn.Module):
    """network/flow.py:549-641 (Block) with the pwquad defaults: PE(y_n) (7) + feature (37)
    -> Reshift -> 3x(Linear 64 + LeakyReLU) -> Linear 21 -> spline on the other coordinate.
    Parameter names match the reference (`nn.1`, `nn.3`, `nn.5`, `nn.7`; `nn.0` = Reshift)."""

    class Reshift(nn.Module):
        def __init__(self):
            super().__init__()
            self.scale = nn.Parameter(torch.scalar_tensor(2.0), requires_grad=False)
            self.offset = nn.Parameter(torch.scalar_tensor(-1.0), requires_grad=False)

        def forward(self, x):
            return x * self.scale + self.offset

    def __init__(self, cond_index: int, feature_dim: int = 37, d_hidden: int = 64, n_bins: int = 21, multires: int = 3):
        super().__init__()
        self.cond = cond_index                 # coordinate that conditions (mask == True)
        self.embed, d_in = _O.get_embedder(multires, 1)
        layers = [FlowBlock.Reshift()]
        last = d_in + feature_dim
        for _ in range(3):
            layers += [nn.Linear(last, d_hidden), nn.LeakyReLU()]
            last = d_hidden
        layers.append(nn.Linear(last, n_bins))
        self.nn = nn.Sequential(*layers)

    def _st(self, y, feature):
        y_n = y[:, self.cond:self.cond + 1]
        return self.nn(torch.cat([self.embed(y_n), feature], -1))

    def flow(self, y, logj, feature):           # sampling direction (flow.py:600-616)
        st = self._st(y, feature)
        t = 1 - self.cond
        xt, lj = pwquad_inverse(y[:, t], st)
        x = torch.zeros_like(y)
        x[:, self.cond] = y[:, self.cond]
        x[:, t] = xt
        return x, logj + lj[:, None]

    def flow_inv(self, y, logj, feature):       # density direction (flow.py:623-641)
        st = self._st(y, feature)
        t = 1 - self.cond
        xt, lj = pwquad_forward(y[:, t], st)
        x = torch.zeros_like(y)
        x[:, self.cond] = y[:, self.cond]
        x[:, t] = xt
        return x, logj + lj[:, None]


class TensoFlow(nn.Module):
    """network/flow.py:643-855 with flow='pwquad' (every shipped config)."""

    def __init__(self, aabb, gridSize=(512, 512, 512), nis_n_comp=12, nis_dim=64, nis_feature_dim=16, dtype=torch.float32):
        super().__init__()
        self.register_buffer("aabb", torch.as_tensor(aabb, dtype=dtype).clone())
        self.gridSize = [int(g) for g in gridSize]
        self.n_levels = 3
        planes, lines = [], []
        for i in range(3):                                                    # flow.py:755-764
            m0, m1 = _O.MAT_MODE[i]
            planes.append(nn.Parameter(1e-4 * (2 * torch.rand(1, nis_n_comp, self.gridSize[m0], self.gridSize[m1], dtype=dtype) - 1)))
            lines.append(nn.Parameter(torch.full((1, nis_n_comp, self.gridSize[_O.VEC_MODE[i]], 1), 1.0 / (nis_n_comp * 3), dtype=dtype)))
        self.nis_plane = nn.ParameterList(planes)
        self.nis_line = nn.ParameterList(lines)
        self.embed_xyz, xyz_ch = _O.get_embedder(3, 3)
        self.nis_mat = nn.Sequential(nn.Linear(3 * nis_n_comp + xyz_ch, nis_dim), nn.Softplus(beta=100),
                                     nn.Linear(nis_dim, nis_feature_dim)).to(dtype)
        self.embed_refl, refl_ch = _O.get_embedder(3, 2)
        self.embed_rough, rough_ch = _O.get_embedder(3, 1)
        fdim = nis_feature_dim + refl_ch + rough_ch
        self.flows = nn.ModuleList([FlowBlock(0, fdim), FlowBlock(1, fdim)]).to(dtype)   # masks [T,F], [F,T]: flow.py:668-674

    def tenso_feature(self, pts):                                             # flow.py:709-744
        feat = _O.vm_feature(self.nis_plane, self.nis_line, pts, None, self.aabb, self.n_levels)
        return self.nis_mat(torch.cat([feat, self.embed_xyz(pts)], -1))

    def condition(self, pts, reflections, roughness):                         # flow.py:803-815 / 836-848
        rough = torch.zeros_like(self.embed_rough(roughness))                 # computed then zeroed (flow.py:814,847)
        return torch.cat([self.tenso_feature(pts), self.embed_refl(reflections), rough], -1)

    def sample(self, pts, reflections, roughness, n_samples, phi_shift=None):
        """flow.py:833-855 -> angles [pn,sn,2], logj [pn,sn,1] (= -log q)."""
        pn = pts.shape[0]
        x, logj = sphere_prior_sample(pn, n_samples, phi_shift, pts.dtype, pts.device)
        feature = self.condition(pts, reflections, roughness)
        feature = feature[:, None, :].expand(-1, n_samples, -1).reshape(pn * n_samples, -1)
        x, logj = x.reshape(-1, 2), logj.reshape(-1, 1)
        for f in self.flows:
            x, logj = f.flow(x, logj, feature)
        return x.reshape(pn, n_samples, 2), logj.reshape(pn, n_samples, 1)

    def forward(self, pts, reflections, roughness, x, rays_id=None):
        """flow.py:801-831 -> z, log q(x); x [pn,sn,2] or ragged [M,2] with rays_id."""
        x = x.clamp(1e-6, 1 - 1e-6)
        feature = self.condition(pts, reflections, roughness)
        if rays_id is not None:
            feature = feature[rays_id]
        pre = x.shape[:-1]
        if x.dim() == 3:
            feature = feature[:, None, :].expand(-1, x.shape[1], -1)
        x = x.reshape(-1, 2)
        feature = feature.reshape(-1, feature.shape[-1])
        logj = torch.zeros(x.shape[0], 1, dtype=x.dtype, device=x.device)
        for f in list(self.flows)[::-1]:
            x, logj = f.flow_inv(x, logj, feature)
        logq = logj + torch.cos(x[:, 1:] * (0.5 * np.pi)).log()              # + prior log-prob (flow.py:78-80,825)
        return x.reshape(*pre, 2), logq.reshape(*pre, 1)
