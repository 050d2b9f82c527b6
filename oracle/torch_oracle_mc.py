"""Pure-PyTorch restatement of the material-stage Monte-Carlo shading path
(reference network/fields.py:618-1473, network/light.py:125-162, utils/ref_utils.py:53-117,
network/other_field.py:12-121).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  In-scope branches only:
shade_mixed, use_nis_diffuse + use_nis_specular with the half-vector parametrisation,
outer_light_version='envlight', geometry_type='schlick', no human lights.

Every random draw of the reference (torch.rand at fields.py:838,876 and flow.py:87) is an
explicit argument so that CUDA and oracle consume identical noise.

Parity status: the Python composition is pinned to the reference's own MCShadingNetwork
(tests/test_oracle_mc_cpu.py, via oracle/ref_shim.py); the cube lookup, segment sums and the
ray tracer underneath are un-vendored third-party ops ("parity unpinned").
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import torch_oracle as O
from . import torch_oracle_mat as OM

EPS = 1e-6


def saturate_dot(a, b):                                                       # utils/network_utils.py:63-64
    return torch.clamp(torch.sum(a * b, dim=-1, keepdim=True), min=0.0, max=1.0)


def sample_sphere(num_samples, begin_elevation=0):                            # utils/base_utils.py:869-882
    ratio = (begin_elevation + 90) / 180
    num_points = int(num_samples // (1 - ratio))
    phi = (np.sqrt(5) - 1.0) / 2.
    az, el = [], []
    for n in range(num_points - num_samples, num_points):
        z = 2. * n / num_points - 1.
        az.append(2 * np.pi * n * phi % (2 * np.pi))
        el.append(np.arcsin(z))
    return np.array(az), np.array(el)


def direction_samples(n: int, dtype=torch.float32) -> torch.Tensor:
    """fields.py:734-737: [n,2] = (az/2pi, 1-2el/pi), stored float32 by the reference."""
    az, el = sample_sphere(n, 0)
    az, el = az * 0.5 / np.pi, 1 - 2 * el / np.pi
    return torch.from_numpy(np.stack([az, el], -1).astype(np.float32)).to(dtype)


# ---- integrated directional encoding (utils/ref_utils.py:8-117) ------------------------
def _ide_tables(deg_view=5):
    ml = []
    for i in range(deg_view):
        l = 2 ** i
        for m in range(l + 1):
            ml.append((m, l))
    ml = np.array(ml).T
    l_max = 2 ** (deg_view - 1)

    def gbc(a, k):
        return np.prod(a - np.arange(k)) / math.factorial(k)

    def alc(l, m, k):
        return ((-1) ** m * 2 ** l * math.factorial(l) / math.factorial(k) / math.factorial(l - k - m)
                * gbc(0.5 * (l + k + m - 1.0), l))

    def shc(l, m, k):
        return np.sqrt((2.0 * l + 1.0) * math.factorial(l - m) / (4.0 * np.pi * math.factorial(l + m))) * alc(l, m, k)

    mat = np.zeros((l_max + 1, ml.shape[1]))
    for i, (m, l) in enumerate(ml.T):
        for k in range(l - m + 1):
            mat[k, i] = shc(l, m, k)
    return ml, mat


_IDE = None


def ide_encode(xyz: torch.Tensor, kappa_inv) -> torch.Tensor:
    """generate_ide_fn(5)(xyz, kappa_inv) -> [...,72]; tables built in float32 like the reference."""
    global _IDE
    if _IDE is None:
        _IDE = _ide_tables(5)
    ml, mat = _IDE
    matt = torch.from_numpy(mat.astype(np.float32)).to(xyz.dtype).to(xyz.device)
    mlt = torch.from_numpy(ml.astype(np.float32)).to(xyz.dtype).to(xyz.device)
    x, y, z = xyz[..., 0:1], xyz[..., 1:2], xyz[..., 2:3]
    vmz = torch.cat([z ** i for i in range(matt.shape[0])], dim=-1)
    vmxy = torch.cat([(x + 1j * y) ** m for m in mlt[0, :]], dim=-1)
    sph = vmxy * torch.matmul(vmz, matt)
    sigma = 0.5 * mlt[1, :] * (mlt[1, :] + 1)
    ide = sph * torch.exp(-sigma * kappa_inv)
    return torch.cat([torch.real(ide), torch.imag(ide)], dim=-1)


# ---- predictor factories (network/other_field.py:12-121) ---------------------------------
class ExpActivation(nn.Module):
    def __init__(self, max_light=5.0):
        super().__init__()
        self.max_light = max_light

    def forward(self, x):
        return torch.exp(torch.clamp(x, max=self.max_light))


def make_predictor(n_layers, feats_dim, output_dim, activation='sigmoid', exp_max=0.0, run_dim=None):
    run_dim = run_dim or (256 if n_layers == 4 else 128)
    act = {'sigmoid': nn.Sigmoid(), 'exp': ExpActivation(exp_max), 'none': nn.Identity()}[activation]
    wn = nn.utils.parametrizations.weight_norm
    layers, last = [], feats_dim
    for _ in range(n_layers - 1):
        layers += [wn(nn.Linear(last, run_dim)), nn.ReLU()]
        last = run_dim
    layers += [wn(nn.Linear(last, output_dim)), act]
    return nn.Sequential(*layers)


class EnvLight(nn.Module):
    """network/light.py:8-31,125-162: trainable log-radiance cubemap, `direct_light` only."""

    def __init__(self, max_res=128, dtype=torch.float32):
        super().__init__()
        self.base = nn.Parameter(torch.full((6, max_res, max_res, 3), math.log(0.5), dtype=dtype))

    def direct_light(self, l, roughness=None):
        return torch.exp(OM.texture_cube(self.base, l.reshape(-1, 3))).reshape(*l.shape[:-1], 3)


def get_orthogonal_directions(d):                                             # fields.py:812-822
    x, y, z = torch.split(d, 1, dim=-1)
    o0 = torch.cat([y, -x, torch.zeros_like(x)], -1)
    o1 = torch.cat([-z, torch.zeros_like(x), x], -1)
    mask0 = (torch.norm(o0, dim=-1) > torch.norm(o1, dim=-1))[:, None]
    return F.normalize(torch.where(mask0, o0, o1), dim=-1)


def direction_to_angle(normals, directions):                                  # fields.py:1035-1048
    z = normals
    x = get_orthogonal_directions(normals)
    y = torch.cross(z, x, dim=-1)
    cx = torch.sum(x.unsqueeze(1) * directions, -1, keepdim=True)
    cy = torch.sum(y.unsqueeze(1) * directions, -1, keepdim=True)
    cz = torch.sum(z.unsqueeze(1) * directions, -1, keepdim=True).clamp(-1 + EPS, 1 - EPS)
    phi = (torch.atan2(cy, cx) + 2 * np.pi) % (2 * np.pi)
    return torch.cat([phi, torch.acos(cz)], dim=-1)


def distribution_ggx(NoH, roughness):                                         # fields.py:1019-1024
    a2 = roughness ** 2
    denom = NoH ** 2 * (a2 - 1.0) + 1.0
    return a2 / (np.pi * denom ** 2).clamp_min(EPS)


def geometry_schlick(NoV, NoL, roughness):                                    # fields.py:987-998
    k = roughness / 2
    return (NoV / (NoV * (1 - k) + k + 1e-5)) * (NoL / (NoL * (1 - k) + k + 1e-5))


class MCShadingNetwork(nn.Module):
    """Restates reference MCShadingNetwork with cfg = configs/mat/syn/compressor.yaml:15-21."""

    def __init__(self, ray_trace_fun: Callable, aabb, gridSize=(512, 512, 512), mat_n_comp=36, flow_grid=(512, 512, 512),
                 light_reso=128, diffuse_sample_num=512, specular_sample_num=256, nis_diffuse_sample_num=64,
                 nis_specular_sample_num=32, exp_max=5.0, dtype=torch.float32, outer_light_version='envlight'):
        super().__init__()
        self.register_buffer("aabb", torch.as_tensor(aabb, dtype=dtype).clone())
        self.n_levels = 3
        planes, lines = [], []
        for i in range(3):                                                    # fields.py:765-774
            m0, m1 = O.MAT_MODE[i]
            planes.append(nn.Parameter(1e-4 * (2 * torch.rand(1, mat_n_comp, gridSize[m0], gridSize[m1], dtype=dtype) - 1)))
            lines.append(nn.Parameter(torch.full((1, mat_n_comp, gridSize[O.VEC_MODE[i]], 1), 1. / (mat_n_comp * 3), dtype=dtype)))
        self.mat_plane, self.mat_line = nn.ParameterList(planes), nn.ParameterList(lines)
        fd = mat_n_comp * 3
        self.metallic_predictor = make_predictor(2, fd, 1).to(dtype)
        self.roughness_predictor = make_predictor(2, fd, 1).to(dtype)
        self.albedo_predictor = make_predictor(2, fd, 3).to(dtype)
        self.outer_light_version = outer_light_version
        if outer_light_version == 'envlight':                                 # fields.py:716-723
            self.outer_light = EnvLight(light_reso, dtype)
        else:
            self.outer_light = make_predictor(4, 72 if outer_light_version == 'direction' else 144, 3, 'exp', exp_max).to(dtype)
            nn.init.constant_(self.outer_light[-2].bias, math.log(0.5))
        self.inner_light = make_predictor(4, 51 + 72, 3, 'exp', exp_max).to(dtype)
        nn.init.constant_(self.inner_light[-2].bias, math.log(0.5))
        self.register_buffer("diffuse_direction_samples", direction_samples(diffuse_sample_num, dtype))
        self.register_buffer("specular_direction_samples", direction_samples(specular_sample_num, dtype))
        self.ray_trace_fun = ray_trace_fun
        self.nis_dn, self.nis_sn = nis_diffuse_sample_num, nis_specular_sample_num
        self.diffuse_sample_num = diffuse_sample_num
        self.flow_diffuse = OM.TensoFlow(aabb, flow_grid, dtype=dtype)
        self.flow_diffuse_copy = OM.TensoFlow(aabb, flow_grid, dtype=dtype)
        self.flow_specular = OM.TensoFlow(aabb, flow_grid, dtype=dtype)
        self.flow_specular_copy = OM.TensoFlow(aabb, flow_grid, dtype=dtype)
        for f in (self.flow_diffuse_copy, self.flow_specular_copy):
            for p in f.parameters():
                p.requires_grad = False
        self.use_flow_diffuse_copy = True
        self.use_flow_specular_copy = True
        self.nis_loss_iter = 500
        self.pos_enc, _ = O.get_embedder(8, 3)

    # ---- materials (fields.py:776-810, 1010-1017) -----------------------------------------
    def predict_materials(self, pts):
        feats = O.vm_feature(self.mat_plane, self.mat_line, pts, None, self.aabb, self.n_levels)
        metallic = self.metallic_predictor(feats)
        roughness = self.roughness_predictor(feats) * (1.0 - 0.04 ** 2) + 0.04 ** 2
        albedo = self.albedo_predictor(feats)
        return metallic, roughness, albedo

    # ---- lights (fields.py:905-975) -----------------------------------------------------------
    def predict_outer_lights(self, points, directions):                       # fields.py:913-933
        if self.outer_light_version == 'envlight':
            return self.outer_light.direct_light(directions)
        enc = ide_encode(directions, 0)
        if self.outer_light_version == 'direction':
            return self.outer_light(enc)
        pts = torch.where((torch.norm(points, dim=-1) > 0.999)[:, None], points * 0.999, points)
        dtx = torch.sum(pts * directions, dim=-1, keepdim=True)                # utils/network_utils.py:108-114
        dist = -dtx + torch.sqrt(dtx ** 2 - torch.sum(pts ** 2, dim=-1, keepdim=True) + 1 + 1e-6)
        return self.outer_light(torch.cat([enc, ide_encode(pts + directions * dist, 0)], -1))

    def get_lights(self, points, directions):
        shape = points.shape[:-1]
        eps = 1e-5
        inters, normals, depth, hit = self.ray_trace_fun(points.reshape(-1, 3) + directions.reshape(-1, 3) * eps,
                                                         directions.reshape(-1, 3))
        inters, normals, depth, hit = inters.reshape(*shape, 3), normals.reshape(*shape, 3), depth.reshape(*shape, 1), hit.reshape(*shape)
        lights = torch.zeros(*shape, 3, dtype=points.dtype, device=points.device)
        miss = ~hit
        if miss.any():
            lights[miss] = self.predict_outer_lights(points[miss], directions[miss])
        if hit.any():
            p, v, n = inters[hit], -directions[hit], normals[hit]
            n, v = F.normalize(n, dim=-1), F.normalize(v, dim=-1)
            refl = torch.sum(v * n, -1, keepdim=True) * n * 2 - v
            lights[hit] = self.inner_light(torch.cat([self.pos_enc(p), ide_encode(refl, 0)], -1))
        lights = lights * (depth > eps).to(lights.dtype)
        return lights, inters, hit

    # ---- direction sets ---------------------------------------------------------------------
    def _frame(self, normals):
        z = normals
        x = get_orthogonal_directions(normals)
        return x, torch.cross(z, x, dim=-1), z

    def sample_diffuse_directions(self, normals, view_dirs, az_shift):      # fields.py:824-856
        pn = normals.shape[0]
        x, y, z = self._frame(normals)
        az, el = torch.split(self.diffuse_direction_samples, 1, dim=1)
        el, az = el.unsqueeze(0), az.unsqueeze(0)
        az = az * np.pi * 2
        el_sqrt = torch.sqrt(el + 1e-7)
        if az_shift is not None:
            az = (az + az_shift * np.pi * 2) % (2 * np.pi)
        cz, cx, cy = torch.sqrt(1 - el + 1e-7), el_sqrt * torch.cos(az), el_sqrt * torch.sin(az)
        directions = cx * x.unsqueeze(1) + cy * y.unsqueeze(1) + cz * z.unsqueeze(1)
        prob = saturate_dot(directions, normals.unsqueeze(1)) / np.pi * (torch.cos((1 - el) * np.pi / 2) * np.pi / 2)
        H = F.normalize(directions + view_dirs.unsqueeze(1), dim=-1)
        cxh = torch.sum(x.unsqueeze(1) * H, -1, keepdim=True)
        cyh = torch.sum(y.unsqueeze(1) * H, -1, keepdim=True)
        czh = torch.sum(z.unsqueeze(1) * H, -1, keepdim=True).clamp(-1 + EPS, 1 - EPS)
        angles_half = torch.cat([(torch.atan2(cyh, cxh) + 2 * np.pi) % (2 * np.pi), torch.acos(czh)], dim=-1)
        return directions, prob, angles_half

    def sample_specular_directions(self, normals, view_dirs, roughness, az_shift):   # fields.py:858-903
        pn = normals.shape[0]
        x, y, z = self._frame(normals)
        a = roughness
        az, el = torch.split(self.specular_direction_samples, 1, dim=1)
        phi = np.pi * 2 * az
        a, el = a.unsqueeze(1), el.unsqueeze(0)
        cos_t = ((1.0 - el) / (1.0 + (a ** 2 - 1.0) * el).clamp_min(EPS)).clamp_min(EPS).sqrt()
        sin_t = (1 - cos_t ** 2).clamp_min(EPS).sqrt()
        phi = phi.unsqueeze(0)
        if az_shift is not None:
            phi = (phi + az_shift * np.pi * 2) % (2 * np.pi)
        H = torch.cos(phi) * sin_t * x.unsqueeze(1) + torch.sin(phi) * sin_t * y.unsqueeze(1) + cos_t * z.unsqueeze(1)
        VoH = saturate_dot(view_dirs.unsqueeze(1), H)
        directions = VoH * H * 2 - view_dirs.unsqueeze(1)
        NoH = cos_t.clamp_min(0.0)
        prob = distribution_ggx(NoH, roughness.unsqueeze(1)) * NoH / (4 * VoH).clamp_min(EPS) * (torch.cos((1 - el) * np.pi / 2) * np.pi / 2)
        angles_H = torch.cat([phi.expand(pn, -1, -1), torch.arcsin(sin_t).expand(pn, -1, -1)], dim=-1)
        return directions, prob, angles_H

    def flow_directions(self, flow, pts, normals, view_dirs, view_angles, roughness, sn, phi_shift):
        """fields.py:1085-1108 / 1164-1188 (half-vector parametrisation)."""
        angles_, logqx = flow.sample(pts, view_angles, roughness, sn, phi_shift)
        ah = torch.cat([angles_[..., :1] * (2 * np.pi), angles_[..., 1:2] * (0.5 * np.pi)], -1)
        phi, theta = torch.split(ah, 1, dim=-1)
        x, y, z = self._frame(normals)
        H = torch.sin(theta) * torch.cos(phi) * x.unsqueeze(1) + torch.sin(theta) * torch.sin(phi) * y.unsqueeze(1) + torch.cos(theta) * z.unsqueeze(1)
        HoV = saturate_dot(view_dirs.unsqueeze(1), H)
        directions = HoV * H * 2 - view_dirs.unsqueeze(1)
        prob = torch.exp(-logqx.clamp(-8, 8)) / (4 * np.pi ** 2 * HoV * torch.sin(theta)).clamp_min(EPS)
        return directions, prob, ah, HoV

    # ---- shade_mixed (fields.py:1075-1335) -----------------------------------------------------
    def shade_mixed(self, pts, normals, view_dirs, metallic, roughness, albedo, noise: Dict[str, Optional[torch.Tensor]],
                    step: Optional[int]):
        pn = pts.shape[0]
        view_angles = direction_to_angle(normals, view_dirs.unsqueeze(1)).squeeze(1)
        view_angles = view_angles / torch.tensor([2 * np.pi, 0.5 * np.pi], dtype=pts.dtype, device=pts.device)
        d2, p2, ah2 = self.sample_diffuse_directions(normals, view_dirs, noise.get("az_diffuse"))
        if self.use_flow_diffuse_copy:
            d1, p1, ah1, _ = self.flow_directions(self.flow_diffuse_copy, pts, normals, view_dirs, view_angles, roughness,
                                                  self.nis_dn, noise.get("phi_diffuse"))
            ddir, dprob, dah = torch.cat([d1, d2], 1), torch.cat([p1, p2], 1), torch.cat([ah1, ah2], 1)
        else:
            ddir, dprob, dah = d2, p2, ah2
        dnum = ddir.shape[1]
        H_d = F.normalize(view_dirs.unsqueeze(1) + ddir, dim=-1)
        HoV_d = torch.clamp(torch.sum(H_d * view_dirs.unsqueeze(1), dim=-1, keepdim=True), min=0.0, max=1.0)
        kd = 1 - metallic.unsqueeze(1)
        dlights, _, _ = self.get_lights(pts.unsqueeze(1).repeat(1, dnum, 1), ddir)
        dweights = albedo.unsqueeze(1) * kd * (saturate_dot(ddir, normals.unsqueeze(1)) / np.pi)
        diffuse_colors = torch.mean(dweights * dlights / dprob.clamp_min(EPS), 1)

        if self.use_flow_specular_copy:
            sdir, sprob, sah, _ = self.flow_directions(self.flow_specular_copy, pts, normals, view_dirs, view_angles, roughness,
                                                       self.nis_sn, noise.get("phi_specular"))
        else:
            sdir, sprob, sah = self.sample_specular_directions(normals, view_dirs, roughness, noise.get("az_specular"))
        snum = sdir.shape[1]
        smask = torch.sum(sdir * normals.unsqueeze(1), dim=-1) > 0
        rid = torch.arange(pn, device=pts.device)[:, None].repeat(1, snum)[smask]
        sdir, sprob, sah = sdir[smask], sprob[smask], sah[smask]
        F0 = 0.04 * (1 - metallic) + metallic * albedo
        Hs = F.normalize(view_dirs[rid] + sdir, dim=-1)
        HoV_s = torch.clamp(torch.sum(Hs * view_dirs[rid], dim=-1, keepdim=True), min=0.0, max=1.0)
        fresnel = F0[rid] + (1.0 - F0[rid]) * torch.clamp(1.0 - HoV_s, min=0.0, max=1.0) ** 5.0
        NoV = saturate_dot(normals, view_dirs)[rid]
        NoL = saturate_dot(normals[rid], sdir)
        geometry = geometry_schlick(NoV, NoL, roughness[rid])
        NoH = saturate_dot(normals[rid], Hs)
        distribution = distribution_ggx(NoH, roughness[rid])
        slights, sinter, shit = self.get_lights(pts[rid], sdir)
        sweights = distribution * fresnel * geometry / (4 * NoV).clamp_min(EPS)

        def seg(v):
            out = torch.zeros(pn, v.shape[-1], dtype=v.dtype, device=v.device)
            return out.index_add(0, rid, v)

        specular_colors = seg(sweights * slights / sprob.clamp_min(EPS)) / snum
        colors = O.linear_to_srgb(diffuse_colors + specular_colors)
        out = {
            'albedo': albedo, 'roughness': roughness, 'metallic': metallic, 'normal': (normals + 1) / 2,
            'diffuse_light': torch.clamp(O.linear_to_srgb(torch.mean(dlights, dim=1)), min=0, max=1),
            'specular_light': torch.clamp(O.linear_to_srgb(seg(slights) / snum), min=0, max=1),
            'diffuse_color': torch.clamp(O.linear_to_srgb(diffuse_colors), min=0, max=1),
            'specular_color': torch.clamp(O.linear_to_srgb(specular_colors), min=0, max=1),
            'visibility': 1 - seg(shit.to(pts.dtype)[:, None]) / snum,
            'indirect_light': seg(slights * shit[..., None].to(pts.dtype)) / snum,
            'diffuse_colors_linear': diffuse_colors, 'specular_colors_linear': specular_colors,
        }
        fx = dweights * dlights
        zero = torch.zeros((), dtype=pts.dtype, device=pts.device)
        if self.use_flow_diffuse_copy and step is not None and step >= self.nis_loss_iter:       # fields.py:1257-1286
            phi, theta = torch.split(dah[:, :self.nis_dn], 1, dim=-1)
            x = torch.cat([phi / (2 * np.pi), theta / (0.5 * np.pi)], -1).clamp(EPS, 1 - EPS)
            _, logq_ = self.flow_diffuse(pts, view_angles, roughness, x)
            logq = logq_ - (4 * np.pi ** 2 * HoV_d[:, :self.nis_dn] * torch.sin(theta)).clamp_min(EPS).log()
            out['loss_nis_diffuse'] = -(fx[:, :self.nis_dn] * logq / dprob[:, :self.nis_dn].clamp_min(EPS)).mean()
        else:
            out['loss_nis_diffuse'] = zero
        fxs = sweights * slights
        if self.use_flow_specular_copy and step is not None and step >= self.nis_loss_iter:      # fields.py:1295-1321
            phi, theta = torch.split(sah, 1, dim=-1)
            x = torch.cat([phi / (2 * np.pi), theta / (0.5 * np.pi)], -1).clamp(EPS, 1 - EPS)
            _, logq_ = self.flow_specular(pts, view_angles, roughness, x, rays_id=rid)
            logq = logq_ - (4 * np.pi ** 2 * HoV_s * torch.sin(theta)).clamp_min(EPS).log()
            out['loss_nis_specular'] = -(fxs * logq / sprob.clamp_min(EPS)).mean()
        else:
            out['loss_nis_specular'] = zero
        out['loss_nis'] = out['loss_nis_diffuse'] + out['loss_nis_specular']
        return colors, out

    def forward(self, pts, view_dirs, normals, noise, step):                 # fields.py:1453-1467
        view_dirs, normals = F.normalize(view_dirs, dim=-1), F.normalize(normals, dim=-1)
        metallic, roughness, albedo = self.predict_materials(pts)
        return self.shade_mixed(pts, normals, view_dirs, metallic, roughness, albedo, noise, step)


def analytic_sphere_tracer(radius: float = 0.5, offset: float = 0.0):
    """Stand-in for the BVH callback (materialRenderer.py:253-263 semantics): closest hit on a
    sphere, miss -> depth 10; returns (inters, unit normals, depth [M,1], hit [M,1])."""

    def trace(o, d):
        o = o + offset * d
        b = (o * d).sum(-1)
        c = (o * o).sum(-1) - radius * radius
        disc = b * b - c
        sq = torch.sqrt(disc.clamp_min(0))
        t0, t1 = -b - sq, -b + sq
        t = torch.where(t0 > 1e-9, t0, t1)
        hit = (disc > 0) & (t > 1e-9)
        depth = torch.where(hit, t, torch.full_like(t, 10.0))
        inter = o + depth[:, None] * d
        n = F.normalize(inter, dim=-1)
        return inter, n, depth[:, None], hit[:, None]
    return trace
