"""Pure-PyTorch restatement of the shape-stage shader and its prefiltered environment light
(reference network/fields.py:320-575 ShapeShadingNetwork, network/light.py:8-122 EnvLight,
network/light_utils.py:66-82 cubemap_mip).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned to the reference's own classes by
tests/test_oracle_shader_cpu.py (through oracle/ref_shim.py); the cube / 2-D texture lookups
underneath restate nvdiffrast's documented semantics ("parity unpinned").
"""
from __future__ import annotations

import math
import os
from typing import Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import torch_oracle as O
from . import torch_oracle_mat as OM
from .torch_oracle_mc import make_predictor, ide_encode


def _texel_center_dirs(res, dtype, device):
    c = torch.linspace(-1.0 + 1.0 / res, 1.0 - 1.0 / res, res, dtype=dtype, device=device)
    gy, gx = torch.meshgrid(c, c, indexing="ij")
    return [F.normalize(OM.cube_to_dir(s, gx, gy), dim=-1, eps=1e-20) for s in range(6)]


class CubemapMip(torch.autograd.Function):
    """network/light_utils.py:66-82: forward = 2x2 average pool; backward = seamless bilinear
    cube lookup of (dout * 0.25) at the fine texel centres (NOT the exact adjoint)."""

    @staticmethod
    def forward(ctx, cubemap):
        return F.avg_pool2d(cubemap.permute(0, 3, 1, 2), (2, 2)).permute(0, 2, 3, 1).contiguous()

    @staticmethod
    def backward(ctx, dout):
        res = dout.shape[1] * 2
        dirs = _texel_center_dirs(res, dout.dtype, dout.device)
        out = torch.stack([OM.texture_cube(dout * 0.25, d.reshape(-1, 3)).reshape(res, res, -1) for d in dirs], 0)
        return out


class EnvLight(nn.Module):
    def __init__(self, max_res=128, min_res=16, min_roughness=0.08, max_roughness=0.5, dtype=torch.float32):
        super().__init__()
        self.min_res, self.max_res = min_res, max_res
        self.min_roughness, self.max_roughness = min_roughness, max_roughness
        self.base = nn.Parameter(torch.full((6, max_res, max_res, 3), math.log(0.5), dtype=dtype))

    def build_mips(self, cutoff=0.99):                                        # light.py:52-64
        self.specular = [self.base]
        while self.specular[-1].shape[1] > self.min_res:
            self.specular.append(CubemapMip.apply(self.specular[-1]))
        self.diffuse = OM.diffuse_cubemap(self.specular[-1])
        n = len(self.specular)
        for idx in range(n - 1):
            r = (idx / (n - 2)) * (self.max_roughness - self.min_roughness) + self.min_roughness
            self.specular[idx] = OM.specular_cubemap(self.specular[idx], r, cutoff)
        self.specular[-1] = OM.specular_cubemap(self.specular[-1], 1.0, cutoff)

    def get_mip(self, roughness):                                             # light.py:72-80
        n = len(self.specular)
        lo = (torch.clamp(roughness, self.min_roughness, self.max_roughness) - self.min_roughness) / \
            (self.max_roughness - self.min_roughness) * (n - 2)
        hi = (torch.clamp(roughness, self.max_roughness, 1.0) - self.max_roughness) / (1.0 - self.max_roughness) + n - 2
        return torch.where(roughness < self.max_roughness, lo, hi)

    def forward(self, l, roughness=None):                                     # light.py:95-122
        if roughness is None:
            light = OM.texture_cube(self.diffuse, l)
        else:
            light = OM.texture_cube_mip(self.specular, l, self.get_mip(roughness)[..., 0])
        return torch.exp(light)


def load_fg_lut(dtype=torch.float32):
    """The split-sum FG LUT fixture (reference assets/bsdf_256_256.bin, [256,256,2])."""
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tensoflow_b200", "assets", "fg_lut_256.npz")
    return torch.from_numpy(np.load(p)["fg"]).to(dtype)


class ShapeShadingNetwork(nn.Module):
    """network/fields.py:320-575 with the shipped defaults: no human light, envlight direct light,
    mat_pos_multires=-1, optional radiance field."""

    def __init__(self, app_feats_dim=128, has_radiance_field=False, light_exp_max=0.0, inner_init=-0.95, env_res=128,
                 env_min_res=16, dtype=torch.float32):
        super().__init__()
        self.has_radiance_field = has_radiance_field
        if has_radiance_field:
            self.rad_mlp = make_predictor(3, app_feats_dim + 3 + 27 + 3, 3, 'sigmoid', run_dim=128).to(dtype)
        self.mat_mlp = make_predictor(3, app_feats_dim, 5, 'sigmoid', run_dim=128).to(dtype)
        self.register_buffer("FG_LUT", load_fg_lut(dtype)[None])
        self.envlight = EnvLight(env_res, env_min_res, dtype=dtype)
        self.inner_light = make_predictor(3, 51 + 72, 3, 'exp', light_exp_max, run_dim=128).to(dtype)
        nn.init.constant_(self.inner_light[-2].bias, math.log(0.5))
        self.inner_weight = make_predictor(3, 51 + 39, 1, 'none', run_dim=128).to(dtype)
        nn.init.constant_(self.inner_weight[-2].bias, inner_init)
        self.pos_enc, _ = O.get_embedder(8, 3)
        self.dir_enc, _ = O.get_embedder(6, 3)
        self.rad_dir_enc, _ = O.get_embedder(4, 3)

    def forward(self, points, normals, view_dirs, feature_vectors, with_radiance=False):
        normals = F.normalize(normals, dim=-1)
        bad = normals[:, :2].sum(dim=-1) == 0.
        normals = torch.where(bad[:, None], torch.tensor([0.0, 1e-6, 1.0], dtype=normals.dtype, device=normals.device), normals)
        view_dirs = F.normalize(view_dirs, dim=-1)
        reflective = torch.sum(view_dirs * normals, -1, keepdim=True) * normals * 2 - view_dirs
        NoV = torch.sum(normals * view_dirs, -1, keepdim=True)
        mat = self.mat_mlp(feature_vectors)
        albedo, roughness, metallic = mat[..., :3] * 0.77 + 0.03, mat[..., 3:4] * 0.9 + 0.09, mat[..., 4:]
        radiance = None
        if self.has_radiance_field and with_radiance:
            radiance = self.rad_mlp(torch.cat([feature_vectors, points, self.rad_dir_enc(view_dirs), normals], -1))
        diffuse_albedo = (1 - metallic) * albedo
        diffuse_light = self.envlight(normals)
        diffuse_color = diffuse_albedo * diffuse_light
        specular_albedo = 0.04 * (1 - metallic) + metallic * albedo
        ref_roughness = ide_encode(reflective, roughness)
        direct_light = self.envlight(reflective, roughness)
        pts = self.pos_enc(points)
        indirect_light = self.inner_light(torch.cat([pts, ref_roughness], -1))
        occ_prob = self.inner_weight(torch.cat([pts.detach(), self.dir_enc(reflective).detach()], -1)) * 0.5 + 0.5
        occ_ = torch.clamp(occ_prob, min=0, max=1)
        specular_light = indirect_light * occ_ + direct_light * (1 - occ_)
        fg_uv = torch.cat([torch.clamp(NoV, min=0.0, max=1.0), torch.clamp(roughness, min=0.0, max=1.0)], -1)
        fg = O.texture2d(self.FG_LUT[0], fg_uv, None, 1)
        specular_ref = specular_albedo * fg[:, 0:1] + fg[:, 1:2]
        specular_color = specular_ref * specular_light
        color = torch.clamp(O.linear_to_srgb(diffuse_color + specular_color), min=0.0, max=1.0)
        occ_info = {'reflective': reflective, 'occ_prob': occ_prob, 'roughness': roughness}
        return color, radiance, occ_info
