"""Drop-in `network.fields` surface of the reference, backed by the sm_100a kernels.

`TensoSDF` keeps the reference's constructor, method names, parameter names and shapes
(`sdf_plane.{0,1,2}` [1,C,H,W], `sdf_line.{i}` [1,C,G,1], `sdf_mat.{0,2}.{weight,bias}`;
reference network/fields.py:20-317) so reference checkpoints load, but holds the factor
tensors channels-last and evaluates them with the fused stencil kernel.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

MAT_MODE = ((0, 1), (0, 2), (1, 2))   # reference network/fields.py:28
VEC_MODE = (2, 1, 0)                  # reference network/fields.py:29


def _cl(t: torch.Tensor) -> torch.Tensor:
    """channels-last storage for a [1,C,H,W] factor tensor (texel channels contiguous)."""
    n, c, h, w = t.shape
    out = torch.empty_strided((n, c, h, w), (c * h * w, 1, w * c, c), dtype=t.dtype, device=t.device)
    out.copy_(t)
    return out


class TVLoss(nn.Module):
    """reference network/other_field.py:170-191"""

    def __init__(self, TVLoss_weight=1):
        super().__init__()
        self.TVLoss_weight = TVLoss_weight

    def forward(self, x):
        b, c, h, w = x.shape
        if not (x.is_cuda and b == 1 and c % 4 == 0 and x.dtype == torch.float32 and x.permute(0, 2, 3, 1).is_contiguous()):
            raise RuntimeError("tensoflow_b200.TVLoss needs one fp32 CUDA texture [1,C,H,W] stored channels-last with C % 4 == 0 "
                               "(the VM factors of this package); there is no PyTorch fallback")
        return ops.TVFunction.apply(x, self.TVLoss_weight)           # fused kernels on the channels-last factors


class TensoSDF(nn.Module):
    def __init__(self, gridSize, aabb, device='cuda', sdf_n_comp=36, sdf_dim=256, app_dim=128, init_n_levels=3,
                 sdf_multires=3):
        super().__init__()
        if sdf_multires != 0:
            # every shipped shape config uses sdf_multires: 0; the fused kernel
            # concatenates raw xyz exactly as reference network/fields.py:265,298 does then.
            raise NotImplementedError("tensoflow_b200.TensoSDF supports sdf_multires=0 only")
        self.sdf_n_comp, self.sdf_dim, self.app_dim = sdf_n_comp, sdf_dim, app_dim
        self.device = device
        self.matMode = [list(m) for m in MAT_MODE]
        self.vecMode = list(VEC_MODE)
        self.nplane = 3
        self.init_radius = 0.2
        self.sdf_multires = sdf_multires
        self.kernel_size, self.sigma = 5, 0.5
        self.update_gridSize_aabb(torch.as_tensor(gridSize).long().cpu(), torch.as_tensor(aabb, dtype=torch.float32).to(device),
                                  init_n_levels)
        planes, lines = [], []
        for i in range(3):                                          # reference fields.py:101-131
            ps = self.gridSize[self.matMode[i]]
            x = torch.linspace(-1, 1, int(ps[0]))
            y = torch.linspace(-1, 1, int(ps[1]))
            xx, yy = torch.meshgrid(x, y, indexing='ij')
            init = torch.linalg.norm(torch.stack([xx, yy], -1), ord=2, dim=-1) - self.init_radius
            planes.append(nn.Parameter(_cl(init[None, None].repeat(1, sdf_n_comp, 1, 1).to(device))))
            ls = int(self.gridSize[self.vecMode[i]])
            lines.append(nn.Parameter(_cl(torch.full((1, sdf_n_comp, ls, 1), 1. / (sdf_n_comp * 3), device=device))))
        self.sdf_plane = nn.ParameterList(planes)
        self.sdf_line = nn.ParameterList(lines)
        self.sdf_mat = nn.Sequential(nn.Linear(3 * sdf_n_comp + 3, sdf_dim), nn.Softplus(beta=100),
                                     nn.Linear(sdf_dim, 1 + app_dim)).to(device)
        nn.init.constant_(self.sdf_mat[0].bias, 0.0)                # reference fields.py:83-91
        nn.init.normal_(self.sdf_mat[0].weight, 0.0, math.sqrt(2) / math.sqrt(sdf_dim))
        nn.init.constant_(self.sdf_mat[-1].bias, -self.init_radius)
        nn.init.normal_(self.sdf_mat[-1].weight, mean=math.sqrt(math.pi) / math.sqrt(sdf_dim), std=0.0001)

    # ---- bookkeeping (reference fields.py:56-62, 143-178) ---------------------------------
    def update_gridSize_aabb(self, gridSize, aabb, n_levels):
        self.gridSize = gridSize
        self.aabb = aabb
        self.aabbSize = self.aabb[1] - self.aabb[0]
        self.units = self.aabbSize / (self.gridSize.to(self.aabbSize.device) - 1)
        self.n_levels = int(n_levels)

    def get_optparam_groups(self, lr_init_spatialxyz=0.02, lr_init_network=0.001):
        return [{'params': self.sdf_line, 'lr': lr_init_spatialxyz}, {'params': self.sdf_plane, 'lr': lr_init_spatialxyz},
                {'params': self.sdf_mat.parameters(), 'lr': lr_init_network}]

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):
        new_levels = self.n_levels + 1
        res_target = torch.as_tensor(res_target)
        res_target = ((res_target / 2 ** (new_levels - 1)).int() * 2 ** (new_levels - 1))
        for i in range(3):
            m0, m1 = self.matMode[i]
            self.sdf_plane[i] = nn.Parameter(_cl(F.interpolate(
                self.sdf_plane[i].data.contiguous(), size=(int(res_target[m1]), int(res_target[m0])), mode='bilinear',
                align_corners=True)))
            self.sdf_line[i] = nn.Parameter(_cl(F.interpolate(
                self.sdf_line[i].data.contiguous(), size=(int(res_target[self.vecMode[i]]), 1), mode='bilinear',
                align_corners=True)))
        self.update_gridSize_aabb(res_target.long().cpu(), self.aabb, new_levels)
        return res_target, self.n_levels

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # reference checkpoints may hold factors at another resolution (after upsampling):
        # adopt their shapes, keeping channels-last storage.
        for name, plist in (('sdf_plane', self.sdf_plane), ('sdf_line', self.sdf_line)):
            for i in range(3):
                k = f'{prefix}{name}.{i}'
                if k in state_dict and state_dict[k].shape != plist[i].shape:
                    plist[i] = nn.Parameter(_cl(state_dict[k].to(plist[i].device, torch.float32)))
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    # ---- the hot path -----------------------------------------------------------------
    def stencil(self, xyz, level):
        """Fused forward + finite-difference taps: (sdf [N], feat [N,A], grad [N,3], hess [N])."""
        units = [float(u) for u in self.units]
        return ops.SdfStencilFunction.apply(xyz, level, units, self.aabb, self.n_levels, self.sdf_mat[0].weight,
                                            self.sdf_mat[0].bias, self.sdf_mat[2].weight, self.sdf_mat[2].bias,
                                            *self.sdf_plane, *self.sdf_line)

    def forward(self, xyz_sampled, level_vol):
        """reference fields.py:262-299 -> [N, 1+app_dim]"""
        if not (torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())):
            # no gradient wanted: one query per point (the stencil would evaluate and discard the six FD taps)
            out = ops.sdf_point(xyz_sampled, level_vol, self.aabb, self.n_levels, self.sdf_mat[0].weight, self.sdf_mat[0].bias,
                                self.sdf_mat[2].weight, self.sdf_mat[2].bias, list(self.sdf_plane), list(self.sdf_line))
            if out is not None:
                return torch.cat([out[0][:, None], out[1]], -1)
        sdf, feat, _, _ = self.stencil(xyz_sampled.reshape(-1, 3), level_vol)
        return torch.cat([sdf[:, None], feat], -1)

    def sdf(self, xyz_sampled, level_vol=None):
        """reference fields.py:148 -> [N,1]"""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self.stencil(xyz_sampled.reshape(-1, 3), level_vol)[0][:, None]
        return ops.sdf_only(xyz_sampled, level_vol, self.aabb, self.n_levels, self.sdf_mat[0].weight, self.sdf_mat[0].bias,
                            self.sdf_mat[2].weight, self.sdf_mat[2].bias, list(self.sdf_plane), list(self.sdf_line))[:, None]

    def gradient(self, x, level_vol, training=False, sdf=None):
        """reference fields.py:227-260 -> (gradients [N,3], normal_hessian [N] or None)"""
        if x.shape[0] == 0:
            z = torch.zeros(0, 3, device=x.device)
            return (z, torch.zeros(0, 3, device=x.device)) if training else (z, None)
        _, _, grad, hess = self.stencil(x, level_vol)
        return grad, (hess if training else None)

    # ---- regularisers (dense passes over the factors; reference fields.py:133-138, 301-309) ----
    def TV_loss_sdf(self, reg):
        total = 0
        for i in range(3):
            total = total + reg(self.sdf_plane[i]) + reg(self.sdf_line[i])
        return total

    def grid_gaussian_loss(self):
        """reference fields.py:301-309 (GaussianBlur2D / 1D of other_field.py:142-168 with kernel_size 5, sigma 0.5)"""
        return gaussian_loss(self.sdf_plane, self.sdf_line, self.kernel_size, self.sigma)


def gaussian_loss(planes, lines, kernel_size, sigma):
    """sum_i sum_interior (plane_i - blur2d(plane_i))^2 + (line_i - blur1d(line_i))^2 on the fused residual kernels."""
    k1, k2 = ops.gaussian_taps(kernel_size, sigma)
    total = 0.
    for i in range(len(planes)):
        total = total + ops.GaussResidualFunction.apply(planes[i], k2, kernel_size, kernel_size)
        total = total + ops.GaussResidualFunction.apply(lines[i], k1, kernel_size, 1)
    return total


class SingleVarianceNetwork(nn.Module):
    """reference network/other_field.py:193-201 (act='exp')"""

    def __init__(self, init_val, activation='exp'):
        super().__init__()
        assert activation == 'exp'
        self.register_parameter('variance', nn.Parameter(torch.tensor(init_val)))

    def forward(self, x):
        return torch.ones([*x.shape[:-1], 1], device=x.device) * torch.exp(self.variance * 10.0)
