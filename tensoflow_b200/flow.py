"""Drop-in `network.flow.TensoFlow` (reference network/flow.py:643-855, flow='pwquad'):
the tensorial-feature-conditioned normalizing-flow sampler of light directions.

Same constructor / method names / parameter names as the reference
(`nis_plane.{i}`, `nis_line.{i}`, `nis_mat.{0,2}`, `flows.{b}.nn.{1,3,5,7}`), evaluated by the
sm_100a kernels: VM feature gather, fused linear layers for the coupling conditioners and the
piecewise-quadratic spline kernels.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .fields import _cl, MAT_MODE, VEC_MODE


def posenc(x: torch.Tensor, multires: int) -> torch.Tensor:
    """reference utils/network_utils.py:38-50: [x, sin(2^k x), cos(2^k x)]_{k<multires}"""
    out = [x]
    for k in range(multires):
        out += [torch.sin(x * (2.0 ** k)), torch.cos(x * (2.0 ** k))]
    return torch.cat(out, -1)


class Reshift(nn.Module):
    """reference network/flow.py:146-164"""

    def __init__(self, scale=2., offset=-1.):
        super().__init__()
        self.scale = nn.Parameter(torch.scalar_tensor(scale), requires_grad=False)
        self.offset = nn.Parameter(torch.scalar_tensor(offset), requires_grad=False)

    def forward(self, x):
        return x * self.scale + self.offset


class SphereSampler(nn.Module):
    """reference network/flow.py:52-90"""

    def __init__(self, d=2):
        super().__init__()
        self.d = d
        self.angle = None

    def set_angle(self, num_samples, device):
        ratio = (1 + 90) / 180
        num_points = int(num_samples // (1 - ratio))
        g = (np.sqrt(5) - 1.0) / 2.
        phis, thetas = [], []
        for n in range(num_points - num_samples, num_points):
            z = 2. * n / num_points - 1.
            phis.append(2 * np.pi * n * g % (2 * np.pi))
            thetas.append(np.arcsin(z))
        phi = torch.tensor(phis, dtype=torch.float32, device=device) / (2 * np.pi)
        theta = torch.tensor(thetas, dtype=torch.float32, device=device) / (0.5 * np.pi)
        self.angle = torch.stack([phi, theta], dim=-1)

    def log_prob(self, x):
        return torch.cos(x[..., 1:] * (0.5 * np.pi)).log()

    def forward(self, shape, device, phi_shift=None):
        if self.angle is None or self.angle.shape[0] != shape[1] or self.angle.device != device:
            self.set_angle(shape[1], device)
        x = self.angle.expand(*shape, 2)
        if self.training or phi_shift is not None:
            if phi_shift is None:
                phi_shift = torch.rand(*shape, 1, device=device)
            x = torch.cat([(x[..., :1] + phi_shift) % 1, x[..., 1:]], dim=-1)
        x = x.clamp(1e-6, 1 - 1e-6)
        return x, -self.log_prob(x)


class Block(nn.Module):
    """Coupling layer (reference network/flow.py:549-641): conditioner MLP on
    [PE(y_n), feature] -> 21 spline parameters for the other coordinate."""

    def __init__(self, d, mask, d_hidden=64, n_hidden=3, feature_dim=37, multires=3, n_bins=21):
        super().__init__()
        assert d == 2 and sum(mask) == 1
        self.mask = list(mask)
        self.cond = self.mask.index(True)
        self.multires = multires
        d_in = 1 + 2 * multires
        layers = [Reshift()]
        last = d_in + feature_dim
        for _ in range(n_hidden):
            layers += [nn.Linear(last, d_hidden), nn.LeakyReLU()]
            last = d_hidden
        layers.append(nn.Linear(last, n_bins))
        self.nn = nn.Sequential(*layers)

    def _st(self, y, feature):
        y_n = y[:, self.cond:self.cond + 1]
        h = self.nn[0](torch.cat([posenc(y_n, self.multires), feature], -1))
        h = ops.linear(h, self.nn[1].weight, self.nn[1].bias, "leaky")
        h = ops.linear(h, self.nn[3].weight, self.nn[3].bias, "leaky")
        h = ops.linear(h, self.nn[5].weight, self.nn[5].bias, "leaky")
        return ops.linear(h, self.nn[7].weight, self.nn[7].bias, "none")

    def _couple(self, y, logj, feature, inverse):
        st = self._st(y, feature)
        t = 1 - self.cond
        xt, lj = ops.PwquadFunction.apply(y[:, t], st, inverse)
        cols = [None, None]
        cols[self.cond] = y[:, self.cond]
        cols[t] = xt
        return torch.stack(cols, -1), logj + lj[:, None]

    def flow(self, y, logj, feature, return_jacobian=True):          # sampling direction
        return self._couple(y, logj, feature, True)

    def flow_inv(self, y, logj, feature, return_jacobian=True):      # density direction
        return self._couple(y, logj, feature, False)

    # ---- fused path: one kernel per block (tf_flow_block_*), one conditioning vector per point ---------------------
    def _reshift_consts(self):
        r = self.nn[0]
        key = (r.scale._version, r.offset._version, r.scale.data_ptr())
        if getattr(self, '_reshift_key', None) != key:
            self._reshift_key, self._reshift_val = key, (float(r.scale), float(r.offset))     # one host read per (re)load
        return self._reshift_val

    def couple_points(self, y, logj, feature_pp, sn, inverse):
        """y [pn*sn,2], logj [pn*sn,1], feature_pp [pn,F] (shared by the sn directions of a point)"""
        scale, offset = self._reshift_consts()
        yo, lj = ops.FlowBlockFunction.apply(y, logj, feature_pp, sn, self.cond, inverse, scale, offset, self.nn[1].weight, self.nn[1].bias,
                                             self.nn[3].weight, self.nn[3].bias, self.nn[5].weight, self.nn[5].bias, self.nn[7].weight,
                                             self.nn[7].bias)
        return yo, lj[:, None]


class TensoFlow(nn.Module):
    def __init__(self, d, aabb, device='cuda', gridSize=[512, 512, 512], nis_n_comp=12, nis_dim=64, nis_feature_dim=16,
                 nis_multires=3, refl_multires=3, roughness_multires=3, angle_multires=3, flow='pwquad', n_bins=10,
                 disable_tensorial=False, disable_reflected=False):
        super().__init__()
        if flow != 'pwquad':
            raise NotImplementedError("only flow='pwquad' (every shipped config) has kernels")
        assert d == 2 and n_bins == 10 and nis_multires == 3 and refl_multires == 3 and roughness_multires == 3
        self.nis_n_comp, self.nis_dim, self.nis_feature_dim = nis_n_comp, nis_dim, nis_feature_dim
        self.device = device
        self.matMode = [list(m) for m in MAT_MODE]
        self.vecMode = list(VEC_MODE)
        self.nplane = 3
        self.gridSize = torch.tensor(gridSize)
        self.aabb = torch.as_tensor(aabb, dtype=torch.float32).to(device)
        self.n_levels = 3
        planes, lines = [], []
        for i in range(3):                                          # reference flow.py:755-764
            ps = self.gridSize[self.matMode[i]]
            planes.append(nn.Parameter(_cl((1e-4 * (2 * torch.rand(1, nis_n_comp, int(ps[0]), int(ps[1])) - 1)).to(device))))
            ls = int(self.gridSize[self.vecMode[i]])
            lines.append(nn.Parameter(_cl(torch.full((1, nis_n_comp, ls, 1), 1. / (nis_n_comp * 3), device=device))))
        self.nis_plane = nn.ParameterList(planes)
        self.nis_line = nn.ParameterList(lines)
        self.nis_mat = nn.Sequential(nn.Linear(3 * nis_n_comp + 21, nis_dim), nn.Softplus(beta=100),
                                     nn.Linear(nis_dim, nis_feature_dim)).to(device)
        self.refl_input_ch, self.roughness_input_ch = 14, 7
        fdim = nis_feature_dim + self.refl_input_ch + self.roughness_input_ch
        self.flows = nn.ModuleList([Block(d, [True, False], feature_dim=fdim, multires=angle_multires, n_bins=2 * n_bins + 1),
                                    Block(d, [False, True], feature_dim=fdim, multires=angle_multires, n_bins=2 * n_bins + 1)]).to(device)
        self.latent_prior = SphereSampler(d)
        self.disable_tensorial, self.disable_reflected = disable_tensorial, disable_reflected

    def get_optparam_groups(self, lr_init_spatialxyz=0.01, lr_init_network=0.001):
        return [{'params': self.nis_line, 'lr': lr_init_spatialxyz}, {'params': self.nis_plane, 'lr': lr_init_spatialxyz},
                {'params': self.nis_mat.parameters(), 'lr': lr_init_network}, {'params': self.flows.parameters(), 'lr': lr_init_network}]

    def tenso_feature(self, xyz_sampled, level_vol=None):
        """reference flow.py:709-744 -> [N, nis_feature_dim]"""
        feat = ops.VMFeatureFunction.apply(xyz_sampled, level_vol, self.aabb, self.n_levels, *self.nis_plane, *self.nis_line)
        h = ops.linear(torch.cat([feat, posenc(xyz_sampled, 3)], -1), self.nis_mat[0].weight, self.nis_mat[0].bias, "softplus100")
        return ops.linear(h, self.nis_mat[2].weight, self.nis_mat[2].bias, "none")

    def _condition(self, pts, reflections, roughness):
        feature = self.tenso_feature(pts)
        if self.disable_tensorial:
            feature = torch.zeros_like(feature)
        refl = posenc(reflections, 3)
        if self.disable_reflected:
            refl = torch.zeros_like(refl)
        rough = torch.zeros(pts.shape[0], self.roughness_input_ch, device=pts.device)   # zeroed in the reference (flow.py:814,847)
        return torch.cat([feature, refl, rough], -1)

    @staticmethod
    def _fused_ok(feature, sn):
        """the fused block kernels take one conditioning vector per point shared by >= 16 consecutive directions"""
        return feature.is_cuda and sn >= 16 and feature.shape[1] <= 40

    def sample(self, pts, reflections, roughness, n_samples, return_jacobian=False, phi_shift=None):
        """reference flow.py:833-855 -> angles [pn,sn,2] (, logj [pn,sn,1] = -log q)"""
        pn = pts.shape[0]
        x, logj = self.latent_prior((pn, n_samples), pts.device, phi_shift)
        feature = self._condition(pts, reflections, roughness)
        x, logj = x.reshape(-1, 2), logj.reshape(-1, 1)
        if self._fused_ok(feature, n_samples):
            for f in self.flows:
                x, logj = f.couple_points(x, logj, feature, n_samples, True)
        else:
            feature = feature[:, None, :].expand(-1, n_samples, -1).reshape(pn * n_samples, -1)
            for f in self.flows:
                x, logj = f.flow(x, logj, feature)
        x, logj = x.reshape(pn, n_samples, 2), logj.reshape(pn, n_samples, 1)
        return (x, logj) if return_jacobian else x

    def forward(self, pts, reflections, roughness, x, return_jacobian=False, rays_id=None):
        """reference flow.py:801-831 -> z (, log q(x))"""
        x = x.clamp(1e-6, 1 - 1e-6)
        feature = self._condition(pts, reflections, roughness)
        if rays_id is not None:
            feature = feature[rays_id]
        pre = x.shape[:-1]
        if x.dim() == 3 and rays_id is None and self._fused_ok(feature, x.shape[1]) and feature.shape[0] == x.shape[0]:
            sn = x.shape[1]
            x = x.reshape(-1, 2)
            logj = None
            for f in list(self.flows)[::-1]:
                x, logj = f.couple_points(x, logj, feature, sn, False)
        else:
            if x.dim() == 3:
                feature = feature[:, None, :].expand(-1, x.shape[1], -1)
            x = x.reshape(-1, 2)
            feature = feature.reshape(-1, feature.shape[-1])
            logj = torch.zeros(x.shape[0], 1, device=x.device)
            for f in list(self.flows)[::-1]:
                x, logj = f.flow_inv(x, logj, feature)
        z = x.reshape(*pre, 2)
        if not return_jacobian:
            return z
        logq = logj + self.latent_prior.log_prob(x)
        return z, logq.reshape(*pre, 1)
