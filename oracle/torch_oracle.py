"""Pure-PyTorch, device-agnostic restatement of the TensoFlow shape-stage hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the
reference file:line (relative to /root/reference) whose arithmetic it restates.

Parity status
-------------
* The Python composition (TensoSDF, finite-difference gradient, NeuS alpha,
  compositing, losses) is PINNED: tests/test_oracle_vs_reference.py runs the
  reference's own classes through oracle/ref_shim.py in the build container and
  tests/golden/*.npz holds their outputs for the GPU box.
* The third-party arithmetic underneath (nvdiffrast.texture, nerfacc,
  torch_scatter) is NOT vendored by the reference and has no reference test:
  the functions marked "parity unpinned" below restate the libraries' documented
  semantics.

Works in float32 or float64 (tests use float64 as the high-precision arbiter).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

MAT_MODE = ((0, 1), (0, 2), (1, 2))   # network/fields.py:28
VEC_MODE = (2, 1, 0)                  # network/fields.py:29


# --------------------------------------------------------------------------
# nvdiffrast.torch.texture, 2-D, boundary_mode='clamp'  (parity unpinned)
# --------------------------------------------------------------------------
def build_mip_chain(tex: torch.Tensor, n_levels: int) -> List[torch.Tensor]:
    """tex [H,W,C] -> list of n_levels tensors; level l+1 is the 2x2 box average
    of level l (2x1 when one extent is already 1).  SURVEY appendix C."""
    chain = [tex]
    for _ in range(1, n_levels):
        t = chain[-1]
        H, W, C = t.shape
        if H > 1:
            assert H % 2 == 0, "mip chain needs even extents"
            t = 0.5 * (t[0::2] + t[1::2])
        if W > 1:
            assert W % 2 == 0, "mip chain needs even extents"
            t = 0.5 * (t[:, 0::2] + t[:, 1::2])
        chain.append(t)
    return chain


def _bilinear_clamp(tex: torch.Tensor, uv: torch.Tensor) -> torch.Tensor:
    """tex [H,W,C], uv [N,2] with (u,v)<->(W,H); texel centres at (i+.5)/W;
    indices clamped to the edge texel."""
    H, W, C = tex.shape
    x = uv[:, 0] * W - 0.5
    y = uv[:, 1] * H - 0.5
    x0 = torch.floor(x)
    y0 = torch.floor(y)
    fx = (x - x0).unsqueeze(-1)
    fy = (y - y0).unsqueeze(-1)
    x0i = x0.long().clamp(0, W - 1)
    x1i = (x0.long() + 1).clamp(0, W - 1)
    y0i = y0.long().clamp(0, H - 1)
    y1i = (y0.long() + 1).clamp(0, H - 1)
    t00 = tex[y0i, x0i]
    t01 = tex[y0i, x1i]
    t10 = tex[y1i, x0i]
    t11 = tex[y1i, x1i]
    return (t00 * (1 - fx) * (1 - fy) + t01 * fx * (1 - fy)
            + t10 * (1 - fx) * fy + t11 * fx * fy)


def texture2d(tex: torch.Tensor, uv: torch.Tensor, level: Optional[torch.Tensor],
              n_levels: int) -> torch.Tensor:
    """dr.texture(tex[None], uv, mip_level_bias=level, boundary_mode='clamp',
    max_mip_level=n_levels-1) for tex [H,W,C], uv [N,2], level [N] -> [N,C].
    level None == filter_mode 'linear' on level 0."""
    if level is None or n_levels == 1:
        return _bilinear_clamp(tex, uv)
    chain = build_mip_chain(tex, n_levels)
    lv = level.reshape(-1).clamp(0.0, float(n_levels - 1))
    l0 = torch.floor(lv)
    f = (lv - l0).unsqueeze(-1)
    l0 = l0.long()
    l1 = (l0 + 1).clamp(max=n_levels - 1)
    out = torch.zeros(uv.shape[0], tex.shape[-1], dtype=tex.dtype, device=tex.device)
    for l in range(n_levels):
        s = _bilinear_clamp(chain[l], uv)
        w = (l0 == l).unsqueeze(-1) * (1 - f) + ((l1 == l) & (l0 != l)).unsqueeze(-1) * f
        out = out + s * w
    return out


# --------------------------------------------------------------------------
# VM feature gather  (network/fields.py:262-293, 776-806; flow.py:709-740)
# --------------------------------------------------------------------------
def vm_feature(planes: Sequence[torch.Tensor], lines: Sequence[torch.Tensor], xyz: torch.Tensor,
               level: Optional[torch.Tensor], aabb: torch.Tensor, n_levels: int) -> torch.Tensor:
    """planes[i] [1,C,H,W], lines[i] [1,C,G,1], xyz [N,3] -> [N,3C].

    u = (xyz-aabb0)/(aabb1-aabb0) (utils/network_utils.py:90), detached; plane i
    is sampled at (u[m0]->W, u[m1]->H), line i at (0, u[vec]) on a width-1 texture."""
    u = ((xyz - aabb[0]) / (aabb[1] - aabb[0])).reshape(-1, 3).detach()
    lv = None if level is None else level.reshape(-1)
    feats = []
    for i in range(3):
        m0, m1 = MAT_MODE[i]
        ptex = planes[i][0].permute(1, 2, 0)           # [H,W,C]
        ltex = lines[i][0].permute(1, 2, 0)            # [G,1,C]
        p = texture2d(ptex, torch.stack([u[:, m0], u[:, m1]], -1), lv, n_levels)
        l = texture2d(ltex, torch.stack([torch.zeros_like(u[:, 0]), u[:, VEC_MODE[i]]], -1), lv, n_levels)
        feats.append(p * l)
    return torch.cat(feats, -1)


def softplus100(x: torch.Tensor) -> torch.Tensor:
    """nn.Softplus(beta=100): identity above beta*x > 20 (fields.py:79)."""
    return F.softplus(x, beta=100.0)


class TensoSDF(nn.Module):
    """Restates network/fields.py:20-317 with sdf_multires=0 (every shipped config)."""

    def __init__(self, gridSize, aabb, sdf_n_comp=36, sdf_dim=256, app_dim=128, init_n_levels=3,
                 dtype=torch.float32):
        super().__init__()
        self.sdf_n_comp, self.sdf_dim, self.app_dim = sdf_n_comp, sdf_dim, app_dim
        self.init_radius = 0.2
        self.register_buffer("aabb", torch.as_tensor(aabb, dtype=dtype).clone())
        self.update_gridSize(torch.as_tensor(gridSize).long(), init_n_levels)
        planes, lines = [], []
        for i in range(3):                                     # fields.py:101-111
            ps = self.gridSize[list(MAT_MODE[i])]
            x = torch.linspace(-1, 1, int(ps[0]), dtype=dtype)
            y = torch.linspace(-1, 1, int(ps[1]), dtype=dtype)
            xx, yy = torch.meshgrid(x, y, indexing="ij")
            init = torch.sqrt(xx * xx + yy * yy) - self.init_radius          # fields.py:125-131
            planes.append(nn.Parameter(init[None, None].repeat(1, sdf_n_comp, 1, 1)))
            ls = int(self.gridSize[VEC_MODE[i]])
            lines.append(nn.Parameter(torch.full((1, sdf_n_comp, ls, 1), 1.0 / (3 * sdf_n_comp), dtype=dtype)))
        self.sdf_plane = nn.ParameterList(planes)
        self.sdf_line = nn.ParameterList(lines)
        self.sdf_mat = nn.Sequential(nn.Linear(3 * sdf_n_comp + 3, sdf_dim), nn.Softplus(beta=100),
                                     nn.Linear(sdf_dim, 1 + app_dim)).to(dtype)
        nn.init.constant_(self.sdf_mat[0].bias, 0.0)                          # fields.py:83-91
        nn.init.normal_(self.sdf_mat[0].weight, 0.0, math.sqrt(2) / math.sqrt(sdf_dim))
        nn.init.constant_(self.sdf_mat[-1].bias, -self.init_radius)
        nn.init.normal_(self.sdf_mat[-1].weight, mean=math.sqrt(math.pi) / math.sqrt(sdf_dim), std=0.0001)

    def update_gridSize(self, gridSize, n_levels):                            # fields.py:56-62
        self.gridSize = gridSize
        self.n_levels = int(n_levels)
        self.units = (self.aabb[1] - self.aabb[0]) / (gridSize.to(self.aabb) - 1)

    @torch.no_grad()
    def upsample_volume_grid(self, res_target):                               # fields.py:155-178
        new_levels = self.n_levels + 1
        res_target = torch.as_tensor(res_target)
        res_target = ((res_target / 2 ** (new_levels - 1)).int() * 2 ** (new_levels - 1)).long()
        for i in range(3):
            m0, m1 = MAT_MODE[i]
            self.sdf_plane[i] = nn.Parameter(F.interpolate(
                self.sdf_plane[i].data, size=(int(res_target[m1]), int(res_target[m0])),
                mode="bilinear", align_corners=True))
            self.sdf_line[i] = nn.Parameter(F.interpolate(
                self.sdf_line[i].data, size=(int(res_target[VEC_MODE[i]]), 1),
                mode="bilinear", align_corners=True))
        self.update_gridSize(res_target, new_levels)
        return res_target, self.n_levels

    def forward(self, xyz, level):                                            # fields.py:262-299
        feat = vm_feature(self.sdf_plane, self.sdf_line, xyz, level, self.aabb, self.n_levels)
        return self.sdf_mat(torch.cat([feat, xyz.reshape(-1, 3)], -1))

    def sdf(self, xyz, level=None):                                           # fields.py:148
        return self.forward(xyz, level)[..., :1]

    def gradient(self, x, level, training=False, sdf=None):                   # fields.py:227-260
        if x.shape[0] == 0:
            z = x.new_zeros(0, 3)
            return (z, x.new_zeros(0, 3)) if training else (z, None)
        eps = self.units
        taps = []
        for k in range(3):
            e = torch.zeros(3, dtype=x.dtype, device=x.device)
            e[k] = eps[k]
            taps.append((self.sdf(x + e, level), self.sdf(x - e, level)))
        g = torch.cat([(p - n) / (2 * eps[k]) for k, (p, n) in enumerate(taps)], -1)
        if not training:
            return g, None
        h = torch.cat([(p + n - 2 * sdf) / (eps[k] ** 2) for k, (p, n) in enumerate(taps)], -1)
        nh = (g * h).sum(-1) / ((g ** 2).sum(-1) + 1e-5)
        return g, nh

    def TV_loss_sdf(self, reg):                                               # fields.py:133-138
        return sum(reg(self.sdf_plane[i]) + reg(self.sdf_line[i]) for i in range(3))


def tv_loss(x: torch.Tensor, weight: float = 1.0) -> torch.Tensor:
    """network/other_field.py:170-191 (TVLoss.forward)."""
    b, c, h, w = x.shape
    count_h = c * (h - 1) * w
    count_w = c * h * (w - 1)
    total = x.new_zeros(())
    if count_h != 0:
        total = total + ((x[:, :, 1:, :] - x[:, :, :h - 1, :]) ** 2).sum() / count_h
    if count_w != 0:
        total = total + ((x[:, :, :, 1:] - x[:, :, :, :w - 1]) ** 2).sum() / count_w
    return weight * 2 * total / b


# --------------------------------------------------------------------------
# NeuS alpha  (network/shapeRenderer.py:995-1025, other_field.py:199-201)
# --------------------------------------------------------------------------
def compute_ball_radii(distance, radiis, cos):                                # shapeRenderer.py:966-970
    inverse_cos = 1.0 / cos
    tmp = (inverse_cos * inverse_cos - 1).sqrt() - radiis
    return distance * radiis * cos / (tmp * tmp + 1.0).sqrt()


def neus_alpha(sdf, gradients, dists, dirs, variance, cos_anneal_ratio):
    """sdf [N], gradients [N,3], dists [N], dirs [N,3], variance scalar tensor."""
    inv_s = torch.exp(variance * 10.0).clip(1e-6, 1e6)
    true_cos = (dirs * gradients).sum(-1)
    iter_cos = -(F.relu(-true_cos * 0.5 + 0.5) * (1.0 - cos_anneal_ratio)
                 + F.relu(-true_cos) * cos_anneal_ratio)
    est_next = sdf + iter_cos * dists * 0.5
    est_prev = sdf - iter_cos * dists * 0.5
    prev_cdf = torch.sigmoid(est_prev * inv_s)
    next_cdf = torch.sigmoid(est_next * inv_s)
    alpha = ((prev_cdf - next_cdf + 1e-5) / (prev_cdf + 1e-5)).clip(0.0, 1.0)
    return alpha, inv_s


# --------------------------------------------------------------------------
# nerfacc.render_weight_from_alpha / accumulate_along_rays  (parity unpinned)
# call sites network/shapeRenderer.py:1166-1206
# --------------------------------------------------------------------------
def render_weight_from_alpha(alpha: torch.Tensor, ray_indices: torch.Tensor, n_rays: int):
    """Packed exclusive cumprod of (1-alpha) per ray; ray_indices sorted ascending."""
    N = alpha.shape[0]
    if N == 0:
        return alpha.clone(), alpha.clone()
    counts = torch.bincount(ray_indices, minlength=n_rays)
    starts = torch.cumsum(counts, 0) - counts
    # exact sequential definition, vectorised over rays through a padded matrix
    S = int(counts.max())
    pos = torch.arange(N, device=alpha.device) - starts[ray_indices]
    pad = alpha.new_zeros(n_rays, S)
    pad[ray_indices, pos] = alpha
    trans = torch.cumprod(torch.cat([pad.new_ones(n_rays, 1), 1.0 - pad], -1), -1)[:, :-1]
    T = trans[ray_indices, pos]
    return alpha * T, T


def accumulate_along_rays(weights, values, ray_indices, n_rays):
    src = weights[:, None] if values is None else weights[:, None] * values
    out = src.new_zeros(n_rays, src.shape[-1])
    out.index_add_(0, ray_indices, src)
    return out


def composite(alpha, ray_indices, n_rays, colors, gradients, white_bg=True):
    """shapeRenderer.py:1166-1195: weights, acc, rgb (+white bg), normal."""
    weights, trans = render_weight_from_alpha(alpha, ray_indices, n_rays)
    acc = accumulate_along_rays(weights, None, ray_indices, n_rays)
    rgb = accumulate_along_rays(weights, colors, ray_indices, n_rays)
    if white_bg:
        rgb = rgb + (1 - acc)
    nrm = accumulate_along_rays(weights, gradients, ray_indices, n_rays)
    up = torch.tensor([0.0, 0.0, 1.0], dtype=nrm.dtype, device=nrm.device)
    normal = F.normalize(nrm * acc + (1.0 - acc) * up, dim=-1)
    return weights, acc, rgb, normal


def charbonnier(rgb_pr, rgb_gt, epsilon=0.001):                               # shapeRenderer.py:803-805
    return torch.sqrt(torch.sum((rgb_gt - rgb_pr) ** 2, dim=-1) + epsilon)


def linear_to_srgb(linear):                                                   # utils/raw_utils.py:4-10
    eps = torch.finfo(torch.float32).eps
    srgb0 = 323 / 25 * linear
    srgb1 = (211 * torch.clamp(linear, min=eps) ** (5 / 12) - 11) / 200
    return torch.where(linear <= 0.0031308, srgb0, srgb1)


def get_embedder(multires: int, input_dims: int = 3):
    """utils/network_utils.py:38-50: [x, sin(2^k x), cos(2^k x)]_{k<multires}."""
    freqs = [2.0 ** k for k in range(multires)]

    def embed(x):
        out = [x]
        for f in freqs:
            out += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(out, -1)
    return embed, input_dims * (1 + 2 * multires)


# --------------------------------------------------------------------------
# Shape-stage per-sample core: render_core without the shader
# (network/shapeRenderer.py:1105-1206).  `color_fn(points, normals, view, feat)`
# stands for ShapeShadingNetwork.forward.
# --------------------------------------------------------------------------
def shape_render_core(field: TensoSDF, variance: torch.Tensor, rays_o, dirs, radiis, rays_cos,
                      t_starts, t_ends, ray_indices, color_fn, cos_anneal_ratio=1.0,
                      base_radii: Optional[float] = None, is_train=True):
    n_rays = rays_o.shape[0]
    mid = (t_starts + t_ends) * 0.5
    dists = t_ends - t_starts
    ro, vd = rays_o[ray_indices], dirs[ray_indices]
    pts = ro + vd * mid[:, None]
    if base_radii is None:                                                    # shapeRenderer.py:251
        base_radii = float((field.aabb[1, 0] - field.aabb[0, 0]) / 2 / field.gridSize[0])
    ball = compute_ball_radii(mid[:, None], radiis[ray_indices], rays_cos[ray_indices])
    levels = torch.log2(ball / base_radii)
    out = field(pts, levels)
    sdf, feat = out[..., 0], out[..., 1:]
    grads, hess = field.gradient(pts, levels, training=is_train, sdf=sdf[..., None])
    alpha, inv_s = neus_alpha(sdf, grads, dists, vd, variance, cos_anneal_ratio)
    normals = F.normalize(grads, dim=-1)
    colors = color_fn(pts, normals, -vd, feat)
    weights, acc, rgb, normal = composite(alpha, ray_indices, n_rays, colors, grads)
    res = {
        "ray_rgb": rgb, "acc": acc, "normal": normal,
        "gradient_error": (torch.linalg.norm(grads, ord=2, dim=-1) - 1.0) ** 2,   # :1152
        "sdf": sdf, "alpha": alpha, "weights": weights, "gradients": grads, "feat": feat,
        "levels": levels, "points": pts, "inv_s": inv_s,
    }
    if is_train:
        res["loss_sparse"] = torch.exp(-20.0 * sdf.abs()).mean()               # :1154-1157
        res["loss_hessian"] = hess.abs().mean()                                # :1161-1162
        res["hessian"] = hess
    return res
