"""TEST INFRASTRUCTURE ONLY (never imported by tensoflow_b200/).  CPU restatement of the occupancy-grid marcher the reference
gets from nerfacc.OccGridEstimator.sampling (network/shapeRenderer.py:950-959, 1065-1072) under the rule stated in
include/tensoflow_b200.h (tf_occ_march_*).  PARITY UNPINNED: nerfacc is not vendored by the reference and the reference holds
no test or fixture at this boundary; the rule is a restatement of nerfacc's documented behaviour (fixed-step marching with
empty-cell skipping, packed output)."""
import math

import torch


def occ_march(rays_o, rays_d, near, far, step, aabb, res, binaries):
    """rays_o/rays_d [R,3] fp32, near [R] fp32, far / step floats, aabb [6], res [3] ints, binaries bool [rx,ry,rz]
    -> (ray_indices int64 [N], t_starts [N], t_ends [N]) packed ray after ray.  Every float op is one rounded fp32 op in the
    order the kernel uses, so keep / drop decisions are bit-identical."""
    o, d, near = rays_o.float(), rays_d.float(), near.float()
    lo, hi = torch.tensor(aabb[:3], dtype=torch.float32), torch.tensor(aabb[3:], dtype=torch.float32)
    resf = torch.tensor([float(r) for r in res], dtype=torch.float32)
    step32 = torch.tensor(step, dtype=torch.float32)
    far32 = torch.tensor(far, dtype=torch.float32)
    reach = float((o.norm(dim=-1) + (hi - lo).norm() + lo.abs().max() + hi.abs().max()).max())
    K = int(math.ceil((min(float(far32), reach) - float(near.min())) / float(step32))) + 4
    k = torch.arange(max(K, 0), dtype=torch.float32)
    t0 = near[:, None] + k[None, :] * step32
    t1 = t0 + step32
    mid = (t0 + t1) * 0.5
    keep = mid < far32
    p = o[:, None, :] + d[:, None, :] * mid[..., None]
    u = (p - lo) / (hi - lo) * resf
    keep &= ((u >= 0) & (u < resf)).all(-1)
    c = torch.floor(u).long().clamp_min(0)
    c = torch.minimum(c, torch.tensor([r - 1 for r in res]))
    keep &= binaries[c[..., 0], c[..., 1], c[..., 2]]
    ray = torch.arange(o.shape[0])[:, None].expand_as(keep)
    return ray[keep], t0[keep], t1[keep]
