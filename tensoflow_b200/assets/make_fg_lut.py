"""Generate the split-sum FG ("DFG") look-up table used by the shape-stage shader
(reference network/fields.py:346,520-523 loads assets/bsdf_256_256.bin, a [256,256,2] fp32
table indexed by (N.V along x, roughness along y)).

The table is the pre-integrated specular BRDF with the GGX distribution (alpha = roughness^2),
the height-correlated Smith visibility term and Schlick's Fresnel weight split into
scale (1-Fc) and bias (Fc):  A,B = 4/N sum_i V(NoV,NoL,alpha) NoL VoH/NoH * {(1-Fc), Fc}
over GGX-importance-sampled half vectors.  We integrate on a dense midpoint grid in the
(u1,u2) sampling square instead of a random sequence, which converges much faster.

    python -m tensoflow_b200.assets.make_fg_lut [n_side]   -> tensoflow_b200/assets/fg_lut_256.npy
"""
import os
import sys

import numpy as np
import torch


def make(res=256, n_side=512, chunk=8):
    torch.set_grad_enabled(False)
    xs = (torch.arange(res, dtype=torch.float64) + 0.5) / res
    u = (torch.arange(n_side, dtype=torch.float64) + 0.5) / n_side
    u1, u2 = torch.meshgrid(u, u, indexing="ij")
    u1, u2 = u1.reshape(-1), u2.reshape(-1)
    cphi = torch.cos(2 * np.pi * u1)
    out = torch.zeros(res, res, 2, dtype=torch.float64)
    for r0 in range(0, res, chunk):
        R = xs[r0:r0 + chunk][:, None, None]
        NoV = xs[None, :, None]
        a = R * R
        cos2 = (1 - u2) / (1 + (a * a - 1) * u2)
        cosT, sinT = torch.sqrt(cos2), torch.sqrt(1 - cos2)
        Hx, Hz = sinT * cphi, cosT
        Vx, Vz = torch.sqrt(1 - NoV ** 2), NoV
        VoH = (Vx * Hx + Vz * Hz).clamp(0, 1)
        NoL = (2 * VoH * Hz - Vz).clamp(0, 1)
        NoH = Hz.clamp(0, 1)
        a2 = a * a
        ggxl = NoV * torch.sqrt((NoL - NoL * a2) * NoL + a2)
        ggxv = NoL * torch.sqrt((NoV - NoV * a2) * NoV + a2)
        v = 0.5 / (ggxv + ggxl + 1e-30) * NoL * (VoH / (NoH + 1e-30))
        Fc = (1 - VoH) ** 5
        m = (NoL > 0).to(v.dtype)
        out[r0:r0 + chunk, :, 0] = 4 * (m * v * (1 - Fc)).mean(-1)
        out[r0:r0 + chunk, :, 1] = 4 * (m * v * Fc).mean(-1)
    return out.float().numpy()


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    lut = make(n_side=n)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fg_lut_256.npy")
    np.save(path, lut)
    print(path, lut.shape, lut[0, 0], lut[255, 0])
