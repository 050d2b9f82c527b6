"""TEST INFRASTRUCTURE ONLY.  Golden fixtures of the cubemap prefilter from the REFERENCE's own CUDA kernels.

Runs on the GPU box (needs a CUDA device):

    python oracle/gen_golden_prefilter.py gpurun_out/prefilter.npz      # then copied to tests/golden/prefilter.npz

Drives oracle/_ref/libref_cubemap.so (the reference's network/renderutils/c_src/cubemap.cu compiled by oracle/build_ref.py)
through the Python logic of the reference's network/renderutils/ops.py:391-458 (restated here because that module imports
its torch plugin at call time): diffuse_cubemap(x), specular_cubemap(x, roughness, cutoff=0.99) = filtered[..., :3] /
filtered[..., 3:] with the GGX-lobe cutoff of __ndfBounds, forward and backward, at the resolutions and roughness values
EnvLight.build_mips uses (network/light.py:52-66: 128 -> 0.08, 64 -> 0.29, 32 -> 0.5, 16 -> 1.0, diffuse at 16).
Inputs are regenerated from the seed by the test; outputs / gradients of the 64^2 and 128^2 faces are stored at every
4th texel to keep the fixture small."""
import ctypes as C
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [(128, 0.08), (64, 0.29), (32, 0.5), (16, 1.0)]


def make_input(res, seed):
    """log-radiance cubemap the test regenerates bit for bit (CPU generator)."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(6, res, res, 3, generator=g) * 0.7 - 0.5).float()


def ndf_cutoff(roughness, cutoff):
    """ops.py:427-441 (__ndfBounds): cos(theta) keeping `cutoff` of the GGX lobe's energy."""
    def ndf_ggx(a2, c):
        c = np.clip(c, 0.0, 1.0)
        d = (c * a2 - c) * c + 1.0
        return a2 / (d * d * np.pi)
    n = 1000000
    costheta = np.cos(np.linspace(0, np.pi / 2.0, n))
    D = np.cumsum(ndf_ggx(roughness ** 4, costheta))
    idx = np.argmax(D >= D[..., -1] * cutoff)
    return float(costheta[idx])


def main(out_path):
    lib = C.CDLL(os.path.join(HERE, "_ref", "libref_cubemap.so"))
    P = C.c_void_p
    dev = torch.device("cuda:0")
    ptr = lambda t: P(t.data_ptr())

    def check(rc):
        assert rc == 0, f"reference kernel failed: cuda error {rc}"

    data = {}
    for res, rough in CASES:
        x = make_input(res, 1000 + res).to(dev)
        sub = 4 if res >= 64 else 1
        g = torch.Generator().manual_seed(2000 + res)
        # ---- specular_cubemap (ops.py:443-458, 413-425) --------------------------------------------------
        cos_cut = ndf_cutoff(rough, 0.99)
        bounds = torch.zeros(6, res, res, 24, device=dev)
        check(lib.ref_specular_bounds(res, C.c_float(cos_cut), ptr(bounds)))
        filt = torch.empty(6, res, res, 4, device=dev)
        check(lib.ref_specular_cubemap_fwd(ptr(x), ptr(bounds), C.c_float(rough), C.c_float(cos_cut), res, ptr(filt)))
        filt.requires_grad_()
        spec = filt[..., 0:3] / filt[..., 3:]
        u = torch.randn(6, res, res, 3, generator=g).to(dev)
        (spec * u).sum().backward()
        dx = torch.zeros(6, res, res, 3, device=dev)
        check(lib.ref_specular_cubemap_bwd(ptr(x), ptr(bounds), ptr(filt.grad.contiguous()), C.c_float(rough), C.c_float(cos_cut), res, ptr(dx)))
        data[f"spec{res}_out"] = spec.detach()[:, ::sub, ::sub].cpu().numpy()
        data[f"spec{res}_dx"] = dx[:, ::sub, ::sub].cpu().numpy()
        data[f"spec{res}_meta"] = np.array([res, rough, cos_cut, sub], dtype=np.float64)
        # ---- diffuse_cubemap (ops.py:391-411) -----------------------------------------------------------------
        if res <= 32:
            out = torch.empty(6, res, res, 3, device=dev)
            check(lib.ref_diffuse_cubemap_fwd(ptr(x), res, ptr(out)))
            ud = torch.randn(6, res, res, 3, generator=g).to(dev)
            dxd = torch.zeros(6, res, res, 3, device=dev)
            check(lib.ref_diffuse_cubemap_bwd(ptr(x), ptr(ud.contiguous()), res, ptr(dxd)))
            data[f"diff{res}_out"] = out.cpu().numpy()
            data[f"diff{res}_dx"] = dxd.cpu().numpy()
    np.savez(out_path, **data)
    print("wrote", out_path, {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prefilter.npz")
