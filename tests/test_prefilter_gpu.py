"""Cubemap prefilter (EnvLight.build_mips, reference network/light.py:52-66) against fixtures produced by the REFERENCE's
own CUDA kernels (network/renderutils/c_src/cubemap.cu:110-350 compiled by oracle/build_ref.py, driven by
oracle/gen_golden_prefilter.py on the B200 box -> tests/golden/prefilter.npz): diffuse_cubemap at 16^2 / 32^2 and
specular_cubemap at 128^2 / 64^2 / 32^2 / 16^2 with the roughness schedule of build_mips, forward and backward."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prefilter.npz")


def _setup():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(GOLDEN):
        pytest.skip("tests/golden/prefilter.npz not generated yet (oracle/gen_golden_prefilter.py on the GPU box)")
    return torch.device("cuda:0"), np.load(GOLDEN)


def _input(res, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(6, res, res, 3, generator=g) * 0.7 - 0.5).float()


@pytest.mark.parametrize("res", [128, 64, 32, 16])
def test_specular_cubemap_matches_reference_kernels(res):
    from tensoflow_b200.shape_shader import specular_cubemap
    dev, gold = _setup()
    _, rough, cos_cut, sub = gold[f"spec{res}_meta"]
    sub = int(sub)
    x = _input(res, 1000 + res).to(dev).requires_grad_()
    g = torch.Generator().manual_seed(2000 + res)
    u = torch.randn(6, res, res, 3, generator=g).to(dev)
    out = specular_cubemap(x, float(rough), 0.99)
    (out * u).sum().backward()
    want, want_dx = torch.from_numpy(gold[f"spec{res}_out"]), torch.from_numpy(gold[f"spec{res}_dx"])
    got, got_dx = out[:, ::sub, ::sub].detach().cpu(), x.grad[:, ::sub, ::sub].cpu()
    if res < 128:
        assert rel_err(got, want) < 1e-4
        assert rel_err(got_dx, want_dx) < 1e-3
    else:
        # roughness 0.08 at 128^2 is ill-conditioned in fp32 for BOTH implementations: alpha^2 = 0.08^4 = 4.1e-5 and the GGX
        # denominator d = (c alpha^2 - c) c + 1 of the few texels inside the lobe cancels down to ~alpha^2, so one ulp of
        # the half-vector cosine c (6e-8; the reference normalises H with its safeNormalize inside the kernel, the operator
        # here with F.normalize) moves d by ~3e-3 relative and the weight ~ 1/d^2 by ~6e-3.  The reference's own fp32
        # result carries that noise; the comparison is therefore made at that level (measured: 2.4e-3 max, ~5e-4 typical),
        # and the well-conditioned levels above (64^2 / 32^2 / 16^2, roughness >= 0.29) at 1e-4 / 1e-3.
        assert rel_err(got, want) < 8e-3
        assert rel_err(got_dx, want_dx) < 2e-2


@pytest.mark.parametrize("res", [32, 16])
def test_diffuse_cubemap_matches_reference_kernels(res):
    from tensoflow_b200.shape_shader import diffuse_cubemap
    dev, gold = _setup()
    x = _input(res, 1000 + res).to(dev).requires_grad_()
    g = torch.Generator().manual_seed(2000 + res)
    torch.randn(6, res, res, 3, generator=g)                  # the generator drew the specular upstream gradient first
    ud = torch.randn(6, res, res, 3, generator=g).to(dev)
    out = diffuse_cubemap(x)
    (out * ud).sum().backward()
    assert rel_err(out, torch.from_numpy(gold[f"diff{res}_out"])) < 1e-4
    assert rel_err(x.grad, torch.from_numpy(gold[f"diff{res}_dx"])) < 1e-3
