"""Tier-A oracle: import the reference's OWN classes from /root/reference on CPU.

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (the
build container).  It is used by oracle/gen_golden.py to produce tests/golden/*
and by tests/test_oracle_vs_reference.py to pin oracle/torch_oracle*.py against
the reference's own Python composition.  No reference source is copied: the
reference is imported at run time.

What is shimmed:
  * pure-IO imports the hot path never calls -> empty stub modules;
  * the four un-vendored native ops -> the oracle's restatements
    (nvdiffrast.torch.texture, nerfacc.render_weight_from_alpha /
    accumulate_along_rays, torch_scatter.segment_coo, raytracing);
  * numpy-2 (`np.math`) and hard-coded `.cuda()` / device='cuda'.
"""
from __future__ import annotations

import math
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("TENSOFLOW_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "network"))


_installed = False


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so "import a.b" works
    sys.modules[name] = m
    return m


def install():
    """Idempotently install the shims and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    from . import torch_oracle as O
    from . import torch_oracle_mat as OM

    if not hasattr(np, "math"):
        np.math = math
    import torchvision  # noqa: F401  (must be imported before the stubs exist)

    class _Any:
        def __init__(self, *a, **k):
            pass

        def __getattr__(self, k):
            return _Any()

        def __call__(self, *a, **k):
            return _Any()

    for name in ["mcubes", "plyfile", "skimage", "skimage.measure", "skimage.io", "skimage.metrics",
                 "h5py", "ghalton", "transforms3d", "transforms3d.axangles", "transforms3d.euler",
                 "transforms3d.quaternions", "open3d", "imageio", "_raytracing", "trimesh",
                 "matplotlib", "matplotlib.pyplot", "tensorboardX", "humanfriendly", "omegaconf",
                 "lpips", "kornia"]:
        if name not in sys.modules:
            m = _stub(name)
            def _ga(k, _n=name):
                if k.startswith("__"):
                    raise AttributeError(k)
                return _Any()
            m.__getattr__ = _ga  # type: ignore[attr-defined]
    sys.modules["skimage.io"].imread = _Any()
    sys.modules["skimage.io"].imsave = _Any()

    # ---- nvdiffrast.torch -------------------------------------------------
    def texture(tex, uv, uv_da=None, mip_level_bias=None, mip=None, filter_mode="auto",
                boundary_mode="wrap", max_mip_level=None):
        if boundary_mode == "cube":
            # tex [1,6,R,R,C], uv [1,h,w,3]
            prefix = uv.shape[:-1]
            d = uv.reshape(-1, 3)
            if mip is not None:
                stack = [tex[0]] + [m[0] for m in mip]
                out = OM.texture_cube_mip(stack, d, mip_level_bias.reshape(-1))
            else:
                out = OM.texture_cube(tex[0], d)
            return out.reshape(*prefix, -1)
        assert boundary_mode == "clamp"
        B, h, w, _ = uv.shape
        assert B == 1 and tex.shape[0] == 1
        lv = None if mip_level_bias is None else mip_level_bias.reshape(-1)
        n_levels = 1 if (lv is None or max_mip_level is None) else max_mip_level + 1
        out = O.texture2d(tex[0], uv.reshape(-1, 2), lv, n_levels)
        return out.reshape(1, h, w, -1)

    nv = _stub("nvdiffrast")
    nvt = _stub("nvdiffrast.torch", texture=texture)
    nv.torch = nvt

    # ---- nerfacc ------------------------------------------------------------
    def render_weight_from_alpha(alpha, ray_indices=None, n_rays=None, **k):
        return O.render_weight_from_alpha(alpha, ray_indices, n_rays)

    def accumulate_along_rays(weights, values=None, ray_indices=None, n_rays=None):
        return O.accumulate_along_rays(weights, values, ray_indices, n_rays)

    _stub("nerfacc", render_weight_from_alpha=render_weight_from_alpha,
          accumulate_along_rays=accumulate_along_rays, OccGridEstimator=_Any)

    # ---- torch_scatter ------------------------------------------------------
    def segment_coo(src, index, out=None, dim_size=None, reduce="sum"):
        assert reduce == "sum"
        if out is None:
            out = src.new_zeros((dim_size,) + tuple(src.shape[1:]))
        return out.index_add(0, index, src)

    _stub("torch_scatter", segment_coo=segment_coo)

    # ---- CPU placement --------------------------------------------------------
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self

    def _dev(v):
        return "cpu" if (isinstance(v, str) and v.startswith("cuda")) or \
            (isinstance(v, torch.device) and v.type == "cuda") else v

    def _wrap_factory(fn):
        def g(*a, **k):
            if "device" in k:
                k["device"] = _dev(k["device"])
            return fn(*a, **k)
        return g

    for n in ["tensor", "full", "zeros", "ones", "rand", "randn", "linspace", "arange", "empty",
              "scalar_tensor", "randperm", "zeros_like", "ones_like", "full_like", "eye"]:
        setattr(torch, n, _wrap_factory(getattr(torch, n)))

    _t_to = torch.Tensor.to

    def tensor_to(self, *a, **k):
        a = tuple(_dev(x) for x in a)
        if "device" in k:
            k["device"] = _dev(k["device"])
        return _t_to(self, *a, **k)
    torch.Tensor.to = tensor_to

    _m_to = torch.nn.Module.to

    def module_to(self, *a, **k):
        a = tuple(_dev(x) for x in a)
        if "device" in k:
            k["device"] = _dev(k["device"])
        return _m_to(self, *a, **k)
    torch.nn.Module.to = module_to

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # assets/bsdf_256_256.bin is opened with a relative path (network/fields.py:346)
    os.chdir(REFERENCE_ROOT)

    # cubemap prefilter kernels (renderutils CUDA ext) -> oracle restatement of cubemap.cu
    import network.renderutils as ru  # noqa: E402
    ru.diffuse_cubemap = OM.diffuse_cubemap
    ru.specular_cubemap = OM.specular_cubemap
    import network.renderutils.ops as ruops
    ruops.diffuse_cubemap = OM.diffuse_cubemap
    ruops.specular_cubemap = OM.specular_cubemap
    _installed = True


def reference_modules():
    """Return the reference's (fields, flow, shapeRenderer-free helpers) modules."""
    install()
    import network.fields as fields
    import network.flow as flow
    import network.other_field as other_field
    import utils.network_utils as network_utils
    return fields, flow, other_field, network_utils
