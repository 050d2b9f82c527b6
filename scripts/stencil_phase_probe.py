"""Phase timing of the tcgen05 stencil kernels at the config-2 shape (512^3 field, C=36, H=256, 3 mip levels).

    python -m tensoflow_b200.build --experiment          # libtensoflow_b200_exp.so with -DTF_TC_DEBUG_SWITCHES
    python scripts/stencil_phase_probe.py [--rays 2048] [--fwd 0,1,2,4] [--bwd 0,1,2,4,8,16]

Each switch REMOVES one phase of the kernel (the results are then wrong by construction): the time that disappears is
what the phase costs on the critical path.  Forward: 1 no gather, 2 no MMAs, 4 no epilogue math.  Backward: 1 no gather,
2 no workspace stores, 4 no scatter, 8 no dW1 reduction, 16 no chunk loop.  Without the experiment library (or with
--product) only the unmodified kernels are timed.  Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--rays", type=int, default=2048)
ap.add_argument("--fwd", default="0,1,2,4,3,5,6")
ap.add_argument("--bwd", default="0,1,2,4,8,16,5")
ap.add_argument("--product", action="store_true")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--levels", type=int, default=0)
args = ap.parse_args()
exp = os.path.join(ROOT, "tensoflow_b200", "libtensoflow_b200_exp.so")
if not args.product and os.path.exists(exp):
    os.environ["TENSOFLOW_B200_LIB"] = exp
else:
    args.fwd = args.bwd = "0"

import torch  # noqa: E402

import bench  # noqa: E402
import ctypes as C  # noqa: E402

from tensoflow_b200 import _lib, synthetic  # noqa: E402

dev = torch.device("cuda:0")
cfg = dict(bench.SHAPE_CFG)
cfg["rays"] = args.rays
if args.levels:
    cfg["L"] = args.levels
field, variance = bench.build_shape(cfg, dev)
params = list(field.parameters()) + [variance]
rays = synthetic.make_rays(cfg["rays"], seed=50, device=dev)


def timed(env_fwd, env_bwd):
    os.environ["TF_TC_DEBUG"] = str(env_fwd)
    os.environ["TF_TC_BWD_DEBUG"] = str(env_bwd)
    for _ in range(2):
        for p in params:
            p.grad = None
        bench.shape_step(field, variance, rays, cfg)
    torch.cuda.synchronize()
    lib = _lib.load()
    lib.tf_kernel_timing_reset()
    lib.tf_kernel_timing_enable(1)
    for _ in range(args.reps):
        for p in params:
            p.grad = None
        bench.shape_step(field, variance, rays, cfg)
    torch.cuda.synchronize()
    t = {}
    for name in ("sdf_stencil_fwd_tc", "sdf_stencil_bwd_tc", "xty_tc", "linear_tc"):
        t_ms, n = C.c_double(0.0), C.c_int32(0)
        if lib.tf_kernel_timing_read(name.encode(), C.byref(t_ms), C.byref(n)) == 0 and n.value:
            t[name] = round(t_ms.value / args.reps, 3)
    lib.tf_kernel_timing_enable(0)
    lib.tf_kernel_timing_reset()
    return t


out = {"rays": args.rays, "samples": args.rays * cfg["samples"], "lib": os.environ.get("TENSOFLOW_B200_LIB", "product"), "fwd": {}, "bwd": {}}
for m in [int(x) for x in args.fwd.split(",")]:
    out["fwd"][m] = timed(m, 0).get("sdf_stencil_fwd_tc")
for m in [int(x) for x in args.bwd.split(",")]:
    if m == 0 and 0 in out["fwd"]:
        continue
    out["bwd"][m] = timed(0, m).get("sdf_stencil_bwd_tc")
out["bwd"][0] = timed(0, 0).get("sdf_stencil_bwd_tc")
print(json.dumps(out), flush=True)
if os.environ.get("TF_TC_BWD_PROF"):      # per-phase cycle counts of CTA 0 (stderr lines "[bwd prof]"), one more step
    timed(0, 0)
