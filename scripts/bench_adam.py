"""Optimizer-step measurement: Adam over the config-2 VM factors + decoder MLP (3 planes 36 x 512^2, 3 lines
36 x 512, MLP 111-256-129: 28.5 M fp32 parameters, 114 MB) with tf_adam_step against torch.optim.Adam (foreach, the reference's
optimizer) and torch's fused=True variant.  HBM-bound: 28 B/element algorithmic (read p, g, m, v; write p, m, v).
Gradients are refreshed between steps by a copy that also evicts L2 (the working set, 456 MB, exceeds the 126 MB L2 anyway).

    python scripts/bench_adam.py [--steps 20]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--comps", type=int, default=36)
    args = ap.parse_args()
    import bench as B
    from tensoflow_b200 import _lib
    from tensoflow_b200.optim import FusedAdam
    dev = torch.device("cuda:0")
    _lib.load()
    cfg = dict(B.SHAPE_CFG)
    cfg.update(G=args.grid, C=args.comps)
    res = {}
    n_elem = 0
    for name in ("tf_adam_step", "torch_foreach", "torch_fused"):
        field, variance = B.build_shape(cfg, dev)
        params = list(field.parameters()) + [variance]
        n_elem = sum(p.numel() for p in params)
        groups = [{'params': params[:6], 'lr': 1e-2}, {'params': params[6:], 'lr': 1e-3}]
        if name == "tf_adam_step":
            opt = FusedAdam(groups, betas=(0.9, 0.99))
        else:
            opt = torch.optim.Adam(groups, betas=(0.9, 0.99), foreach=(name == "torch_foreach"), fused=(name == "torch_fused"))
        for p in params:
            p.grad = torch.randn_like(p) * 1e-3
        for _ in range(3):
            opt.step()
        torch.cuda.synchronize()
        # back to back: the launches of step i+1 are queued while step i runs, so the figure is the device's, not the host's
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            opt.step()
        b.record()
        torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / args.steps
        del opt, field, params
        torch.cuda.empty_cache()
    pk = B.peaks()
    by = 28 * n_elem
    ms = res["tf_adam_step"]
    print(json.dumps({"kernel": "adam_kernel (tf_adam_step)", "elements": n_elem, "algorithmic_bytes": by, "ms": res,
                      "roofline": {"bound": "hbm", "achieved": by / (ms / 1e3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                                   "frac": by / (ms / 1e3) / 1e9 / pk["hbm"], "peak_source": pk["src"]},
                      "speedup_vs_torch_foreach": res["torch_foreach"] / ms, "speedup_vs_torch_fused": res["torch_fused"] / ms}), flush=True)


if __name__ == "__main__":
    main()
