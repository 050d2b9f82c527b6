"""Shape-stage train step through the module API: ShapeRenderer.forward = hierarchical NeuS
sampler (64 + 4 x 16 samples, no-grad SDF-only queries) -> fused field stencil -> shading network (prefiltered env
light, material / light MLPs) -> NeuS alpha + compositing -> eikonal / sparse / hessian / TV / occlusion losses, then
backward.  Compressor-scale field (512^3, C=36, H=256, A=128, 3 mip levels).

    python scripts/bench_shape_renderer.py [--rays 4096] [--steps 5]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--grid", type=int, default=512)
    args = ap.parse_args()
    from tensoflow_b200 import _lib, ops, synthetic
    from tensoflow_b200.shape_renderer import ShapeRenderer
    dev = torch.device("cuda:0")
    torch.manual_seed(6033)
    G0 = args.grid >> 2
    cfg = dict(device=dev, gridSize=[G0] * 3, sdf_n_comp=36, sdf_dim=256, app_dim=128, max_levels=1, train_ray_num=args.rays,
               has_radiance_field=True, radiance_field_step=20000, occ_loss_step=20000, n_samples=64, n_importance=64, up_sample_steps=4)
    m = ShapeRenderer(cfg)
    for l in (1, 2):                       # the reference's upsampling schedule: 128 -> 256 -> 512, one more mip level each time
        m.sdf_network.upsample_volume_grid(torch.tensor([G0 << l] * 3))
        m.update_stepSize(torch.tensor([G0 << l] * 3), l + 1)
    synthetic.perturb_field(m.sdf_network, seed=1, noise=1e-2)
    rays = synthetic.make_rays(args.rays * 4, seed=1)
    m.set_train_batch({k: v.pin_memory() for k, v in dict(rays_o=rays["rays_o"], rays_d=rays["dirs"], dirs=rays["dirs"], radiis=rays["radiis"],
                                                           rays_cos=rays["rays_cos"], rgbs=rays["rgbs"]).items()})
    params = [p for p in m.parameters() if p.requires_grad]
    step = 30000

    def one_step():
        for p in params:
            p.grad = None
        out = m({'step': step})
        loss = out['loss_rgb'].mean() + 0.1 * out['gradient_error'].mean() + 0.1 * out['loss_tv_sdf']
        for k in ('loss_sparse', 'loss_hessian', 'loss_occ', 'loss_radiance'):
            if k in out and out[k] is not None:
                loss = loss + 0.01 * out[k].mean()
        loss.backward()
        return float(loss.detach().cpu()), float(out['sample_num']) if 'sample_num' in out else 0.0

    for _ in range(3):
        one_step()
    torch.cuda.synchronize()
    ops.KernelTimers.reset(True)
    l0 = _lib.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if os.environ.get('TF_PROFILE_RANGE'):          # ncu --profile-from-start off: only the timed steps are captured
        torch.cuda.profiler.start()
    a.record()
    for _ in range(args.steps):
        loss, ns = one_step()
    b.record()
    torch.cuda.synchronize()
    if os.environ.get('TF_PROFILE_RANGE'):
        torch.cuda.profiler.stop()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    calls = {k: (round(v[0] / args.steps, 3), v[1] // args.steps) for k, v in ops.KernelTimers.totals_ms().items()}
    ops.KernelTimers.reset(False)
    print(json.dumps({"metric": "shape-stage train rays/sec through ShapeRenderer.forward (fwd+bwd, e2e from pinned host rays)",
                      "value": args.rays / (ms / 1e3), "unit": "rays/s", "ms_per_step": ms, "rays": args.rays,
                      "samples_per_ray_after_culling": ns, "grid": args.grid,
                      "calls_ms_per_step": dict(sorted(calls.items(), key=lambda kv: -kv[1][0])[:14]),
                      "gpu_launches_per_step": (_lib.launch_count() - l0) / args.steps, "loss": loss}), flush=True)


if __name__ == "__main__":
    main()
