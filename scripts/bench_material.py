"""Material-stage measurement (BASELINE config 3): 8192 surface points on a 1M-triangle bumpy
sphere, diffuse 512 cosine + 64 flow-sampled directions, specular 32 flow-sampled directions, trainable env
cubemap 128^2, occlusion rays through the BVH, NIS losses on; fwd + bwd of MaterialRenderer.forward.

    python scripts/bench_material.py [--points 8192] [--steps 5] [--tris-u 1000 --tris-v 500] [--cpu-points 64]

Prints one JSON line (points/s, per-call CUDA-event times, launches) and, with --cpu-points > 0, the oracle
port timed on the host cores on a bounded sample of the same workload."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bumpy_sphere(nu, nv, r=0.5):
    u = torch.linspace(0, 2 * np.pi, nu + 1)[:-1]
    v = torch.linspace(0.05, np.pi - 0.05, nv)
    vv, uu = torch.meshgrid(v, u, indexing="ij")
    rad = r * (1 + 0.1 * torch.sin(5 * uu) * torch.sin(4 * vv))
    verts = torch.stack([rad * torch.sin(vv) * torch.cos(uu), rad * torch.sin(vv) * torch.sin(uu), rad * torch.cos(vv)], -1).reshape(-1, 3)
    i = torch.arange(nv - 1)[:, None]
    j = torch.arange(nu)[None, :]
    a, b = i * nu + j, i * nu + (j + 1) % nu
    c, d = (i + 1) * nu + j, (i + 1) * nu + (j + 1) % nu
    tris = torch.cat([torch.stack([a, c, b], -1).reshape(-1, 3), torch.stack([b, c, d], -1).reshape(-1, 3)], 0)
    return verts.float(), tris.to(torch.int32)


def make_batch(verts, pn, seed=0):
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, verts.shape[0], (pn,), generator=g)
    pts = verts[idx] * 1.001
    normals = F.normalize(pts, dim=-1)
    cams = F.normalize(torch.randn(pn, 3, generator=g) + 2 * normals, dim=-1) * 2.0
    rays_d = F.normalize(pts - cams, dim=-1)
    rgb = torch.rand(pn, 3, generator=g)
    noise = dict(az_diffuse=torch.rand(pn, 1, 1, generator=g), az_specular=torch.rand(pn, 1, 1, generator=g),
                 phi_diffuse=torch.rand(pn, 64, 1, generator=g), phi_specular=torch.rand(pn, 32, 1, generator=g))
    return dict(inters=pts, normals=normals, rays_d=rays_d, rgb=rgb), noise


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tris-u", type=int, default=1000)
    ap.add_argument("--tris-v", type=int, default=501)
    ap.add_argument("--grid", type=int, default=512)
    args = ap.parse_args()
    from tensoflow_b200 import _lib, ops
    from tensoflow_b200.material import MaterialRenderer
    dev = torch.device("cuda:0")
    torch.manual_seed(6033)
    t0 = time.perf_counter()
    verts, tris = bumpy_sphere(args.tris_u, args.tris_v)
    cfg = dict(train_ray_num=args.points, device=dev, gridSize=[args.grid] * 3,
               shader_cfg=dict(diffuse_sample_num=512, specular_sample_num=256, nis_diffuse_sample_num=64, nis_specular_sample_num=32,
                               light_reso=128, gridSize=[args.grid] * 3, mat_grid=args.grid))
    r = MaterialRenderer(cfg, verts, tris)
    t_build = time.perf_counter() - t0
    sh = r.shader_network
    with torch.no_grad():
        for p in list(sh.mat_plane) + list(sh.flow_diffuse.parameters()) + list(sh.flow_specular.parameters()):
            if p.dim() == 4:
                p.add_(1e-2 * torch.randn_like(p))
        sh.outer_light.base.add_(0.5 * torch.randn_like(sh.outer_light.base))
    step = 2000                     # flows active, NIS losses on (fields.py:1050-1068 thresholds)
    sh.update_step(step)
    sh.use_flow_diffuse_copy = sh.use_flow_specular_copy = True
    batch, noise = make_batch(verts, args.points)
    host = {k: v.pin_memory() for k, v in batch.items()}
    noise = {k: v.to(dev) for k, v in noise.items()}
    params = [p for p in r.parameters() if p.requires_grad]

    def one_step():
        for p in params:
            p.grad = None
        r.set_train_batch(host)
        out = r({"step": step, "noise": noise})
        loss = out["loss_rgb"].mean()
        for k in ("loss_mat_reg", "loss_diffuse_light", "loss_nis_diffuse", "loss_nis_specular"):
            if k in out and out[k] is not None:
                loss = loss + out[k].mean()
        loss.backward()
        return float(loss.detach().cpu())

    for _ in range(max(args.warmup, 3)):
        one_step()
    torch.cuda.synchronize()
    ops.KernelTimers.reset(True)
    l0 = _lib.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if os.environ.get('TF_PROFILE_RANGE'):          # ncu --profile-from-start off: only the timed steps are captured
        torch.cuda.profiler.start()
    a.record()
    for _ in range(args.steps):
        loss = one_step()
    b.record()
    torch.cuda.synchronize()
    if os.environ.get('TF_PROFILE_RANGE'):
        torch.cuda.profiler.stop()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / args.steps
    calls = {k: (round(v[0] / args.steps, 3), v[1] // args.steps) for k, v in ops.KernelTimers.totals_ms().items()}
    ops.KernelTimers.reset(False)
    D = 512 + 64 + 32
    line = {"metric": "material-stage train points/sec (fwd+bwd, e2e from pinned host batch)", "value": args.points / (ms / 1e3), "unit": "points/s",
            "pairs_per_s": args.points * D / (ms / 1e3), "ms_per_step": ms, "steps": args.steps,
            "config": {"workload": f"material stage: {args.points} surface points x (512 cosine + 64 flow + 32 flow) directions, "
                                   f"{tris.shape[0]} triangles, env cubemap 128^2, mat/flow grids {args.grid}^2 x 36/16"},
            "bvh_build_s": round(t_build, 2), "calls_ms_per_step": calls, "gpu_launches_per_step": (_lib.launch_count() - l0) / args.steps,
            "loss": loss}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
