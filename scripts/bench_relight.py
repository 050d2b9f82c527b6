"""Full-image inference sharded over the GPUs of one box (BASELINE config 5).

    python scripts/bench_relight.py [--res 800]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_relight.py

One 800 x 800 view (640 000 rays), forward only, two passes as in the reference's evaluation:
  * shape `nvs` (ShapeRenderer.nvs, reference shapeRenderer.py:569-668): hierarchical NeuS sampler + fused field stencil +
    shading network + compositing, in chunks of --shape-chunk rays;
  * material `nvs` (MaterialRenderer.nvs, reference materialRenderer.py:641-752): BVH trace of the camera rays against a
    1 M-triangle mesh, 32+9-query SDF refinement of the hit + FD normal, then the MC shading network (plain + NIS estimator
    passes: (512+256) + (512+64+32) directions per hit pixel with occlusion rays), in chunks of --mat-chunk rays.
Every rank renders a contiguous slice of the pixels; the [rays, C] image tiles are all-gathered in rank order (no gradient
traffic).  Prints one JSON line: rays/s per pass and for both, max over ranks of the CUDA-event time."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", type=int, default=800)
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--shape-chunk", type=int, default=8192)
    ap.add_argument("--mat-chunk", type=int, default=8192, help="the reference hard-codes 512 rays per chunk (materialRenderer.py:705)")
    ap.add_argument("--tris-u", type=int, default=1000)
    ap.add_argument("--tris-v", type=int, default=501)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--skip-shape", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from tensoflow_b200 import _lib, synthetic
    from tensoflow_b200.shape_renderer import ShapeRenderer
    from tensoflow_b200.material import MaterialRenderer
    _lib.load()
    torch.manual_seed(6033)
    G = args.grid
    shape = ShapeRenderer(dict(device=dev, gridSize=[G // 4] * 3, sdf_n_comp=36, sdf_dim=256, app_dim=128, max_levels=1,
                               has_radiance_field=False, test_ray_num=args.shape_chunk, n_samples=64, n_importance=64))
    shape.upsample_sdf_grid([G // 2] * 3)
    shape.upsample_sdf_grid([G] * 3)
    synthetic.perturb_field(shape.sdf_network, seed=1, noise=2e-3)
    shape.color_network.envlight.build_mips()
    verts, tris = synthetic.bumpy_sphere(args.tris_u, args.tris_v, r=0.2, bump=0.05)     # near the zero set of the initial SDF
    mat = MaterialRenderer(dict(device=dev, gridSize=[G] * 3, nvs_ray_num=args.mat_chunk,
                                shader_cfg=dict(diffuse_sample_num=512, specular_sample_num=256, nis_diffuse_sample_num=64,
                                                nis_specular_sample_num=32, light_reso=128, gridSize=[G] * 3, mat_grid=G)),
                           verts, tris)
    mat.init_sdf(shape.ckpt_to_save())
    sh = mat.shader_network
    with torch.no_grad():
        sh.outer_light.base.add_(0.5 * torch.randn_like(sh.outer_light.base))
    sh.update_step(2000)
    sh.use_flow_diffuse_copy = sh.use_flow_specular_copy = True
    sh.outer_light.build_mips_direct()
    h = w = args.res
    f = 1.6 * w                                   # the object (radius 0.2 at distance 1.2) fills ~half of the frame
    K = np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1]], np.float32)
    pose = np.array([[1, 0, 0, 0.0], [0, 1, 0, 0.0], [0, 0, 1, 1.2]], np.float32)

    def timed(fn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = _lib.launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.reps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / args.reps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t), out, (_lib.launch_count() - l0) / args.reps

    # warm-up on a small view (module / allocator initialisation)
    shape.nvs(pose, K * np.array([[0.08], [0.08], [1]], np.float32), 64, 64, rank=rank, world=world, perturb_overwrite=0)
    mat.nvs(pose, K * np.array([[0.08], [0.08], [1]], np.float32), 64, 64, rank=rank, world=world)
    ms_shape, img_s, l_s = (0.0, None, 0) if args.skip_shape else timed(
        lambda: shape.nvs(pose, K, h, w, rank=rank, world=world, perturb_overwrite=0))
    ms_mat, img_m, l_m = timed(lambda: mat.nvs(pose, K, h, w, rank=rank, world=world))
    if rank == 0:
        hit = float((np.abs(img_m["color"] - 1.0).max(-1) > 0).mean())
        line = {"metric": "inference rays/sec (forward only, full image)", "value": h * w / ((ms_shape + ms_mat) / 1e3), "unit": "rays/s",
                "n_gpus": world, "higher_is_better": True, "scaling": "strong",
                "shape_nvs": {"ms": ms_shape, "rays_per_s": h * w / (ms_shape / 1e3) if ms_shape else None, "gpu_launches": l_s,
                              "acc_mean": None if img_s is None else float(img_s["acc"].mean())},
                "material_nvs": {"ms": ms_mat, "rays_per_s": h * w / (ms_mat / 1e3), "gpu_launches": l_m, "hit_fraction": hit},
                "config": {"workload": f"relight inference: {h}x{w} image ({h * w} rays), shape nvs (64+64 samples/ray, VM field {G}^3) + "
                                       f"material nvs (trace vs {tris.shape[0]} triangles + SDF refinement + MC shading), "
                                       f"{(h * w + world - 1) // world} rays per GPU, tiles all-gathered",
                           "chunks": {"shape": args.shape_chunk, "material": args.mat_chunk}}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
