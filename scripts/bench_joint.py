"""Joint shape + material train step sharded over the GPUs of one box (BASELINE config 4).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_joint.py [--rays 65536]
    torchrun ... scripts/bench_joint.py --check          # gradient equivalence with the single-GPU full batch (small sizes)

R rays (default 65 536) are split into contiguous slices of R/N rays per rank; every rank runs the config-2 shape step on
its rays and the config-3 material step on as many surface points (micro-batches of --micro units, gradients accumulate), then
ONE flat fp32 sum-allreduce (dist.FlatGradBucket) combines VM-factor, MLP, env-map and flow gradients.  Strong scaling: the
total work is fixed, so `value` = R / step time.  Sample-normalised means are normalised by the GLOBAL counts (one small
allreduce before the backward), so the summed gradient equals the single-GPU gradient of the full batch: `--check` verifies
that to 1e-3 relative on reduced sizes (every rank also runs the full batch alone)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def shape_loss_sums(field, variance, rays, cfg):
    """Returns (sum over rays of the charbonnier colour loss, sum over samples of the eikonal error, #samples)."""
    from tensoflow_b200 import synthetic
    from tensoflow_b200.shape_renderer import render_core, charbonnier
    t0, t1, idx = synthetic.uniform_samples(rays["rays_o"], rays["dirs"], field.aabb, cfg["samples"])
    out = render_core(field, variance, synthetic.simple_color_fn, rays["rays_o"], rays["dirs"], rays["radiis"], rays["rays_cos"],
                      t0, t1, idx, cos_anneal_ratio=1.0)
    return charbonnier(out["ray_rgb"], rays["rgbs"]).sum(), out["gradient_error"].sum(), idx.shape[0]


def material_loss_sums(mr, batch, noise, step, full_losses):
    """Sum over points of the colour loss (+ the per-point regularisers when `full_losses`)."""
    mr.shader_network.update_step(step)
    out = mr.shade(batch["inters"], -batch["rays_d"], batch["normals"], None, True, step, noise=noise)
    loss = mr.compute_rgb_loss(out["rgb_pr"], batch["rgb"]).sum()
    extra = 0.0
    if full_losses:
        pn = batch["inters"].shape[0]
        loss = loss + mr.compute_diffuse_light_regularization(out["diffuse_light"]).sum()
        # NIS losses are means over the local points: weight by the local point count (exact when ranks hold equal counts)
        extra = (out["loss_nis_diffuse"] + out["loss_nis_specular"]) * pn
    return loss + extra


def build(args, dev):
    import bench as B
    from bench_material import make_batch
    from tensoflow_b200.material import MaterialRenderer
    from tensoflow_b200.synthetic import bumpy_sphere
    cfg = dict(B.SHAPE_CFG)
    cfg.update(G=args.grid, samples=args.samples)
    if args.check:
        cfg.update(C=8, H=64, L=2)
    field, variance = B.build_shape(cfg, dev)
    torch.manual_seed(6033)
    verts, tris = bumpy_sphere(args.tris_u, args.tris_v)
    mg = args.mat_grid
    mcfg = dict(train_ray_num=args.micro, device=dev, gridSize=[mg] * 3,
                shader_cfg=dict(diffuse_sample_num=args.diffuse, specular_sample_num=256, nis_diffuse_sample_num=64,
                                nis_specular_sample_num=32, light_reso=128 if not args.check else 16, gridSize=[mg] * 3, mat_grid=mg))
    mr = MaterialRenderer(mcfg, verts, tris)
    sh = mr.shader_network
    g = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for p in list(sh.mat_plane) + list(sh.flow_diffuse.parameters()) + list(sh.flow_specular.parameters()):
            if p.dim() == 4:
                p.add_((1e-2 * torch.randn(p.shape, generator=g)).to(dev))
        sh.outer_light.base.add_((0.5 * torch.randn(sh.outer_light.base.shape, generator=g)).to(dev))
    sh.update_step(2000)
    sh.use_flow_diffuse_copy = sh.use_flow_specular_copy = True
    sh.outer_light.build_mips_direct()
    return cfg, field, variance, mr, verts, make_batch


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--samples", type=int, default=512)
    ap.add_argument("--micro", type=int, default=8192, help="rays / points per micro-batch on a rank")
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--mat-grid", type=int, default=512)
    ap.add_argument("--diffuse", type=int, default=512)
    ap.add_argument("--tris-u", type=int, default=1000)
    ap.add_argument("--tris-v", type=int, default=501)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--check", action="store_true")
    return ap.parse_args(argv)


def main():
    line = run_joint(parse_args(), own_process_group=True)
    if line is not None:
        print(json.dumps(line), flush=True)


def run_joint(args, own_process_group=True):
    """One measurement of the joint step; returns the JSON line (a dict) on rank 0, None elsewhere.  With
    own_process_group=False the caller (bench.py under torchrun) has initialised torch.distributed already."""
    if args.check:
        args.rays, args.samples, args.micro, args.grid, args.mat_grid = 512, 48, 128, 64, 32
        args.tris_u, args.tris_v, args.diffuse = 64, 33, 64
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and own_process_group:
        dist.init_process_group("nccl", device_id=dev)
    from tensoflow_b200 import _lib, synthetic
    from tensoflow_b200.dist import FlatGradBucket, shard_slice
    _lib.load()
    cfg, field, variance, mr, verts, make_batch = build(args, dev)
    params = list(field.parameters()) + [variance] + [p for p in mr.parameters() if p.requires_grad]
    bucket = FlatGradBucket(params)
    R, step = args.rays, 2000
    rays_all = synthetic.make_rays(R, seed=11)
    pts_all, noise_all = make_batch(verts, R, seed=12)
    sl = shard_slice(R, rank, world)

    def run(sl, allreduce):
        """fwd + bwd of the joint step over the rays / points of `sl`, micro-batched; losses normalised by GLOBAL counts."""
        if allreduce:
            bucket.begin_step()        # the flat buffer is the gradient storage: the first micro-batch's kernels scatter into its views,
        else:                          # the following ones accumulate in place
            for p in params:
                p.grad = None
        # pass 0: global sample count for the eikonal mean (uniform sampling: the count is known before the forward)
        n_loc = torch.zeros(1, device=dev)
        chunks = [slice(a, min(a + args.micro, sl.stop)) for a in range(sl.start, sl.stop, args.micro)]
        staged = []
        for c in chunks:
            rays = {k: v[c].to(dev, non_blocking=True) for k, v in rays_all.items()}
            _, _, idx = synthetic.uniform_samples(rays["rays_o"], rays["dirs"], field.aabb, cfg["samples"])
            n_loc += idx.shape[0]
            staged.append(rays)
        if allreduce and world > 1:
            dist.all_reduce(n_loc)
        n_glob = float(n_loc)
        total = torch.zeros((), device=dev)
        for c, rays in zip(chunks, staged):
            col, eik, _ = shape_loss_sums(field, variance, rays, cfg)
            loss = col / R + 0.1 * eik / n_glob
            loss.backward()
            total += loss.detach()
            batch = {k: v[c].to(dev, non_blocking=True) for k, v in pts_all.items()}
            noise = {k: v[c].to(dev, non_blocking=True) for k, v in noise_all.items()}
            loss = material_loss_sums(mr, batch, noise, step, full_losses=not args.check) / R
            loss.backward()
            total += loss.detach()
        if allreduce:
            bucket.allreduce(async_op=True)
            bucket.finish()
            if world > 1:
                dist.all_reduce(total)              # reported loss = the full batch's
        return total

    if args.check:
        run(sl, True)
        got = [p.grad.clone() if p.grad is not None else None for p in params]
        run(slice(0, R), False)                       # the single-GPU full batch, on every rank
        worst, name = 0.0, ""
        names = [n for n, _ in field.named_parameters()] + ["variance"] + [n for n, p in mr.named_parameters() if p.requires_grad]
        for n, g, p in zip(names, got, params):
            if p.grad is None or g is None:
                continue
            ref = p.grad
            if float(ref.abs().max()) < 1e-12:
                continue
            e = float((g - ref).abs().max() / ref.abs().max())
            if e > worst:
                worst, name = e, n
        t = torch.tensor([worst], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"check": "sharded + allreduced gradients vs single-GPU full batch", "n_gpus": world, "rays": R,
                              "max_rel_err": float(t), "worst_param": name, "n_params": len(params), "tol": 1e-3,
                              "ok": bool(float(t) < 1e-3)}), flush=True)
        if world > 1 and own_process_group:
            dist.destroy_process_group()
        sys.exit(0 if float(t) < 1e-3 else 1)

    for _ in range(args.warmup):
        run(sl, True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        loss = run(sl, True)
    b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([a.elapsed_time(b)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t) / args.steps
    line = None
    if rank == 0:
        line = ({
            "metric": "joint shape+material train rays/sec (fwd+bwd)", "value": R / (ms / 1e3), "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "scaling": "strong", "higher_is_better": True,
            "config": {"workload": f"joint step: {R} rays x {cfg['samples']} samples (VM field {cfg['G']}^3, C={cfg['C']}, H={cfg['H']}) + "
                                   f"{R} surface points x ({args.diffuse}+64+32) directions vs {2 * args.tris_u * (args.tris_v - 1)} triangles, "
                                   f"{R // world} rays+points per GPU in micro-batches of {args.micro}",
                       "parallelism": f"dp{world}: ray / point slices, one flat fp32 allreduce of {bucket.numel * 4 / 1e6:.0f} MB"},
            "gpu_launches_per_step": (_lib.launch_count() - l0) / args.steps, "loss": float(loss)})
    if world > 1 and own_process_group:
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
