"""Occupancy-grid marcher (tf_occ_march_*; nerfacc.OccGridEstimator of the reference's `*_occ` configs,
shapeRenderer.py:211-215, 950-959, 1285-1290) against the CPU restatement, and the ShapeRenderer wiring."""
import pytest
import torch
import torch.nn.functional as F

from oracle import torch_oracle_occ as OO

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _rays(n, seed):
    g = torch.Generator().manual_seed(seed)
    o = F.normalize(torch.randn(n, 3, generator=g), dim=-1) * (1.2 + 1.5 * torch.rand(n, 1, generator=g))
    o[: n // 8] *= 0.2                                           # some origins inside the box
    tgt = (torch.rand(n, 3, generator=g) - 0.5) * 2.4             # some rays miss the box
    d = F.normalize(tgt - o, dim=-1)
    d[n // 2, 0] = 0.0                                            # an axis-parallel component
    d[n // 2] = F.normalize(d[n // 2], dim=-1)
    return o, d


@pytest.mark.parametrize("res,step,stratified,fill", [((16, 16, 16), 0.031, False, 0.3), ((32, 24, 40), 0.0123, True, 0.1),
                                                      ((8, 8, 8), 0.05, True, 1.0), ((16, 16, 16), 0.02, False, 0.0)])
def test_march_matches_oracle(res, step, stratified, fill):
    from tensoflow_b200.occ_grid import OccGridEstimator
    dev = _cuda()
    aabb = [-1.0, -0.9, -1.1, 1.0, 0.8, 1.05]
    est = OccGridEstimator(aabb, resolution=list(res)).to(dev)
    g = torch.Generator().manual_seed(1)
    binaries = torch.rand(*res, generator=g) < fill
    est.binaries.copy_(binaries[None].to(dev))
    o, d = _rays(301, 2)
    noise = torch.rand(301, generator=g)
    near_plane, far_plane = 0.05, 3.3
    ri, t0, t1 = est.sampling(o.to(dev), d.to(dev), near_plane=near_plane, far_plane=far_plane, render_step_size=step,
                              stratified=stratified, noise=noise)
    near = torch.full((301,), near_plane) + (noise * step if stratified else 0.0)
    ri_o, t0_o, t1_o = OO.occ_march(o, d, near, far_plane, step, aabb, list(res), binaries)
    assert ri.dtype == torch.int64
    assert torch.equal(ri.cpu(), ri_o)                           # same samples kept, same order
    assert torch.equal(t0.cpu(), t0_o) and torch.equal(t1.cpu(), t1_o)
    offs = est.last_ray_offsets.cpu()
    assert int(offs[-1]) == ri_o.shape[0] and torch.equal(offs[1:] - offs[:-1], torch.bincount(ri_o, minlength=301).int())
    if fill == 0.0:
        assert ri.numel() == 0
    if fill == 1.0:                                              # a full grid: contiguous runs, every mid-point inside the box
        mid = (t0 + t1) * 0.5
        p = o.to(dev)[ri] + d.to(dev)[ri] * mid[:, None]
        lo, hi = torch.tensor(aabb[:3], device=dev), torch.tensor(aabb[3:], device=dev)
        assert bool(((p >= lo - 1e-5) & (p <= hi + 1e-5)).all()) and ri.numel() > 1000


def test_update_and_state_dict():
    from tensoflow_b200.occ_grid import OccGridEstimator
    dev = _cuda()
    est = OccGridEstimator([-1, -1, -1, 1, 1, 1], resolution=16).to(dev)
    occ_fn = lambda x: (x.norm(dim=-1) < 0.5).float()            # a ball of radius 0.5
    jitter = torch.full((16 ** 3, 3), 0.5)
    est._update(0, occ_fn, warmup_steps=10, jitter=jitter)
    centres = (est.grid_coords.float() + 0.5) / 16 * 2 - 1
    assert torch.equal(est.binaries.reshape(-1), centres.norm(dim=-1) < 0.5)
    n_occ = int(est.binaries.sum())
    est.train()
    est.update_every_n_steps(100, occ_fn, n=100, warmup_steps=10)     # partial refresh: EMA keeps cells occupied
    assert int(est.binaries.sum()) >= n_occ * 0.9
    est.update_every_n_steps(101, occ_fn, n=100, warmup_steps=10)     # not a multiple of n: untouched
    sd = est.state_dict()
    assert set(sd) == {"resolution", "aabbs", "occs", "binaries"}
    est2 = OccGridEstimator([-1, -1, -1, 1, 1, 1], resolution=16).to(dev)
    est2.load_state_dict(sd)
    assert torch.equal(est2.binaries, est.binaries)


def test_shape_renderer_with_occ_grid():
    """`use_occ_grid: true` end to end: grid refresh from compute_alpha, marched samples through the fused field / compositor,
    occlusion loss through the grid, gradients reach the factors; with a full grid the image equals a dense fixed-step march."""
    from tensoflow_b200.shape_renderer import ShapeRenderer, near_far_from_sphere
    from tensoflow_b200 import synthetic
    dev = _cuda()
    torch.manual_seed(0)
    cfg = dict(gridSize=[32, 32, 32], sdf_n_comp=8, sdf_dim=32, app_dim=128, max_levels=1, has_radiance_field=True, radiance_field_step=100,
               occ_loss_step=0, occ_loss_max_pn=100000, use_occ_grid=True, occ_grid_reso=16, device=dev,
               shader_config=dict(env_res=16, env_min_res=4))
    m = ShapeRenderer(cfg)
    rays = synthetic.make_rays(96, seed=4, device=dev, radii_jitter=False)
    rays['rays_d'] = rays['dirs']
    m.set_train_batch({**{k: v.cpu() for k, v in rays.items()}, 'human_poses': torch.zeros(96, 3, 4)})
    m.cfg['train_ray_num'] = 96
    out = m({'step': 30000})                                     # warm-up refresh (step < 10000 is all cells; here partial) + step
    assert out['sample_num'] >= 0 and torch.isfinite(out['ray_rgb']).all()
    m.occ_grid._update(0, m.compute_alpha, warmup_steps=10)      # full refresh
    frac = float(m.occ_grid.binaries.float().mean())
    assert 0.0 < frac < 0.9                                      # the initial sphere occupies part of the grid
    m.set_train_batch({**{k: v.cpu() for k, v in rays.items()}, 'human_poses': torch.zeros(96, 3, 4)})
    out = m({'step': 30001})
    loss = out['loss_rgb'].mean() + out['gradient_error'].mean() * 0.1 + out['loss_occ'].sum()
    loss.backward()
    assert m.sdf_network.sdf_plane[0].grad is not None and float(m.sdf_network.sdf_plane[0].grad.abs().sum()) > 0
    assert float(out['acc'].max()) > 0.5
    # skipping empty cells must not change the picture much: compare with the full grid.  With the initial inv_s = 20 the NeuS
    # opacity has long tails (alpha ~ 0.01 a fifth of the box away from the surface), so sharpen the surface first and refresh
    # the grid a few times (one random probe per cell and refresh; the EMA keeps the maximum)
    near, far = near_far_from_sphere(rays['rays_o'], rays['dirs'], float(m.radius))
    with torch.no_grad():
        m.deviation_network.variance.fill_(0.5)
        for _ in range(6):
            m.occ_grid._update(0, m.compute_alpha, warmup_steps=10)
        assert 0.0 < float(m.occ_grid.binaries.float().mean()) < 0.9
        a = m.render(rays, near, far, None, perturb_overwrite=0, is_train=False, step=30001)['ray_rgb']
        m.occ_grid.mark_all_occupied()
        b = m.render(rays, near, far, None, perturb_overwrite=0, is_train=False, step=30001)['ray_rgb']
    assert float((a - b).abs().max()) < 2e-2
    ck = m.ckpt_to_save()
    assert 'occ_grid_state_dict' in ck
