"""Inference side of the path (BASELINE config 5): surface points = BVH hit refined on the frozen SDF
(reference materialRenderer.py:265-343) and MaterialRenderer.nvs (:641-752), plus the shape-stage checkpoint hand-over
(shapeRenderer.py:326-363 -> materialRenderer.py:148-179)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from oracle import torch_oracle as O
from oracle import torch_oracle_renderer as RR

SHAPE_CFG = dict(gridSize=[32, 32, 32], sdf_n_comp=8, sdf_dim=32, app_dim=128, max_levels=1, has_radiance_field=False,
                 n_samples=16, n_importance=16, sdf_multires=0)
MAT_SHADER = dict(gridSize=[16, 16, 16], light_reso=16, mat_grid=24)


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _renderers(dev, **mat_cfg):
    from tensoflow_b200.shape_renderer import ShapeRenderer
    from tensoflow_b200.material import MaterialRenderer
    from tensoflow_b200.synthetic import bumpy_sphere, perturb_field
    torch.manual_seed(0)
    shape = ShapeRenderer(dict(device=dev, shader_config=dict(env_res=16, env_min_res=4), **SHAPE_CFG))
    perturb_field(shape.sdf_network, seed=2, noise=5e-3)
    verts, tris = bumpy_sphere(48, 24, r=0.22, bump=0.05)          # close to the zero set of the initial sphere SDF
    mat = MaterialRenderer(dict(device=dev, shader_cfg=dict(MAT_SHADER), gridSize=SHAPE_CFG['gridSize'], **mat_cfg), verts, tris)
    mat.init_sdf(shape.ckpt_to_save())
    return shape, mat


def test_ckpt_layout_cpu():
    """reference checkpoint dictionary layout; load at another grid resolution; bit-packed alpha mask round trip."""
    from tensoflow_b200.shape_renderer import ShapeRenderer, AlphaGridMask
    a = ShapeRenderer(dict(device='cpu', shader_config=dict(env_res=16, env_min_res=4), **SHAPE_CFG))
    a.upsample_sdf_grid([64, 64, 64])
    vol = (torch.rand(8, 8, 8) > 0.5).float()
    a.alphaMask = AlphaGridMask('cpu', a.aabb, vol)
    ck = a.ckpt_to_save()
    assert set(ck) == {'kwargs', 'network_state_dict', 'alphaMask.shape', 'alphaMask.mask', 'alphaMask.aabb'}
    for k in ('aabb', 'gridSize', 'sdf_n_comp', 'sdf_dim', 'app_dim', 'sdf_multires', 'max_levels'):      # materialRenderer.py:151-160
        assert k in ck['kwargs']
    b = ShapeRenderer(dict(device='cpu', shader_config=dict(env_res=16, env_min_res=4), **SHAPE_CFG))
    b.load_ckpt(ck)
    assert b.gridSize.tolist() == [64, 64, 64] and b.max_levels == 2 and b.sdf_network.n_levels == 2
    assert float(b.stepSize) == float(a.stepSize)
    assert torch.equal(b.alphaMask.alpha_volume.reshape(-1), vol.reshape(-1))
    for (n, p), (_, q) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(p, q), n


@pytest.mark.gpu
def test_surface_points_match_oracle():
    """trace_sdf_with_mesh: refined depth and FD normal against the fp64 oracle fed with the same mesh depths."""
    dev = _cuda()
    shape, mat = _renderers(dev)
    from tensoflow_b200.synthetic import make_rays
    rays = make_rays(600, seed=3, device=dev, radii_jitter=False)
    o, d = rays['rays_o'], rays['dirs']
    _, _, depth0, hit0 = mat.trace(o, d)
    inters, normals, depth, hit = mat.trace_sdf_with_mesh(o, d, 32, 9)
    hit = hit.squeeze(-1)
    assert torch.equal(hit, hit0.squeeze(-1)) and 50 < int(hit.sum()) < 600
    assert torch.equal(depth[~hit], depth0[~hit])                       # misses keep the tracer's output

    f64 = O.TensoSDF([32] * 3, torch.tensor([[-1., -1, -1], [1, 1, 1]]), sdf_n_comp=8, sdf_dim=32, app_dim=128, init_n_levels=1)
    f64.load_state_dict({k: v.detach().cpu() for k, v in shape.sdf_network.state_dict().items()}, strict=False)
    f64 = f64.double()
    oh, dh, m_depth = o[hit].cpu().double(), d[hit].cpu().double(), depth0[hit].cpu().double()
    inv_s = float(mat.deviation_net(torch.zeros(1, 3, device=dev))[0, 0])
    unit = float(mat.unit_size)
    # oracle/torch_oracle_renderer.surface_refine is pinned to the reference's own trace_sdf_with_mesh (tests/test_oracle_cpu.py)
    dep, pts, n = RR.surface_refine(f64, inv_s, oh, dh, m_depth, unit, float(mat.radius), 32, 9)
    assert rel_err(depth[hit], dep) < 1e-4
    assert rel_err(inters[hit], pts) < 1e-4
    # FD normals divide by the voxel size: same bar as the stencil tests (a few 1e-4 absolute on unit vectors)
    assert float((normals[hit].cpu().double() - n).abs().max()) < 2e-3
    assert float(((normals[hit] * d[hit]).sum(-1)).max()) <= 0


@pytest.mark.gpu
def test_material_nvs_image():
    """MaterialRenderer.nvs: chunking does not change the image; misses are white with normal (0,0,1); a 2-way pixel split
    (what two ranks would render) concatenates to the full image."""
    dev = _cuda()
    h, w = 12, 14
    K = np.array([[18.0, 0, w / 2], [0, 18.0, h / 2], [0, 0, 1]], np.float32)
    pose = np.array([[1, 0, 0, 0.02], [0, 1, 0, -0.03], [0, 0, 1, 1.5]], np.float32)
    imgs = []
    for trn in (512, 53):
        _, mat = _renderers(dev, nvs_ray_num=trn)
        mat.shader_network.outer_light.build_mips_direct()
        imgs.append(mat.nvs(pose, K, h, w))
    a, b = imgs
    assert set(a) == set(mat.NVS_KEYS)
    for k, dch in mat.NVS_KEYS.items():
        assert a[k].shape == (h, w, dch) and np.isfinite(a[k]).all()
        assert np.abs(a[k] - b[k]).max() < 1e-4, k
    hit = np.abs(a['normal'] - np.array([0, 0, 1.0])).max(-1) > 1e-6
    assert 10 < hit.sum() < h * w
    assert np.allclose(a['color'][~hit], 1.0) and np.allclose(a['albedo'][~hit], 0.0)
    assert np.allclose(np.linalg.norm(a['normal'], axis=-1), 1.0, atol=1e-5)
    # rank split: surface points of the strided pixel sets of two "ranks", put back in pixel order, equal the full image's
    from tensoflow_b200.dist import interleaved_ids
    rays = mat.image_rays(pose, K, h, w, dev)
    full = mat._get_trace_ray_batch_info(rays, is_train=False)['inters']
    stitched = torch.empty_like(full)
    for r in range(2):
        ids = interleaved_ids(h * w, r, 2, dev)
        stitched[ids] = mat._get_trace_ray_batch_info({k: v[ids] for k, v in rays.items()}, is_train=False)['inters']
    assert torch.equal(stitched, full)
