"""ShapeRenderer.render end to end (hierarchical NeuS sampler -> fused field stencil -> shader ->
NeuS alpha / compositing -> occlusion, sparse, hessian, TV losses) against the reference's own
ShapeRenderer.render output (tests/golden/renderer.npz, from oracle/gen_golden.py)."""
import pytest
import torch

from conftest import rel_err
from test_golden import load
from oracle import torch_oracle_renderer as RR

KEYS = ['ray_rgb', 'acc', 'normal', 'radiance', 'roughness_weights', 'gradient_error', 'loss_sparse', 'loss_hessian', 'std',
        'loss_tv_sdf', 'loss_occ']
CFG = dict(gridSize=[32, 32, 32], sdf_n_comp=8, sdf_dim=32, app_dim=128, max_levels=1, predict_BG=False, has_radiance_field=True,
           radiance_field_step=100, occ_loss_step=0, occ_loss_max_pn=100000, n_samples=16, n_importance=16, up_sample_steps=4,
           sdf_multires=0)


def total_loss(o):
    return (o['ray_rgb'].sum() + o['radiance'].sum() * 0.5 + o['gradient_error'].mean() * 0.1 + o['loss_sparse']
            + o['loss_hessian'] * 1e-3 + o['loss_tv_sdf'] + o['loss_occ'].sum())


def _oracle(g, dtype=torch.float32):
    m = RR.ShapeRenderer([32] * 3, sdf_n_comp=8, sdf_dim=32, app_dim=128, max_levels=1, has_radiance_field=True,
                         radiance_field_step=100, n_samples=16, n_importance=16, occ_loss_step=0, occ_loss_max_pn=100000,
                         env_res=16, env_min_res=4, dtype=dtype)
    res = m.load_state_dict({k: v.to(dtype) for k, v in g["state"].items()}, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    return m


def _render_oracle(m, i, dt):
    m.color_network.envlight.build_mips()
    return m.render(i["rays_o"].to(dt), i["dirs"].to(dt), i["radiis"].to(dt), i["rays_cos"].to(dt), i["near"].to(dt), i["far"].to(dt),
                    0.7, 30000, t_rand=i["t_rand"].to(dt))


def test_oracle_renderer_golden():
    g = load("renderer.npz")
    m = _oracle(g)
    out = _render_oracle(m, g["inputs"], torch.float32)
    assert abs(out["sample_num"] - float(g["outputs"]["sample_num"])) < 1e-6
    for k in KEYS:
        assert rel_err(torch.as_tensor(out[k]), g["outputs"][k]) < 1e-5, k
    total_loss(out).backward()
    for n, p in m.named_parameters():
        if n in g["grads"] and p.grad is not None:
            assert rel_err(p.grad, g["grads"][n]) < 1e-4, n


@pytest.mark.gpu
def test_cuda_renderer_golden():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tensoflow_b200.shape_renderer import ShapeRenderer
    dev = torch.device("cuda:0")
    g = load("renderer.npz")
    i = g["inputs"]
    m = ShapeRenderer(dict(device=dev, shader_config=dict(env_res=16, env_min_res=4), **CFG))
    res = m.load_state_dict(g["state"], strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    m64 = _oracle(g, torch.float64)
    o64 = _render_oracle(m64, i, torch.float64)
    batch = {k: i[k].to(dev) for k in ("rays_o", "rays_d", "dirs", "radiis", "rays_cos")}
    m.color_network.envlight.build_mips()
    out = m.render(batch, i["near"].to(dev), i["far"].to(dev), None, -1, 0.7, is_train=True, step=30000, t_rand=i["t_rand"].to(dev))
    assert abs(out["sample_num"] - float(g["outputs"]["sample_num"])) < 0.05   # a sample on the aabb face may flip

    def close(got, ref32, ref64, tol, what):
        e, e_ref = rel_err(got, ref64), rel_err(ref32, ref64)
        assert e <= max(tol, 4 * e_ref), f"{what}: rel err {e:.3e} (reference fp32 vs fp64 oracle {e_ref:.3e})"

    if abs(out["sample_num"] - float(g["outputs"]["sample_num"])) < 1e-6:
        for k in KEYS:
            close(torch.as_tensor(out[k]), g["outputs"][k], torch.as_tensor(o64[k]), 1e-4 if k != "loss_hessian" else 1e-3, k)
        total_loss(out).backward()
        total_loss(o64).backward()
        p64 = dict(m64.named_parameters())
        for n, p in m.named_parameters():
            if n in g["grads"] and p.grad is not None and n in p64:
                close(p.grad, g["grads"][n], p64[n].grad, 1e-3, f"d {n}")
    else:   # per-ray outputs still have to agree
        for k in ('ray_rgb', 'acc', 'radiance'):
            assert rel_err(out[k], g["outputs"][k]) < 1e-3, k


@pytest.mark.gpu
def test_alpha_mask_and_train_step():
    """updateAlphaMask + a train_step through set_train_batch (host-resident rays, H2D per step)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tensoflow_b200.shape_renderer import ShapeRenderer
    from tensoflow_b200 import synthetic
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = ShapeRenderer(dict(device=dev, shader_config=dict(env_res=16, env_min_res=4), train_ray_num=64, **CFG))
    rays = synthetic.make_rays(256, seed=1)
    m.set_train_batch(dict(rays_o=rays["rays_o"], rays_d=rays["dirs"], dirs=rays["dirs"], radiis=rays["radiis"],
                           rays_cos=rays["rays_cos"], rgbs=rays["rgbs"]))
    out = m({'step': 30000})
    loss = out['loss_rgb'].mean() + 0.1 * out['gradient_error'].mean() + out['loss_tv_sdf']
    loss.backward()
    assert torch.isfinite(loss) and m.sdf_network.sdf_plane[0].grad is not None
    n_before = out['sample_num']
    new_aabb = m.updateAlphaMask((32, 32, 32))
    assert new_aabb.shape == (2, 3) and float(m.alphaMask.alpha_volume.mean()) < 1.0
    out2 = m({'step': 30001})
    assert out2['sample_num'] <= n_before + 1e-6       # the mask only removes samples
    assert torch.isfinite(out2['ray_rgb']).all()


@pytest.mark.gpu
def test_nvs_full_image_chunked():
    """ShapeRenderer.nvs (reference shapeRenderer.py:569-668): chunking must not change the image."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import numpy as np
    from tensoflow_b200.shape_renderer import ShapeRenderer
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    h, w = 10, 12
    K = np.array([[20.0, 0, w / 2], [0, 20.0, h / 2], [0, 0, 1]], np.float32)
    pose = np.array([[1, 0, 0, 0.1], [0, 1, 0, -0.05], [0, 0, 1, 2.0]], np.float32)      # camera at z = 2 looking down -z
    imgs = []
    for trn in (4096, 37):
        torch.manual_seed(0)
        m = ShapeRenderer(dict(device=dev, shader_config=dict(env_res=16, env_min_res=4), test_ray_num=trn, **CFG))
        m.color_network.envlight.build_mips()
        imgs.append(m.nvs(pose, K, h, w, perturb_overwrite=0))
    for k in ('color', 'normal', 'acc', 'radiance'):
        assert imgs[0][k].shape[:2] == (h, w)
        assert np.isfinite(imgs[0][k]).all()
        assert np.abs(imgs[0][k] - imgs[1][k]).max() < 1e-4, k
    assert imgs[0]['acc'].max() > 0.5            # the initial sphere is visible
