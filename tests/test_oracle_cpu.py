"""CPU tests of the oracle itself: texture semantics, autograd consistency, and (where the
reference tree is present) agreement with the reference's own classes through the shim."""
import pytest
import torch
import torch.nn.functional as F

from oracle import torch_oracle as O, ref_shim
from conftest import rel_err


def test_texture2d_matches_grid_sample():
    """Level-0 clamp-mode bilinear == grid_sample(align_corners=False, border)."""
    torch.manual_seed(0)
    tex = torch.randn(12, 20, 5, dtype=torch.float64)
    uv = torch.rand(500, 2, dtype=torch.float64) * 1.4 - 0.2
    a = O.texture2d(tex, uv, None, 1)
    b = F.grid_sample(tex.permute(2, 0, 1)[None], (uv * 2 - 1)[None, :, None, :], mode="bilinear",
                      padding_mode="border", align_corners=False)[0, :, :, 0].T
    assert rel_err(a, b) < 1e-12


def test_texture2d_mip_levels():
    torch.manual_seed(1)
    tex = torch.randn(16, 16, 3, dtype=torch.float64)
    uv = torch.rand(200, 2, dtype=torch.float64)
    chain = O.build_mip_chain(tex, 3)
    assert chain[1].shape == (8, 8, 3) and chain[2].shape == (4, 4, 3)
    assert torch.allclose(chain[1][0, 0], tex[:2, :2].mean((0, 1)))
    for lv, (l0, l1, f) in {-1.0: (0, 0, 0.0), 0.25: (0, 1, 0.25), 1.0: (1, 1, 0.0), 1.5: (1, 2, 0.5), 7.0: (2, 2, 0.0)}.items():
        got = O.texture2d(tex, uv, torch.full((200,), lv, dtype=torch.float64), 3)
        want = (1 - f) * O._bilinear_clamp(chain[l0], uv) + f * O._bilinear_clamp(chain[l1], uv)
        assert rel_err(got, want) < 1e-12, lv
    line = torch.randn(16, 1, 3, dtype=torch.float64)
    lc = O.build_mip_chain(line, 3)
    assert lc[2].shape == (4, 1, 3)
    assert torch.allclose(lc[1][0, 0], line[:2, 0].mean(0))


def test_weights_and_accumulate():
    torch.manual_seed(2)
    alpha = torch.rand(10, dtype=torch.float64)
    idx = torch.tensor([0, 0, 0, 2, 2, 3, 3, 3, 3, 3])
    w, T = O.render_weight_from_alpha(alpha, idx, 5)
    assert torch.allclose(T[0], torch.tensor(1.0, dtype=torch.float64)) and torch.allclose(T[3], torch.tensor(1.0, dtype=torch.float64))
    assert torch.allclose(T[2], (1 - alpha[0]) * (1 - alpha[1]))
    acc = O.accumulate_along_rays(w, None, idx, 5)
    assert acc.shape == (5, 1) and float(acc[1]) == 0.0 and float(acc[4]) == 0.0


def test_sphere_init_is_a_sphere():
    """Known-answer from the reference's deterministic initialisers: eikonal ~ 1."""
    torch.manual_seed(0)
    f = O.TensoSDF([128] * 3, [[-1.0] * 3, [1.0] * 3], sdf_n_comp=16, sdf_dim=128, app_dim=128, init_n_levels=1)
    x = torch.rand(4096, 3) * 1.6 - 0.8
    out = f(x, None)
    g, _ = f.gradient(x, None)
    assert abs(float(out[:, 0].mean()) - 0.587) < 0.05
    assert abs(float(g.norm(dim=-1).mean()) - 0.99) < 0.05


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")
def test_oracle_matches_reference_tensosdf():
    ref_shim.install()
    import network.fields as RF
    torch.manual_seed(0)
    G = torch.tensor([16, 16, 16]); aabb = torch.tensor([[-1., -1, -1], [1, 1, 1]])
    ref = RF.TensoSDF(G, aabb, device='cpu', sdf_n_comp=8, sdf_dim=32, app_dim=16, init_n_levels=1, sdf_multires=0)
    ref.upsample_volume_grid(torch.tensor([33, 33, 33]))
    ref.upsample_volume_grid(torch.tensor([66, 66, 66]))
    mine = O.TensoSDF(G, aabb, sdf_n_comp=8, sdf_dim=32, app_dim=16, init_n_levels=1)
    mine.upsample_volume_grid(torch.tensor([33, 33, 33])); mine.upsample_volume_grid(torch.tensor([66, 66, 66]))
    assert list(mine.gridSize) == list(ref.gridSize) and mine.n_levels == ref.n_levels == 3
    with torch.no_grad():
        for p in list(ref.sdf_plane) + list(ref.sdf_line):
            p.add_(0.05 * torch.randn_like(p))
    mine.load_state_dict({k: v for k, v in ref.state_dict().items() if 'gaussian' not in k}, strict=False)
    x = torch.rand(777, 3) * 2.2 - 1.1
    lv = torch.rand(777, 1) * 4 - 1
    a, b = ref(x, lv), mine(x, lv)
    assert rel_err(b, a) < 1e-6
    ga, ha = ref.gradient(x, lv, training=True, sdf=a[:, :1])
    gb, hb = mine.gradient(x, lv, training=True, sdf=b[:, :1])
    assert rel_err(gb, ga) < 1e-5 and rel_err(hb, ha) < 1e-4
    (a.sum() + ga.sum()).backward(); (b.sum() + gb.sum()).backward()
    for (n, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
        assert rel_err(q.grad, p.grad) < 1e-5, n
