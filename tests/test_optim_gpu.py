"""FusedAdam (tf_adam_step) against torch.optim.Adam with the reference trainer's settings (train/trainer_inv.py:112:
betas=(0.9, 0.99), per-group learning rates, lr schedule by multiplying param_group['lr'], :247-248)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def _params(dev, seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1, 36, 64, 64), (1, 36, 64, 1), (259, 111), (256,), (1,), (4097,), (3, 5, 7), (8192 + 3,)] + [(17, 3)] * 40
    ps = []
    for i, s in enumerate(shapes):
        t = torch.randn(s, generator=g).to(dev)
        if len(s) == 4:                                   # VM factors are stored channels-last
            t = t.contiguous(memory_format=torch.channels_last)
        ps.append(torch.nn.Parameter(t))
    return ps


def test_fused_adam_matches_torch_adam():
    dev = _cuda()
    from tensoflow_b200.optim import FusedAdam
    a, b = _params(dev, 0), _params(dev, 0)
    groups = lambda ps: [{'params': ps[:2], 'lr': 1e-2}, {'params': ps[2:], 'lr': 1e-3}]
    ref = torch.optim.Adam(groups(a), betas=(0.9, 0.99))
    ours = FusedAdam(groups(b), betas=(0.9, 0.99))
    g = torch.Generator().manual_seed(5)
    for it in range(7):
        for i, (p, q) in enumerate(zip(a, b)):
            if it == 3 and i == 4:
                p.grad = q.grad = None                     # a parameter without a gradient this step is skipped by both
                continue
            gr = (torch.randn(p.shape, generator=g) * (10.0 ** ((i % 5) - 3))).to(dev)
            p.grad = gr.clone()
            # the autograd functions hand back channels-last OR contiguous gradients: both must work
            q.grad = gr.clone().contiguous(memory_format=torch.channels_last) if (p.dim() == 4 and it % 2 == 0) else gr.clone()
        ref.step()
        ours.step()
        for grp_r, grp_o in zip(ref.param_groups, ours.param_groups):      # the trainer's cosine schedule
            grp_r['lr'] *= 0.97
            grp_o['lr'] *= 0.97
    for i, (p, q) in enumerate(zip(a, b)):
        assert q.stride() == p.stride()
        assert rel_err(q, p) < 2e-6, i
        assert rel_err(ours.state[q]['exp_avg'], ref.state[p]['exp_avg']) < 2e-6, i
        assert rel_err(ours.state[q]['exp_avg_sq'], ref.state[p]['exp_avg_sq']) < 2e-6, i


def test_fused_adam_refuses_cpu():
    _cuda()
    from tensoflow_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdam([p]).step()


def test_fused_adam_with_multilevel_field_rebuilds_mips():
    """FusedAdam writes parameters through raw pointers (no autograd version bump).  The VM descriptor cache of
    tensoflow_b200.ops holds the box mip chain of the factors, so a stale entry would make levels >= 1 lag one step
    behind level 0.  Three optimizer steps on a 3-level TensoSDF sampled at mip levels > 0 (reference trainer:
    train/trainer_inv.py:112,212; level-dependent lookups network/fields.py:262-299): (a) the FusedAdam run equals the
    torch.optim.Adam run of the same kernels, (b) after the steps the field equals the oracle evaluated on the SAME
    parameters -- which fails if any cached level survived an update."""
    dev = _cuda()
    from tensoflow_b200.fields import TensoSDF
    from tensoflow_b200.optim import FusedAdam
    from tensoflow_b200 import synthetic
    from oracle import torch_oracle as O

    def make():
        torch.manual_seed(3)
        f = TensoSDF(torch.tensor([16] * 3), torch.tensor([[-1.0] * 3, [1.0] * 3]), device=dev, sdf_n_comp=8, sdf_dim=64, app_dim=16,
                     init_n_levels=1, sdf_multires=0)
        for r in (32, 64):
            f.upsample_volume_grid(torch.tensor([r] * 3))
        synthetic.perturb_field(f, seed=4, noise=0.05)
        return f

    fa, fb = make(), make()
    assert fa.n_levels == 3
    groups = lambda f: [{'params': list(f.sdf_plane) + list(f.sdf_line), 'lr': 2e-2}, {'params': f.sdf_mat.parameters(), 'lr': 1e-3}]
    opt_a = torch.optim.Adam(groups(fa), betas=(0.9, 0.99))
    opt_b = FusedAdam(groups(fb), betas=(0.9, 0.99))
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(3000, 3, generator=g) * 1.9 - 0.95).to(dev)
    lv = (torch.rand(3000, 1, generator=g) * 1.8 + 0.1).to(dev)          # every sample blends mip levels (0,1) or (1,2)
    u = torch.randn(3000, generator=g).to(dev)
    losses = []
    for step in range(3):
        row = []
        for f, opt in ((fa, opt_a), (fb, opt_b)):
            opt.zero_grad(set_to_none=True)
            sdf, feat, grad, hess = f.stencil(x, lv)
            loss = (sdf * u).mean() + 0.1 * ((grad.norm(dim=-1) - 1) ** 2).mean() + 0.01 * feat.square().mean()
            loss.backward()
            opt.step()
            row.append(float(loss))
        losses.append(row)
    for step, (la, lb) in enumerate(losses):
        assert abs(la - lb) <= 2e-5 * max(abs(la), 1e-3), (step, la, lb)
    for (n, pa), (_, pb) in zip(fa.named_parameters(), fb.named_parameters()):
        assert rel_err(pb, pa) < 1e-4, n
    # (b) the field after the last FusedAdam step against the oracle on the same parameters, at levels > 0
    o = O.TensoSDF([16] * 3, [[-1.0] * 3, [1.0] * 3], sdf_n_comp=8, sdf_dim=64, app_dim=16, init_n_levels=1, dtype=torch.float64)
    for r in (32, 64):
        o.upsample_volume_grid(torch.tensor([r] * 3))
    synthetic.copy_field_params(fb, o)
    with torch.no_grad():
        want = o(x.cpu().double(), lv.cpu().double())
        got = fb(x, lv)
    assert rel_err(got, want) < 1e-4
