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
