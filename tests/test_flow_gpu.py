"""GPU parity tests: fused linear layers, piecewise-quadratic spline kernels and the
TensoFlow sampler (through the C ABI) against the oracle (oracle/torch_oracle_mat.py, which
tests/test_oracle_cpu.py pins to the reference's own network/flow.py)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

from oracle import torch_oracle_mat as OM  # noqa: E402


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def close_as_fp32(got, o64, o32, tol, what, slack=4.0):
    e_got, e_ref = rel_err(got, o64), rel_err(o32, o64)
    assert e_got <= max(tol, slack * e_ref), f"{what}: rel err {e_got:.3e} (fp32 oracle {e_ref:.3e}, tol {tol:.1e})"


def _act_ref(x, act, p):
    return {"none": lambda v: v, "relu": F.relu, "leaky": lambda v: F.leaky_relu(v, 0.01),
            "softplus100": lambda v: F.softplus(v, beta=100), "sigmoid": torch.sigmoid,
            "exp": lambda v: torch.exp(torch.clamp(v, max=p))}[act](x)


# the last four shapes take the output-head kernels (wide input, N <= 8 outputs, >= 4096 rows: one warp per row)
@pytest.mark.parametrize("M,K,N", [(1000, 44, 64), (777, 123, 256), (513, 64, 21), (300, 57, 64), (129, 256, 3), (1, 8, 5),
                                   (5003, 128, 3), (9001, 256, 5), (4100, 96, 1), (6000, 132, 8)])
@pytest.mark.parametrize("act", ["none", "relu", "leaky", "softplus100", "sigmoid", "exp"])
def test_linear(M, K, N, act):
    from tensoflow_b200 import ops
    dev = _cuda()
    g = torch.Generator().manual_seed(M + K + N)
    X, W, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) * (0.3 if act != "softplus100" else 0.02), torch.randn(N, generator=g) * 0.1
    U = torch.randn(M, N, generator=g)
    p = 1.0

    def ref(dt):
        x, w, bb = (t.detach().clone().to(dt).requires_grad_() for t in (X, W, b))
        y = _act_ref(x @ w.T + bb, act, p)
        (y * U.to(dt)).sum().backward()
        return y, x.grad, w.grad, bb.grad

    r64, r32 = ref(torch.float64), ref(torch.float32)
    x, w, bb = (t.detach().clone().to(dev).requires_grad_() for t in (X, W, b))
    y = ops.linear(x, w, bb, act, p)
    (y * U.to(dev)).sum().backward()
    for got, a, c, nm in zip((y, x.grad, w.grad), r64, r32, ("y", "dX", "dW")):
        close_as_fp32(got, a, c, 1e-5, f"linear[{act}] {nm}")
    # db[n] = sum_m dPre[m][n] is a signed sum that can cancel (one element when N = 1) and is accumulated with atomics in
    # an arbitrary order: bound its error by fp32 summation error on sum |dPre| rather than on |sum dPre| alone
    pre = (X.double() @ W.double().T + b.double()).requires_grad_()
    dpre64, = torch.autograd.grad(_act_ref(pre, act, p), pre, U.double())
    err = float((bb.grad.detach().cpu().double() - r64[3]).abs().max())
    err32 = float((r32[3].double() - r64[3]).abs().max())
    allowed = max(1e-5 * float(r64[3].abs().max()), 4.0 * err32, 3e-7 * float(dpre64.abs().sum(0).max()))
    assert err <= allowed, f"linear[{act}] db: abs err {err:.3e} > {allowed:.3e}"


def _spline_inputs(M, seed):
    g = torch.Generator().manual_seed(seed)
    st = torch.randn(M, 21, generator=g)
    y = torch.rand(M, generator=g).clamp(1e-6, 1 - 1e-6)
    y[:4] = torch.tensor([1e-6, 1 - 1e-6, 0.5, 0.25])
    return y, st


def test_pwquad_forward_and_backward():
    from tensoflow_b200 import ops
    dev = _cuda()
    M = 5000
    y, st = _spline_inputs(M, 3)
    g = torch.Generator().manual_seed(4)
    ux, ul = torch.randn(M, generator=g), torch.randn(M, generator=g)

    def ref(dt):
        yy, ss = y.detach().clone().to(dt).requires_grad_(), st.detach().clone().to(dt).requires_grad_()
        x, lj = OM.pwquad_forward(yy, ss)
        ((x * ux.to(dt)).sum() + (lj * ul.to(dt)).sum()).backward()
        return x, lj, yy.grad, ss.grad

    r64, r32 = ref(torch.float64), ref(torch.float32)
    yy, ss = y.detach().clone().to(dev).requires_grad_(), st.detach().clone().to(dev).requires_grad_()
    x, lj = ops.PwquadFunction.apply(yy, ss, False)
    ((x * ux.to(dev)).sum() + (lj * ul.to(dev)).sum()).backward()
    for got, a, c, nm in zip((x, lj, yy.grad, ss.grad), r64, r32, ("x", "logj", "d y", "d st")):
        close_as_fp32(got, a, c, 1e-5, f"pwquad forward {nm}")


def test_pwquad_inverse_and_round_trip():
    from tensoflow_b200 import ops
    dev = _cuda()
    M = 5000
    y, st = _spline_inputs(M, 5)
    x64, l64 = OM.pwquad_inverse(y.double(), st.double())
    x32, l32 = OM.pwquad_inverse(y, st)
    x, lj = ops.PwquadFunction.apply(y.to(dev), st.to(dev), True)
    close_as_fp32(x, x64, x32, 1e-5, "pwquad inverse x")
    close_as_fp32(lj, l64, l32, 1e-5, "pwquad inverse logj")
    # known-answer identity of the reference: forward(inverse(y)) ~ y, logj's cancel.
    # the forward spline clamps its widths at 1e-6 and the inverse does not, so restrict to benign rows.
    back, lj2 = ops.PwquadFunction.apply(x, st.to(dev), False)
    ok = (st[:, 11:].max(-1).values - st[:, 11:].min(-1).values < 8).to(dev)
    # (fp32 quadratic-formula cancellation makes a few rows ~1e-4; the bulk is at rounding level)
    err = (back - y.to(dev)).abs()[ok]
    assert float(err.max()) < 2e-3 and float(err.median()) < 1e-6
    assert float((lj + lj2).abs()[ok].median()) < 1e-5


def _make_flows(G=32, seed=0):
    from tensoflow_b200.flow import TensoFlow
    dev = _cuda()
    torch.manual_seed(seed)
    aabb = torch.tensor([[-1., -1, -1], [1, 1, 1]])
    o32 = OM.TensoFlow(aabb, gridSize=(G, G, G))
    with torch.no_grad():
        for p in o32.nis_plane:
            p.mul_(2000.0)
        for blk in o32.flows:
            blk.nn[7].weight.mul_(3.0)
    o64 = OM.TensoFlow(aabb, gridSize=(G, G, G), dtype=torch.float64)
    o64.load_state_dict({k: v.double() for k, v in o32.state_dict().items()})
    cu = TensoFlow(2, aabb, device=dev, gridSize=[G, G, G])
    missing = cu.load_state_dict({k: v for k, v in o32.state_dict().items() if k != "aabb"}, strict=False)
    assert not missing.unexpected_keys and not missing.missing_keys, missing
    return o32, o64, cu


def test_tensoflow_sample():
    dev = _cuda()
    o32, o64, cu = _make_flows()
    pn, sn = 200, 64
    g = torch.Generator().manual_seed(7)
    pts, va, rough = torch.rand(pn, 3, generator=g) * 1.9 - 0.95, torch.rand(pn, 2, generator=g), torch.rand(pn, 1, generator=g)
    shift = torch.rand(pn, sn, 1, generator=g)
    with torch.no_grad():
        a64, l64 = o64.sample(pts.double(), va.double(), rough.double(), sn, shift.double())
        a32, l32 = o32.sample(pts, va, rough, sn, shift)
        a, l = cu.sample(pts.to(dev), va.to(dev), rough.to(dev), sn, return_jacobian=True, phi_shift=shift.to(dev))
        close_as_fp32(a, a64, a32, 1e-4, "sampled angles")
        close_as_fp32(l, l64, l32, 1e-4, "sample logj")
        for sn2 in (80, 8):               # 200 x 80: points straddle the 128-pair tiles of the fused kernel; 8 < 16: per-layer path
            sh2 = torch.rand(pn, sn2, 1, generator=g)
            b64, m64 = o64.sample(pts.double(), va.double(), rough.double(), sn2, sh2.double())
            b32, m32 = o32.sample(pts, va, rough, sn2, sh2)
            b, m = cu.sample(pts.to(dev), va.to(dev), rough.to(dev), sn2, return_jacobian=True, phi_shift=sh2.to(dev))
            close_as_fp32(b, b64, b32, 1e-4, f"sampled angles ({sn2})")
            close_as_fp32(m, m64, m32, 1e-4, f"sample logj ({sn2})")
        cu.eval()
        a_eval = cu.sample(pts.to(dev), va.to(dev), rough.to(dev), 32)
        e64, _ = o64.sample(pts.double(), va.double(), rough.double(), 32, None)
        e32, _ = o32.sample(pts, va, rough, 32, None)
        close_as_fp32(a_eval, e64, e32, 1e-4, "sampled angles (eval, 32)")


@pytest.mark.parametrize("ragged,pn,sn", [(False, 150, 64), (True, 150, 64), (False, 37, 48), (False, 5, 16), (False, 9, 8)])
def test_tensoflow_logq_forward_backward(ragged, pn, sn):
    """dense [pn, sn] direction sets run on the fused coupling-block kernels (tf_flow_block_*: sn >= 16; 37 x 48 = 13.9 tiles of 128
    pairs with points straddling tile boundaries, 5 x 16 = less than one tile), ragged sets (rays_id) and sn < 16 on the
    per-layer kernels"""
    dev = _cuda()
    o32, o64, cu = _make_flows(seed=1)
    g = torch.Generator().manual_seed(9)
    pts, va, rough = torch.rand(pn, 3, generator=g) * 1.9 - 0.95, torch.rand(pn, 2, generator=g), torch.rand(pn, 1, generator=g)
    if ragged:
        rid = torch.sort(torch.randint(0, pn, (3000,), generator=g))[0]
        x = torch.rand(3000, 2, generator=g)
    else:
        rid = None
        x = torch.rand(pn, sn, 2, generator=g)
    u = torch.randn(*x.shape[:-1], 1, generator=g)

    def ref(f, dt):
        z, lq = f(pts.to(dt), va.to(dt), rough.to(dt), x.to(dt), rays_id=rid)
        (lq * u.to(dt)).sum().backward()
        return z, lq

    z64, q64 = ref(o64, torch.float64)
    z32, q32 = ref(o32, torch.float32)
    z, lq = cu(pts.to(dev), va.to(dev), rough.to(dev), x.to(dev), return_jacobian=True, rays_id=None if rid is None else rid.to(dev))
    (lq * u.to(dev)).sum().backward()
    close_as_fp32(z, z64, z32, 1e-4, "z")
    close_as_fp32(lq, q64, q32, 1e-4, "log q")
    p64, p32, pc = dict(o64.named_parameters()), dict(o32.named_parameters()), dict(cu.named_parameters())
    for name, p in pc.items():
        if not p.requires_grad:
            continue
        assert p.grad is not None, name
        close_as_fp32(p.grad, p64[name].grad, p32[name].grad, 1e-3, f"d {name}")
