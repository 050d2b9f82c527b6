"""CPU checks of the C-ABI boundary: the library builds, loads without a GPU, and exports
every symbol include/tensoflow_b200.h declares.  No compute calls."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from tensoflow_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from tensoflow_b200 import _lib
    header = (ROOT / "include" / "tensoflow_b200.h").read_text()
    declared = set(re.findall(r"TF_API\s+[\w\s\*]+?\b(tf_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(str(_lib.lib_path()))
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())


def test_abi_version_and_error_string(lib):
    assert lib.tf_abi_version() == 1
    assert isinstance(lib.tf_last_error(), bytes)


def test_no_cpu_fallback():
    """Product ops refuse CPU tensors instead of silently computing elsewhere."""
    import torch
    from tensoflow_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(4))


def test_product_does_not_import_oracle():
    for p in (ROOT / "tensoflow_b200").glob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p


def test_new_entry_points_refuse_cpu_tensors():
    """FusedAdam, the occupancy marcher and the cube lookup have no CPU path either: CPU tensors raise before any launch."""
    import torch
    from tensoflow_b200.optim import FusedAdam
    from tensoflow_b200.occ_grid import OccGridEstimator
    from tensoflow_b200.shape_shader import cube_lookup
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        FusedAdam([p]).step()
    est = OccGridEstimator([-1, -1, -1, 1, 1, 1], resolution=4)
    est.mark_all_occupied()
    with pytest.raises(RuntimeError):
        est.sampling(torch.zeros(2, 3), torch.ones(2, 3), near_plane=0.1, far_plane=2.0, render_step_size=0.1)
    with pytest.raises(RuntimeError):
        cube_lookup([torch.zeros(6, 4, 4, 3)], torch.ones(5, 3))


def test_occ_grid_refresh_host_logic():
    """OccGridEstimator._update / update_every_n_steps / state_dict are tensor logic: check them on CPU (nerfacc semantics:
    occs = max(occs * decay, occ), binaries = occs > min(mean, thre); refresh only when step % n == 0 in training mode)."""
    import torch
    from tensoflow_b200.occ_grid import OccGridEstimator
    est = OccGridEstimator([-1, -1, -1, 1, 1, 1], resolution=8)
    ball = lambda x: (x.norm(dim=-1) < 0.6).float()
    est._update(0, ball, warmup_steps=10, jitter=torch.full((512, 3), 0.5))
    centres = (est.grid_coords.float() + 0.5) / 8 * 2 - 1
    assert torch.equal(est.binaries.reshape(-1), centres.norm(dim=-1) < 0.6)
    before = est.occs.clone()
    est._update(0, lambda x: torch.zeros(x.shape[0]), ema_decay=0.5, warmup_steps=10)
    assert torch.allclose(est.occs, before * 0.5)                    # decay only
    est.eval()
    est.update_every_n_steps(100, ball, n=100, warmup_steps=0)        # not training: untouched
    assert torch.allclose(est.occs, before * 0.5)
    est.train()
    est.update_every_n_steps(101, ball, n=100, warmup_steps=0)        # not a multiple of n: untouched
    assert torch.allclose(est.occs, before * 0.5)
    est.update_every_n_steps(200, ball, n=100, warmup_steps=0)        # partial refresh (a quarter + the occupied cells)
    assert float(est.occs.max()) == 1.0
    sd = est.state_dict()
    assert set(sd) == {"resolution", "aabbs", "occs", "binaries"}
    est2 = OccGridEstimator([-1, -1, -1, 1, 1, 1], resolution=8)
    est2.load_state_dict(sd)
    assert torch.equal(est2.binaries, est.binaries) and torch.equal(est2.occs, est.occs)
