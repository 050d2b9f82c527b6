"""Known-answer checks of the occupancy-marcher restatement (oracle/torch_oracle_occ.py) that the GPU parity test uses."""
import torch

from oracle import torch_oracle_occ as OO


def test_full_grid_axis_ray():
    o = torch.tensor([[0.0, 0.0, -2.0], [0.3, 0.2, -3.0], [5.0, 5.0, 5.0]])
    d = torch.tensor([[0.0, 0.0, 1.0], [0.0, 0.0, 1.0], [0.0, 0.0, 1.0]])
    full = torch.ones(4, 4, 4, dtype=torch.bool)
    ri, t0, t1 = OO.occ_march(o, d, torch.full((3,), 0.1), 10.0, 0.25, [-1, -1, -1, 1, 1, 1], [4, 4, 4], full)
    assert ri.tolist() == [0] * 8 + [1] * 8                      # 2 / 0.25 lattice samples per hit ray, none for the miss
    assert abs(float(t0[0]) - 1.1) < 1e-6 and abs(float(t1[7]) - 3.1) < 1e-6
    assert torch.allclose(t1 - t0, torch.full_like(t0, 0.25), atol=1e-6)


def test_empty_cells_are_skipped_and_far_clips():
    o = torch.tensor([[0.0, 0.0, -2.0]])
    d = torch.tensor([[0.0, 0.0, 1.0]])
    b = torch.zeros(4, 4, 4, dtype=torch.bool)
    b[2, 2, 3] = True                                            # only the last cell along +z on this ray (x = y = 0 -> cell 2)
    ri, t0, t1 = OO.occ_march(o, d, torch.zeros(1), 10.0, 0.25, [-1, -1, -1, 1, 1, 1], [4, 4, 4], b)
    assert ri.numel() == 2 and abs(float(t0[0]) - 2.5) < 1e-6    # z in [0.5, 1): t in [2.5, 3)
    ri, _, _ = OO.occ_march(o, d, torch.zeros(1), 2.7, 0.25, [-1, -1, -1, 1, 1, 1], [4, 4, 4], b)
    assert ri.numel() == 1                                       # the second mid-point (2.875) lies beyond far
