"""Occupancy-grid marcher of the `*_occ` shape configs (reference network/shapeRenderer.py:211-215, 950-959, 1285-1290:
`nerfacc.OccGridEstimator(roi_aabb, resolution).sampling(...)` / `.update_every_n_steps(...)`), SURVEY.md 8f-2.

nerfacc is not vendored by the reference and no reference test pins its arithmetic: this is a restatement of its documented
behaviour ("parity unpinned", like every nerfacc boundary of the path):
  * state: `occs [res^3]` fp32 (EMA of the evaluated occupancy), `binaries [1,res,res,res]` bool, `aabbs [1,6]`,
    `resolution [3]`; flat cell index = (x * res_y + y) * res_z + z.  Same buffer names as nerfacc, so an
    `occ_grid_state_dict` of a reference checkpoint (shapeRenderer.py:351-352, 360-361) loads.
  * `sampling`: every ray marches the lattice t_k = near + k * render_step_size (near jittered by U[0,1) * step when
    `stratified`); the interval [t_k, t_k + step] is emitted when its mid-point lies before `far_plane`, inside the aabb and
    in an occupied cell.  Output = nerfacc's packed wire format (ray_indices int64, t_starts, t_ends), ray after ray.
  * `update_every_n_steps`: every n steps evaluate `occ_eval_fn` at one random point per cell (all cells during warm-up, else a
    uniform quarter + the occupied cells), occs = max(occs * ema_decay, occ), binaries = occs > min(mean(occs), occ_thre).
The march is a CUDA kernel pair (`tf_occ_march_count` / `tf_occ_march_write`, warp per ray, ballot compaction); there is no CPU path."""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr


class OccGridEstimator(torch.nn.Module):
    def __init__(self, roi_aabb, resolution=128, levels: int = 1):
        super().__init__()
        if levels != 1:
            raise NotImplementedError("the reference builds a single-level grid (shapeRenderer.py:213-215)")
        aabb = torch.as_tensor(roi_aabb, dtype=torch.float32).reshape(-1)
        res = torch.as_tensor([resolution] * 3 if isinstance(resolution, int) else list(resolution), dtype=torch.int32)
        self.levels = 1
        self.cells_per_lvl = int(res.prod())
        self.register_buffer("resolution", res)
        self.register_buffer("aabbs", aabb.reshape(1, 6).clone())
        self.register_buffer("occs", torch.zeros(self.cells_per_lvl))
        self.register_buffer("binaries", torch.zeros([1] + res.tolist(), dtype=torch.bool))
        gx, gy, gz = (torch.arange(int(r)) for r in res)
        self.register_buffer("grid_coords", torch.stack(torch.meshgrid(gx, gy, gz, indexing="ij"), -1).reshape(-1, 3), persistent=False)
        self.last_ray_offsets: Optional[torch.Tensor] = None

    # ---- marching ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sampling(self, rays_o, rays_d, sigma_fn=None, alpha_fn=None, near_plane: float = 0.0, far_plane: float = 1e10,
                 t_min=None, t_max=None, render_step_size: float = 1e-3, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0,
                 stratified: bool = False, cone_angle: float = 0.0, noise: Optional[torch.Tensor] = None):
        """-> (ray_indices [N] int64, t_starts [N], t_ends [N]).  `noise` [R] in [0,1) replaces the stratified draw (tests)."""
        if sigma_fn is not None or alpha_fn is not None or alpha_thre > 0.0:
            raise NotImplementedError("visibility filtering by sigma_fn / alpha_fn is not used by the reference (shapeRenderer.py:953)")
        if cone_angle != 0.0 or t_min is not None or t_max is not None:
            raise NotImplementedError("cone_angle / per-ray t_min,t_max are not used by the reference call")
        lib = _lib.load()
        dev = rays_o.device
        R = rays_o.shape[0]
        step = float(render_step_size)
        o = rays_o.contiguous().float()
        d = rays_d.contiguous().float()
        near = torch.full((R,), float(near_plane), device=dev, dtype=torch.float32)
        if stratified:
            near = near + (torch.rand(R, device=dev) if noise is None else noise.to(dev).float().reshape(-1)) * step
        bits = self.binaries.reshape(-1).view(torch.uint8)
        aabb = (C.c_float * 6)(*self.aabbs.reshape(-1).tolist())
        res = (C.c_int32 * 3)(*self.resolution.tolist())
        counts = torch.empty(R, device=dev, dtype=torch.int32)
        if R > 0:
            check(lib.tf_occ_march_count(ptr(o), ptr(d), ptr(near), R, float(far_plane), step, aabb, res, ptr(bits), ptr(counts),
                                         stream_ptr()), "tf_occ_march_count")
        offsets = torch.zeros(R + 1, device=dev, dtype=torch.int32)
        torch.cumsum(counts, 0, out=offsets[1:])
        n = int(offsets[-1]) if R > 0 else 0        # the packed size is data dependent: one D2H read, as in nerfacc
        ray_indices = torch.empty(n, device=dev, dtype=torch.int64)
        t_starts = torch.empty(n, device=dev, dtype=torch.float32)
        t_ends = torch.empty(n, device=dev, dtype=torch.float32)
        if n > 0:
            check(lib.tf_occ_march_write(ptr(o), ptr(d), ptr(near), R, float(far_plane), step, aabb, res, ptr(bits), ptr(offsets),
                                         ptr(ray_indices), ptr(t_starts), ptr(t_ends), stream_ptr()), "tf_occ_march_write")
        self.last_ray_offsets = offsets
        return ray_indices, t_starts, t_ends

    # ---- refresh ---------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _sample_cells(self, n: int):
        uniform = torch.randint(self.cells_per_lvl, (n,), device=self.occs.device)
        occupied = torch.nonzero(self.binaries.reshape(-1))[:, 0]
        if n < occupied.shape[0]:
            occupied = occupied[torch.randint(occupied.shape[0], (n,), device=self.occs.device)]
        return torch.cat([uniform, occupied], 0)

    @torch.no_grad()
    def _update(self, step: int, occ_eval_fn: Callable, occ_thre: float = 0.01, ema_decay: float = 0.95, warmup_steps: int = 256,
                jitter: Optional[torch.Tensor] = None):
        dev = self.occs.device
        cells = torch.arange(self.cells_per_lvl, device=dev) if step < warmup_steps else self._sample_cells(self.cells_per_lvl // 4)
        coords = self.grid_coords[cells].float()
        u = torch.rand_like(coords) if jitter is None else jitter.to(dev)[cells]
        x = (coords + u) / self.resolution.float()
        x = self.aabbs[0, :3] + x * (self.aabbs[0, 3:] - self.aabbs[0, :3])
        occ = occ_eval_fn(x).reshape(-1)
        self.occs[cells] = torch.maximum(self.occs[cells] * ema_decay, occ)
        thre = torch.clamp(self.occs[self.occs >= 0].mean(), max=occ_thre)
        self.binaries = (self.occs > thre).view(self.binaries.shape)

    @torch.no_grad()
    def update_every_n_steps(self, step: int, occ_eval_fn: Callable, occ_thre: float = 1e-2, ema_decay: float = 0.95,
                             warmup_steps: int = 256, n: int = 16):
        if self.training and step % n == 0:
            self._update(step, occ_eval_fn, occ_thre, ema_decay, warmup_steps)

    @torch.no_grad()
    def mark_all_occupied(self):
        self.occs.fill_(1.0)
        self.binaries.fill_(True)
