"""Material-stage operators over the C ABI: mesh ray tracer, direction sets, env-light
lookup and the BRDF / Monte-Carlo estimator (reference network/fields.py:824-1335,
network/light.py:125-162, raytracing/raytracer.py)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr
from .ops import _f32c, _timed


class RayTracer:
    """Drop-in for the reference's raytracing.RayTracer (raytracing/raytracer.py:8-54):
    RayTracer(vertices, triangles).trace(rays_o, rays_d) -> positions, face_normals, depth."""

    def __init__(self, vertices, triangles):
        if torch.is_tensor(vertices):
            vertices = vertices.detach().cpu().numpy()
        if torch.is_tensor(triangles):
            triangles = triangles.detach().cpu().numpy()
        assert triangles.shape[0] > 8, "BVH needs at least 8 triangles."
        v = np.ascontiguousarray(vertices, dtype=np.float32)
        t = np.ascontiguousarray(triangles, dtype=np.int32)
        self._h = C.c_void_p()
        lib = _lib.load()
        check(lib.tf_bvh_create(v.ctypes.data_as(C.c_void_p), v.shape[0], t.ctypes.data_as(C.c_void_p), t.shape[0],
                                C.byref(self._h)), "tf_bvh_create")
        self.n_triangles = int(t.shape[0])

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _lib.load().tf_bvh_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def trace(self, rays_o, rays_d, inplace=False):
        rays_o = _f32c(rays_o)
        rays_d = _f32c(rays_d)
        prefix = rays_o.shape[:-1]
        o, d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
        n = o.shape[0]
        positions = torch.empty_like(o)
        normals = torch.empty_like(d)
        depth = torch.empty(n, device=o.device, dtype=torch.float32)
        with _timed("bvh_trace"):
            check(_lib.load().tf_bvh_trace(self._h, ptr(o), ptr(d), n, ptr(positions), ptr(normals), ptr(depth), stream_ptr()),
                  "tf_bvh_trace")
        return positions.reshape(*prefix, 3), normals.reshape(*prefix, 3), depth.reshape(*prefix)


def mc_directions(mode: int, normals, view_dirs, src, aux, roughness, n_dirs: int, dirs_out, prob_out, offset: int):
    """Fill dirs_out[:, offset:offset+n_dirs], prob_out[:, offset:offset+n_dirs] (no autograd:
    direction sets are geometry + frozen-flow samples in the reference)."""
    pn, D = prob_out.shape
    with _timed("mc_directions"):
        check(_lib.load().tf_mc_directions(mode, ptr(_f32c(normals)), ptr(_f32c(view_dirs)), ptr(_f32c(src)), ptr(_f32c(aux)),
                                           ptr(_f32c(roughness)), pn, n_dirs, ptr(dirs_out), ptr(prob_out), D, offset, stream_ptr()),
              "tf_mc_directions")


def hit_encode(inters, dirs, hit_normals, idx, ide_mat, ide_m, ldx: int = 128):
    """[posenc(hit point, 8) | IDE(mirrored view direction)] rows of the occluded pairs `idx`, zero-padded to ldx columns
    (the inner-light MLP input of reference fields.py:951-975) in one kernel; no autograd (hit records are geometry)."""
    M = int(idx.shape[0])
    X = torch.empty(M, ldx, device=inters.device, dtype=torch.float32)
    with _timed("hit_encode"):
        check(_lib.load().tf_hit_encode(ptr(_f32c(inters.reshape(-1, 3))), ptr(_f32c(dirs.reshape(-1, 3))), ptr(_f32c(hit_normals.reshape(-1, 3))),
                                        ptr(idx.contiguous()), M, ptr(ide_mat), ptr(ide_m), ldx, ptr(X), stream_ptr()), "tf_hit_encode")
    return X


class CubeLightFunction(torch.autograd.Function):
    """EnvLight.direct_light: exp(cube bilinear(base, dirs)) on the masked pairs, 0 elsewhere."""

    @staticmethod
    def forward(ctx, base, dirs, mask):
        basec, dirsc = _f32c(base), _f32c(dirs.reshape(-1, 3))
        maskc = None if mask is None else mask.reshape(-1).to(torch.uint8).contiguous()
        n = dirsc.shape[0]
        out = torch.empty(n, 3, device=dirsc.device, dtype=torch.float32)
        with _timed("cube_light_fwd"):
            check(_lib.load().tf_cube_light_fwd(ptr(basec), int(basec.shape[1]), ptr(dirsc), ptr(maskc), n, ptr(out), stream_ptr()),
                  "tf_cube_light_fwd")
        ctx.save_for_backward(dirsc, maskc, out)
        ctx.base_shape = basec.shape
        return out.reshape(*dirs.shape)

    @staticmethod
    def backward(ctx, g):
        dirsc, maskc, out = ctx.saved_tensors
        d_base = torch.zeros(ctx.base_shape, device=out.device, dtype=torch.float32)
        gc = _f32c(g.reshape(-1, 3))
        with _timed("cube_light_bwd"):
            check(_lib.load().tf_cube_light_bwd(int(ctx.base_shape[1]), ptr(dirsc), ptr(maskc), dirsc.shape[0], ptr(out), ptr(gc),
                                                ptr(d_base), stream_ptr()), "tf_cube_light_bwd")
        return d_base, None, None


class McEstimateFunction(torch.autograd.Function):
    """BRDF weights + diffuse / specular Monte-Carlo estimators per surface point, plus the per-point sums of the two
    neural-importance-sampling losses (reference fields.py:1254-1333) when the flows' log q are given.
    returns out [pn,19]: diffuse(3) specular(3) mean diffuse light(3) mean specular light(3) visibility(1) indirect(3)
    nis_diffuse_sum(1) nis_specular_sum(1) nis_specular_count(1)."""

    @staticmethod
    def forward(ctx, normals, view_dirs, albedo, metallic, roughness, dirs, prob, lights, hit, n_diffuse,
                logq_d=None, ang_d=None, logq_s=None, ang_s=None):
        t = [_f32c(x) for x in (normals, view_dirs, albedo, metallic.reshape(-1), roughness.reshape(-1), dirs, prob, lights)]
        hitc = hit.to(torch.uint8).contiguous()
        pn, D = t[6].shape
        nis = [None if x is None else _f32c(x) for x in (logq_d, ang_d, logq_s, ang_s)]
        nd = 0 if nis[0] is None else int(nis[0].numel() // max(pn, 1))
        out = torch.empty(pn, 19, device=t[0].device, dtype=torch.float32)
        with _timed("mc_estimate_fwd"):
            check(_lib.load().tf_mc_estimate_fwd(*(ptr(x) for x in t), ptr(hitc), pn, int(n_diffuse), int(D - n_diffuse), ptr(nis[0]),
                                                 ptr(nis[1]), nd, ptr(nis[2]), ptr(nis[3]), ptr(out), stream_ptr()), "tf_mc_estimate_fwd")
        ctx.save_for_backward(*t, hitc, *[x for x in nis if x is not None])
        ctx.has_nis = [x is not None for x in nis]
        ctx.n_diffuse, ctx.nd = int(n_diffuse), nd
        ctx.m_shape, ctx.r_shape = metallic.shape, roughness.shape
        ctx.q_shapes = (None if logq_d is None else logq_d.shape, None if logq_s is None else logq_s.shape)
        return out

    @staticmethod
    def backward(ctx, g_out):
        saved = list(ctx.saved_tensors)
        t, hitc, rest = saved[:8], saved[8], saved[9:]
        nis = [rest.pop(0) if h else None for h in ctx.has_nis]
        pn, D = t[6].shape
        dev = t[0].device
        d_alb = torch.empty(pn, 3, device=dev, dtype=torch.float32)
        d_met = torch.empty(pn, device=dev, dtype=torch.float32)
        d_rgh = torch.empty(pn, device=dev, dtype=torch.float32)
        d_lights = torch.empty(pn, D, 3, device=dev, dtype=torch.float32)
        d_qd = None if nis[0] is None else torch.empty_like(nis[0])
        d_qs = None if nis[2] is None else torch.empty_like(nis[2])
        go = _f32c(g_out)
        with _timed("mc_estimate_bwd"):
            check(_lib.load().tf_mc_estimate_bwd(*(ptr(x) for x in t), ptr(hitc), pn, ctx.n_diffuse, D - ctx.n_diffuse, ptr(nis[0]),
                                                 ptr(nis[1]), ctx.nd, ptr(nis[2]), ptr(nis[3]), ptr(go), ptr(d_alb), ptr(d_met), ptr(d_rgh),
                                                 ptr(d_lights), ptr(d_qd), ptr(d_qs), stream_ptr()), "tf_mc_estimate_bwd")
        return (None, None, d_alb, d_met.reshape(ctx.m_shape), d_rgh.reshape(ctx.r_shape), None, None, d_lights, None, None,
                None if d_qd is None else d_qd.reshape(ctx.q_shapes[0]), None, None if d_qs is None else d_qs.reshape(ctx.q_shapes[1]), None)
