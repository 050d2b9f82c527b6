// Shared-stencil sampling of the VM field for the tensor-core decoder kernels.
//
// The 7 finite-difference queries of a sample (centre, +-x, +-y, +-z; reference network/fields.py:239-244)
// move one coordinate at a time, and plane i only sees two of the three coordinates while line i sees the
// third.  So of the 7 (plane, line) fetch pairs per plane index only 5 plane positions and 3 line positions
// are distinct: 78 texel taps per sample and level instead of 126.  The values are bit-identical to sampling
// every query on its own because the shared positions have identical coordinates.
//
// MMA tiles are therefore sample-major: row r = s * 7 + q for the 18 samples of a tile (126 rows, 2 zero rows),
// so one thread produces / consumes all 7 rows of a (sample, plane, channel group) site.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace site {

constexpr int NQ = 7;
constexpr int SPT = 18;                 // samples per 128-row MMA tile (stencil mode)
constexpr uint32_t A_LBO = 144;         // K-chunk stride of the A operand: 128-byte core matrix + 16 bytes (bank spread)

__device__ __forceinline__ uint32_t a_sbo(int KT) { return (uint32_t)(KT / 4) * A_LBO; }
__device__ __forceinline__ uint32_t a_part_bytes(int KT) { return 16u * a_sbo(KT); }
// byte offset of the 16-byte unit (row r, column group g) in the A operand
__device__ __forceinline__ uint32_t a_off(int r, int g, int KT) { return (uint32_t)(r >> 3) * a_sbo(KT) + (uint32_t)g * A_LBO + (uint32_t)(r & 7) * 16; }

struct Levels {
    const float* pt0; const float* pt1; const float* lt0; const float* lt1;
    int W0, H0, W1, H1, G0, G1;
    float fl;                           // weight of the second level (0 = single level)
};

// per-kernel table of the mip levels (pointers / extents of every plane and line level) in shared memory: the
// per-task level lookup is two indexed reads instead of loops over the mip chain and constant-bank indexing
constexpr int MAXL = 8;
struct LevelTab {
    const float* plane[3 * MAXL]; const float* line[3 * MAXL];
    int pw[3 * MAXL], ph[3 * MAXL], lg[3 * MAXL];
};
// all threads of the CTA call this before the first gather (followed by a CTA-wide barrier)
__device__ __forceinline__ void build_level_tab(const tf_vm_field_t& f, LevelTab* t) {
    for (int e = threadIdx.x; e < 3 * MAXL; e += blockDim.x) {
        const int i = e / MAXL, l = e % MAXL;
        if (l < f.n_levels) {
            int H, W, G;
            t->plane[e] = plane_level_ptr(f, i, l, H, W);
            t->ph[e] = H; t->pw[e] = W;
            t->line[e] = line_level_ptr(f, i, l, G);
            t->lg[e] = G;
        }
    }
}

__device__ __forceinline__ Levels levels(const tf_vm_field_t& f, const LevelTab& tab, float level, bool has_level, int i) {
    Levels L;
    int l0 = 0, l1 = 0;
    L.fl = 0.f;
    if (has_level && f.n_levels > 1) mip_levels(level, f.n_levels, l0, l1, L.fl);
    const int e0 = i * MAXL + l0;
    L.pt0 = tab.plane[e0]; L.H0 = tab.ph[e0]; L.W0 = tab.pw[e0];
    L.lt0 = tab.line[e0]; L.G0 = tab.lg[e0];
    L.pt1 = L.pt0; L.lt1 = L.lt0; L.W1 = L.W0; L.H1 = L.H0; L.G1 = L.G0;
    if (L.fl > 0.f) {
        const int e1 = i * MAXL + l1;
        L.pt1 = tab.plane[e1]; L.H1 = tab.ph[e1]; L.W1 = tab.pw[e1];
        L.lt1 = tab.line[e1]; L.G1 = tab.lg[e1];
    }
    return L;
}

__device__ __forceinline__ float4 mix(float fl, float4 a, float4 b) {
    const float w = 1.f - fl;
    return make_float4(w * a.x + fl * b.x, w * a.y + fl * b.y, w * a.z + fl * b.z, w * a.w + fl * b.w);
}

// 1-D sampling plan (same arithmetic as make_bitap / make_litap: texel centres at i + 0.5, clamp addressing)
struct Tap1 { int i0, i1; float w0, w1; };
__device__ __forceinline__ Tap1 tap1(float u, int N) {
    float x = u * (float)N - 0.5f;
    x = fminf(fmaxf(x, -2.f), (float)N + 1.f);
    const float x0f = floorf(x);
    const float fx = x - x0f;
    const int x0 = (int)x0f;
    Tap1 t;
    t.i1 = min(max(x0 + 1, 0), N - 1);
    t.i0 = min(max(x0, 0), N - 1);
    t.w0 = 1.f - fx;
    t.w1 = fx;
    return t;
}
// the 4 texels of a bilinear footprint / the 2 texels of a linear one (loads only, no arithmetic: callers batch
// the loads of several positions before combining them so that many requests are in flight)
// element offsets are 32-bit unsigned (a level holds < 2^32 floats): one IMAD.WIDE.U32 per address instead of
// sign extensions and 64-bit multiplies (the gather / scatter are issue-bound as much as latency-bound)
__device__ __forceinline__ uint32_t texel_off(int iy, int ix, int W, int C, int c) { return ((uint32_t)(iy * W) + (uint32_t)ix) * (uint32_t)C + (uint32_t)c; }
__device__ __forceinline__ void bi_load(const float* T, const Tap1& tx, const Tap1& ty, int W, int C, int c, float4 t[4]) {
    const uint32_t r0 = (uint32_t)(ty.i0 * W), r1 = (uint32_t)(ty.i1 * W), uc = (uint32_t)C, cc = (uint32_t)c;
    t[0] = ldg4(T + ((r0 + (uint32_t)tx.i0) * uc + cc));
    t[1] = ldg4(T + ((r0 + (uint32_t)tx.i1) * uc + cc));
    t[2] = ldg4(T + ((r1 + (uint32_t)tx.i0) * uc + cc));
    t[3] = ldg4(T + ((r1 + (uint32_t)tx.i1) * uc + cc));
}
__device__ __forceinline__ float4 bi_combine(const Tap1& tx, const Tap1& ty, const float4 t[4]) {
    float4 r = f4_scale(tx.w0 * ty.w0, t[0]);
    r = f4_fma(tx.w1 * ty.w0, t[1], r);
    r = f4_fma(tx.w0 * ty.w1, t[2], r);
    return f4_fma(tx.w1 * ty.w1, t[3], r);
}
__device__ __forceinline__ void li_load(const float* T, const Tap1& tl, int C, int c, float4 t[2]) {
    t[0] = ldg4(T + ((uint32_t)tl.i0 * (uint32_t)C + (uint32_t)c));
    t[1] = ldg4(T + ((uint32_t)tl.i1 * (uint32_t)C + (uint32_t)c));
}
__device__ __forceinline__ float4 li_combine(const Tap1& tl, const float4 t[2]) { return f4_fma(tl.w1, t[1], f4_scale(tl.w0, t[0])); }

// normalised coordinate of value v along axis ax (same arithmetic as vm_coords)
__device__ __forceinline__ float coord(const tf_vm_field_t& f, float v, int ax) { return (v - f.aabb_min[ax]) / (f.aabb_max[ax] - f.aabb_min[ax]); }

struct Axes { int m0, m1, vm; };
__device__ __forceinline__ Axes axes(int i) { Axes a; a.m0 = (i == 2) ? 1 : 0; a.m1 = (i == 0) ? 1 : 2; a.vm = 2 - i; return a; }

// coordinates of the stencil of sample x for plane i: index 0 centre, 1 = +unit, 2 = -unit along the axis
struct Coords { float pu[3], pv[3], lv[3]; };
__device__ __forceinline__ float pick(const float v[3], int ax) { return ax == 0 ? v[0] : (ax == 1 ? v[1] : v[2]); }
__device__ __forceinline__ Coords coords(const tf_vm_field_t& f, const float x[3], const float units[3], const Axes& a) {
    Coords k;
    const float x0 = pick(x, a.m0), x1 = pick(x, a.m1), x2 = pick(x, a.vm);
    const float e0 = pick(units, a.m0), e1 = pick(units, a.m1), e2 = pick(units, a.vm);
    k.pu[0] = coord(f, x0, a.m0); k.pu[1] = coord(f, x0 + e0, a.m0); k.pu[2] = coord(f, x0 + (-e0), a.m0);
    k.pv[0] = coord(f, x1, a.m1); k.pv[1] = coord(f, x1 + e1, a.m1); k.pv[2] = coord(f, x1 + (-e1), a.m1);
    k.lv[0] = coord(f, x2, a.vm); k.lv[1] = coord(f, x2 + e2, a.vm); k.lv[2] = coord(f, x2 + (-e2), a.vm);
    return k;
}

__device__ __forceinline__ float4 tf32_hi(float4 v) { return make_float4(tc::tf32_rn(v.x), tc::tf32_rn(v.y), tc::tf32_rn(v.z), tc::tf32_rn(v.w)); }
__device__ __forceinline__ float4 tf32_lo(float4 v, float4 h) {
    return make_float4(tc::tf32_rn(v.x - h.x), tc::tf32_rn(v.y - h.y), tc::tf32_rn(v.z - h.z), tc::tf32_rn(v.w - h.w));
}
// a_lo == nullptr: the tile is an fp32 staging buffer (the forward kernel splits when it moves the rows to tensor memory)
__device__ __forceinline__ void put(uint8_t* a_hi, uint8_t* a_lo, uint32_t off, float4 v) {
    if (a_lo == nullptr) { *reinterpret_cast<float4*>(a_hi + off) = v; return; }
    const float4 h = tf32_hi(v);
    *reinterpret_cast<float4*>(a_hi + off) = h;
    *reinterpret_cast<float4*>(a_lo + off) = tf32_lo(v, h);
}

__device__ __forceinline__ float* twin(const float* p, const float* base0, float* g0, const float* basem, float* gm) {
    return p == base0 ? g0 : gm + (p - basem);
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4_fma4(float4 a, float4 b, float4 c) { return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w)); }

// ---- register-lean variants (the warp-specialised backward runs its memory group next to a math group and cannot
// ---- afford the ~200 registers of the batched fetch): sampling plans are built position by position (16 registers
// ---- each), 8 texel loads in flight per thread, latency hidden by the group's 8 warps
template <bool TWO> struct PlanePos { Tap1 x0, y0, x1, y1; };
template <bool TWO> struct LinePos { Tap1 l0, l1; };
template <bool TWO>
__device__ __forceinline__ PlanePos<TWO> plane_pos(const Levels& L, float pu, float pv) {
    PlanePos<TWO> p;
    p.x0 = tap1(pu, L.W0); p.y0 = tap1(pv, L.H0);
    if (TWO) { p.x1 = tap1(pu, L.W1); p.y1 = tap1(pv, L.H1); }
    return p;
}
template <bool TWO>
__device__ __forceinline__ LinePos<TWO> line_pos(const Levels& L, float lv) {
    LinePos<TWO> p;
    p.l0 = tap1(lv, L.G0);
    if (TWO) p.l1 = tap1(lv, L.G1);
    return p;
}
template <bool TWO>
__device__ __forceinline__ float4 fetch_plane(const Levels& L, const PlanePos<TWO>& p, int C, int c) {
    float4 t0[4], t1[4];
    bi_load(L.pt0, p.x0, p.y0, L.W0, C, c, t0);
    if (TWO) bi_load(L.pt1, p.x1, p.y1, L.W1, C, c, t1);
    float4 P = bi_combine(p.x0, p.y0, t0);
    if (TWO) P = mix(L.fl, P, bi_combine(p.x1, p.y1, t1));
    return P;
}
template <bool TWO>
__device__ __forceinline__ float4 fetch_line(const Levels& L, const LinePos<TWO>& p, int C, int c) {
    float4 t0[2], t1[2];
    li_load(L.lt0, p.l0, C, c, t0);
    if (TWO) li_load(L.lt1, p.l1, C, c, t1);
    float4 V = li_combine(p.l0, t0);
    if (TWO) V = mix(L.fl, V, li_combine(p.l1, t1));
    return V;
}
template <bool TWO>
__device__ __forceinline__ void scatter_plane_pos(float* t0, float* t1, const Levels& L, const PlanePos<TWO>& p, int C, int c, float4 d) {
    {
        const float w0 = 1.f - L.fl;
        red_add_v4(t0 + texel_off(p.y0.i0, p.x0.i0, L.W0, C, c), f4_scale(w0 * (p.x0.w0 * p.y0.w0), d));
        red_add_v4(t0 + texel_off(p.y0.i0, p.x0.i1, L.W0, C, c), f4_scale(w0 * (p.x0.w1 * p.y0.w0), d));
        red_add_v4(t0 + texel_off(p.y0.i1, p.x0.i0, L.W0, C, c), f4_scale(w0 * (p.x0.w0 * p.y0.w1), d));
        red_add_v4(t0 + texel_off(p.y0.i1, p.x0.i1, L.W0, C, c), f4_scale(w0 * (p.x0.w1 * p.y0.w1), d));
    }
    if (TWO && L.fl > 0.f) {
        red_add_v4(t1 + texel_off(p.y1.i0, p.x1.i0, L.W1, C, c), f4_scale(L.fl * (p.x1.w0 * p.y1.w0), d));
        red_add_v4(t1 + texel_off(p.y1.i0, p.x1.i1, L.W1, C, c), f4_scale(L.fl * (p.x1.w1 * p.y1.w0), d));
        red_add_v4(t1 + texel_off(p.y1.i1, p.x1.i0, L.W1, C, c), f4_scale(L.fl * (p.x1.w0 * p.y1.w1), d));
        red_add_v4(t1 + texel_off(p.y1.i1, p.x1.i1, L.W1, C, c), f4_scale(L.fl * (p.x1.w1 * p.y1.w1), d));
    }
}
template <bool TWO>
__device__ __forceinline__ void scatter_line_pos(float* t0, float* t1, const Levels& L, const LinePos<TWO>& p, int C, int c, float4 d) {
    const float w0 = 1.f - L.fl;
    red_add_v4(t0 + texel_off(0, p.l0.i0, 0, C, c), f4_scale(w0 * p.l0.w0, d));
    red_add_v4(t0 + texel_off(0, p.l0.i1, 0, C, c), f4_scale(w0 * p.l0.w1, d));
    if (TWO && L.fl > 0.f) {
        red_add_v4(t1 + texel_off(0, p.l1.i0, 0, C, c), f4_scale(L.fl * p.l1.w0, d));
        red_add_v4(t1 + texel_off(0, p.l1.i1, 0, C, c), f4_scale(L.fl * p.l1.w1, d));
    }
}

template <bool TWO>
__device__ __forceinline__ void gather_tile_lean_t(const tf_vm_field_t& f, const LevelTab& tab, const float* __restrict__ xyz, const float* __restrict__ level,
                                                   int64_t n_total, const float units[3], int64_t s_base, int KT, uint8_t* a_hi, uint8_t* a_lo,
                                                   float* arow, int nthreads, int tid0) {
    const int C = f.n_comp, C4 = C / 4, G = KT / 4;
    const bool has_level = level != nullptr;
    const int n_tasks = SPT * 3 * C4;
    auto emit = [&](int r, int g, float4 v) {
        put(a_hi, a_lo, a_off(r, g, KT), v);
        if (arow) *reinterpret_cast<float4*>(arow + (size_t)r * KT + g * 4) = v;
    };
    for (int task = tid0; task < n_tasks; task += nthreads) {
        const int c4 = task % C4, si = task / C4, i = si % 3, s = si / 3;
        const int64_t n = s_base + s;
        const int g = i * C4 + c4, r0 = s * NQ, c = c4 * 4;
        const Axes a = axes(i);
        const int r_m0 = r0 + 1 + 2 * a.m0, r_m1 = r0 + 1 + 2 * a.m1, r_vm = r0 + 1 + 2 * a.vm;
        if (n >= n_total) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) emit(r0 + j, g, f4_zero());
            continue;
        }
        const float x[3] = {xyz[n * 3 + 0], xyz[n * 3 + 1], xyz[n * 3 + 2]};
        const Levels L = levels(f, tab, has_level ? level[n] : 0.f, has_level, i);
        const Coords k = coords(f, x, units, a);
        const float4 L0 = fetch_line<TWO>(L, line_pos<TWO>(L, k.lv[0]), C, c);
        const float4 P0 = fetch_plane<TWO>(L, plane_pos<TWO>(L, k.pu[0], k.pv[0]), C, c);
        emit(r0, g, f4_mul(P0, L0));
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float pu = v == 0 ? k.pu[1] : (v == 1 ? k.pu[2] : k.pu[0]);
            const float pv = v == 2 ? k.pv[1] : (v == 3 ? k.pv[2] : k.pv[0]);
            const int r = v < 2 ? r_m0 + v : r_m1 + (v - 2);
            emit(r, g, f4_mul(fetch_plane<TWO>(L, plane_pos<TWO>(L, pu, pv), C, c), L0));
        }
        emit(r_vm, g, f4_mul(P0, fetch_line<TWO>(L, line_pos<TWO>(L, k.lv[1]), C, c)));
        emit(r_vm + 1, g, f4_mul(P0, fetch_line<TWO>(L, line_pos<TWO>(L, k.lv[2]), C, c)));
    }
    // raw stencil points, zero padding groups and the two zero rows of the tile (as in gather_tile_t)
    const int tail_g = G - 3 * C4;
    for (int it = tid0; it < 128 * tail_g; it += nthreads) {
        const int r = it % 128, g = 3 * C4 + it / 128;
        const int s = r / NQ, q = r - s * NQ;
        const int64_t n = s_base + s;
        float4 v = f4_zero(), vh = f4_zero();
        if (g == 3 * C4 && s < SPT && n < n_total) {
            const float x[3] = {xyz[n * 3 + 0], xyz[n * 3 + 1], xyz[n * 3 + 2]};
            float pt[3];
            stencil_point(x, units, q, pt);
            v = make_float4(pt[0], pt[1], pt[2], 0.f);
            vh = make_float4(pt[0], pt[1], pt[2], 1.f);
        }
        put(a_hi, a_lo, a_off(r, g, KT), v);
        if (arow) *reinterpret_cast<float4*>(arow + (size_t)r * KT + g * 4) = vh;
    }
    for (int it = tid0; it < 2 * 3 * C4; it += nthreads) {
        const int r = SPT * NQ + it / (3 * C4), g = it % (3 * C4);
        put(a_hi, a_lo, a_off(r, g, KT), f4_zero());
        if (arow) *reinterpret_cast<float4*>(arow + (size_t)r * KT + g * 4) = f4_zero();
    }
}
__device__ __forceinline__ void gather_tile_lean(const tf_vm_field_t& f, const LevelTab& tab, const float* __restrict__ xyz, const float* __restrict__ level, int64_t n_total,
                                                 const float units[3], int64_t s_base, int KT, uint8_t* a_hi, uint8_t* a_lo, float* arow, int nthreads,
                                                 int tid0) {
    if (level != nullptr && f.n_levels > 1) gather_tile_lean_t<true>(f, tab, xyz, level, n_total, units, s_base, KT, a_hi, a_lo, arow, nthreads, tid0);
    else gather_tile_lean_t<false>(f, tab, xyz, level, n_total, units, s_base, KT, a_hi, a_lo, arow, nthreads, tid0);
}

// dA may live in shared memory or in a global scratch tile written by other warps of the same CTA (read with ld.global.cg)
template <bool TWO>
__device__ __forceinline__ void scatter_tile_lean_t(const tf_vm_field_t& f, const LevelTab& tab, const tf_vm_mut_t& gm, const float* __restrict__ xyz,
                                                    const float* __restrict__ level, int64_t n_total, const float units[3], int64_t s_base,
                                                    const float* dA, int ld, int nthreads, int tid0, bool da_global) {
    const int C = f.n_comp, C4 = C / 4;
    const bool has_level = level != nullptr;
    const int n_tasks = SPT * 3 * C4;
    for (int task = tid0; task < n_tasks; task += nthreads) {
        const int c4 = task % C4, si = task / C4, i = si % 3, s = si / 3;
        const int64_t n = s_base + s;
        if (n >= n_total) continue;
        const int g = i * C4 + c4, r0 = s * NQ, c = c4 * 4;
        const Axes a = axes(i);
        const float x[3] = {xyz[n * 3 + 0], xyz[n * 3 + 1], xyz[n * 3 + 2]};
        const Levels L = levels(f, tab, has_level ? level[n] : 0.f, has_level, i);
        const Coords k = coords(f, x, units, a);
        float* pm0 = twin(L.pt0, f.plane[i], gm.plane[i], f.plane_mip[i], gm.plane_mip[i]);
        float* pm1 = twin(L.pt1, f.plane[i], gm.plane[i], f.plane_mip[i], gm.plane_mip[i]);
        float* lm0 = twin(L.lt0, f.line[i], gm.line[i], f.line_mip[i], gm.line_mip[i]);
        float* lm1 = twin(L.lt1, f.line[i], gm.line[i], f.line_mip[i], gm.line_mip[i]);
        const float* dcol = dA + g * 4;
        auto drow = [&](int r) {
            const float4* q = reinterpret_cast<const float4*>(dcol + (size_t)r * ld);
            return da_global ? __ldcg(q) : *q;
        };
        const int r_m0 = r0 + 1 + 2 * a.m0, r_m1 = r0 + 1 + 2 * a.m1, r_vm = r0 + 1 + 2 * a.vm;
        const LinePos<TWO> l0 = line_pos<TWO>(L, k.lv[0]);
        const PlanePos<TWO> p0 = plane_pos<TWO>(L, k.pu[0], k.pv[0]);
        const float4 L0 = fetch_line<TWO>(L, l0, C, c), P0 = fetch_plane<TWO>(L, p0, C, c);
        const float4 d0 = drow(r0);
        // plane gradient at the centre position: centre and +-vm queries share it
        float4 dP = f4_mul(d0, L0);
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const LinePos<TWO> lp = line_pos<TWO>(L, k.lv[1 + v]);
            const float4 dv = drow(r_vm + v);
            dP = f4_fma4(dv, fetch_line<TWO>(L, lp, C, c), dP);
            scatter_line_pos<TWO>(lm0, lm1, L, lp, C, c, f4_mul(dv, P0));
        }
        scatter_plane_pos<TWO>(pm0, pm1, L, p0, C, c, dP);
        // line gradient at the centre position: centre and the four in-plane queries share it
        float4 dL = f4_mul(d0, P0);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const float pu = v == 0 ? k.pu[1] : (v == 1 ? k.pu[2] : k.pu[0]);
            const float pv = v == 2 ? k.pv[1] : (v == 3 ? k.pv[2] : k.pv[0]);
            const int r = v < 2 ? r_m0 + v : r_m1 + (v - 2);
            const PlanePos<TWO> pp = plane_pos<TWO>(L, pu, pv);
            const float4 d = drow(r);
            dL = f4_fma4(d, fetch_plane<TWO>(L, pp, C, c), dL);
            scatter_plane_pos<TWO>(pm0, pm1, L, pp, C, c, f4_mul(d, L0));
        }
        scatter_line_pos<TWO>(lm0, lm1, L, l0, C, c, dL);
    }
}
__device__ __forceinline__ void scatter_tile_lean(const tf_vm_field_t& f, const LevelTab& tab, const tf_vm_mut_t& gm, const float* __restrict__ xyz,
                                                  const float* __restrict__ level, int64_t n_total, const float units[3], int64_t s_base,
                                                  const float* dA, int ld, int nthreads, int tid0, bool da_global = false) {
    if (level != nullptr && f.n_levels > 1) scatter_tile_lean_t<true>(f, tab, gm, xyz, level, n_total, units, s_base, dA, ld, nthreads, tid0, da_global);
    else scatter_tile_lean_t<false>(f, tab, gm, xyz, level, n_total, units, s_base, dA, ld, nthreads, tid0, da_global);
}

}  // namespace site
