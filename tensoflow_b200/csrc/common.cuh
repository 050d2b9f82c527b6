// Shared device helpers for the tensoflow_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tensoflow_b200.h"

void tf_set_error(const char* fmt, ...);
void tf_count_launches(int n);   // bumps the counter behind tf_launch_count()
// scoped CUDA-event timer around a kernel launch; records only while tf_kernel_timing_enable(1) is in effect
struct TfKernelTimer {
    TfKernelTimer(const char* name, cudaStream_t stream);
    ~TfKernelTimer();
    const char* name_; cudaStream_t stream_; cudaEvent_t start_, stop_;
};

#define TF_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            tf_set_error(__VA_ARGS__);        \
            return 1;                         \
        }                                     \
    } while (0)

#define TF_CHECK_LAUNCH(name)                                                      \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            tf_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));  \
            return 2;                                                              \
        }                                                                          \
    } while (0)

static inline int tf_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// nn.Softplus(beta=100): x when 100x > 20 (network/fields.py:79)
__device__ __forceinline__ float softplus100(float x) {
    float bx = 100.f * x;
    return bx > 20.f ? x : log1pf(expf(bx)) * 0.01f;
}
// d softplus100 / dx = sigmoid(100 x) (1 in the linear branch)
__device__ __forceinline__ float softplus100_grad(float x) {
    float bx = 100.f * x;
    return bx > 20.f ? 1.f : 1.f / (1.f + expf(-bx));
}

// Fast Softplus(beta=100) for the tensor-core epilogues:
//   softplus(x) = max(x,0) + log1p(exp(-|100 x|)) / 100
// exp through ex2.approx (rel. error 2^-22), log1p(t) = t*g(t) with a degree-9 interpolant of
// g on [0,1] (max abs error 1.1e-7 in fp32 Horner) -> absolute error of the result ~1e-9, i.e.
// fp32-rounding level for the hidden activations; 1 MUFU + ~14 FP ops instead of expf+log1pf (~50).
__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float log1p_01(float t) {
    float g = -3.176057010e-03f;
    g = fmaf(g, t, 1.954252722e-02f);
    g = fmaf(g, t, -5.637361275e-02f);
    g = fmaf(g, t, 1.054362379e-01f);
    g = fmaf(g, t, -1.526966707e-01f);
    g = fmaf(g, t, 1.966327426e-01f);
    g = fmaf(g, t, -2.495161626e-01f);
    g = fmaf(g, t, 3.332971050e-01f);
    g = fmaf(g, t, -4.999989265e-01f);
    g = fmaf(g, t, 9.999999947e-01f);
    return g * t;
}
__device__ __forceinline__ float softplus100_fast(float x) {
    const float t = ex2_approx(-fabsf(x) * 144.26950408889634f);   // exp(-|100x|) = 2^(-|x| * 100*log2(e))
    return fmaf(log1p_01(t), 0.01f, fmaxf(x, 0.f));
}
// softplus and its derivative sigmoid(100x) from one exponential
__device__ __forceinline__ void softplus100_fast_both(float x, float& sp, float& sg) {
    const float t = ex2_approx(-fabsf(x) * 144.26950408889634f);
    sp = fmaf(log1p_01(t), 0.01f, fmaxf(x, 0.f));
    const float r = __fdividef(1.f, 1.f + t);
    sg = x >= 0.f ? r : t * r;
}

// ---- packed fp32 pairs (sm_100a FFMA2 / FMUL2 / FADD2): one issue slot for two lanes of work; a pair built from the same
// ---- register or from a constant is a free broadcast operand in SASS.  Used where the epilogues are issue-bound.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fmul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t bcast2(float c) { return pack2(c, c); }
// log1p_01 of a pair (same polynomial, same operation order as the scalar version: bit-identical results)
__device__ __forceinline__ uint64_t log1p_01_2(uint64_t t) {
    uint64_t g = bcast2(-3.176057010e-03f);
    g = ffma2(g, t, bcast2(1.954252722e-02f));
    g = ffma2(g, t, bcast2(-5.637361275e-02f));
    g = ffma2(g, t, bcast2(1.054362379e-01f));
    g = ffma2(g, t, bcast2(-1.526966707e-01f));
    g = ffma2(g, t, bcast2(1.966327426e-01f));
    g = ffma2(g, t, bcast2(-2.495161626e-01f));
    g = ffma2(g, t, bcast2(3.332971050e-01f));
    g = ffma2(g, t, bcast2(-4.999989265e-01f));
    g = ffma2(g, t, bcast2(9.999999947e-01f));
    return fmul2(g, t);
}
// softplus100_fast of two values
__device__ __forceinline__ void softplus100_fast2(float x0, float x1, float& s0, float& s1) {
    const float t0 = ex2_approx(-fabsf(x0) * 144.26950408889634f), t1 = ex2_approx(-fabsf(x1) * 144.26950408889634f);
    const uint64_t r = ffma2(log1p_01_2(pack2(t0, t1)), bcast2(0.01f), pack2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
    unpack2(r, s0, s1);
}
// softplus100_fast_both of two values
__device__ __forceinline__ void softplus100_fast_both2(float x0, float x1, float& s0, float& s1, float& g0, float& g1) {
    const float t0 = ex2_approx(-fabsf(x0) * 144.26950408889634f), t1 = ex2_approx(-fabsf(x1) * 144.26950408889634f);
    const uint64_t t = pack2(t0, t1);
    const uint64_t r = ffma2(log1p_01_2(t), bcast2(0.01f), pack2(fmaxf(x0, 0.f), fmaxf(x1, 0.f)));
    unpack2(r, s0, s1);
    float d0, d1;
    unpack2(fadd2(t, bcast2(1.f)), d0, d1);
    const float r0 = __fdividef(1.f, d0), r1 = __fdividef(1.f, d1);
    g0 = x0 >= 0.f ? r0 : t0 * r0;
    g1 = x1 >= 0.f ? r1 : t1 * r1;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// vector reduction into global memory (sm_90+): 4 floats, 16-byte aligned
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// ---------------------------------------------------------------------------
// dr.texture(..., boundary_mode='clamp') sampling plan for one level of one
// 2-D texture: 4 texel offsets (in texels, row-major y*W+x) and weights.
// Texel centres at (i+.5)/W; indices clamped to the edge.
// ---------------------------------------------------------------------------
struct BiTap {
    int o00, o01, o10, o11;
    float w00, w01, w10, w11;
};

__device__ __forceinline__ BiTap make_bitap(float u, float v, int W, int H) {
    float x = u * (float)W - 0.5f;
    float y = v * (float)H - 0.5f;
    x = fminf(fmaxf(x, -2.f), (float)W + 1.f);
    y = fminf(fmaxf(y, -2.f), (float)H + 1.f);
    float x0f = floorf(x), y0f = floorf(y);
    float fx = x - x0f, fy = y - y0f;
    int x0 = (int)x0f, y0 = (int)y0f;
    int x1 = min(max(x0 + 1, 0), W - 1), y1 = min(max(y0 + 1, 0), H - 1);
    x0 = min(max(x0, 0), W - 1);
    y0 = min(max(y0, 0), H - 1);
    BiTap t;
    t.o00 = y0 * W + x0; t.o01 = y0 * W + x1; t.o10 = y1 * W + x0; t.o11 = y1 * W + x1;
    t.w00 = (1.f - fx) * (1.f - fy); t.w01 = fx * (1.f - fy);
    t.w10 = (1.f - fx) * fy;         t.w11 = fx * fy;
    return t;
}

// 1-D plan along a [G][C] line texture sampled at uv=(0,v) on a width-1 texture:
// the horizontal pair collapses onto the single column, leaving a linear blend.
struct LiTap {
    int o0, o1;
    float w0, w1;
};
__device__ __forceinline__ LiTap make_litap(float v, int G) {
    float y = v * (float)G - 0.5f;
    y = fminf(fmaxf(y, -2.f), (float)G + 1.f);
    float y0f = floorf(y);
    float fy = y - y0f;
    int y0 = (int)y0f;
    LiTap t;
    t.o1 = min(max(y0 + 1, 0), G - 1);
    t.o0 = min(max(y0, 0), G - 1);
    t.w0 = 1.f - fy;
    t.w1 = fy;
    return t;
}

// trilinear level selection: level clamped to [0, L-1]; l1 = min(l0+1, L-1)
__device__ __forceinline__ void mip_levels(float level, int L, int& l0, int& l1, float& f) {
    float lv = fminf(fmaxf(level, 0.f), (float)(L - 1));
    float fl = floorf(lv);
    l0 = (int)fl;
    f = lv - fl;
    l1 = min(l0 + 1, L - 1);
    if (l1 == l0) f = 0.f;
}

// pointer to level `l` of plane i / line i inside the field description
__device__ __forceinline__ const float* plane_level_ptr(const tf_vm_field_t& f, int i, int l, int& H, int& W) {
    H = f.plane_h[i]; W = f.plane_w[i];
    if (l == 0) return f.plane[i];
    size_t off = 0;
    int h = H, w = W;
    for (int k = 1; k < l; ++k) { h >>= 1; w >>= 1; off += (size_t)h * w; }
    H = H >> l; W = W >> l;
    return f.plane_mip[i] + off * f.n_comp;
}
__device__ __forceinline__ const float* line_level_ptr(const tf_vm_field_t& f, int i, int l, int& G) {
    G = f.line_g[i];
    if (l == 0) return f.line[i];
    size_t off = 0;
    int g = G;
    for (int k = 1; k < l; ++k) { g >>= 1; off += (size_t)g; }
    G = G >> l;
    return f.line_mip[i] + off * f.n_comp;
}
__device__ __forceinline__ float* plane_level_ptr_mut(const tf_vm_field_t& f, const tf_vm_mut_t& g, int i, int l, int& H, int& W) {
    H = f.plane_h[i]; W = f.plane_w[i];
    if (l == 0) return g.plane[i];
    size_t off = 0;
    int h = H, w = W;
    for (int k = 1; k < l; ++k) { h >>= 1; w >>= 1; off += (size_t)h * w; }
    H = H >> l; W = W >> l;
    return g.plane_mip[i] + off * f.n_comp;
}
__device__ __forceinline__ float* line_level_ptr_mut(const tf_vm_field_t& f, const tf_vm_mut_t& g, int i, int l, int& G) {
    G = f.line_g[i];
    if (l == 0) return g.line[i];
    size_t off = 0;
    int gg = G;
    for (int k = 1; k < l; ++k) { gg >>= 1; off += (size_t)gg; }
    G = G >> l;
    return g.line_mip[i] + off * f.n_comp;
}

__device__ __forceinline__ float4 f4_fma(float a, float4 b, float4 c) {
    return make_float4(fmaf(a, b.x, c.x), fmaf(a, b.y, c.y), fmaf(a, b.z, c.z), fmaf(a, b.w, c.w));
}
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_scale(float a, float4 b) { return make_float4(a * b.x, a * b.y, a * b.z, a * b.w); }
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// bilinear fetch of 4 channels [c4*4 .. c4*4+3] of a channels-last texture level
__device__ __forceinline__ float4 fetch_bi(const float* tex, const BiTap& t, int C, int c) {
    float4 r = f4_scale(t.w00, ldg4(tex + (size_t)t.o00 * C + c));
    r = f4_fma(t.w01, ldg4(tex + (size_t)t.o01 * C + c), r);
    r = f4_fma(t.w10, ldg4(tex + (size_t)t.o10 * C + c), r);
    r = f4_fma(t.w11, ldg4(tex + (size_t)t.o11 * C + c), r);
    return r;
}
__device__ __forceinline__ float4 fetch_li(const float* tex, const LiTap& t, int C, int c) {
    float4 r = f4_scale(t.w0, ldg4(tex + (size_t)t.o0 * C + c));
    return f4_fma(t.w1, ldg4(tex + (size_t)t.o1 * C + c), r);
}

// u = contraction(x) for plane/line i (utils/network_utils.py:90, fields.py:268-270)
__device__ __forceinline__ void vm_coords(const tf_vm_field_t& f, const float q[3], int i, float& pu, float& pv, float& lv) {
    const int m0 = (i == 2) ? 1 : 0;          // matMode = (0,1),(0,2),(1,2)
    const int m1 = (i == 0) ? 1 : 2;
    const int vm = 2 - i;                     // vecMode = 2,1,0
    pu = (q[m0] - f.aabb_min[m0]) / (f.aabb_max[m0] - f.aabb_min[m0]);
    pv = (q[m1] - f.aabb_min[m1]) / (f.aabb_max[m1] - f.aabb_min[m1]);
    lv = (q[vm] - f.aabb_min[vm]) / (f.aabb_max[vm] - f.aabb_min[vm]);
}

// plane_i(x) and line_i(x) for channels c..c+3, trilinear over mip levels
__device__ __forceinline__ void vm_sample(const tf_vm_field_t& f, const float q[3], float level, bool has_level,
                                          int i, int c, float4& P, float4& Lv) {
    float pu, pv, lv;
    vm_coords(f, q, i, pu, pv, lv);
    const int C = f.n_comp;
    int l0 = 0, l1 = 0;
    float fl = 0.f;
    if (has_level && f.n_levels > 1) mip_levels(level, f.n_levels, l0, l1, fl);
    int H, W, G;
    const float* pt = plane_level_ptr(f, i, l0, H, W);
    P = fetch_bi(pt, make_bitap(pu, pv, W, H), C, c);
    const float* lt = line_level_ptr(f, i, l0, G);
    Lv = fetch_li(lt, make_litap(lv, G), C, c);
    if (fl > 0.f) {
        pt = plane_level_ptr(f, i, l1, H, W);
        float4 P1 = fetch_bi(pt, make_bitap(pu, pv, W, H), C, c);
        lt = line_level_ptr(f, i, l1, G);
        float4 L1 = fetch_li(lt, make_litap(lv, G), C, c);
        const float a = 1.f - fl;
        P = make_float4(a * P.x + fl * P1.x, a * P.y + fl * P1.y, a * P.z + fl * P1.z, a * P.w + fl * P1.w);
        Lv = make_float4(a * Lv.x + fl * L1.x, a * Lv.y + fl * L1.y, a * Lv.z + fl * L1.z, a * Lv.w + fl * L1.w);
    }
}

// Sampling plan of plane i / line i at one point, computed once and reused for every channel group
struct VmTaps {
    const float* pt0; const float* pt1; const float* lt0; const float* lt1;
    BiTap b0, b1;
    LiTap t0, t1;
    float fl;
};
__device__ __forceinline__ VmTaps vm_taps(const tf_vm_field_t& f, const float q[3], float level, bool has_level, int i) {
    VmTaps t;
    float pu, pv, lv;
    vm_coords(f, q, i, pu, pv, lv);
    int l0 = 0, l1 = 0;
    t.fl = 0.f;
    if (has_level && f.n_levels > 1) mip_levels(level, f.n_levels, l0, l1, t.fl);
    int H, W, G;
    t.pt0 = plane_level_ptr(f, i, l0, H, W);
    t.b0 = make_bitap(pu, pv, W, H);
    t.lt0 = line_level_ptr(f, i, l0, G);
    t.t0 = make_litap(lv, G);
    t.pt1 = t.pt0; t.lt1 = t.lt0; t.b1 = t.b0; t.t1 = t.t0;
    if (t.fl > 0.f) {
        t.pt1 = plane_level_ptr(f, i, l1, H, W);
        t.b1 = make_bitap(pu, pv, W, H);
        t.lt1 = line_level_ptr(f, i, l1, G);
        t.t1 = make_litap(lv, G);
    }
    return t;
}
__device__ __forceinline__ void vm_fetch(const VmTaps& t, int C, int c, float4& P, float4& Lv) {
    P = fetch_bi(t.pt0, t.b0, C, c);
    Lv = fetch_li(t.lt0, t.t0, C, c);
    if (t.fl > 0.f) {
        const float4 P1 = fetch_bi(t.pt1, t.b1, C, c);
        const float4 L1 = fetch_li(t.lt1, t.t1, C, c);
        const float a = 1.f - t.fl, fl = t.fl;
        P = make_float4(a * P.x + fl * P1.x, a * P.y + fl * P1.y, a * P.z + fl * P1.z, a * P.w + fl * P1.w);
        Lv = make_float4(a * Lv.x + fl * L1.x, a * Lv.y + fl * L1.y, a * Lv.z + fl * L1.z, a * Lv.w + fl * L1.w);
    }
}

// scatter with a precomputed plan (mutable twins of the read pointers are passed by the caller)
__device__ __forceinline__ void vm_scatter_taps(const VmTaps& t, float* pt0, float* pt1, float* lt0, float* lt1, int C, int c,
                                                float4 dP, float4 dL) {
    const float w0 = 1.f - t.fl;
    red_add_v4(pt0 + (size_t)t.b0.o00 * C + c, f4_scale(w0 * t.b0.w00, dP));
    red_add_v4(pt0 + (size_t)t.b0.o01 * C + c, f4_scale(w0 * t.b0.w01, dP));
    red_add_v4(pt0 + (size_t)t.b0.o10 * C + c, f4_scale(w0 * t.b0.w10, dP));
    red_add_v4(pt0 + (size_t)t.b0.o11 * C + c, f4_scale(w0 * t.b0.w11, dP));
    red_add_v4(lt0 + (size_t)t.t0.o0 * C + c, f4_scale(w0 * t.t0.w0, dL));
    red_add_v4(lt0 + (size_t)t.t0.o1 * C + c, f4_scale(w0 * t.t0.w1, dL));
    if (t.fl > 0.f) {
        const float w1 = t.fl;
        red_add_v4(pt1 + (size_t)t.b1.o00 * C + c, f4_scale(w1 * t.b1.w00, dP));
        red_add_v4(pt1 + (size_t)t.b1.o01 * C + c, f4_scale(w1 * t.b1.w01, dP));
        red_add_v4(pt1 + (size_t)t.b1.o10 * C + c, f4_scale(w1 * t.b1.w10, dP));
        red_add_v4(pt1 + (size_t)t.b1.o11 * C + c, f4_scale(w1 * t.b1.w11, dP));
        red_add_v4(lt1 + (size_t)t.t1.o0 * C + c, f4_scale(w1 * t.t1.w0, dL));
        red_add_v4(lt1 + (size_t)t.t1.o1 * C + c, f4_scale(w1 * t.t1.w1, dL));
    }
}

// scatter dP (grad wrt plane_i(x)) and dL (grad wrt line_i(x)) into the factor grads
__device__ __forceinline__ void vm_scatter(const tf_vm_field_t& f, const tf_vm_mut_t& g, const float q[3], float level,
                                           bool has_level, int i, int c, float4 dP, float4 dL) {
    float pu, pv, lv;
    vm_coords(f, q, i, pu, pv, lv);
    const int C = f.n_comp;
    int l0 = 0, l1 = 0;
    float fl = 0.f;
    if (has_level && f.n_levels > 1) mip_levels(level, f.n_levels, l0, l1, fl);
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int l = pass == 0 ? l0 : l1;
        const float wl = pass == 0 ? 1.f - fl : fl;
        if (pass == 1 && !(fl > 0.f)) break;
        int H, W, G;
        float* pt = plane_level_ptr_mut(f, g, i, l, H, W);
        BiTap t = make_bitap(pu, pv, W, H);
        red_add_v4(pt + (size_t)t.o00 * C + c, f4_scale(wl * t.w00, dP));
        red_add_v4(pt + (size_t)t.o01 * C + c, f4_scale(wl * t.w01, dP));
        red_add_v4(pt + (size_t)t.o10 * C + c, f4_scale(wl * t.w10, dP));
        red_add_v4(pt + (size_t)t.o11 * C + c, f4_scale(wl * t.w11, dP));
        float* lt = line_level_ptr_mut(f, g, i, l, G);
        LiTap s = make_litap(lv, G);
        red_add_v4(lt + (size_t)s.o0 * C + c, f4_scale(wl * s.w0, dL));
        red_add_v4(lt + (size_t)s.o1 * C + c, f4_scale(wl * s.w1, dL));
    }
}

// the 7 stencil queries: 0 centre, 1/2 +-x, 3/4 +-y, 5/6 +-z (fields.py:239-244)
__device__ __forceinline__ void stencil_point(const float x[3], const float units[3], int qi, float q[3]) {
    q[0] = x[0]; q[1] = x[1]; q[2] = x[2];
    if (qi > 0) {
        const int ax = (qi - 1) >> 1;
        const float e = ((qi - 1) & 1) ? -units[ax] : units[ax];
        q[ax] = x[ax] + e;
    }
}
