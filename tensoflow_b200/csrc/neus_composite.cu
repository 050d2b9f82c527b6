// NeuS SDF -> alpha and per-ray compositing, fused, one warp per ray.
// Replaces the tail of ShapeRenderer.compute_sdf_alpha (network/shapeRenderer.py:1004-1024),
// nerfacc.render_weight_from_alpha (exclusive cumprod of 1-alpha per ray) and the
// nerfacc.accumulate_along_rays index_adds (network/shapeRenderer.py:1166-1206).
// Samples are packed ray after ray; CSR ray_offsets replace the int64 ray_indices.
#include "common.cuh"

namespace {

constexpr int MAXD = 16;

__device__ __forceinline__ float ldg_stream(const float* p) { return __ldg(p); }   // read-only path, each byte used once per pass

struct AlphaTerms {
    float alpha, raw, P, Nx, est_prev, est_next, tc, inv_s;
};

// shapeRenderer.py:1010-1024
__device__ __forceinline__ AlphaTerms neus_alpha(float sdf, const float g[3], const float d[3], float dist, float inv_s, float r) {
    AlphaTerms t;
    t.inv_s = inv_s;
    t.tc = d[0] * g[0] + d[1] * g[1] + d[2] * g[2];
    const float ic = -(fmaxf(-t.tc * 0.5f + 0.5f, 0.f) * (1.f - r) + fmaxf(-t.tc, 0.f) * r);
    t.est_next = sdf + ic * dist * 0.5f;
    t.est_prev = sdf - ic * dist * 0.5f;
    t.P = 1.f / (1.f + expf(-t.est_prev * inv_s));
    t.Nx = 1.f / (1.f + expf(-t.est_next * inv_s));
    t.raw = (t.P - t.Nx + 1e-5f) / (t.P + 1e-5f);
    t.alpha = fminf(fmaxf(t.raw, 0.f), 1.f);
    return t;
}

__device__ __forceinline__ float inv_s_from_variance(const float* variance) {
    return fminf(fmaxf(expf(__ldg(variance) * 10.f), 1e-6f), 1e6f);   // other_field.py:199-201, shapeRenderer.py:1004
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// inclusive scans across the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v *= u;
    }
    return v;
}
__device__ __forceinline__ float warp_scan_add(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// One warp per ray.  The prefix product makes the 32-sample chunks of a ray a serial chain, and a chunk's loads used to be
// issued only after the previous chunk's scan: with every warp resident at once the kernel lasted (chunks per ray) x (one
// DRAM round trip).  Now U chunks are loaded up front (U x (5 + D) independent loads per lane in flight) and scanned from
// registers, which shortens the chain U-fold; DMAX bounds the register arrays (D <= 8 covers colour + SDF gradient).
template <int DMAX, int U>
__global__ void __launch_bounds__(256) neus_composite_fwd_kernel(
    const float* __restrict__ sdf, const float* __restrict__ grad, const float* __restrict__ dists,
    const float* __restrict__ dirs, const int32_t* __restrict__ offs, int n_rays, const float* __restrict__ variance,
    float cos_anneal, const float* __restrict__ vals, int D, float* __restrict__ alpha_out, float* __restrict__ weights,
    float* __restrict__ acc_out, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const float inv_s = inv_s_from_variance(variance);
    for (int ray = blockIdx.x * wpb + (threadIdx.x >> 5); ray < n_rays; ray += gridDim.x * wpb) {
        const int b = offs[ray], e = offs[ray + 1];
        const float d[3] = {dirs[ray * 3 + 0], dirs[ray * 3 + 1], dirs[ray * 3 + 2]};
        float carry = 1.f, acc = 0.f;
        float o[DMAX];
#pragma unroll
        for (int k = 0; k < DMAX; ++k) o[k] = 0.f;
        for (int base = b; base < e; base += 32 * U) {
            float s_[U], g_[U][3], dist_[U], v_[U][DMAX];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * 32 + lane;
                const bool ok = i < e;
                s_[u] = ok ? ldg_stream(sdf + i) : 0.f;
                dist_[u] = ok ? ldg_stream(dists + i) : 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) g_[u][k] = ok ? ldg_stream(grad + (size_t)i * 3 + k) : 0.f;
#pragma unroll
                for (int k = 0; k < DMAX; ++k) v_[u][k] = (ok && k < D) ? ldg_stream(vals + (size_t)i * D + k) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * 32 + lane;
                const bool ok = i < e;
                const float a = ok ? neus_alpha(s_[u], g_[u], d, dist_[u], inv_s, cos_anneal).alpha : 0.f;
                const float incl = warp_scan_mul(1.f - a, lane);
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = 1.f;
                const float T = carry * excl;
                const float w = a * T;
                carry *= __shfl_sync(0xffffffffu, incl, 31);
                if (ok) {
                    alpha_out[i] = a;
                    weights[i] = w;
                    acc += w;
#pragma unroll
                    for (int k = 0; k < DMAX; ++k) o[k] = fmaf(w, v_[u][k], o[k]);
                }
            }
        }
        acc = warp_sum(acc);
#pragma unroll
        for (int k = 0; k < DMAX; ++k)
            if (k < D) o[k] = warp_sum(o[k]);
        if (lane == 0) {
            acc_out[ray] = acc;
#pragma unroll
            for (int k = 0; k < DMAX; ++k)
                if (k < D) out[(size_t)ray * D + k] = o[k];
        }
    }
}

// Backward.  With u_i = g_acc + g_out . vals_i + g_w_i (= dL/dw_i):
//   dL/dalpha_i = u_i T_i - (sum_{j>i} u_j w_j) / max(1-alpha_i, 1e-10)
// (the running "total - prefix" form nerfacc's own backward uses), then through the
// alpha formula into sdf, the SDF gradient (via true_cos) and inv_s.
template <int DMAX, int U>
__global__ void __launch_bounds__(256) neus_composite_bwd_kernel(
    const float* __restrict__ sdf, const float* __restrict__ grad, const float* __restrict__ dists,
    const float* __restrict__ dirs, const int32_t* __restrict__ offs, int n_rays, const float* __restrict__ variance,
    float cos_anneal, const float* __restrict__ vals, int D, const float* __restrict__ alpha_in,
    const float* __restrict__ weights, const float* __restrict__ acc_in, const float* __restrict__ out_in,
    const float* __restrict__ g_acc, const float* __restrict__ g_out,
    const float* __restrict__ g_w, float* __restrict__ d_sdf, float* __restrict__ d_grad, float* __restrict__ d_vals,
    float* __restrict__ d_variance) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const float inv_s = inv_s_from_variance(variance);
    float ds_total = 0.f;
    for (int ray = blockIdx.x * wpb + (threadIdx.x >> 5); ray < n_rays; ray += gridDim.x * wpb) {
        const int b = offs[ray], e = offs[ray + 1];
        const float d[3] = {dirs[ray * 3 + 0], dirs[ray * 3 + 1], dirs[ray * 3 + 2]};
        const float ga = g_acc ? g_acc[ray] : 0.f;
        float go[DMAX];
#pragma unroll
        for (int k = 0; k < DMAX; ++k) go[k] = (k < D && g_out) ? g_out[(size_t)ray * D + k] : 0.f;
        // total = sum_i u_i w_i with u_i = g_acc + g_out . vals_i + g_w_i.  The first two terms are the forward's own per-ray
        // results, sum_i w_i = acc and sum_i w_i vals_i = out, so no pass over the samples is needed for them; only an upstream
        // gradient on the weights themselves (g_w) needs one (8 B/sample, U chunks in flight).
        float total = ga * acc_in[ray];
#pragma unroll
        for (int k = 0; k < DMAX; ++k)
            if (k < D) total = fmaf(go[k], out_in[(size_t)ray * D + k], total);
        if (g_w) {
            float part = 0.f;
            for (int base = b; base < e; base += 32 * U) {
                float w_[U], gw_[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = base + u * 32 + lane;
                    const bool ok = i < e;
                    w_[u] = ok ? weights[i] : 0.f;
                    gw_[u] = ok ? g_w[i] : 0.f;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) part = fmaf(gw_[u], w_[u], part);
            }
            total += warp_sum(part);
        }
        // pass 2
        float carryT = 1.f, carryS = 0.f;
        for (int base = b; base < e; base += 32 * U) {
            float a_[U], w_[U], u_[U], s_[U], dist_[U], g_[U][3];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * 32 + lane;
                const bool ok = i < e;
                a_[u] = ok ? ldg_stream(alpha_in + i) : 0.f;
                w_[u] = ok ? ldg_stream(weights + i) : 0.f;
                s_[u] = ok ? ldg_stream(sdf + i) : 0.f;
                dist_[u] = ok ? ldg_stream(dists + i) : 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) g_[u][k] = ok ? ldg_stream(grad + (size_t)i * 3 + k) : 0.f;
                float uu = ok ? ga + (g_w ? ldg_stream(g_w + i) : 0.f) : 0.f;
                float v[DMAX];
#pragma unroll
                for (int k = 0; k < DMAX; ++k) v[k] = (ok && k < D) ? ldg_stream(vals + (size_t)i * D + k) : 0.f;
#pragma unroll
                for (int k = 0; k < DMAX; ++k) uu = fmaf(go[k], v[k], uu);
                u_[u] = uu;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = base + u * 32 + lane;
                const bool ok = i < e;
                const float a = a_[u], w = w_[u], uu = u_[u];
                const float incl = warp_scan_mul(1.f - a, lane);
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = 1.f;
                const float T = carryT * excl;
                carryT *= __shfl_sync(0xffffffffu, incl, 31);
                const float pre = warp_scan_add(uu * w, lane);
                const float S = total - (carryS + pre);
                carryS += __shfl_sync(0xffffffffu, pre, 31);
                if (ok) {
                    const float dalpha = uu * T - S / fmaxf(1.f - a, 1e-10f);
                    if (d_vals) {
#pragma unroll
                        for (int k = 0; k < DMAX; ++k)
                            if (k < D) d_vals[(size_t)i * D + k] = w * go[k];
                    }
                    const float dist = dist_[u];
                    const AlphaTerms t = neus_alpha(s_[u], g_[u], d, dist, inv_s, cos_anneal);
                    float dsdf = 0.f, dtc = 0.f;
                    if (t.raw >= 0.f && t.raw <= 1.f) {   // torch.clip passes the gradient inside [0,1]
                        const float den = t.P + 1e-5f;
                        const float dP = dalpha * t.Nx / (den * den);     // d alpha / d prev_cdf
                        const float dN = -dalpha / den;                   // d alpha / d next_cdf
                        const float dzp = dP * t.P * (1.f - t.P);         // wrt est_prev*inv_s
                        const float dzn = dN * t.Nx * (1.f - t.Nx);       // wrt est_next*inv_s
                        dsdf = (dzp + dzn) * inv_s;
                        const float dic = (dzn - dzp) * inv_s * dist * 0.5f;
                        const float dic_dtc = ((-t.tc * 0.5f + 0.5f) > 0.f ? 0.5f * (1.f - cos_anneal) : 0.f) +
                                              ((-t.tc) > 0.f ? cos_anneal : 0.f);
                        dtc = dic * dic_dtc;
                        ds_total += dzp * t.est_prev + dzn * t.est_next;
                    }
                    d_sdf[i] = dsdf;
                    d_grad[(size_t)i * 3 + 0] = dtc * d[0];
                    d_grad[(size_t)i * 3 + 1] = dtc * d[1];
                    d_grad[(size_t)i * 3 + 2] = dtc * d[2];
                }
            }
        }
    }
    if (d_variance) {
        ds_total = warp_sum(ds_total);
        __shared__ float red[8];
        if (lane == 0) red[threadIdx.x >> 5] = ds_total;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
            for (int k = 0; k < wpb; ++k) s += red[k];
            const float raw = expf(__ldg(variance) * 10.f);
            if (raw >= 1e-6f && raw <= 1e6f && s != 0.f) atomicAdd(d_variance, s * 10.f * inv_s);
        }
    }
}

}  // namespace

extern "C" TF_API int tf_neus_composite_fwd(const float* sdf, const float* grad, const float* dists, const float* dirs,
                                     const int32_t* ray_offsets, int32_t n_rays, const float* variance, float cos_anneal,
                                     const float* vals, int32_t D, float* alpha, float* weights, float* acc, float* out,
                                     tf_stream_t stream) {
    if (n_rays == 0) return 0;
    TF_REQUIRE(sdf && grad && dists && dirs && ray_offsets && variance, "an input pointer is NULL");
    TF_REQUIRE(alpha && weights && acc, "an output pointer is NULL");
    TF_REQUIRE(D >= 0 && D <= MAXD, "D must be in [0,%d] (got %d)", MAXD, D);
    TF_REQUIRE(D == 0 || (vals && out), "vals/out is NULL with D > 0");
    const int wpb = 8;
    int grid = (n_rays + wpb - 1) / wpb;
    const int cap = tf_num_sms() * 16;
    if (grid > cap) grid = cap;
    if (D <= 8)
        neus_composite_fwd_kernel<8, 4><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(sdf, grad, dists, dirs, ray_offsets, n_rays, variance,
                                                                                    cos_anneal, vals, D, alpha, weights, acc, out);
    else
        neus_composite_fwd_kernel<MAXD, 2><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(sdf, grad, dists, dirs, ray_offsets, n_rays, variance,
                                                                                       cos_anneal, vals, D, alpha, weights, acc, out);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_neus_composite_fwd");
    return 0;
}

extern "C" TF_API int tf_neus_composite_bwd(const float* sdf, const float* grad, const float* dists, const float* dirs,
                                     const int32_t* ray_offsets, int32_t n_rays, const float* variance, float cos_anneal,
                                     const float* vals, int32_t D, const float* alpha, const float* weights,
                                     const float* acc, const float* out,
                                     const float* g_acc, const float* g_out, const float* g_weights, float* d_sdf,
                                     float* d_grad, float* d_vals, float* d_variance, tf_stream_t stream) {
    if (n_rays == 0) return 0;
    TF_REQUIRE(sdf && grad && dists && dirs && ray_offsets && variance && alpha && weights, "an input pointer is NULL");
    TF_REQUIRE(d_sdf && d_grad, "an output pointer is NULL");
    TF_REQUIRE(D >= 0 && D <= MAXD, "D must be in [0,%d] (got %d)", MAXD, D);
    TF_REQUIRE(D == 0 || vals, "vals is NULL with D > 0");
    TF_REQUIRE(acc && (D == 0 || out), "the forward's acc / out are needed by the backward");
    const int wpb = 8;
    int grid = (n_rays + wpb - 1) / wpb;
    const int cap = tf_num_sms() * 16;
    if (grid > cap) grid = cap;
    if (D <= 8)
        neus_composite_bwd_kernel<8, 4><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(sdf, grad, dists, dirs, ray_offsets, n_rays, variance,
                                                                                    cos_anneal, vals, D, alpha, weights, acc, out, g_acc, g_out,
                                                                                    g_weights, d_sdf, d_grad, d_vals, d_variance);
    else
        neus_composite_bwd_kernel<MAXD, 2><<<grid, wpb * 32, 0, (cudaStream_t)stream>>>(sdf, grad, dists, dirs, ray_offsets, n_rays, variance,
                                                                                       cos_anneal, vals, D, alpha, weights, acc, out, g_acc, g_out,
                                                                                       g_weights, d_sdf, d_grad, d_vals, d_variance);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_neus_composite_bwd");
    return 0;
}
