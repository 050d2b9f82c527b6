// Multi-tensor Adam step for the VM factors, MLP weights and the env map (reference train/trainer_inv.py:112,212:
// torch.optim.Adam(grad_vars, betas=(0.9, 0.99)), one learning rate per parameter group).
// The reference's optimizer.step() runs PyTorch's foreach Adam: ~8 elementwise passes over every tensor list; at
// compressor scale the factors are 113 MB (shape) / 260 MB (material) of fp32, so the update is pure HBM streaming.
// Here ONE launch per <= 32 tensors reads p, g, m, v once and writes p, m, v once: 28 B/element, the minimum.
// Memory order does not matter (elementwise), so channels-last factors are updated in their storage order.
#include "common.cuh"
#include <cmath>

namespace {

constexpr int ADAM_MAX_TENSORS = 32;
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC_PER_THREAD = 4;                                       // float4 per thread per block-chunk
constexpr int ADAM_CHUNK = ADAM_THREADS * ADAM_VEC_PER_THREAD * 4;           // 4096 elements per CTA chunk

struct AdamTable {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    long long numel[ADAM_MAX_TENSORS];
    int chunk_end[ADAM_MAX_TENSORS];        // exclusive prefix of chunk counts: chunks of tensor t are [chunk_end[t-1], chunk_end[t])
    float step_size[ADAM_MAX_TENSORS];      // lr / (1 - beta1^t)
    int n;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float b1c, float b2, float b2c, float inv_bc2_sqrt,
                                          float eps, float step_size) {
    // torch.optim.Adam (single/foreach path): exp_avg.lerp_(g, 1-b1); exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2);
    // denom = sqrt(exp_avg_sq) / sqrt(bias_correction2) + eps; p.addcdiv_(exp_avg, denom, -step_size)
    m = m + b1c * (g - m);
    v = v * b2 + b2c * g * g;
    const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
    p = p - step_size * (m / denom);
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(const __grid_constant__ AdamTable tab, int total_chunks, float b1c, float b2,
                                                            float b2c, float inv_bc2_sqrt, float eps) {
    for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        int t = 0;
#pragma unroll 1
        while (t + 1 < tab.n && chunk >= tab.chunk_end[t]) ++t;
        const long long base = (long long)(chunk - (t ? tab.chunk_end[t - 1] : 0)) * ADAM_CHUNK;
        const long long n = tab.numel[t];
        float* __restrict__ p = tab.p[t];
        const float* __restrict__ g = tab.g[t];
        float* __restrict__ m = tab.m[t];
        float* __restrict__ v = tab.v[t];
        const float ss = tab.step_size[t];
        const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0) && base + ADAM_CHUNK <= n;
        if (vec) {
            float4 P[ADAM_VEC_PER_THREAD], G[ADAM_VEC_PER_THREAD], M[ADAM_VEC_PER_THREAD], V[ADAM_VEC_PER_THREAD];
#pragma unroll
            for (int k = 0; k < ADAM_VEC_PER_THREAD; ++k) {                  // all 16 loads of the thread in flight before any math
                const long long i = base + ((long long)k * ADAM_THREADS + threadIdx.x) * 4;
                G[k] = __ldcs(reinterpret_cast<const float4*>(g + i));      // gradient: read once, evict first
                P[k] = *reinterpret_cast<const float4*>(p + i);
                M[k] = *reinterpret_cast<const float4*>(m + i);
                V[k] = *reinterpret_cast<const float4*>(v + i);
            }
#pragma unroll
            for (int k = 0; k < ADAM_VEC_PER_THREAD; ++k) {
                const long long i = base + ((long long)k * ADAM_THREADS + threadIdx.x) * 4;
                adam_elem(P[k].x, G[k].x, M[k].x, V[k].x, b1c, b2, b2c, inv_bc2_sqrt, eps, ss);
                adam_elem(P[k].y, G[k].y, M[k].y, V[k].y, b1c, b2, b2c, inv_bc2_sqrt, eps, ss);
                adam_elem(P[k].z, G[k].z, M[k].z, V[k].z, b1c, b2, b2c, inv_bc2_sqrt, eps, ss);
                adam_elem(P[k].w, G[k].w, M[k].w, V[k].w, b1c, b2, b2c, inv_bc2_sqrt, eps, ss);
                *reinterpret_cast<float4*>(p + i) = P[k];
                *reinterpret_cast<float4*>(m + i) = M[k];
                *reinterpret_cast<float4*>(v + i) = V[k];
            }
        } else {
            const long long end = base + ADAM_CHUNK < n ? base + ADAM_CHUNK : n;
            for (long long i = base + threadIdx.x; i < end; i += ADAM_THREADS) {
                float pp = p[i], mm = m[i], vv = v[i];
                adam_elem(pp, g[i], mm, vv, b1c, b2, b2c, inv_bc2_sqrt, eps, ss);
                p[i] = pp; m[i] = mm; v[i] = vv;
            }
        }
    }
}

}  // namespace

extern "C" TF_API int tf_adam_step(int32_t n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                                   float* const* exp_avg_sq, const int64_t* numel, const float* lr, float beta1, float beta2, float eps,
                                   int32_t step, tf_stream_t stream) {
    TF_REQUIRE(n_tensors >= 0 && step >= 1, "tf_adam_step: n_tensors >= 0 and step >= 1 required");
    if (n_tensors == 0) return 0;
    TF_REQUIRE(params && grads && exp_avg && exp_avg_sq && numel && lr, "tf_adam_step: NULL table");
    TF_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "tf_adam_step: betas in [0,1), eps >= 0");
    const double bc1 = 1.0 - std::pow((double)beta1, (double)step);
    const double bc2 = 1.0 - std::pow((double)beta2, (double)step);
    const float inv_bc2_sqrt = (float)(1.0 / std::sqrt(bc2));
    int launches = 0;
    for (int t0 = 0; t0 < n_tensors;) {
        AdamTable tab;
        tab.n = 0;
        long long chunks = 0;
        while (t0 < n_tensors && tab.n < ADAM_MAX_TENSORS) {
            const int64_t n = numel[t0];
            TF_REQUIRE(n >= 0, "tf_adam_step: negative numel");
            if (n == 0) { ++t0; continue; }
            TF_REQUIRE(params[t0] && grads[t0] && exp_avg[t0] && exp_avg_sq[t0], "tf_adam_step: NULL tensor pointer");
            const long long c = (n + ADAM_CHUNK - 1) / ADAM_CHUNK;
            if (chunks + c > 0x7fffffffLL) break;
            const int k = tab.n++;
            tab.p[k] = params[t0]; tab.g[k] = grads[t0]; tab.m[k] = exp_avg[t0]; tab.v[k] = exp_avg_sq[t0];
            tab.numel[k] = n;
            chunks += c;
            tab.chunk_end[k] = (int)chunks;
            tab.step_size[k] = (float)((double)lr[t0] / bc1);
            ++t0;
        }
        if (tab.n == 0) { TF_REQUIRE(t0 >= n_tensors, "tf_adam_step: tensor too large"); break; }
        for (int k = tab.n; k < ADAM_MAX_TENSORS; ++k) { tab.p[k] = nullptr; tab.g[k] = nullptr; tab.m[k] = nullptr; tab.v[k] = nullptr;
                                                         tab.numel[k] = 0; tab.chunk_end[k] = (int)chunks; tab.step_size[k] = 0.f; }
        // 28 B/element streaming: 8 resident CTAs per SM x 16 x 16-byte loads per thread keep ~64 KB per SM in flight
        const long long cap = (long long)tf_num_sms() * 8;
        const int grid = (int)(chunks < cap ? chunks : cap);
        adam_kernel<<<grid, ADAM_THREADS, 0, (cudaStream_t)stream>>>(tab, (int)chunks, 1.f - beta1, beta2, 1.f - beta2, inv_bc2_sqrt, eps);
        ++launches;
    }
    tf_count_launches(launches);
    TF_CHECK_LAUNCH("tf_adam_step");
    return 0;
}
