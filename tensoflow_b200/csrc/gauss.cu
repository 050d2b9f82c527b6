// Gaussian-smoothness regulariser of the VM planes / lines (reference network/fields.py:301-309 and :1537-1545:
// sum over the interior texels of (x - GaussianBlur(x))^2, GaussianBlur2D / GaussianBlur1D = F.conv2d / F.conv1d with a
// normalised KS x KS / KS kernel, stride 1, zero padding; network/other_field.py:121-168) on channels-last [H,W,C]
// textures.  The reference permutes the factor to [C,1,H,W] and runs a cuDNN convolution per texture; here one pass
// writes the interior residual r = x - k (*) x (zero outside the interior) and its squared sum, and one pass turns the
// residual into the gradient 2 u (r - k^T (*) r).  A line [G,C] is the same kernel with W = 1, KW = 1.
// HBM-bound: 4 B read + 4 B written per element forward (the KH x KW neighbourhood is served by L1/L2), 8 B backward.
#include "common.cuh"

namespace {

constexpr int MAX_TAPS = 81;   // up to 9 x 9

struct Taps { float w[MAX_TAPS]; };

__global__ void __launch_bounds__(256) gauss_residual_fwd_kernel(const float* __restrict__ x, int H, int W, int C4, Taps taps, int KH, int KW,
                                                                 float* __restrict__ r, float* __restrict__ sum) {
    const int kh = KH / 2, kw = KW / 2;
    const int64_t total = (int64_t)H * W * C4;
    float acc = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hw = i / C4;
        const int w = (int)(hw % W), h = (int)(hw / W);
        float4 res = f4_zero();
        if (h >= kh && h < H - kh && w >= kw && w < W - kw) {
            float4 y = f4_zero();
            for (int a = 0; a < KH; ++a)
                for (int b = 0; b < KW; ++b)
                    y = f4_fma(taps.w[a * KW + b], ldg4(x + (i + ((int64_t)(a - kh) * W + (b - kw)) * C4) * 4), y);
            const float4 c = ldg4(x + i * 4);
            res = make_float4(c.x - y.x, c.y - y.y, c.z - y.z, c.w - y.w);
            acc += (res.x * res.x + res.y * res.y) + (res.z * res.z + res.w * res.w);
        }
        *reinterpret_cast<float4*>(r + i * 4) = res;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float red[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[k];
        if (t != 0.f) atomicAdd(sum, t);
    }
}

// g[p] += 2 u (r[p] - sum_{a,b} k[a,b] r[p - (a-kh, b-kw)])   (r is zero outside the interior and outside the texture)
__global__ void __launch_bounds__(256) gauss_residual_bwd_kernel(const float* __restrict__ r, int H, int W, int C4, Taps taps, int KH, int KW,
                                                                 const float* __restrict__ upstream, float* __restrict__ g) {
    const int kh = KH / 2, kw = KW / 2;
    const int64_t total = (int64_t)H * W * C4;
    const float u2 = 2.f * (upstream ? __ldg(upstream) : 1.f);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hw = i / C4;
        const int w = (int)(hw % W), h = (int)(hw / W);
        float4 y = f4_zero();
        for (int a = 0; a < KH; ++a) {
            const int hh = h - (a - kh);
            if (hh < 0 || hh >= H) continue;
            for (int b = 0; b < KW; ++b) {
                const int ww = w - (b - kw);
                if (ww < 0 || ww >= W) continue;
                y = f4_fma(taps.w[a * KW + b], ldg4(r + (((int64_t)hh * W + ww) * C4 + (i % C4)) * 4), y);
            }
        }
        const float4 c = ldg4(r + i * 4);
        float4* dst = reinterpret_cast<float4*>(g + i * 4);
        const float4 o = *dst;
        *dst = make_float4(fmaf(u2, c.x - y.x, o.x), fmaf(u2, c.y - y.y, o.y), fmaf(u2, c.z - y.z, o.z), fmaf(u2, c.w - y.w, o.w));
    }
}

int gauss_grid(int64_t total) {
    int64_t b = (total + 255) / 256;
    const int64_t cap = (int64_t)tf_num_sms() * 8;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

int load_taps(const float* taps, int KH, int KW, Taps& t) {
    TF_REQUIRE(taps && KH >= 1 && KW >= 1 && (KH & 1) && (KW & 1) && KH * KW <= MAX_TAPS, "gaussian taps: odd KH, KW with KH*KW <= %d required", MAX_TAPS);
    for (int i = 0; i < KH * KW; ++i) t.w[i] = taps[i];
    return 0;
}

}  // namespace

extern "C" TF_API int tf_gauss_residual_fwd(const float* x, int32_t H, int32_t W, int32_t C, const float* taps_host, int32_t KH, int32_t KW,
                                            float* r, float* sum, tf_stream_t stream) {
    TF_REQUIRE(x && r && sum, "tf_gauss_residual_fwd: NULL pointer");
    TF_REQUIRE(H > 0 && W > 0 && C > 0 && C % 4 == 0 && (((uintptr_t)x | (uintptr_t)r) & 15) == 0,
               "tf_gauss_residual_fwd: C must be a multiple of 4, buffers 16-byte aligned");
    Taps t;
    if (int e = load_taps(taps_host, KH, KW, t)) return e;
    gauss_residual_fwd_kernel<<<gauss_grid((int64_t)H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(x, H, W, C / 4, t, KH, KW, r, sum);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_gauss_residual_fwd");
    return 0;
}

extern "C" TF_API int tf_gauss_residual_bwd(const float* r, int32_t H, int32_t W, int32_t C, const float* taps_host, int32_t KH, int32_t KW,
                                            const float* upstream, float* g, tf_stream_t stream) {
    TF_REQUIRE(r && g, "tf_gauss_residual_bwd: NULL pointer");
    TF_REQUIRE(H > 0 && W > 0 && C > 0 && C % 4 == 0 && (((uintptr_t)r | (uintptr_t)g) & 15) == 0,
               "tf_gauss_residual_bwd: C must be a multiple of 4, buffers 16-byte aligned");
    Taps t;
    if (int e = load_taps(taps_host, KH, KW, t)) return e;
    gauss_residual_bwd_kernel<<<gauss_grid((int64_t)H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(r, H, W, C / 4, t, KH, KW, upstream, g);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_gauss_residual_bwd");
    return 0;
}
