// Total-variation regulariser of the VM planes / lines (reference network/other_field.py:170-191, called from
// network/fields.py:133-138 and :1525-1530 every training step) on channels-last [H,W,C] textures.
// The PyTorch formulation slices the permuted tensor four times per texture and runs strided elementwise + reduce
// kernels over 38 MB planes; here one pass reads each texel once (float4 over channels, coalesced) for the two sums
// of squared differences, and one pass writes the gradient.  HBM-bound: 4 B/element forward, 8 B/element backward.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) tv_fwd_kernel(const float* __restrict__ x, int H, int W, int C4, float* __restrict__ sums) {
    const int64_t total = (int64_t)H * W * C4;
    float sh = 0.f, sw = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hw = i / C4;
        const int w = (int)(hw % W), h = (int)(hw / W);
        const float4 a = ldg4(x + i * 4);
        if (h + 1 < H) {
            const float4 b = ldg4(x + (i + (int64_t)W * C4) * 4);
            const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z, dw = b.w - a.w;
            sh += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
        if (w + 1 < W) {
            const float4 b = ldg4(x + (i + C4) * 4);
            const float dx = b.x - a.x, dy = b.y - a.y, dz = b.z - a.z, dw = b.w - a.w;
            sw += (dx * dx + dy * dy) + (dz * dz + dw * dw);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sh += __shfl_xor_sync(0xffffffffu, sh, o); sw += __shfl_xor_sync(0xffffffffu, sw, o); }
    __shared__ float red[2][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = sh; red[1][warp] = sw; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) t += red[threadIdx.x][k];
        if (t != 0.f) atomicAdd(sums + threadIdx.x, t);
    }
}

// g += sh * d/dx sum_h (x[h+1]-x[h])^2 + sw * d/dx sum_w (x[w+1]-x[w])^2
__global__ void __launch_bounds__(256) tv_bwd_kernel(const float* __restrict__ x, int H, int W, int C4, float sh, float sw,
                                                     const float* __restrict__ upstream, float* __restrict__ g) {
    const int64_t total = (int64_t)H * W * C4;
    if (upstream) { const float u = __ldg(upstream); sh *= u; sw *= u; }      // the loss gradient stays on the device
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hw = i / C4;
        const int w = (int)(hw % W), h = (int)(hw / W);
        const float4 a = ldg4(x + i * 4);
        float4 acc = f4_zero();
        const int64_t rs = (int64_t)W * C4;
        if (h > 0) { const float4 b = ldg4(x + (i - rs) * 4); acc = f4_fma(2.f * sh, make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w), acc); }
        if (h + 1 < H) { const float4 b = ldg4(x + (i + rs) * 4); acc = f4_fma(2.f * sh, make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w), acc); }
        if (w > 0) { const float4 b = ldg4(x + (i - C4) * 4); acc = f4_fma(2.f * sw, make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w), acc); }
        if (w + 1 < W) { const float4 b = ldg4(x + (i + C4) * 4); acc = f4_fma(2.f * sw, make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w), acc); }
        float4* dst = reinterpret_cast<float4*>(g + i * 4);
        const float4 o = *dst;
        *dst = make_float4(o.x + acc.x, o.y + acc.y, o.z + acc.z, o.w + acc.w);
    }
}

int tv_grid(int64_t total) {
    int64_t b = (total + 255) / 256;
    const int64_t cap = (int64_t)tf_num_sms() * 8;
    return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" TF_API int tf_tv_fwd(const float* x, int32_t H, int32_t W, int32_t C, float* sums, tf_stream_t stream) {
    TF_REQUIRE(x && sums, "tf_tv_fwd: NULL pointer");
    TF_REQUIRE(H > 0 && W > 0 && C > 0 && C % 4 == 0 && ((uintptr_t)x & 15) == 0, "tf_tv_fwd: C must be a multiple of 4, x 16-byte aligned");
    tv_fwd_kernel<<<tv_grid((int64_t)H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(x, H, W, C / 4, sums);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_tv_fwd");
    return 0;
}

extern "C" TF_API int tf_tv_bwd(const float* x, int32_t H, int32_t W, int32_t C, float scale_h, float scale_w, const float* upstream, float* g,
                                tf_stream_t stream) {
    TF_REQUIRE(x && g, "tf_tv_bwd: NULL pointer");
    TF_REQUIRE(H > 0 && W > 0 && C > 0 && C % 4 == 0 && (((uintptr_t)x | (uintptr_t)g) & 15) == 0, "tf_tv_bwd: C must be a multiple of 4, buffers 16-byte aligned");
    tv_bwd_kernel<<<tv_grid((int64_t)H * W * (C / 4)), 256, 0, (cudaStream_t)stream>>>(x, H, W, C / 4, scale_h, scale_w, upstream, g);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_tv_bwd");
    return 0;
}
