// Cubemap prefilter as a cached sparse operator.
// The reference's DiffuseCubemap / SpecularCubemap kernels (network/renderutils/c_src/cubemap.cu:
// 110-350) convolve the cubemap with weights that depend only on (resolution, roughness, cutoff):
//   diffuse : w = clamp(N.L,0,0.999) * texel_area / 3.141592
//   specular: w = max(N.L,0) * D_ggx(alpha^2, N.H) * texel_area / 4 inside the cone N.L >= cos_cutoff,
//             result divided by sum(w)
// so out = W x is linear in the cubemap.  W is built once per (res, roughness, cutoff) (the
// reference caches its per-texel bounds the same way, ops.py:427-444) and applied here as a CSR
// product: fwd y = W x (one warp per output texel), bwd dx += W^T dy (atomic scatter).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) csr_spmm3_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                            const float* __restrict__ val, const float* __restrict__ x, int n_rows,
                                                            float* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    float a = 0.f, b = 0.f, c = 0.f;
    for (int k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) {
        const float w = __ldg(val + k);
        const float* px = x + (size_t)__ldg(col + k) * 3;
        a = fmaf(w, __ldg(px), a); b = fmaf(w, __ldg(px + 1), b); c = fmaf(w, __ldg(px + 2), c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { y[(size_t)row * 3] = a; y[(size_t)row * 3 + 1] = b; y[(size_t)row * 3 + 2] = c; }
}

__global__ void __launch_bounds__(256) csr_spmm3_bwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                            const float* __restrict__ val, const float* __restrict__ gy, int n_rows,
                                                            float* __restrict__ gx) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const float a = gy[(size_t)row * 3], b = gy[(size_t)row * 3 + 1], c = gy[(size_t)row * 3 + 2];
    if (a == 0.f && b == 0.f && c == 0.f) return;
    for (int k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) {
        const float w = __ldg(val + k);
        float* px = gx + (size_t)__ldg(col + k) * 3;
        atomicAdd(px, w * a); atomicAdd(px + 1, w * b); atomicAdd(px + 2, w * c);
    }
}

}  // namespace

extern "C" TF_API int tf_csr_spmm3_fwd(const int32_t* rowptr, const int32_t* col, const float* val, const float* x, int32_t n_rows,
                                       float* y, tf_stream_t stream) {
    if (n_rows == 0) return 0;
    TF_REQUIRE(rowptr && col && val && x && y, "tf_csr_spmm3_fwd: NULL pointer");
    const int64_t threads = (int64_t)n_rows * 32;
    csr_spmm3_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rowptr, col, val, x, n_rows, y);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_csr_spmm3_fwd");
    return 0;
}

extern "C" TF_API int tf_csr_spmm3_bwd(const int32_t* rowptr, const int32_t* col, const float* val, const float* gy, int32_t n_rows,
                                       float* gx, tf_stream_t stream) {
    if (n_rows == 0) return 0;
    TF_REQUIRE(rowptr && col && val && gy && gx, "tf_csr_spmm3_bwd: NULL pointer");
    const int64_t threads = (int64_t)n_rows * 32;
    csr_spmm3_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rowptr, col, val, gy, n_rows, gx);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_csr_spmm3_bwd");
    return 0;
}
