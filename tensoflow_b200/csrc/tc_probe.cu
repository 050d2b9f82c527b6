// tcgen05 bring-up / self-test kernel: D[128,N] = A[128,K] * B[N,K]^T on the 5th-gen tensor cores
// with fp32-level accuracy from tf32 operand splitting (passes = 3: hi*hi + hi*lo + lo*hi;
// passes = 1: plain tf32).  It validates the descriptor / TMEM / mbarrier plumbing of
// tc_common.cuh that the fused decoder kernels build on, and doubles as a throughput probe
// (repeat > 1 re-issues the MMA sequence).
#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int PM = 128;

template <int N>
__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const float* __restrict__ A, const float* __restrict__ B, int K, int passes,
                                                          int repeat, float* __restrict__ D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int kch = K / 4;                       // 16-byte chunks along K
    const uint32_t a_bytes = PM * K * 4, b_bytes = N * K * 4;
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + a_bytes;
    uint8_t* b_hi = a_lo + a_bytes;
    uint8_t* b_lo = b_hi + b_bytes;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) tc::tmem_alloc<(N < 32 ? 32 : N)>(&tmem_base_s);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }

    for (int i = tid; i < PM * K; i += 128) {
        const int r = i / K, k = i % K;
        const float x = A[i];
        const float h = tc::tf32_rn(x);
        const uint32_t off = tc::tile_off_b32(r, k, kch);
        *reinterpret_cast<float*>(a_hi + off) = h;
        *reinterpret_cast<float*>(a_lo + off) = tc::tf32_rn(x - h);
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        const float x = B[i];
        const float h = tc::tf32_rn(x);
        const uint32_t off = tc::tile_off_b32(r, k, kch);
        *reinterpret_cast<float*>(b_hi + off) = h;
        *reinterpret_cast<float*>(b_lo + off) = tc::tf32_rn(x - h);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc(2, 2, PM, N);
        const uint32_t sbo = (uint32_t)kch * 128, lbo = 128;
        const uint64_t adh = tc::make_smem_desc(tc::smem_u32(a_hi), lbo, sbo), adl = tc::make_smem_desc(tc::smem_u32(a_lo), lbo, sbo);
        const uint64_t bdh = tc::make_smem_desc(tc::smem_u32(b_hi), lbo, sbo), bdl = tc::make_smem_desc(tc::smem_u32(b_lo), lbo, sbo);
        for (int rep = 0; rep < repeat; ++rep) {
            uint32_t acc = 0;
            for (int p = 0; p < passes; ++p) {
                const uint64_t ad0 = (p == 2) ? adl : adh;        // pass 0: hi*hi, 1: hi*lo, 2: lo*hi
                const uint64_t bd0 = (p == 1) ? bdl : bdh;
#pragma unroll 4
                for (int ks = 0; ks < K / 8; ++ks) {
                    tc::mma_tf32_ss(tmem_d, tc::desc_add(ad0, ks * 256), tc::desc_add(bd0, ks * 256), idesc, acc);
                    acc = 1;
                }
            }
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();

    // epilogue: warp q owns TMEM lanes 32q..32q+31 (= output rows), thread t -> row 32q+t
    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<(N < 32 ? 32 : N)>(tmem_d);
}

}  // namespace

extern "C" TF_API int tf_tc_probe(const float* A, const float* B, int32_t N, int32_t K, int32_t passes, int32_t repeat, float* D,
                                  tf_stream_t stream) {
    TF_REQUIRE(A && B && D, "tf_tc_probe: NULL pointer");
    TF_REQUIRE((N == 128 || N == 256) && K % 8 == 0 && K >= 8, "tf_tc_probe: N must be 128 or 256, K a multiple of 8");
    TF_REQUIRE(passes == 1 || passes == 3, "tf_tc_probe: passes must be 1 or 3");
    const size_t smem = (size_t)2 * (PM + N) * K * 4;
    TF_REQUIRE(smem <= 220 * 1024, "tf_tc_probe: tile does not fit shared memory (K too large)");
    if (N == 128) {
        cudaFuncSetAttribute(tc_probe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_probe_kernel<128><<<1, 128, smem, (cudaStream_t)stream>>>(A, B, K, passes, repeat < 1 ? 1 : repeat, D);
    } else {
        cudaFuncSetAttribute(tc_probe_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_probe_kernel<256><<<1, 128, smem, (cudaStream_t)stream>>>(A, B, K, passes, repeat < 1 ? 1 : repeat, D);
    }
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_tc_probe");
    return 0;
}
