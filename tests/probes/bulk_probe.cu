// Test infrastructure (not linked into the product library): how fast can every SM stream the SAME pre-tiled weight
// slices L2 -> shared memory through a cp.async.bulk ring?  One driver thread per CTA, 148 CTAs, each consumes
// `rounds` passes over `S` slots of `slot_bytes` through `nst` stages (a slot is released as soon as it has landed).
//   bulk_probe <slot_bytes> <nst> <S> <rounds> <mode> <wait>    mode 0: all CTAs in phase, 1: CTA b starts at slot b % S,
//                                                         2: every CTA streams a private copy of the weights
// Prints JSON: cycles per slot (median over CTAs), bytes/clk/SM.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../tensoflow_b200/csrc/tc_common.cuh"

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void wait_spin(uint64_t* bar, uint32_t parity) { while (!tc::mbar_test(bar, parity)) {} }
__device__ __forceinline__ void wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP_H:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_H;\n\t"
        "bra WAIT_LOOP_H;\n\t"
        "DONE_H:\n\t}" ::"r"(tc::smem_u32(bar)), "r"(parity), "r"(ns) : "memory");
}
__global__ void __launch_bounds__(512, 1) stream_kernel(const uint8_t* W, uint32_t slot_bytes, int nst, int S, int rounds, int mode, int wmode, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem_all[];
    // every issuing warp (lane 0) runs its own ring in its own slice of shared memory
    const int nwarps = blockDim.x / 32, warp = threadIdx.x / 32;
    uint8_t* smem = smem_all + (size_t)warp * nst * slot_bytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_all + (size_t)nwarps * nst * slot_bytes) + warp * nst;
    if ((threadIdx.x & 31) == 0) {
        for (int i = 0; i < nst; ++i) tc::mbar_init(&full[i], 1);
        tc::mbar_fence_init();
    }
    __syncthreads();
    if ((threadIdx.x & 31) != 0) return;
    const uint8_t* base = W + (mode == 2 ? (size_t)blockIdx.x * S * slot_bytes : 0);
    const int first = mode == 1 ? blockIdx.x % S : 0;
    const long long total = (long long)S * rounds;
    long long issued = 0;
    const long long t0 = clock64();
    for (long long i = 0; i < total; ++i) {
        while (issued < total && issued < i + nst) {
            const int st = (int)(issued % nst);
            expect_tx(&full[st], slot_bytes);
            bulk_g2s(smem + (size_t)st * slot_bytes, base + (size_t)((first + issued) % S) * slot_bytes, slot_bytes, &full[st]);
            ++issued;
        }
        const int st = (int)(i % nst);
        const uint32_t par = (uint32_t)((i / nst) & 1);
        if (wmode == 0) tc::mbar_wait(&full[st], par);
        else if (wmode == 1) wait_spin(&full[st], par);
        else wait_hint(&full[st], par, 20);
    }
    if (warp == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main(int argc, char** argv) {
    const uint32_t slot_bytes = argc > 1 ? atoi(argv[1]) : 32768;
    const int nst = argc > 2 ? atoi(argv[2]) : 2, S = argc > 3 ? atoi(argv[3]) : 7, rounds = argc > 4 ? atoi(argv[4]) : 200, mode = argc > 5 ? atoi(argv[5]) : 0, wmode = argc > 6 ? atoi(argv[6]) : 0, nwarps = argc > 7 ? atoi(argv[7]) : 1;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t wbytes = (size_t)S * slot_bytes * (mode == 2 ? sms : 1);
    uint8_t* W; long long* cyc;
    cudaMalloc(&W, wbytes); cudaMemset(W, 1, wbytes);
    cudaMalloc(&cyc, sms * sizeof(long long));
    const size_t smem = (size_t)nwarps * nst * slot_bytes + 8 * nwarps * nst + 64;
    cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 2; ++rep) stream_kernel<<<sms, 32 * nwarps, smem>>>(W, slot_bytes, nst, S, rounds, mode, wmode, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double per_slot = (double)h[sms / 2] / ((double)S * rounds);
    printf("{\"slot_bytes\": %u, \"nst\": %d, \"S\": %d, \"mode\": %d, \"wait\": %d, \"issuing_warps\": %d, \"cycles_per_slot_median\": %.1f, \"cycles_per_slot_max\": %.1f, \"bytes_per_clk_per_sm\": %.2f}\n",
           slot_bytes, nst, S, mode, wmode, nwarps, per_slot, (double)h[sms - 1] / ((double)S * rounds), (double)nwarps * slot_bytes / per_slot);
    return 0;
}
