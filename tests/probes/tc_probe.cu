// tcgen05 bring-up / self-test / issue-rate probe (TEST INFRASTRUCTURE: built into tests/probes/libtf_probe.so, never into
// the product library).  One CTA computes D[128,N] = A[128,K] * B[N,K]^T on the 5th-gen tensor cores with the operands
// staged in the layouts the fused kernels use or consider:
//   a_mode / b_mode: 0 K-major no-swizzle (K-chunk stride = lbo bytes), 1 MN-major no-swizzle (the transposed view of a
//                    K-major tile: 16-byte units of 4 MN elements, 8 K rows per core matrix), 2 K-major SWIZZLE_128B,
//                    3 (A only) tensor memory (TS-mode MMA)
//   passes = 1 plain tf32, 3 = hi*hi + hi*lo + lo*hi (fp32-level accuracy)
// and reports the clock64 cycles between the first issue and the completion of reps * passes * K/8 MMAs.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../tensoflow_b200/csrc/tc_common.cuh"

namespace {

constexpr int PM = 128;

struct ProbeParams {
    const float* A; const float* B; float* D; long long* cycles;
    int N, K, passes, reps;
    int a_mode, b_mode;
    int a_lbo;          // K-major no-swizzle: K-chunk stride of A in bytes (128 dense, 144 padded)
    int mn_sbo;         // MN-major: stride between 4-element MN groups in bytes (128 dense, 144 = padded A tile)
    int swap_mn;        // MN-major: exchange the LBO / SBO descriptor fields (layout-semantics probe)
};

__device__ __forceinline__ uint32_t off_kmajor(int r, int k, int kch, int lbo) { return (uint32_t)((r >> 3) * (kch * lbo) + (k >> 2) * lbo + (r & 7) * 16 + (k & 3) * 4); }
// MN-major: MN group (4 elements) stride mn_sbo, K group (8 rows) stride = groups * mn_sbo
__device__ __forceinline__ uint32_t off_mnmajor(int r, int k, int rows, int mn_sbo) { return (uint32_t)((r >> 2) * mn_sbo + (k >> 3) * ((rows / 4) * mn_sbo) + (k & 7) * 16 + (r & 3) * 4); }
// K-major SWIZZLE_128B: 8-row x 128-byte atoms (1024 B), 16-byte chunk index XOR row-in-atom; K blocks of 32 elements
__device__ __forceinline__ uint32_t off_sw128(int r, int k, int rows) {
    const int kb = k >> 5, kk = k & 31;
    return (uint32_t)(kb * (rows / 8) * 1024 + (r >> 3) * 1024 + (r & 7) * 128 + (((kk >> 2) ^ (r & 7)) * 16) + (kk & 3) * 4);
}

__device__ __forceinline__ uint32_t operand_off(int mode, int r, int k, int rows, int K, int lbo, int mn_sbo) {
    if (mode == 1) return off_mnmajor(r, k, rows, mn_sbo);
    if (mode == 2) return off_sw128(r, k, rows);
    return off_kmajor(r, k, K / 4, lbo);
}
__device__ __forceinline__ uint32_t operand_bytes(int mode, int rows, int K, int lbo, int mn_sbo) {
    if (mode == 1) return (uint32_t)(K / 8) * (rows / 4) * mn_sbo;
    if (mode == 2) return (uint32_t)((K + 31) / 32) * (rows / 8) * 1024;
    return (uint32_t)(rows / 8) * (K / 4) * lbo;
}
__device__ __forceinline__ uint64_t operand_desc(int mode, uint32_t addr, int rows, int K, int lbo, int mn_sbo, int swap_mn) {
    if (mode == 1) {
        const uint32_t kgrp = (uint32_t)(rows / 4) * mn_sbo;             // stride between 8-row K groups
        return swap_mn ? tc::make_smem_desc(addr, (uint32_t)mn_sbo, kgrp) : tc::make_smem_desc(addr, kgrp, (uint32_t)mn_sbo);
    }
    if (mode == 2) return tc::make_smem_desc(addr, 16, 1024) | ((uint64_t)2 << 61);
    return tc::make_smem_desc(addr, (uint32_t)lbo, (uint32_t)(K / 4) * lbo);
}
// byte advance of the start address for k-step ks (8 tf32 of K)
__device__ __forceinline__ uint32_t operand_kstep(int mode, int ks, int rows, int lbo, int mn_sbo) {
    if (mode == 1) return (uint32_t)ks * (rows / 4) * mn_sbo;
    if (mode == 2) return (uint32_t)(ks >> 2) * (rows / 8) * 1024 + (uint32_t)(ks & 3) * 32;
    return (uint32_t)ks * 2 * lbo;
}

__global__ void __launch_bounds__(128, 1) tc_probe_kernel(ProbeParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int N = p.N, K = p.K;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int a_smem_mode = p.a_mode == 3 ? 0 : p.a_mode;
    const uint32_t a_bytes = (operand_bytes(a_smem_mode, PM, K, p.a_lbo, p.mn_sbo) + 1023u) & ~1023u;
    const uint32_t b_bytes = (operand_bytes(p.b_mode, N, K, 128, p.mn_sbo) + 1023u) & ~1023u;
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + a_bytes;
    uint8_t* b_hi = a_lo + a_bytes;
    uint8_t* b_lo = b_hi + b_bytes;

    if (warp == 0) tc::tmem_alloc<512>(&tmem_base_s);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
    for (int i = tid; i < (int)(2 * a_bytes + 2 * b_bytes) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < PM * K; i += 128) {
        const int r = i / K, k = i % K;
        const float x = p.A[i];
        const float h = tc::tf32_rn(x);
        const uint32_t off = operand_off(a_smem_mode, r, k, PM, K, p.a_lbo, p.mn_sbo);
        *reinterpret_cast<float*>(a_hi + off) = h;
        *reinterpret_cast<float*>(a_lo + off) = tc::tf32_rn(x - h);
    }
    for (int i = tid; i < N * K; i += 128) {
        const int r = i / K, k = i % K;
        const float x = p.B[i];
        const float h = tc::tf32_rn(x);
        const uint32_t off = operand_off(p.b_mode, r, k, N, K, 128, p.mn_sbo);
        *reinterpret_cast<float*>(b_hi + off) = h;
        *reinterpret_cast<float*>(b_lo + off) = tc::tf32_rn(x - h);
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_d = tmem_base_s;               // accumulator: columns [0, N)
    const uint32_t tmem_a = tmem_base_s + 256;         // TS mode: A hi at +0, lo at +K (K <= 128)
    if (p.a_mode == 3) {
        const int row = tid;
        for (int k0 = 0; k0 < K; k0 += 16) {
            float h[16], l[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { const float x = p.A[row * K + k0 + j]; h[j] = tc::tf32_rn(x); l[j] = tc::tf32_rn(x - h[j]); }
            tc::tmem_st16(tmem_a + ((uint32_t)(warp * 32) << 16) + k0, h);
            tc::tmem_st16(tmem_a + ((uint32_t)(warp * 32) << 16) + K + k0, l);
        }
        tc::tmem_st_wait();
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
    }

    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc(2, 2, PM, N) | (p.a_mode == 1 ? (1u << 15) : 0u) | (p.b_mode == 1 ? (1u << 16) : 0u);
        const uint64_t adh = operand_desc(a_smem_mode, tc::smem_u32(a_hi), PM, K, p.a_lbo, p.mn_sbo, p.swap_mn);
        const uint64_t adl = operand_desc(a_smem_mode, tc::smem_u32(a_lo), PM, K, p.a_lbo, p.mn_sbo, p.swap_mn);
        const uint64_t bdh = operand_desc(p.b_mode, tc::smem_u32(b_hi), N, K, 128, p.mn_sbo, p.swap_mn);
        const uint64_t bdl = operand_desc(p.b_mode, tc::smem_u32(b_lo), N, K, 128, p.mn_sbo, p.swap_mn);
        // the timed loop is MMA issue only: fully unrolled groups of 8 k-steps with the descriptor offsets in registers
        const uint32_t a_step = operand_kstep(a_smem_mode, 1, PM, p.a_lbo, p.mn_sbo), b_step = operand_kstep(p.b_mode, 1, N, 128, p.mn_sbo);
        const bool linear_steps = a_smem_mode != 2 && p.b_mode != 2;     // SWIZZLE_128B advances 32 B inside a 128-byte span
        const int nks = K / 8;
        const long long t0 = clock64();
        for (int rep = 0; rep < p.reps; ++rep) {
            uint32_t acc = 0;
            for (int ps = 0; ps < p.passes; ++ps) {                       // pass 0: hi*hi, 1: hi*lo, 2: lo*hi
                const uint64_t a0 = ps == 2 ? adl : adh, b0 = ps == 1 ? bdl : bdh;
                const uint32_t ta = tmem_a + (ps == 2 ? K : 0);
                if (linear_steps) {
                    uint64_t ad = a0, bd = b0;
                    uint32_t tk = ta;
#pragma unroll 8
                    for (int ks = 0; ks < nks; ++ks) {
                        if (p.a_mode == 3) tc::mma_tf32_ts(tmem_d, tk, bd, idesc, acc);
                        else tc::mma_tf32_ss(tmem_d, ad, bd, idesc, acc);
                        acc = 1;
                        ad = tc::desc_add(ad, a_step); bd = tc::desc_add(bd, b_step); tk += 8;
                    }
                } else {
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint64_t bd = tc::desc_add(b0, operand_kstep(p.b_mode, ks, N, 128, p.mn_sbo));
                        if (p.a_mode == 3) tc::mma_tf32_ts(tmem_d, ta + ks * 8, bd, idesc, acc);
                        else tc::mma_tf32_ss(tmem_d, tc::desc_add(a0, operand_kstep(a_smem_mode, ks, PM, p.a_lbo, p.mn_sbo)), bd, idesc, acc);
                        acc = 1;
                    }
                }
            }
        }
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (p.cycles) *p.cycles = t1 - t0;
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after_sync();

    const int row = warp * 32 + (tid & 31);
    for (int c0 = 0; c0 < N; c0 += 8) {
        float v[8];
        tc::tmem_ld8(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) p.D[(size_t)row * N + c0 + j] = v[j];
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem_d);
}

}  // namespace

// Standalone driver: tc_probe N K passes reps a_mode b_mode a_lbo mn_sbo swap_mn  -> one JSON line
// (max relative error of D against an fp64 host product, cycles per MMA).  A faulting configuration only kills this process.
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

static float lcg(uint32_t& s) { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xFFFF) / 65536.f * 2.f - 1.f; }

int main(int argc, char** argv) {
    if (argc < 10) { fprintf(stderr, "usage: tc_probe N K passes reps a_mode b_mode a_lbo mn_sbo swap_mn\n"); return 2; }
    const int N = atoi(argv[1]), K = atoi(argv[2]), passes = atoi(argv[3]), reps = atoi(argv[4]), a_mode = atoi(argv[5]), b_mode = atoi(argv[6]),
              a_lbo = atoi(argv[7]), mn_sbo = atoi(argv[8]), swap_mn = atoi(argv[9]);
    if (N < 8 || N > 256 || N % 8 || K < 8 || K % 8 || K > 128 || (passes != 1 && passes != 3) || (a_mode == 3 && (K % 16 || K > 128)) || a_lbo < 128 || a_lbo % 16 ||
        mn_sbo < 128 || mn_sbo % 16) { fprintf(stderr, "bad arguments\n"); return 2; }
    std::vector<float> A((size_t)PM * K), B((size_t)N * K), D((size_t)PM * N);
    uint32_t seed = 12345u + N * 7 + K;
    for (auto& v : A) v = lcg(seed);
    for (auto& v : B) v = lcg(seed);
    float *dA, *dB, *dD; long long* dC;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xFF, D.size() * 4);
    ProbeParams p = {dA, dB, dD, dC, N, K, passes, reps < 1 ? 1 : reps, a_mode, b_mode, a_lbo, mn_sbo, swap_mn};
    const size_t smem = (size_t)(PM + N) * K * 4 * 2 * 9 / 8 + 8192;    // hi + lo of both operands, 144-byte padded chunks included
    if (smem > 220 * 1024) { fprintf(stderr, "tile does not fit shared memory\n"); return 2; }
    cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long cyc = 0;
    for (int it = 0; it < 2; ++it) {                                  // second run is the timed one (warm instruction cache)
        tc_probe_kernel<<<1, 128, smem>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double max_err = 0, max_ref = 0;
    for (int r = 0; r < PM; ++r)
        for (int n = 0; n < N; ++n) {
            double ref = 0;
            for (int k = 0; k < K; ++k) ref += (double)A[(size_t)r * K + k] * (double)B[(size_t)n * K + k];
            const double d = fabs((double)D[(size_t)r * N + n] - ref);
            if (!(d <= max_err)) max_err = d;                           // NaN-propagating max
            if (fabs(ref) > max_ref) max_ref = fabs(ref);
        }
    const long long n_mma = (long long)p.reps * passes * (K / 8);
    printf("{\"N\": %d, \"K\": %d, \"passes\": %d, \"reps\": %d, \"a_mode\": %d, \"b_mode\": %d, \"a_lbo\": %d, \"mn_sbo\": %d, \"swap_mn\": %d, "
           "\"rel_err\": %.3e, \"cycles_per_mma\": %.1f}\n",
           N, K, passes, p.reps, a_mode, b_mode, a_lbo, mn_sbo, swap_mn, max_err / max_ref, (double)cyc / (double)n_mma);
    return 0;
}
