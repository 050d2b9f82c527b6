// Fused coupling block of the TensoFlow sampler on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as flow_block_fwd_kernel in flow.cu (reference network/flow.py:549-641), but the conditioner chain
// [PE(y_c) | feature] -> 64 -> 64 -> 64 -> 21 of a tile of 128 (point, direction) pairs runs as four small GEMMs whose A operand
// never leaves the SM: thread t owns pair t = TMEM lane t; after each layer the accumulator D [128 x 64] is read back
// (tcgen05.ld), bias + LeakyReLU applied, split into tf32 hi | lo and written to TENSOR MEMORY (tcgen05.st) as the A
// operand of the next layer's TS-mode MMAs (3xTF32: hi.hi + hi.lo + lo.hi, fp32-level accuracy).  The weights sit in shared
// memory pre-split in the K-major no-swizzle UMMA layout for the whole kernel (84 KB), the per-point feature part of the first
// layer is added as a per-row bias, the spline runs on the 21 outputs of the last layer in registers.
// Two CTAs share an SM (256 TMEM columns each) so that one CTA's epilogue overlaps the other's MMAs.
// The FFMA kernel this replaces spent its time on broadcast LDS.128 of the weights (one per 4 FMAs, ~21 % of the FP32 peak).
#include "common.cuh"
#include "tc_common.cuh"
#include "flow_spline.cuh"

namespace {

using namespace flowsp;

constexpr int FH = 64, FPE = 7, FSTP = 24, FTILE = 128, FMAXP = 10;
constexpr int K1 = 8;        // first-layer K: 7 positional-encoding inputs + one zero column
constexpr int N4 = 32;       // last-layer N: 21 spline parameters padded to 32

struct FlowTcParams {
    const float *W1, *b1, *W2, *b2, *W3, *b3, *W4, *b4;
    int F; float scale, offset;
    const float* y_in; const float* logj_in; const float* feat;
    int sn, cond, inverse; int64_t M;
    float* y_out; float* logj_out;
    float* save_h;           // [M][3][64] post-activation h1, h2, h3 (NULL: not kept)
    float* save_st;          // [M][24] spline parameters (NULL: not kept)
};

// W[n][koff + k] (n < n_valid, k < k_valid, row stride ldw) -> K-major no-swizzle B tile [N][K], split hi | lo
__device__ __forceinline__ void fill_b_tile(float* hi, float* lo, const float* __restrict__ W, int ldw, int koff, int N, int K, int n_valid,
                                            int k_valid, int tid, int nth) {
    for (int i = tid; i < N * K; i += nth) {
        const int n = i / K, k = i % K;
        const float v = (n < n_valid && k < k_valid) ? W[(size_t)n * ldw + koff + k] : 0.f;
        const float h = tc::tf32_rn(v);
        const uint32_t off = tc::tile_off_b32(n, k, K / 4) / 4;
        hi[off] = h;
        lo[off] = tc::tf32_rn(v - h);
    }
}

// D[128 x N] = A[128 x K] (TMEM, hi at a_hi, lo at a_lo) * B[N x K]^T (shared memory, hi | lo): 3 * K / 8 MMAs by one thread
__device__ __forceinline__ void issue_layer(uint32_t d, uint32_t a_hi, uint32_t a_lo, const float* b_hi, const float* b_lo, int K, int N) {
    const uint32_t idesc = tc::make_idesc(2, 2, FTILE, N);
    const uint32_t sbo = (uint32_t)(K / 4) * 128;
    const uint64_t bdh = tc::make_smem_desc(tc::smem_u32(b_hi), 128, sbo), bdl = tc::make_smem_desc(tc::smem_u32(b_lo), 128, sbo);
    for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t h = tc::desc_add(bdh, ks * 256), l = tc::desc_add(bdl, ks * 256);
        tc::mma_tf32_ts(d, a_hi + ks * 8, h, idesc, ks > 0);
        tc::mma_tf32_ts(d, a_hi + ks * 8, l, idesc, 1);
        tc::mma_tf32_ts(d, a_lo + ks * 8, h, idesc, 1);
    }
}

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.01f * x; }

__global__ void __launch_bounds__(FTILE, 2) flow_block_fwd_tc_kernel(FlowTcParams p) {
    extern __shared__ __align__(1024) uint8_t fsm_raw[];
    float* sp = reinterpret_cast<float*>(fsm_raw);
    auto carve = [&](int n) { float* r = sp; sp += n; return r; };
    float* b1_hi = carve(FH * K1); float* b1_lo = carve(FH * K1);
    float* b2_hi = carve(FH * FH); float* b2_lo = carve(FH * FH);
    float* b3_hi = carve(FH * FH); float* b3_lo = carve(FH * FH);
    float* b4_hi = carve(N4 * FH); float* b4_lo = carve(N4 * FH);
    float* bias1 = carve(FH); float* bias2 = carve(FH); float* bias3 = carve(FH); float* bias4 = carve(N4);
    float* h1p = carve(FMAXP * FH);
    uint64_t* bar = reinterpret_cast<uint64_t*>(carve(2));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(carve(2));

    const int tid = threadIdx.x, warp = tid >> 5;
    const int F = p.F, KW1 = FPE + F;
    fill_b_tile(b1_hi, b1_lo, p.W1, KW1, 0, FH, K1, FH, FPE, tid, FTILE);
    fill_b_tile(b2_hi, b2_lo, p.W2, FH, 0, FH, FH, FH, FH, tid, FTILE);
    fill_b_tile(b3_hi, b3_lo, p.W3, FH, 0, FH, FH, FH, FH, tid, FTILE);
    fill_b_tile(b4_hi, b4_lo, p.W4, FH, 0, N4, FH, NST, FH, tid, FTILE);
    for (int i = tid; i < FH; i += FTILE) { bias1[i] = p.b1[i]; bias2[i] = p.b2[i]; bias3[i] = p.b3[i]; }
    for (int i = tid; i < N4; i += FTILE) bias4[i] = i < NST ? p.b4[i] : 0.f;
    if (warp == 0) tc::tmem_alloc<256>(tmem_slot);
    if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t d_col = tmem, a_hi = tmem + 64, a_lo = tmem + 128;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    uint32_t parity = 0;

    const int64_t ntiles = (p.M + FTILE - 1) / FTILE;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t i0 = tile * FTILE;
        const int64_t i_last = (i0 + FTILE - 1 < p.M ? i0 + FTILE - 1 : p.M - 1);
        const int64_t p_first = i0 / p.sn;
        const int np = (int)(i_last / p.sn - p_first) + 1;
        // per-point part of the first layer: h1p[lp][j] = b1[j] + sum_k W1[j][7 + k] (scale feat[p][k] + offset)
        for (int e = tid; e < np * FH; e += FTILE) {
            const int lp = e / FH, j = e % FH;
            const float* f = p.feat + (p_first + lp) * F;
            const float* wr = p.W1 + (size_t)j * KW1 + FPE;
            float acc = bias1[j];
            for (int k = 0; k < F; ++k) acc = fmaf(__ldg(wr + k), __ldg(f + k) * p.scale + p.offset, acc);
            h1p[lp * FH + j] = acc;
        }
        const int64_t i = i0 + tid;
        const bool live = i < p.M;
        float yc = 0.f, yt = 0.f;
        {   // ---- layer-1 A operand: Reshift(PE(y_c)) (7 columns + zeros) -> tensor memory ---------------------------
            float xh[16], xl[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) { xh[k] = 0.f; xl[k] = 0.f; }
            if (live) {
                const float y0 = p.y_in[i * 2], y1 = p.y_in[i * 2 + 1];
                yc = p.cond ? y1 : y0; yt = p.cond ? y0 : y1;
                float x[FPE];
                x[0] = yc; x[1] = sinf(yc); x[2] = cosf(yc); x[3] = sinf(yc * 2.f); x[4] = cosf(yc * 2.f); x[5] = sinf(yc * 4.f); x[6] = cosf(yc * 4.f);
#pragma unroll
                for (int k = 0; k < FPE; ++k) { const float v = x[k] * p.scale + p.offset; xh[k] = tc::tf32_rn(v); xl[k] = tc::tf32_rn(v - xh[k]); }
            }
            tc::tmem_st16(a_hi + lane_base, xh);
            tc::tmem_st16(a_lo + lane_base, xl);
            tc::tmem_st_wait();
        }
        tc::fence_before_sync();
        __syncthreads();                                        // A operand + h1p visible
        if (tid == 0) { tc::fence_after_sync(); issue_layer(d_col, a_hi, a_lo, b1_hi, b1_lo, K1, FH); tc::mma_commit(bar); }
        const float* hb = h1p + (live ? (int)(i / p.sn - p_first) : 0) * FH;
        // ---- layers 1-3: D -> bias + LeakyReLU -> A operand of the next layer ------------------------------------------
#pragma unroll 1
        for (int layer = 0; layer < 3; ++layer) {
            tc::mbar_wait(bar, parity); parity ^= 1;
            tc::fence_after_sync();
            const float* bias = layer == 0 ? hb : (layer == 1 ? bias2 : bias3);
#pragma unroll
            for (int c0 = 0; c0 < FH; c0 += 16) {
                float v[16], lo[16];
                tc::tmem_ld16(d_col + lane_base + c0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = live ? leaky(v[j] + bias[c0 + j]) : 0.f;
                if (p.save_h && live) {
                    float4* dst = reinterpret_cast<float4*>(p.save_h + ((size_t)i * 3 + layer) * FH + c0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) { const float h = tc::tf32_rn(v[j]); lo[j] = tc::tf32_rn(v[j] - h); v[j] = h; }
                tc::tmem_st16(a_hi + lane_base + c0, v);
                tc::tmem_st16(a_lo + lane_base + c0, lo);
            }
            tc::tmem_st_wait();
            tc::fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tc::fence_after_sync();
                if (layer == 0) issue_layer(d_col, a_hi, a_lo, b2_hi, b2_lo, FH, FH);
                else if (layer == 1) issue_layer(d_col, a_hi, a_lo, b3_hi, b3_lo, FH, FH);
                else issue_layer(d_col, a_hi, a_lo, b4_hi, b4_lo, FH, N4);
                tc::mma_commit(bar);
            }
        }
        // ---- layer 4 -> spline -------------------------------------------------------------------------------------------
        tc::mbar_wait(bar, parity); parity ^= 1;
        tc::fence_after_sync();
        {
            float st[FSTP];
            tc::tmem_ld16(d_col + lane_base, st);
            tc::tmem_ld8(d_col + lane_base + 16, st + 16);
            if (live) {
#pragma unroll
                for (int j = 0; j < FSTP; ++j) st[j] += bias4[j];
                if (p.save_st) {
                    float4* dst = reinterpret_cast<float4*>(p.save_st + (size_t)i * FSTP);
#pragma unroll
                    for (int j = 0; j < FSTP / 4; ++j) dst[j] = make_float4(st[4 * j], st[4 * j + 1], st[4 * j + 2], st[4 * j + 3]);
                }
                float xt, lj;
                if (p.inverse) pwquad_eval_inverse(st, yt, xt, lj);
                else pwquad_eval_forward(st, yt, xt, lj);
                p.y_out[i * 2 + p.cond] = yc;
                p.y_out[i * 2 + 1 - p.cond] = xt;
                p.logj_out[i] = (p.logj_in ? p.logj_in[i] : 0.f) + lj;
            }
        }
        tc::fence_before_sync();
        __syncthreads();                                        // D and h1p are free for the next tile
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

size_t flow_tc_smem() {
    return sizeof(float) * (size_t)(2 * FH * K1 + 4 * FH * FH + 2 * N4 * FH + 3 * FH + N4 + FMAXP * FH + 4) + 1024;
}

}  // namespace

// tensor-core forward of one coupling block (arguments as tf_flow_block_fwd; save_h / save_st keep the activations for the backward)
int tf_internal_flow_block_fwd_tc(const float* y_in, const float* logj_in, const float* feat, int feat_dim, int sn, const float* W1,
                                  const float* b1, const float* W2, const float* b2, const float* W3, const float* b3, const float* W4,
                                  const float* b4, float scale, float offset, int cond, int inverse, int64_t M, float* y_out, float* logj_out,
                                  float* save_h, float* save_st, cudaStream_t stream) {
    FlowTcParams p = {W1, b1, W2, b2, W3, b3, W4, b4, feat_dim, scale, offset, y_in, logj_in, feat, sn, cond, inverse, M, y_out, logj_out,
                      save_h, save_st};
    const size_t smem = flow_tc_smem();
    cudaFuncSetAttribute(flow_block_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t ntiles = (M + FTILE - 1) / FTILE;
    const int64_t cap = (int64_t)tf_num_sms() * 2;
    const int grid = (int)(ntiles < cap ? ntiles : cap);
    {
        TfKernelTimer timer("flow_block_fwd_tc", stream);
        flow_block_fwd_tc_kernel<<<grid, FTILE, smem, stream>>>(p);
    }
    tf_count_launches(1);
    return 0;
}
