// Fused TensoSDF stencil forward on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same contract as the SIMT kernel in sdf_stencil.cu (reference network/fields.py:262-299, 227-260)
// but the decoder's first layer -- 7 x [N,K] x [K,H], 85 % of the decoder FLOPs -- runs as
// tcgen05.mma kind::tf32 with fp32-level accuracy from operand splitting
// (x = hi + lo, D = A_hi W_hi + A_hi W_lo + A_lo W_hi: "3xTF32").
//
// One persistent CTA (256 threads) per SM walks MMA tiles of M = 128 rows.  In stencil mode a tile is
// sample-major: 18 samples x 7 queries (centre, +-x, +-y, +-z) so that the queries of a sample share
// their plane / line fetches (stencil_site.cuh); in SDF-only mode it is 128 samples:
//   gather  : all threads; one (sample, plane, channel group) site per thread and pass, texel reads
//             are 144-byte runs per site; values are split hi/lo and stored in the K-major no-swizzle
//             UMMA layout with a padded K-chunk stride (conflict-free stores)
//   W0      : pre-split / pre-tiled once per call into K-slices of 16 (prep kernel); slices stream
//             L2 -> shared memory through a 3-stage ring with cp.async.bulk + mbarrier (one driver thread)
//   MMA     : driver thread issues 6 tcgen05.mma per slice (2 k-steps x 3 passes), accumulator
//             [128 x H] fp32 in TMEM, double buffered (2 x 256 columns)
//   epilogue: overlaps the next tile's MMAs; tcgen05.ld -> +b0 -> Softplus(beta=100) -> dot with
//             W1[0,:] (the SDF output; taps need nothing else) ; the centre tile also streams its
//             hidden activations to HBM for the appearance head (second layer, [N,H] x [H,A])
//   finalize: per tile, the 7 SDF values of a sample -> sdf7, central-difference gradient, hessian term.
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"
#include "stencil_site.cuh"

namespace {

constexpr int TM = 128;     // rows per MMA tile = samples per block
constexpr int NQ7 = 7;
constexpr int KSL = 16;     // K-slice (2 tf32 MMA k-steps)
constexpr int NST = 2;      // W ring stages
constexpr int NTH = 512;    // 16 warps at <= 128 registers: the gather is latency-bound, thread-level parallelism hides it
constexpr int NCG = NTH / 128; // column groups of the epilogue (warps sharing a TMEM lane quarter)

struct TcParams {
    tf_vm_field_t f;
    const float* xyz;
    const float* level;
    int64_t n;
    const float* W0tc;   // [S][2][H*16] pre-tiled hi/lo slices
    const float* b0;
    const float* w1r0;   // W1 row 0
    const float* b1;
    int K, KT, H, nq;    // nq = 7 (stencil) or 1 (sdf only)
    float units[3];
    float* sdf7; float* grad; float* hess; float* sdf1;
    float* spc;          // [n][H] centre hidden activations (NULL = not needed)
#ifdef TF_TC_DEBUG_SWITCHES
    int debug;           // timing experiments only (never compiled into the product library): 1 skip gathers, 2 skip MMAs, 4 skip epilogue math
#endif
};

// timing-experiment switches exist only in builds made with -DTF_TC_DEBUG_SWITCHES; the product library has none
#ifdef TF_TC_DEBUG_SWITCHES
#define TF_DBG(p, bit) ((p).debug & (bit))
#else
#define TF_DBG(p, bit) 0
#endif

__global__ void tc_prep_w0_kernel(const float* __restrict__ W0, int K, int KT, int H, float* __restrict__ W0tc) {
    const int S = KT / KSL;
    const int total = S * H * KSL;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int kl = i % KSL, h = (i / KSL) % H, s = i / (KSL * H);
        const int k = s * KSL + kl;
        const float v = k < K ? W0[(size_t)h * K + k] : 0.f;
        const float hi = tc::tf32_rn(v);
        const float lo = tc::tf32_rn(v - hi);
        const uint32_t off = tc::tile_off_b32(h, kl, KSL / 4) / 4;
        float* base = W0tc + (size_t)s * 2 * H * KSL;
        base[off] = hi;
        base[(size_t)H * KSL + off] = lo;
    }
}

__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc::smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}

// SDF-only mode: gather the feature rows of 128 samples (one query each) into the A operand
__device__ __forceinline__ void tc_gather_rows(const TcParams& p, int64_t s_base, uint8_t* a_hi, uint8_t* a_lo) {
    const int C = p.f.n_comp, C4 = C / 4, G = p.KT / 4;
    const bool has_level = p.level != nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // unit = (8-row group, plane): lane -> (row = lane%8, channel groups lane/8, lane/8+4, ...); the
    // sampling plan of the (row, plane) pair is computed once and reused for its channel groups
    const int n_units = (TM / 8) * 3;
    for (int u = warp; u < n_units; u += NTH / 32) {
        const int rg = u / 3, i = u % 3;
        const int row = rg * 8 + (lane & 7);
        const int64_t n = s_base + row;
        const bool valid = n < p.n;
        VmTaps taps;
        if (valid) {
            const float x[3] = {p.xyz[n * 3 + 0], p.xyz[n * 3 + 1], p.xyz[n * 3 + 2]};
            taps = vm_taps(p.f, x, has_level ? p.level[n] : 0.f, has_level, i);
        }
        for (int c4 = lane >> 3; c4 < C4; c4 += 4) {
            float4 v = f4_zero();
            if (valid) {
                float4 P, L;
                vm_fetch(taps, C, c4 * 4, P, L);
                v = f4_mul(P, L);
            }
            site::put(a_hi, a_lo, site::a_off(row, i * C4 + c4, p.KT), v);
        }
    }
    // raw xyz (fields.py:265,298) + zero padding groups
    const int tail_g = G - 3 * C4;
    for (int it = threadIdx.x; it < TM * tail_g; it += NTH) {
        const int row = it % TM, g = 3 * C4 + it / TM;
        const int64_t n = s_base + row;
        float4 v = f4_zero();
        if (g == 3 * C4 && n < p.n) v = make_float4(p.xyz[n * 3 + 0], p.xyz[n * 3 + 1], p.xyz[n * 3 + 2], 0.f);
        site::put(a_hi, a_lo, site::a_off(row, g, p.KT), v);
    }
}

__device__ __forceinline__ void tc_gather(const TcParams& p, int64_t tile, uint8_t* a_hi, uint8_t* a_lo) {
    if (p.nq == NQ7) site::gather_tile_lean(p.f, p.xyz, p.level, p.n, p.units, tile * site::SPT, p.KT, a_hi, a_lo, nullptr, NTH, threadIdx.x);
    else tc_gather_rows(p, tile * TM, a_hi, a_lo);
}

__global__ void __launch_bounds__(NTH, 1) sdf_stencil_fwd_tc_kernel(TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int H = p.H, KT = p.KT, S = KT / KSL, nq = p.nq;
    const int spt = nq == NQ7 ? site::SPT : TM;               // samples per tile
    const uint32_t a_part = site::a_part_bytes(KT);           // bytes of one A part
    const uint32_t w_part = (uint32_t)H * KSL * 4;            // bytes of one W slice part
    uint8_t* a_hi = smem;
    uint8_t* a_lo = a_hi + a_part;
    uint8_t* wst = a_lo + a_part;                             // NST stages x (hi, lo)
    float* b0s = reinterpret_cast<float*>(wst + (size_t)NST * 2 * w_part);
    float* w1s = b0s + H;
    float* sdfs = w1s + H;                                    // [2 tiles][NCG column groups][TM]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sdfs + 2 * NCG * TM);
    uint64_t* full = bars;                                    // [NST]
    uint64_t* empty = bars + NST;                             // [NST]
    uint64_t* dfull = bars + 2 * NST;                         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NST + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ntiles = (p.n + spt - 1) / spt;
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 0) tc::tmem_alloc<512>(tmem_slot);
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(&dfull[0], 1); tc::mbar_init(&dfull[1], 1);
        tc::mbar_fence_init();
    }
    for (int i = tid; i < H; i += NTH) { b0s[i] = p.b0[i]; w1s[i] = p.w1r0[i]; }
    if (my_tiles > 0) tc_gather(p, blockIdx.x, a_hi, a_lo);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t idesc = tc::make_idesc(2, 2, TM, H);
    const uint32_t a_sbo = site::a_sbo(KT), a_kstep = 2 * site::A_LBO;
    const uint32_t w_sbo = (KSL / 4) * 128;
    const uint64_t a_desc_hi = tc::make_smem_desc(tc::smem_u32(a_hi), site::A_LBO, a_sbo), a_desc_lo = tc::make_smem_desc(tc::smem_u32(a_lo), site::A_LBO, a_sbo);

    int64_t g_issue = 0, g_mma = 0;                           // driver-thread state (W slice counters)
    const int64_t total_slices = my_tiles * S;

    for (int64_t t = 0; t <= my_tiles; ++t) {
        // ---- driver: stream W slices and issue the MMAs of tile t --------------------------------
        if (tid == 0 && t < my_tiles) {
            const uint32_t dcol = tmem_base + (uint32_t)(t & 1) * 256;
            for (int s = 0; s < S; ++s) {
                while (g_issue < total_slices && g_issue < g_mma + NST) {
                    const int st = (int)(g_issue % NST);
                    tc::mbar_wait(&empty[st], (uint32_t)(((g_issue / NST) & 1) ^ 1));
                    mbar_expect_tx(&full[st], 2 * w_part);
                    bulk_copy_g2s(wst + (size_t)st * 2 * w_part, p.W0tc + (size_t)(g_issue % S) * 2 * H * KSL, 2 * w_part, &full[st]);
                    ++g_issue;
                }
                const int st = (int)(g_mma % NST);
                tc::mbar_wait(&full[st], (uint32_t)((g_mma / NST) & 1));
                tc::fence_after_sync();
                const uint32_t w_hi = tc::smem_u32(wst + (size_t)st * 2 * w_part);
                const uint64_t wdh0 = tc::make_smem_desc(w_hi, 128, w_sbo), wdl0 = tc::make_smem_desc(w_hi + w_part, 128, w_sbo);
                const uint64_t adh0 = tc::desc_add(a_desc_hi, s * (KSL / 8) * a_kstep), adl0 = tc::desc_add(a_desc_lo, s * (KSL / 8) * a_kstep);
#pragma unroll
                for (int ks = 0; ks < KSL / 8; ++ks) {
                    if (TF_DBG(p, 2)) break;
                    const uint64_t adh = tc::desc_add(adh0, ks * a_kstep), adl = tc::desc_add(adl0, ks * a_kstep);
                    const uint64_t wdh = tc::desc_add(wdh0, ks * 256), wdl = tc::desc_add(wdl0, ks * 256);
                    tc::mma_tf32_ss(dcol, adh, wdh, idesc, (s | ks) != 0);
                    tc::mma_tf32_ss(dcol, adh, wdl, idesc, 1);
                    tc::mma_tf32_ss(dcol, adl, wdh, idesc, 1);
                }
                tc::mma_commit(&empty[st]);
                ++g_mma;
            }
            tc::mma_commit(&dfull[t & 1]);
        }
        // ---- epilogue of tile t-1 (overlaps the MMAs of tile t) --------------------------------------
        if (t > 0) {
            const int64_t tp = t - 1;
            const int64_t tile = blockIdx.x + tp * gridDim.x;
            tc::mbar_wait(&dfull[tp & 1], (uint32_t)((tp >> 1) & 1));
            tc::fence_after_sync();
            const int lq = warp & 3, chh = warp >> 2;
            const int row = lq * 32 + lane;
            const int s = row / nq, q = row - s * nq;
            const int64_t n = tile * spt + s;
            const bool centre = q == 0 && s < spt && n < p.n;
            const uint32_t dcol = tmem_base + (uint32_t)(tp & 1) * 256 + ((uint32_t)(lq * 32) << 16);
            float psum = 0.f;
            for (int c0 = chh * 32; c0 < H; c0 += 32 * NCG) {
                if (TF_DBG(p, 4)) break;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {            // two 16-column reads keep the register footprint small
                    const int cc = c0 + hh * 16;
                    float v[16];
                    tc::tmem_ld16(dcol + cc, v);
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 bb = *reinterpret_cast<const float4*>(b0s + cc + j);
                        const float4 ww = *reinterpret_cast<const float4*>(w1s + cc + j);
                        v[j + 0] = softplus100_fast(v[j + 0] + bb.x); psum = fmaf(v[j + 0], ww.x, psum);
                        v[j + 1] = softplus100_fast(v[j + 1] + bb.y); psum = fmaf(v[j + 1], ww.y, psum);
                        v[j + 2] = softplus100_fast(v[j + 2] + bb.z); psum = fmaf(v[j + 2], ww.z, psum);
                        v[j + 3] = softplus100_fast(v[j + 3] + bb.w); psum = fmaf(v[j + 3], ww.w, psum);
                    }
                    if (centre && p.spc) {
                        float4* dst = reinterpret_cast<float4*>(p.spc + (size_t)n * H + cc);
#pragma unroll
                        for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            }
            sdfs[((tp & 1) * NCG + chh) * TM + row] = psum;
            tc::fence_before_sync();
        }
        // ---- wait for the MMAs of tile t, then gather tile t+1 into the (now free) A buffer ---------------
        if (t < my_tiles) {
            tc::mbar_wait(&dfull[t & 1], (uint32_t)((t >> 1) & 1));
            if (t + 1 < my_tiles) {
                if (!TF_DBG(p, 1)) tc_gather(p, blockIdx.x + (t + 1) * gridDim.x, a_hi, a_lo);
                tc::fence_async_smem();
            }
        }
        tc::fence_before_sync();
        __syncthreads();
        tc::fence_after_sync();
        // ---- finalize the samples of tile t-1 ------------------------------------------------------------
        if (t > 0 && tid < spt) {
            const int64_t tp = t - 1;
            const int64_t n = (blockIdx.x + tp * gridDim.x) * spt + tid;
            const float* sp = &sdfs[(tp & 1) * NCG * TM + tid * nq];
            if (n < p.n) {
                const float b1 = __ldg(p.b1);
                if (nq == 1) {
                    float v = b1;
#pragma unroll
                    for (int g = 0; g < NCG; ++g) v += sp[g * TM];
                    p.sdf1[n] = v;
                } else {
                    float sd[NQ7];
#pragma unroll
                    for (int r = 0; r < NQ7; ++r) {
                        float v = b1;
#pragma unroll
                        for (int g = 0; g < NCG; ++g) v += sp[g * TM + r];
                        sd[r] = v; p.sdf7[n * NQ7 + r] = v;
                    }
                    float g[3], h[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float e = p.units[k];
                        g[k] = (sd[1 + 2 * k] - sd[2 + 2 * k]) / (2.f * e);
                        h[k] = (sd[1 + 2 * k] + sd[2 + 2 * k] - 2.f * sd[0]) / (e * e);
                    }
                    if (p.grad) { p.grad[n * 3 + 0] = g[0]; p.grad[n * 3 + 1] = g[1]; p.grad[n * 3 + 2] = g[2]; }
                    if (p.hess) p.hess[n] = (g[0] * h[0] + g[1] * h[1] + g[2] * h[2]) / (g[0] * g[0] + g[1] * g[1] + g[2] * g[2] + 1e-5f);
                }
            }
        }
    }
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem_base);
}

}  // namespace

size_t tf_internal_tc_fwd_smem(int KT, int H) {
    return (size_t)2 * 16 * (KT / 4) * site::A_LBO + (size_t)NST * 2 * H * KSL * 4 + (size_t)2 * H * 4 + (size_t)2 * NCG * TM * 4 + (2 * NST + 2) * 8 + 16;
}

// workspace floats needed in front of spc: the pre-tiled W0
size_t tf_internal_tc_w0_floats(int KT, int H) { return (size_t)KT * H * 2; }

int tf_internal_stencil_fwd_tc(const tf_vm_field_t* f, const tf_sdf_mlp_t* m, const float* xyz, const float* level, int64_t n,
                               const float units[3], int nq, float* sdf7, float* grad, float* hess, float* sdf1, float* spc,
                               float* w0tc, cudaStream_t stream) {
    const int C = f->n_comp, K = 3 * C + 3, KT = (K + KSL - 1) / KSL * KSL, H = m->hidden;
    TcParams p = {};
    p.f = *f; p.xyz = xyz; p.level = level; p.n = n;
    p.W0tc = w0tc; p.b0 = m->b0; p.w1r0 = m->W1; p.b1 = m->b1;
    p.K = K; p.KT = KT; p.H = H; p.nq = nq;
    for (int k = 0; k < 3; ++k) p.units[k] = units ? units[k] : 0.f;
    p.sdf7 = sdf7; p.grad = grad; p.hess = hess; p.sdf1 = sdf1; p.spc = spc;
#ifdef TF_TC_DEBUG_SWITCHES
    { const char* e = getenv("TF_TC_DEBUG"); p.debug = e ? atoi(e) : 0; }
#endif
    tc_prep_w0_kernel<<<64, 256, 0, stream>>>(m->W0, K, KT, H, w0tc);
    const size_t smem = tf_internal_tc_fwd_smem(KT, H);
    if (smem > 227 * 1024) { tf_set_error("tensor-core stencil: tile does not fit shared memory (KT=%d, H=%d)", KT, H); return 1; }
    cudaFuncSetAttribute(sdf_stencil_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int spt = nq == NQ7 ? site::SPT : TM;
    const int64_t ntiles = (n + spt - 1) / spt;
    const int grid = (int)(ntiles < tf_num_sms() ? ntiles : tf_num_sms());
    {
        TfKernelTimer timer("sdf_stencil_fwd_tc", stream);
        sdf_stencil_fwd_tc_kernel<<<grid, NTH, smem, stream>>>(p);
    }
    tf_count_launches(2);
    return 0;
}
