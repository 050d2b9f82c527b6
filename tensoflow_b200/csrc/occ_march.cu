// Occupancy-grid ray marcher: the packed (ray_indices, t_starts, t_ends) wire format nerfacc.OccGridEstimator.sampling
// hands to the compositor in the reference's `*_occ` configs (network/shapeRenderer.py:950-959; grid built at :213-215).
// nerfacc is not vendored: this restates its documented behaviour (tensoflow_b200/occ_grid.py states the exact rule).
//   lattice      t_k = near[r] + k * step                      (near already carries the stratified jitter)
//   keep k  iff  mid = t_k + step/2 < far,  p = o + d * mid inside the aabb,  binaries[cell(p)] != 0
// One warp per ray: lane l tests k = base + l, a ballot compacts the kept samples in order (count pass, host prefix sum,
// write pass).  The slab test only bounds the k range (one step of slack on both sides); the per-sample test above decides.
// Every float operation is a single rounded op in a fixed order (no FMA contraction) so that a plain PyTorch restatement
// makes bit-identical keep / drop decisions.
#include "common.cuh"

namespace {

struct OccGrid {
    float lo[3], hi[3];
    int res[3];
};

__device__ __forceinline__ bool occ_keep(const OccGrid& g, const uint8_t* __restrict__ bits, const float o[3], const float d[3],
                                         float near, float step, float far, int k, float& t0, float& t1) {
    t0 = __fadd_rn(near, __fmul_rn((float)k, step));
    t1 = __fadd_rn(t0, step);
    const float mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
    if (!(mid < far)) return false;
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float p = __fadd_rn(o[a], __fmul_rn(d[a], mid));
        const float u = __fmul_rn(__fdiv_rn(__fsub_rn(p, g.lo[a]), __fsub_rn(g.hi[a], g.lo[a])), (float)g.res[a]);
        if (!(u >= 0.f) || !(u < (float)g.res[a])) return false;
        c[a] = (int)floorf(u);
    }
    return bits[((size_t)c[0] * g.res[1] + c[1]) * g.res[2] + c[2]] != 0;
}

// conservative k range of the ray inside the aabb (empty when the ray misses it)
__device__ __forceinline__ void occ_range(const OccGrid& g, const float o[3], const float d[3], float near, float step, float far,
                                          int& k0, int& k1) {
    float ta = -1e30f, tb = 1e30f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        if (fabsf(d[a]) < 1e-20f) {
            if (o[a] < g.lo[a] || o[a] > g.hi[a]) { ta = 1e30f; tb = -1e30f; }
        } else {
            const float i = 1.f / d[a];
            const float x = (g.lo[a] - o[a]) * i, y = (g.hi[a] - o[a]) * i;
            ta = fmaxf(ta, fminf(x, y));
            tb = fminf(tb, fmaxf(x, y));
        }
    }
    tb = fminf(tb, far);
    if (!(tb >= ta) || ta <= -1e29f) { k0 = 0; k1 = 0; return; }      // miss, or a zero direction (no slab bounds the march)
    const float lo = (ta - near) / step - 2.f, hi = (tb - near) / step + 2.f;
    k0 = lo > 0.f ? (int)fminf(lo, 2.0e9f) : 0;
    k1 = hi > 0.f ? (int)fminf(hi, 2.0e9f) : 0;
}

template <bool WRITE>
__global__ void __launch_bounds__(256) occ_march_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                        const float* __restrict__ near_p, int n_rays, float far, float step,
                                                        const OccGrid g, const uint8_t* __restrict__ bits, int32_t* __restrict__ counts,
                                                        const int32_t* __restrict__ offsets, int64_t* __restrict__ ray_indices,
                                                        float* __restrict__ t_starts, float* __restrict__ t_ends) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int ray = blockIdx.x * wpb + (threadIdx.x >> 5); ray < n_rays; ray += gridDim.x * wpb) {
        const float o[3] = {rays_o[ray * 3 + 0], rays_o[ray * 3 + 1], rays_o[ray * 3 + 2]};
        const float d[3] = {rays_d[ray * 3 + 0], rays_d[ray * 3 + 1], rays_d[ray * 3 + 2]};
        const float near = near_p[ray];
        int k0, k1;
        occ_range(g, o, d, near, step, far, k0, k1);
        int n = 0;
        const int out0 = WRITE ? offsets[ray] : 0;
        for (int base = k0; base < k1; base += 32) {
            const int k = base + lane;
            float t0 = 0.f, t1 = 0.f;
            const bool keep = k < k1 && occ_keep(g, bits, o, d, near, step, far, k, t0, t1);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (WRITE && keep) {
                const int dst = out0 + n + __popc(m & ((1u << lane) - 1u));
                ray_indices[dst] = ray;
                t_starts[dst] = t0;
                t_ends[dst] = t1;
            }
            n += __popc(m);
        }
        if (!WRITE && lane == 0) counts[ray] = n;
    }
}

int occ_check(const float* aabb, const int32_t* res, float step) {
    TF_REQUIRE(aabb && res, "occ march: NULL aabb / resolution");
    TF_REQUIRE(res[0] > 0 && res[1] > 0 && res[2] > 0, "occ march: resolution must be positive");
    TF_REQUIRE(aabb[3] > aabb[0] && aabb[4] > aabb[1] && aabb[5] > aabb[2], "occ march: empty aabb");
    TF_REQUIRE(step > 0.f, "occ march: render_step_size must be > 0");
    return 0;
}

OccGrid occ_grid(const float* aabb, const int32_t* res) {
    OccGrid g;
    for (int a = 0; a < 3; ++a) { g.lo[a] = aabb[a]; g.hi[a] = aabb[3 + a]; g.res[a] = res[a]; }
    return g;
}

int occ_blocks(int n_rays) {
    int grid = (n_rays + 7) / 8;
    const int cap = tf_num_sms() * 8;
    return grid > cap ? cap : grid;
}

}  // namespace

extern "C" TF_API int tf_occ_march_count(const float* rays_o, const float* rays_d, const float* near, int32_t n_rays, float far,
                                         float step, const float* aabb, const int32_t* res, const uint8_t* binaries,
                                         int32_t* counts, tf_stream_t stream) {
    if (n_rays == 0) return 0;
    TF_REQUIRE(rays_o && rays_d && near && binaries && counts, "tf_occ_march_count: NULL pointer");
    if (int rc = occ_check(aabb, res, step)) return rc;
    occ_march_kernel<false><<<occ_blocks(n_rays), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, near, n_rays, far, step, occ_grid(aabb, res),
                                                                                   binaries, counts, nullptr, nullptr, nullptr, nullptr);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_occ_march_count");
    return 0;
}

extern "C" TF_API int tf_occ_march_write(const float* rays_o, const float* rays_d, const float* near, int32_t n_rays, float far,
                                         float step, const float* aabb, const int32_t* res, const uint8_t* binaries,
                                         const int32_t* offsets, int64_t* ray_indices, float* t_starts, float* t_ends,
                                         tf_stream_t stream) {
    if (n_rays == 0) return 0;
    TF_REQUIRE(rays_o && rays_d && near && binaries && offsets && ray_indices && t_starts && t_ends, "tf_occ_march_write: NULL pointer");
    if (int rc = occ_check(aabb, res, step)) return rc;
    occ_march_kernel<true><<<occ_blocks(n_rays), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, near, n_rays, far, step, occ_grid(aabb, res),
                                                                                  binaries, nullptr, offsets, ray_indices, t_starts, t_ends);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_occ_march_write");
    return 0;
}
