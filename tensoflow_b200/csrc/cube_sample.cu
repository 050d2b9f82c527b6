// Seamless (tri)linear cubemap lookup, differentiable in the textures, the direction and the mip level: the
// `dr.texture(..., boundary_mode='cube')` calls of the shape-stage split-sum light (reference network/light.py:95-122:
// diffuse lookup at the normal, specular lookup at the reflected direction with `mip_level_bias` from the roughness over the
// user-supplied prefiltered stack).  One thread per sample:
//   forward   out = (1-f) S_l0(d) + f S_l1(d),  lv = clamp(level, 0, L-1), l0 = floor(lv), f = lv - l0, l1 = min(l0+1, L-1)
//             S_l(d) = sum_k w_k T_l[idx_k] / sum_k w_k over the 4 bilinear taps of the face footprint; taps that leave the face
//             fold onto the neighbouring face, the tap leaving in both axes (cube corner) is dropped
//   backward  d T_l[idx_k] += g w_k / W  (atomics; the textures are <= 128^2 x 6 texels: L2 resident)
//             d d       through the bilinear fractions (d S / d fu = sum_k dw_k/dfu (T_k - S) / W), u = (x/m + 1) R/2 - 1/2
//             d level   = g . (S_l1 - S_l0) inside the clamp range
// In round 1 this was ~150 tensor ops per lookup and level (about 3000 launches per ShapeRenderer step).
#include "common.cuh"

namespace {

constexpr int CS_MAX_LEVELS = 8;

struct CubeStack {
    const float* tex[CS_MAX_LEVELS];
    float* d_tex[CS_MAX_LEVELS];
    int res[CS_MAX_LEVELS];
    int n;
};

// face convention of network/light_utils.py:24-31 / renderutils/c_src/cubemap.cu:32-60
__device__ __forceinline__ void cs_cube_to_dir(int s, float x, float y, float c[3]) {
    switch (s) {
        case 0: c[0] = 1.f; c[1] = -y; c[2] = -x; break;
        case 1: c[0] = -1.f; c[1] = -y; c[2] = x; break;
        case 2: c[0] = x; c[1] = 1.f; c[2] = y; break;
        case 3: c[0] = x; c[1] = -1.f; c[2] = -y; break;
        case 4: c[0] = x; c[1] = -y; c[2] = 1.f; break;
        default: c[0] = -x; c[1] = -y; c[2] = -1.f; break;
    }
}

// x = sx d[ax] / m, y = sy d[ay] / m, m = sm d[am] > 0
struct FaceMap { int face, ax, ay, am; float sx, sy, sm, m, x, y; };

__device__ __forceinline__ FaceMap cs_face(const float d[3]) {
    FaceMap f;
    const float a0 = fabsf(d[0]), a1 = fabsf(d[1]), a2 = fabsf(d[2]);
    if (a0 >= a1 && a0 >= a2) {
        f.am = 0; f.ax = 2; f.ay = 1; f.sy = -1.f;
        if (d[0] >= 0.f) { f.face = 0; f.sx = -1.f; f.sm = 1.f; } else { f.face = 1; f.sx = 1.f; f.sm = -1.f; }
    } else if (a1 >= a2) {
        f.am = 1; f.ax = 0; f.ay = 2; f.sx = 1.f;
        if (d[1] >= 0.f) { f.face = 2; f.sy = 1.f; f.sm = 1.f; } else { f.face = 3; f.sy = -1.f; f.sm = -1.f; }
    } else {
        f.am = 2; f.ax = 0; f.ay = 1; f.sy = -1.f;
        if (d[2] >= 0.f) { f.face = 4; f.sx = 1.f; f.sm = 1.f; } else { f.face = 5; f.sx = -1.f; f.sm = -1.f; }
    }
    f.m = fmaxf(fabsf(d[f.am]), 1e-30f);
    f.x = f.sx * d[f.ax] / f.m;
    f.y = f.sy * d[f.ay] / f.m;
    return f;
}

struct LevelTaps {
    int idx[4];
    float w[4];        // raw bilinear weights (0 for the dropped corner tap)
    float dwu[4], dwv[4];   // d w / d fu, d w / d fv
    float wsum;
};

__device__ __forceinline__ LevelTaps cs_taps(const FaceMap& f, int R) {
    const float u = (f.x + 1.f) * 0.5f * R - 0.5f, v = (f.y + 1.f) * 0.5f * R - 0.5f;
    const float u0f = floorf(u), v0f = floorf(v);
    const float fu = u - u0f, fv = v - v0f;
    const int u0 = (int)u0f, v0 = (int)v0f;
    LevelTaps t;
    t.wsum = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int du = k & 1, dv = k >> 1;
        int iu = u0 + du, iv = v0 + dv;
        const float a = du ? fu : 1.f - fu, b = dv ? fv : 1.f - fv;
        float w = a * b, dwu = (du ? 1.f : -1.f) * b, dwv = (dv ? 1.f : -1.f) * a;
        const bool ou = iu < 0 || iu >= R, ov = iv < 0 || iv >= R;
        int f2 = f.face;
        if (ou && ov) {
            w = 0.f; dwu = 0.f; dwv = 0.f;
            iu = min(max(iu, 0), R - 1); iv = min(max(iv, 0), R - 1);
        } else if (ou || ov) {
            const float fx = 2.f * ((float)iu + 0.5f) / R - 1.f, fy = 2.f * ((float)iv + 0.5f) / R - 1.f;
            float c[3];
            cs_cube_to_dir(f.face, fx, fy, c);
            const int major = f.face >> 1;
            float e = 0.f;
            int over = -1;
#pragma unroll
            for (int a3 = 0; a3 < 3; ++a3)
                if (a3 != major && fabsf(c[a3]) > 1.f) { e = fabsf(c[a3]) - 1.f; over = a3; }
#pragma unroll
            for (int a3 = 0; a3 < 3; ++a3) {
                if (a3 == over) c[a3] = c[a3] > 0.f ? 1.f : -1.f;
                if (a3 == major) c[a3] = (c[a3] > 0.f ? 1.f : -1.f) * (1.f - e);
            }
            const FaceMap g = cs_face(c);
            f2 = g.face;
            iu = min(max((int)floorf((g.x + 1.f) * 0.5f * R), 0), R - 1);
            iv = min(max((int)floorf((g.y + 1.f) * 0.5f * R), 0), R - 1);
        }
        t.idx[k] = (f2 * R + iv) * R + iu;
        t.w[k] = w; t.dwu[k] = dwu; t.dwv[k] = dwv;
        t.wsum += w;
    }
    return t;
}

__device__ __forceinline__ void cs_levels(const float* level, int64_t i, int n, int& l0, int& l1, float& f, bool& inside) {
    l0 = 0; l1 = 0; f = 0.f; inside = false;
    if (!level || n <= 1) return;
    const float raw = level[i];
    const float lv = fminf(fmaxf(raw, 0.f), (float)(n - 1));
    inside = raw >= 0.f && raw <= (float)(n - 1);
    const float fl = floorf(lv);
    l0 = (int)fl;
    f = lv - fl;
    l1 = min(l0 + 1, n - 1);
}

__device__ __forceinline__ void cs_fetch(const float* tex, const LevelTaps& t, float s[3]) {
    s[0] = s[1] = s[2] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float* px = tex + (size_t)t.idx[k] * 3;
        s[0] = fmaf(t.w[k], __ldg(px), s[0]); s[1] = fmaf(t.w[k], __ldg(px + 1), s[1]); s[2] = fmaf(t.w[k], __ldg(px + 2), s[2]);
    }
    const float inv = 1.f / t.wsum;
    s[0] *= inv; s[1] *= inv; s[2] *= inv;
}

__global__ void __launch_bounds__(256) cube_sample_fwd_kernel(const CubeStack st, const float* __restrict__ dirs,
                                                              const float* __restrict__ level, int64_t n, float* __restrict__ out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d[3] = {dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]};
    const FaceMap fm = cs_face(d);
    int l0, l1; float f; bool inside;
    cs_levels(level, i, st.n, l0, l1, f, inside);
    float s0[3];
    cs_fetch(st.tex[l0], cs_taps(fm, st.res[l0]), s0);
    if (l1 != l0 && f != 0.f) {
        float s1[3];
        cs_fetch(st.tex[l1], cs_taps(fm, st.res[l1]), s1);
#pragma unroll
        for (int c = 0; c < 3; ++c) s0[c] = (1.f - f) * s0[c] + f * s1[c];
    }
    out[i * 3] = s0[0]; out[i * 3 + 1] = s0[1]; out[i * 3 + 2] = s0[2];
}

// adjoint of one level's lookup: scatters g * scale into d_tex, returns S and accumulates (dS/dx, dS/dy) . g * scale
__device__ __forceinline__ void cs_level_bwd(const float* tex, float* d_tex, const FaceMap& fm, int R, const float g[3], float scale,
                                             float s[3], float& gx, float& gy) {
    const LevelTaps t = cs_taps(fm, R);
    cs_fetch(tex, t, s);
    const float inv = 1.f / t.wsum;
    float gu = 0.f, gv = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float* px = tex + (size_t)t.idx[k] * 3;
        const float dot = g[0] * (__ldg(px) - s[0]) + g[1] * (__ldg(px + 1) - s[1]) + g[2] * (__ldg(px + 2) - s[2]);
        gu = fmaf(t.dwu[k], dot, gu);
        gv = fmaf(t.dwv[k], dot, gv);
        if (d_tex && t.w[k] != 0.f && scale != 0.f) {
            const float w = t.w[k] * inv * scale;
            float* q = d_tex + (size_t)t.idx[k] * 3;
            atomicAdd(q, w * g[0]); atomicAdd(q + 1, w * g[1]); atomicAdd(q + 2, w * g[2]);
        }
    }
    // d fu / d x = R / 2
    gx += gu * inv * scale * 0.5f * R;
    gy += gv * inv * scale * 0.5f * R;
}

__global__ void __launch_bounds__(256) cube_sample_bwd_kernel(const CubeStack st, const float* __restrict__ dirs,
                                                              const float* __restrict__ level, int64_t n, const float* __restrict__ g_out,
                                                              float* __restrict__ d_dirs, float* __restrict__ d_level) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float d[3] = {dirs[i * 3], dirs[i * 3 + 1], dirs[i * 3 + 2]};
    const float g[3] = {g_out[i * 3], g_out[i * 3 + 1], g_out[i * 3 + 2]};
    const FaceMap fm = cs_face(d);
    int l0, l1; float f; bool inside;
    cs_levels(level, i, st.n, l0, l1, f, inside);
    float gx = 0.f, gy = 0.f, s0[3], s1[3] = {0.f, 0.f, 0.f};
    cs_level_bwd(st.tex[l0], st.d_tex[l0], fm, st.res[l0], g, 1.f - f, s0, gx, gy);
    const bool two = l1 != l0;
    if (two) cs_level_bwd(st.tex[l1], st.d_tex[l1], fm, st.res[l1], g, f, s1, gx, gy);
    if (d_level) {
        // out = (1-f) S_l0 + [l1 != l0] f S_l1 (the formulation of the tensor version this kernel replaces)
        float gl = 0.f;
        if (level && st.n > 1 && inside)
            gl = g[0] * ((two ? s1[0] : 0.f) - s0[0]) + g[1] * ((two ? s1[1] : 0.f) - s0[1]) + g[2] * ((two ? s1[2] : 0.f) - s0[2]);
        d_level[i] = gl;
    }
    if (d_dirs) {
        // x = sx d[ax] / m, y = sy d[ay] / m, m = sm d[am]
        float dd[3] = {0.f, 0.f, 0.f};
        const bool live = fabsf(d[fm.am]) >= 1e-30f;          // clamp_min(1e-30) passes no gradient below the bound
        dd[fm.ax] += gx * fm.sx / fm.m;
        dd[fm.ay] += gy * fm.sy / fm.m;
        if (live) dd[fm.am] += -(gx * fm.x + gy * fm.y) * fm.sm / fm.m;
        d_dirs[i * 3] = dd[0]; d_dirs[i * 3 + 1] = dd[1]; d_dirs[i * 3 + 2] = dd[2];
    }
}

int cs_fill(CubeStack& st, const float* const* tex, float* const* d_tex, const int32_t* res, int32_t n_levels) {
    TF_REQUIRE(tex && res && n_levels >= 1 && n_levels <= CS_MAX_LEVELS, "cube sample: 1..%d levels required", CS_MAX_LEVELS);
    st.n = n_levels;
    for (int l = 0; l < CS_MAX_LEVELS; ++l) {
        st.tex[l] = l < n_levels ? tex[l] : nullptr;
        st.d_tex[l] = (l < n_levels && d_tex) ? d_tex[l] : nullptr;
        st.res[l] = l < n_levels ? res[l] : 1;
        if (l < n_levels) TF_REQUIRE(tex[l] && res[l] > 0, "cube sample: NULL texture / bad resolution at level %d", l);
    }
    return 0;
}

}  // namespace

extern "C" TF_API int tf_cube_sample_fwd(const float* const* tex, const int32_t* res, int32_t n_levels, const float* dirs,
                                         const float* level, int64_t n, float* out, tf_stream_t stream) {
    if (n == 0) return 0;
    TF_REQUIRE(dirs && out, "tf_cube_sample_fwd: NULL pointer");
    CubeStack st;
    if (int rc = cs_fill(st, tex, nullptr, res, n_levels)) return rc;
    cube_sample_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(st, dirs, level, n, out);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_cube_sample_fwd");
    return 0;
}

extern "C" TF_API int tf_cube_sample_bwd(const float* const* tex, const int32_t* res, int32_t n_levels, const float* dirs,
                                         const float* level, int64_t n, const float* g_out, float* const* d_tex, float* d_dirs,
                                         float* d_level, tf_stream_t stream) {
    if (n == 0) return 0;
    TF_REQUIRE(dirs && g_out, "tf_cube_sample_bwd: NULL pointer");
    CubeStack st;
    if (int rc = cs_fill(st, tex, d_tex, res, n_levels)) return rc;
    cube_sample_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(st, dirs, level, n, g_out, d_dirs, d_level);
    tf_count_launches(1);
    TF_CHECK_LAUNCH("tf_cube_sample_bwd");
    return 0;
}
