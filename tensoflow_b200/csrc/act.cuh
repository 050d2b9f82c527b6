// Activations of the small MLPs (shared by the FFMA and the tensor-core dense-layer kernels).
#pragma once
#include "common.cuh"

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2, ACT_SOFTPLUS100 = 3, ACT_SIGMOID = 4, ACT_EXP = 5 };

__device__ __forceinline__ float act_fwd(float x, int act, float p) {
    switch (act) {
        case ACT_RELU: return fmaxf(x, 0.f);
        case ACT_LEAKY: return x > 0.f ? x : 0.01f * x;
        case ACT_SOFTPLUS100: return softplus100(x);
        case ACT_SIGMOID: return 1.f / (1.f + expf(-x));
        case ACT_EXP: return expf(fminf(x, p));          // ExpActivation (other_field.py:12-18)
        default: return x;
    }
}
// derivative expressed through the OUTPUT y (so only Y has to be kept for backward)
__device__ __forceinline__ float act_bwd(float y, int act, float p) {
    switch (act) {
        case ACT_RELU: return y > 0.f ? 1.f : 0.f;
        case ACT_LEAKY: return y > 0.f ? 1.f : 0.01f;
        case ACT_SOFTPLUS100: return 1.f - expf(-100.f * y);
        case ACT_SIGMOID: return y * (1.f - y);
        case ACT_EXP: return y < expf(p) ? y : 0.f;
        default: return 1.f;
    }
}
