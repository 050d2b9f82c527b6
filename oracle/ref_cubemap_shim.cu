// TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): C-ABI launchers around the REFERENCE's own cubemap prefilter kernels.
//
// The kernels are compiled from the reference sources where they lie (/root/reference/network/renderutils/c_src/cubemap.cu,
// common.cpp, *.h -- plain CUDA, no torch) by oracle/build_ref.py (this file + the reference's common.cpp as a second translation unit) into oracle/_ref/libref_cubemap.so; nothing of the
// reference is copied into this repository.  The launchers below restate only the parameter set-up of the reference's torch
// binding (c_src/torch_bindings.cpp:740-889: grid = (res, res, 6), 8x8 blocks, contiguous NHWC tensors) so that
// oracle/gen_golden_prefilter.py can run diffuse_cubemap / specular_cubemap (network/renderutils/ops.py:391-458) on the GPU
// box and store their outputs + gradients as golden fixtures for tensoflow_b200's prefilter operator.
#include <cuda_runtime.h>
#include <string.h>
#include "cubemap.cu"

namespace {

Tensor make_tensor(void* val, int n, int h, int w, int c, dim3 grid) {
    Tensor t;
    memset(&t, 0, sizeof(t));
    t.val = val;
    t.d_val = nullptr;
    t.dims[0] = n; t.dims[1] = h; t.dims[2] = w; t.dims[3] = c;
    t.strides[0] = h * w * c; t.strides[1] = w * c; t.strides[2] = c; t.strides[3] = 1;
    t._dims[0] = grid.z; t._dims[1] = grid.y; t._dims[2] = grid.x; t._dims[3] = c;
    t.fp16 = false;
    return t;
}

template <class P>
int launch(const void* kernel, P& p) {
    dim3 block = getLaunchBlockSize(8, 8, p.gridSize);      // BLOCK_X, BLOCK_Y of torch_bindings.cpp:40-41
    dim3 grid = getLaunchGridSize(block, p.gridSize);
    void* args[] = {&p};
    cudaError_t e = cudaLaunchKernel(kernel, grid, block, args, 0, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    return (int)e;
}

}  // namespace

#define REF_API extern "C" __attribute__((visibility("default")))

// out[6,res,res,3] = diffuse_cubemap_fwd(cubemap[6,res,res,3])            (torch_bindings.cpp:740-767)
REF_API int ref_diffuse_cubemap_fwd(float* cubemap, int res, float* out) {
    DiffuseCubemapKernelParams p;
    memset(&p, 0, sizeof(p));
    p.gridSize = dim3(res, res, 6);
    p.cubemap = make_tensor(cubemap, 6, res, res, 3, p.gridSize);
    p.out = make_tensor(out, 6, res, res, 3, p.gridSize);
    return launch((const void*)DiffuseCubemapFwdKernel, p);
}

// cubemap_grad[6,res,res,3] (zero-initialised by the caller) = diffuse_cubemap_bwd(cubemap, grad)   (:769-797)
REF_API int ref_diffuse_cubemap_bwd(float* cubemap, float* grad, int res, float* cubemap_grad) {
    DiffuseCubemapKernelParams p;
    memset(&p, 0, sizeof(p));
    p.gridSize = dim3(res, res, 6);
    p.cubemap = make_tensor(cubemap, 6, res, res, 3, p.gridSize);
    p.out = make_tensor(grad, 6, res, res, 3, p.gridSize);
    p.cubemap.d_val = cubemap_grad;
    return launch((const void*)DiffuseCubemapBwdKernel, p);
}

// bounds[6,res,res,24] (zero-initialised by the caller) = specular_bounds(res, costheta_cutoff)       (:799-824)
REF_API int ref_specular_bounds(int res, float costheta_cutoff, float* bounds) {
    SpecularBoundsKernelParams p;
    memset(&p, 0, sizeof(p));
    p.costheta_cutoff = costheta_cutoff;
    p.gridSize = dim3(res, res, 6);
    p.out = make_tensor(bounds, 6, res, res, 24, p.gridSize);
    return launch((const void*)SpecularBoundsKernel, p);
}

// out[6,res,res,4] = specular_cubemap_fwd(cubemap, bounds, roughness, costheta_cutoff)                (:826-857)
REF_API int ref_specular_cubemap_fwd(float* cubemap, float* bounds, float roughness, float costheta_cutoff, int res, float* out) {
    SpecularCubemapKernelParams p;
    memset(&p, 0, sizeof(p));
    p.roughness = roughness;
    p.costheta_cutoff = costheta_cutoff;
    p.gridSize = dim3(res, res, 6);
    p.cubemap = make_tensor(cubemap, 6, res, res, 3, p.gridSize);
    p.bounds = make_tensor(bounds, 6, res, res, 24, p.gridSize);
    p.out = make_tensor(out, 6, res, res, 4, p.gridSize);
    return launch((const void*)SpecularCubemapFwdKernel, p);
}

// cubemap_grad[6,res,res,3] (zero-initialised) = specular_cubemap_bwd(cubemap, bounds, grad[6,res,res,4], ...)   (:859-890)
REF_API int ref_specular_cubemap_bwd(float* cubemap, float* bounds, float* grad, float roughness, float costheta_cutoff, int res,
                                     float* cubemap_grad) {
    SpecularCubemapKernelParams p;
    memset(&p, 0, sizeof(p));
    p.roughness = roughness;
    p.costheta_cutoff = costheta_cutoff;
    p.gridSize = dim3(res, res, 6);
    p.cubemap = make_tensor(cubemap, 6, res, res, 3, p.gridSize);
    p.bounds = make_tensor(bounds, 6, res, res, 24, p.gridSize);
    p.out = make_tensor(grad, 6, res, res, 4, p.gridSize);
    p.cubemap.d_val = cubemap_grad;
    return launch((const void*)SpecularCubemapBwdKernel, p);
}
