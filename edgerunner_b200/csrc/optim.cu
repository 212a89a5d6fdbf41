// Optimizer step of the reference's training loop (main.py:133 `torch.optim.AdamW(lr, weight_decay=0.01, betas=(0.9, 0.95))`, :175-177
// `accelerator.clip_grad_norm_(model.parameters(), opt.gradient_clip)`) over ONE flat fp32 buffer per state (parameters, gradients, exp_avg,
// exp_avg_sq) — the layout a flattened data-parallel gradient buffer gives for free.  HBM-bound streaming kernels: the step reads p, g, m, v and
// writes p, m, v (+ the fp16 copy the forward kernels consume): 30 bytes per parameter, 23 GB for the 766.8 M parameters of the ArAE preset.
// The gradient-clipping coefficient is applied to the gradient as it is read (torch scales the gradients in place first: same rounding, one pass
// over 3 GB less).  These entry points are the optimizer half of SURVEY §8 f2; the gradients come from er_train_step (train.cu).
#include "../../include/edgerunner_b200.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "kernels.h"

#define CK(call)                                                                                                          \
    do {                                                                                                                  \
        cudaError_t _e = (call);                                                                                          \
        if (_e != cudaSuccess) return er_set_error(ER_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)

namespace {

constexpr int NORM_BLOCKS = 592;        // 4 per SM

// partial[b] = sum of squares of block b's grid-stride share, accumulated in fp64 per thread and reduced in a fixed order
__global__ void __launch_bounds__(256) grad_sq_partial_kernel(const float* __restrict__ g, size_t n, double* __restrict__ partial) {
    double acc = 0.0;
    const size_t n4 = n >> 2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = reinterpret_cast<const float4*>(g)[i];
        acc += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float t = g[(n4 << 2) + threadIdx.x]; acc += (double)t * t; }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
// norm = sqrt(sum partial); scale = min(1, max_norm / (norm + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void grad_norm_finish_kernel(const double* __restrict__ partial, int nb, float max_norm, float* norm_out, float* scale_out) {
    __shared__ double red[1024];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nb; i += 1024) acc += partial[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) {
        const float norm = (float)sqrt(red[0]);
        *norm_out = norm;
        const float coef = max_norm / (norm + 1e-6f);
        *scale_out = max_norm > 0.f ? fminf(coef, 1.f) : 1.f;
    }
}

// torch.optim.AdamW (multi-tensor path), per element:
//   p *= 1 - lr * wd;  m = m + (1 - b1) (g - m);  v = v * b2 + ((1 - b2) g) g;  p -= (lr / bc1) * (m / (sqrt(v) / sqrt(bc2) + eps))
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                    __half* __restrict__ p16, size_t n, float decay, float omb1, float b2, float omb2, float step_size,
                                                    float bc2_sqrt, float eps, const float* __restrict__ gscale) {
    const float gs = gscale ? *gscale : 1.f;
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    float pp[4], gg[4], mm[4], vv[4];
    const bool vec = i + 4 <= n;
    if (vec) {
        const float4 a = *reinterpret_cast<const float4*>(p + i), b = *reinterpret_cast<const float4*>(g + i), c = *reinterpret_cast<const float4*>(m + i),
                     d = *reinterpret_cast<const float4*>(v + i);
        pp[0] = a.x; pp[1] = a.y; pp[2] = a.z; pp[3] = a.w; gg[0] = b.x; gg[1] = b.y; gg[2] = b.z; gg[3] = b.w;
        mm[0] = c.x; mm[1] = c.y; mm[2] = c.z; mm[3] = c.w; vv[0] = d.x; vv[1] = d.y; vv[2] = d.z; vv[3] = d.w;
    } else {
        for (int k = 0; k < 4; k++) { const bool in = i + k < n; pp[k] = in ? p[i + k] : 0.f; gg[k] = in ? g[i + k] : 0.f; mm[k] = in ? m[i + k] : 0.f; vv[k] = in ? v[i + k] : 0.f; }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float gk = gscale ? __fmul_rn(gg[k], gs) : gg[k];
        pp[k] = __fmul_rn(pp[k], decay);
        mm[k] = fmaf(omb1, gk - mm[k], mm[k]);
        vv[k] = fmaf(__fmul_rn(omb2, gk), gk, __fmul_rn(vv[k], b2));
        const float denom = __fadd_rn(__fdiv_rn(sqrtf(vv[k]), bc2_sqrt), eps);
        pp[k] = fmaf(-step_size, __fdiv_rn(mm[k], denom), pp[k]);
    }
    if (vec) {
        *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
        *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
        *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        if (p16) {
            __align__(8) __half h[4] = {__float2half_rn(pp[0]), __float2half_rn(pp[1]), __float2half_rn(pp[2]), __float2half_rn(pp[3])};
            *reinterpret_cast<uint2*>(p16 + i) = *reinterpret_cast<const uint2*>(h);
        }
    } else {
        for (int k = 0; k < 4 && i + k < n; k++) { p[i + k] = pp[k]; m[i + k] = mm[k]; v[i + k] = vv[k]; if (p16) p16[i + k] = __float2half_rn(pp[k]); }
    }
}

}  // namespace

extern "C" int er_grad_norm_clip(const float* grad_dev, int64_t n, float max_norm, double* scratch_dev, float* norm_out_dev, float* scale_out_dev, void* stream) {
    if (!grad_dev || n <= 0 || !scratch_dev || !norm_out_dev || !scale_out_dev) return er_set_error(ER_ERR_INVALID, "bad argument");
    if ((uintptr_t)grad_dev & 15) return er_set_error(ER_ERR_INVALID, "gradient buffer must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    grad_sq_partial_kernel<<<NORM_BLOCKS, 256, 0, st>>>(grad_dev, (size_t)n, scratch_dev);
    CK(cudaGetLastError());
    grad_norm_finish_kernel<<<1, 1024, 0, st>>>(scratch_dev, NORM_BLOCKS, max_norm, norm_out_dev, scale_out_dev);
    CK(cudaGetLastError());
    return ER_OK;
}

extern "C" int er_adamw_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, void* param16_out_dev, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int32_t step, const float* grad_scale_dev, void* stream) {
    if (!param_dev || !grad_dev || !exp_avg_dev || !exp_avg_sq_dev || n <= 0 || step < 1) return er_set_error(ER_ERR_INVALID, "bad argument");
    if (((uintptr_t)param_dev | (uintptr_t)grad_dev | (uintptr_t)exp_avg_dev | (uintptr_t)exp_avg_sq_dev) & 15 || ((uintptr_t)param16_out_dev & 7))
        return er_set_error(ER_ERR_INVALID, "optimizer buffers must be 16-byte aligned");
    // the scalars torch computes as Python floats (double) and hands to its kernels as fp32
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    const float decay = (float)(1.0 - (double)lr * (double)weight_decay), step_size = (float)((double)lr / bc1), bc2_sqrt = (float)sqrt(bc2);
    const size_t threads = ((size_t)n + 3) / 4;
    adamw_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(param_dev, grad_dev, exp_avg_dev, exp_avg_sq_dev, (__half*)param16_out_dev, (size_t)n,
                                                                                      decay, (float)(1.0 - (double)beta1), beta2, (float)(1.0 - (double)beta2),
                                                                                      step_size, bc2_sqrt, eps, grad_scale_dev);
    CK(cudaGetLastError());
    return ER_OK;
}
