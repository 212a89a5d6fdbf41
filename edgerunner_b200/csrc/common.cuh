// Shared device helpers for the edgerunner_b200 sm_100a kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define ER_WARP 32

namespace er {

// ---- memory-path helpers --------------------------------------------------------------------------
// Weights: read once per token, never written by a kernel -> non-coherent streaming load, no L1 allocate.
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// Data produced by other CTAs inside the same launch (KV cache rows, activations): coherent at L2.
__device__ __forceinline__ uint4 ldg_cg(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_cg_f32(const float* p) {
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned short ldg_cg_u16(const void* p) {
    unsigned short r;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}
// TMA-unit bulk prefetch of a contiguous byte range into L2 (SASS: UBLKPF).  16-byte aligned / sized.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l2_range(const void* p, size_t bytes) {
    const char* c = reinterpret_cast<const char*>(p);
    const size_t CH = 32768;
    for (size_t o = 0; o < bytes; o += CH) {
        size_t n = bytes - o < CH ? bytes - o : CH;
        prefetch_l2_bulk(c + o, (uint32_t)(n & ~size_t(15)));
    }
}

// ---- math helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ float2 h2f2(uint32_t u) {
    __half2 h = *reinterpret_cast<__half2*>(&u);
    return __half22float2(h);
}
// acc += sum_i w[i] * x[i] over 8 fp16 pairs, fp32 products and accumulation
__device__ __forceinline__ float dot8(const uint4& w, const uint4& x, float acc) {
    float2 a, b;
    a = h2f2(w.x); b = h2f2(x.x); acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
    a = h2f2(w.y); b = h2f2(x.y); acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
    a = h2f2(w.z); b = h2f2(x.z); acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
    a = h2f2(w.w); b = h2f2(x.w); acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
    return acc;
}
__device__ __forceinline__ float round_f16(float x) { return __half2float(__float2half_rn(x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace er
