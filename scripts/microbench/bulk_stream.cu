// Microbenchmark: per-CTA HBM->smem streaming rate with cp.async.bulk + mbarrier ring (1 CTA/SM), vs plain LDG.128 streaming.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_addr(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_addr(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_addr(b)), "r"(n) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s_addr(b)), "r"(par) : "memory");
    return ok;
}
__device__ __forceinline__ void bulk(void* d, const void* s, uint32_t n, uint64_t* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s_addr(d)), "l"(s), "r"(n), "r"(s_addr(b)) : "memory");
}
// each CTA streams `bytes_per_cta` contiguous bytes starting at base + cta*bytes_per_cta; consumers touch `touch` fraction
__global__ void __launch_bounds__(544, 1) ring_kernel(const char* base, size_t bytes_per_cta, int stage_bytes, int nstage, int split, int touch, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = (uint64_t*)(smem + (size_t)nstage * stage_bytes);
    uint64_t* empty = full + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { for (int i = 0; i < nstage; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 16); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const char* src = base + (size_t)blockIdx.x * bytes_per_cta;
    const uint32_t n = (uint32_t)(bytes_per_cta / stage_bytes);
    if (warp == 16) {
        if (lane == 0) {
            for (uint32_t it = 0; it < n; ++it) {
                const uint32_t s = it % nstage, par = (it / nstage) & 1;
                while (!mbar_try(&empty[s], par ^ 1)) {}
                mbar_expect(&full[s], stage_bytes);
                const int piece = stage_bytes / split;
                for (int k = 0; k < split; k++) bulk(smem + (size_t)s * stage_bytes + k * piece, src + (size_t)it * stage_bytes + k * piece, piece, &full[s]);
            }
        }
    } else {
        float acc = 0.f;
        for (uint32_t it = 0; it < n; ++it) {
            const uint32_t s = it % nstage, par = (it / nstage) & 1;
            while (!mbar_try(&full[s], par)) {}
            if (touch) {
                const uint4* p = (const uint4*)(smem + (size_t)s * stage_bytes);
                for (int v = tid; v < stage_bytes / 16; v += 512) { uint4 x = p[v]; acc += __uint_as_float(x.x ^ x.y ^ x.z ^ x.w); }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 123.456f) sink[0] = acc;
    }
    __syncthreads();
}
__global__ void __launch_bounds__(512, 1) ldg_kernel(const char* base, size_t bytes_per_cta, int unroll, float* sink) {
    const uint4* src = (const uint4*)(base + (size_t)blockIdx.x * bytes_per_cta);
    const size_t nv = bytes_per_cta / 16;
    float acc = 0.f;
    for (size_t v = threadIdx.x; v + 512 * 7 < nv; v += 512 * 8) {
        uint4 x[8];
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x[j].x), "=r"(x[j].y), "=r"(x[j].z), "=r"(x[j].w) : "l"(src + v + j * 512));
#pragma unroll
        for (int j = 0; j < 8; j++) acc += __uint_as_float(x[j].x ^ x[j].y ^ x[j].z ^ x[j].w);
    }
    if (acc == 123.456f) sink[0] = acc;
}
int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t per_cta = 96ull << 20;           // 96 MB per CTA -> 14 GB total (>> L2)
    char* buf; CK(cudaMalloc(&buf, per_cta * sms)); CK(cudaMemset(buf, 1, per_cta * sms));
    float* sink; CK(cudaMalloc(&sink, 4));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    CK(cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    int cfgs[][4] = {{24576, 7, 1, 1}, {24576, 7, 1, 0}, {24576, 4, 1, 1}, {16384, 12, 1, 1}, {8192, 24, 1, 1}, {32768, 6, 1, 1}, {24576, 7, 4, 1}, {24576, 7, 12, 1}, {4096, 48, 1, 1}};
    for (auto& c : cfgs) {
        size_t smem = (size_t)c[0] * c[1] + 512;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(a);
            ring_kernel<<<sms, 544, smem>>>(buf, per_cta, c[0], c[1], c[2], c[3], sink);
            cudaEventRecord(b); CK(cudaEventSynchronize(b));
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("ring stage=%d nstage=%d split=%d touch=%d : %.1f GB/s\n", c[0], c[1], c[2], c[3], per_cta * sms / ms / 1e6);
    }
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(a); ldg_kernel<<<sms, 512>>>(buf, per_cta, 8, sink); cudaEventRecord(b); CK(cudaEventSynchronize(b)); }
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("ldg.128 x8 unroll, 512 thr/SM : %.1f GB/s\n", per_cta * sms / ms / 1e6);
    return 0;
}
