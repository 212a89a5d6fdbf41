// Microbenchmark: does a TMA-fed ring prefetch during consumer idle time?  Consumers idle `idle_ns`, then drain `burst` stages.
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
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
constexpr int STAGE = 24576;
__global__ void __launch_bounds__(288, 1) burst_kernel(const char* base, size_t bytes_per_cta, int nstage, int burst, int idle_ns, int rounds, int use_gbar, unsigned* gcount, unsigned long long* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = (uint64_t*)(smem + (size_t)nstage * STAGE);
    uint64_t* empty = full + 16;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { for (int i = 0; i < nstage; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 8); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const char* src = base + (size_t)blockIdx.x * bytes_per_cta;
    const uint32_t n = (uint32_t)(rounds * burst);
    if (warp == 8) {
        if (lane == 0) {
            for (uint32_t it = 0; it < n; ++it) {
                const uint32_t s = it % nstage, par = (it / nstage) & 1;
                while (!mbar_try(&empty[s], par ^ 1)) {}
                mbar_expect(&full[s], STAGE);
                bulk(smem + (size_t)s * STAGE, src + (size_t)it * STAGE, STAGE, &full[s]);
            }
        }
    } else {
        unsigned long long drain = 0, firstwait = 0;
        uint32_t it = 0;
        unsigned epoch = 0;
        for (int r = 0; r < rounds; r++) {
            // idle period: either a timed spin or a real grid barrier (+ timed spin)
            if (use_gbar) {
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tid == 0) {
                    epoch++;
                    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(gcount) : "memory");
                    unsigned v;
                    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(gcount) : "memory"); } while (v < epoch * gridDim.x);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const unsigned long long t0 = gtime();
            while (gtime() - t0 < (unsigned long long)idle_ns) {}
            const unsigned long long t1 = gtime();
            for (int b = 0; b < burst; b++, it++) {
                const uint32_t s = it % nstage, par = (it / nstage) & 1;
                while (!mbar_try(&full[s], par)) {}
                if (b == 0 && tid == 0) firstwait += gtime() - t1;
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[s]);
            }
            if (tid == 0) drain += gtime() - t1;
        }
        if (tid == 0) { out[blockIdx.x * 2] = drain / rounds; out[blockIdx.x * 2 + 1] = firstwait / rounds; }
    }
    __syncthreads();
}
int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int rounds = 400;
    const size_t per_cta = (size_t)rounds * 8 * STAGE;
    char* buf; CK(cudaMalloc(&buf, per_cta * sms)); CK(cudaMemset(buf, 1, per_cta * sms));
    unsigned long long* out; CK(cudaMalloc(&out, sms * 16));
    unsigned* gc; CK(cudaMalloc(&gc, 4));
    CK(cudaFuncSetAttribute(burst_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int cfgs[][4] = {{7, 4, 0, 0}, {7, 4, 2000, 0}, {7, 4, 5000, 0}, {7, 6, 5000, 0}, {7, 6, 10000, 0}, {7, 4, 0, 1}, {7, 4, 3000, 1}, {4, 4, 5000, 0}};
    for (auto& c : cfgs) {
        CK(cudaMemset(gc, 0, 4));
        size_t smem = (size_t)c[0] * STAGE + 512;
        void* args[] = {&buf, (void*)&per_cta, &c[0], &c[1], &c[2], (void*)&rounds, &c[3], &gc, &out};
        CK(cudaLaunchCooperativeKernel((void*)burst_kernel, dim3(sms), dim3(288), args, smem, 0));
        CK(cudaDeviceSynchronize());
        unsigned long long h[2 * 148];
        CK(cudaMemcpy(h, out, sms * 16, cudaMemcpyDeviceToHost));
        double d = 0, f = 0; for (int i = 0; i < sms; i++) { d += h[2 * i]; f += h[2 * i + 1]; }
        printf("nstage=%d burst=%d idle_ns=%d gbar=%d : drain %.2f us (first-stage wait %.2f us) per burst of %d KB\n", c[0], c[1], c[2], c[3], d / sms / 1e3, f / sms / 1e3, c[1] * 24);
    }
    return 0;
}
