// Microbenchmark: latency of grid-barrier variants on 148 CTAs (1 per SM), with a small publish/consume around each barrier.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acq(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_rlx(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_vol(const unsigned* p) { unsigned v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
template <int V>
__device__ __forceinline__ void gbar(unsigned* counter, unsigned* flags, unsigned& epoch) {
    const int tid = threadIdx.x;
    if (V == 0) {          // release-red + acquire-load poll by thread 0
        cbar();
        if (tid == 0) {
            epoch++;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            while (ld_acq(counter) < epoch * gridDim.x) {}
        }
        cbar();
    } else if (V == 1) {   // classic: threadfence + relaxed atomic + volatile poll + threadfence
        cbar();
        if (tid == 0) {
            epoch++;
            __threadfence();
            atomicAdd(counter, 1);
            while (ld_vol(counter) < epoch * gridDim.x) {}
            __threadfence();
        }
        cbar();
    } else if (V == 2) {   // release-red + relaxed poll + one acquire fence
        cbar();
        if (tid == 0) {
            epoch++;
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
            while (ld_rlx(counter) < epoch * gridDim.x) {}
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        cbar();
    } else if (V == 3) {   // per-CTA flags, polled in parallel by 148 threads (no atomics)
        cbar();
        epoch++;
        if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
        if (tid < gridDim.x) { while (ld_acq(flags + tid) < epoch) {} }
        cbar();
    } else if (V == 4) {   // per-CTA flags, relaxed polls + fence
        cbar();
        epoch++;
        if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(epoch) : "memory");
        if (tid < gridDim.x) { while (ld_rlx(flags + tid) < epoch) {} asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
        cbar();
    } else if (V == 5) {   // 32-wide flag lines: flags padded to 128 B each
        cbar();
        epoch++;
        if (tid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x * 32), "r"(epoch) : "memory");
        if (tid < gridDim.x) { while (ld_rlx(flags + tid * 32) < epoch) {} asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
        cbar();
    }
}
template <int V>
__global__ void __launch_bounds__(288, 1) bench_kernel(unsigned* counter, unsigned* flags, float* data, int rounds, int publish, float* sink) {
    unsigned epoch = 0;
    float acc = 0.f;
    if (threadIdx.x >= 256) return;   // mimic the producer warp being elsewhere
    for (int r = 0; r < rounds; r++) {
        if (publish && threadIdx.x < 16) data[blockIdx.x * 16 + threadIdx.x] = (float)r;      // phase output (like y rows)
        gbar<V>(counter, flags, epoch);
        if (publish) { float v; asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(data + ((blockIdx.x * 37 + threadIdx.x) % (gridDim.x * 16)))); acc += v; }
    }
    if (acc == 123.f) sink[0] = acc;
}
template <int V>
int run(const char* name, unsigned* counter, unsigned* flags, float* data, float* sink, int sms) {
    const int rounds = 20000;
    for (int publish = 0; publish < 2; publish++) {
        CK(cudaMemset(counter, 0, 4)); CK(cudaMemset(flags, 0, 148 * 128));
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        void* args[] = {&counter, &flags, &data, (void*)&rounds, &publish, &sink};
        cudaEventRecord(a);
        CK(cudaLaunchCooperativeKernel((void*)bench_kernel<V>, dim3(sms), dim3(288), args, 0, 0));
        cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("%-58s publish=%d : %.3f us per barrier round\n", name, publish, ms * 1e3 / rounds);
    }
    return 0;
}
int main() {
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned *counter, *flags; float *data, *sink;
    CK(cudaMalloc(&counter, 4)); CK(cudaMalloc(&flags, 148 * 128)); CK(cudaMalloc(&data, 148 * 16 * 4)); CK(cudaMalloc(&sink, 4));
    run<0>("V0 red.release + ld.acquire poll (current)", counter, flags, data, sink, sms);
    run<1>("V1 threadfence + atomicAdd + volatile poll + threadfence", counter, flags, data, sink, sms);
    run<2>("V2 red.release + ld.relaxed poll + fence", counter, flags, data, sink, sms);
    run<3>("V3 per-CTA flags, 148 parallel ld.acquire polls", counter, flags, data, sink, sms);
    run<4>("V4 per-CTA flags, parallel ld.relaxed polls + fence", counter, flags, data, sink, sms);
    run<5>("V5 per-CTA flags padded to 128 B, relaxed polls + fence", counter, flags, data, sink, sms);
    return 0;
}
