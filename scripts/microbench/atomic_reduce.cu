// Microbenchmark for the round-2 question "can a cross-CTA K-split reduction replace a vector exchange?":
// 148 CTAs (1 per SM) each add a 1536-element partial vector into ONE shared accumulator with L2 reductions, then meet at a grid
// barrier (release-red + acquire poll, the decode kernel's barrier) and read the 1536 sums back.  Reported: microseconds per
// round for several encodings of the partial, next to the barrier + read-back alone.
//   V0  barrier + read-back only (baseline)
//   V1  1536 x red.add.u64      (fixed point, 2^-32 resolution: order-independent, bit-reproducible)
//   V2   768 x red.add.u64      (two biased 32-bit fixed-point fields per word: no carry between fields for <= 148 addends)
//   V3  1536 x red.add.u32      (one biased 32-bit fixed-point field per word)
//   V4   384 x red.add.v4.f32   (fp32 vector reduction: fewer L2 operations, summation order NOT reproducible)
//   V5  1536 x red.add.f32
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomic_reduce atomic_reduce.cu ; run: ./atomic_reduce [rounds]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int kN = 1536;
constexpr int kThreads = 256;

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += 1;
        const unsigned target = epoch * gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory"); } while (v < target);
    }
    __syncthreads();
}

template <int V>
__global__ void __launch_bounds__(kThreads, 1) reduce_kernel(unsigned long long* acc64, unsigned* acc32, float* accf, unsigned* counter, int rounds,
                                                             unsigned long long* t_out, float* sink) {
    const int tid = threadIdx.x;
    unsigned epoch = 0;
    float keep = 0.f;
    grid_barrier(counter, epoch);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int r = 0; r < rounds; ++r) {
        // double-buffered accumulators (round parity): nobody zeroes inside the timed loop, the values just keep growing
        const int buf = (r & 1) * kN;
        const float x = 1.0f + 0.001f * (float)((tid + blockIdx.x + r) & 63);
        if (V == 1) {
            for (int i = tid; i < kN; i += kThreads) {
                const long long fx = (long long)((double)x * 4294967296.0);
                asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(acc64 + buf + i), "l"(fx) : "memory");
            }
        } else if (V == 2) {
            for (int i = tid; i < kN / 2; i += kThreads) {
                const unsigned lo = (unsigned)(int)(x * 65536.0f) + (1u << 24), hi = (unsigned)(int)(-x * 65536.0f) + (1u << 24);
                const unsigned long long w = ((unsigned long long)hi << 32) | lo;
                asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(acc64 + buf + i), "l"(w) : "memory");
            }
        } else if (V == 3) {
            for (int i = tid; i < kN; i += kThreads) {
                const unsigned w = (unsigned)(int)(x * 65536.0f) + (1u << 24);
                asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(acc32 + buf + i), "r"(w) : "memory");
            }
        } else if (V == 4) {
            for (int i = tid; i < kN / 4; i += kThreads)
                asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(accf + buf + 4 * i), "f"(x), "f"(x), "f"(x), "f"(x) : "memory");
        } else if (V == 5) {
            for (int i = tid; i < kN; i += kThreads)
                asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(accf + buf + i), "f"(x) : "memory");
        }
        grid_barrier(counter, epoch);
        // read the whole reduced vector back (what the LayerNorm after the phase needs): 6 elements per thread
        for (int i = tid; i < kN; i += kThreads) {
            if (V == 1 || V == 2) { unsigned long long v; asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(acc64 + buf + (V == 2 ? i / 2 : i))); keep += (float)(v & 0xffff); }
            else if (V == 3) { unsigned v; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(acc32 + buf + i)); keep += (float)(v & 0xffff); }
            else { float v; asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(accf + buf + i)); keep += v; }
        }
    }
    unsigned long long t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    if (tid == 0) t_out[blockIdx.x] = t1 - t0;
    if (keep == 123.456f) sink[0] = keep;
}

template <int V>
static int run(const char* name, int rounds, int sms, void** bufs) {
    unsigned long long* acc64 = (unsigned long long*)bufs[0]; unsigned* acc32 = (unsigned*)bufs[1]; float* accf = (float*)bufs[2];
    unsigned* counter = (unsigned*)bufs[3]; unsigned long long* t_out = (unsigned long long*)bufs[4]; float* sink = (float*)bufs[5];
    CK(cudaMemset(acc64, 0, 2 * kN * 8)); CK(cudaMemset(acc32, 0, 2 * kN * 4)); CK(cudaMemset(accf, 0, 2 * kN * 4)); CK(cudaMemset(counter, 0, 4));
    void* args[] = {&acc64, &acc32, &accf, &counter, &rounds, &t_out, &sink};
    CK(cudaLaunchCooperativeKernel((const void*)reduce_kernel<V>, dim3(sms), dim3(kThreads), args, 0, 0));
    CK(cudaDeviceSynchronize());
    unsigned long long h[256];
    CK(cudaMemcpy(h, t_out, sms * 8, cudaMemcpyDeviceToHost));
    unsigned long long mx = 0;
    for (int i = 0; i < sms; i++) mx = h[i] > mx ? h[i] : mx;
    printf("%-44s %8.3f us per round\n", name, (double)mx / 1e3 / rounds);
    return 0;
}

int main(int argc, char** argv) {
    const int rounds = argc > 1 ? atoi(argv[1]) : 2000;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    void* bufs[6];
    CK(cudaMalloc(&bufs[0], 2 * kN * 8)); CK(cudaMalloc(&bufs[1], 2 * kN * 4)); CK(cudaMalloc(&bufs[2], 2 * kN * 4));
    CK(cudaMalloc(&bufs[3], 256)); CK(cudaMalloc(&bufs[4], 256 * 8)); CK(cudaMalloc(&bufs[5], 256));
    printf("%d CTAs x %d threads, %d-element partial per CTA, %d rounds\n", sms, kThreads, kN, rounds);
    if (run<0>("V0 barrier + read-back only", rounds, sms, bufs)) return 1;
    if (run<1>("V1 1536 x red.add.u64 (fixed point)", rounds, sms, bufs)) return 1;
    if (run<2>("V2  768 x red.add.u64 (2 packed fields)", rounds, sms, bufs)) return 1;
    if (run<3>("V3 1536 x red.add.u32 (biased fixed point)", rounds, sms, bufs)) return 1;
    if (run<4>("V4  384 x red.add.v4.f32", rounds, sms, bufs)) return 1;
    if (run<5>("V5 1536 x red.add.f32", rounds, sms, bufs)) return 1;
    return 0;
}
