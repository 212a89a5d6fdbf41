// Round-2 starting point for the dense path (prefill / encoder / teacher-forced forward): out[M][N] = A[M][K] . W[N][K]^T (+ bias), fp16 in,
// fp32 accumulate in TENSOR MEMORY, written for sm_100a with tcgen05.mma + TMA — the replacement of the mma.sync kernel in csrc/gemm.cu.
// NOT validated on hardware yet (written after round 1 ran out of GPU minutes); self-checking: main() compares with a plain GPU
// reference and prints max |err| and TFLOP/s.
//
// One CTA per 128 x 128 output tile, 192 threads:
//   warp 0      TMA producer: per K block of 64 elements two 2-D bulk tensor copies (A tile 128 x 64, W tile 128 x 64, 128-byte swizzle)
//               into a 4-stage shared-memory ring, completion on an mbarrier (complete_tx)
//   warp 1      MMA issuer (one elected lane): 4 x tcgen05.mma.cta_group::1.kind::f16 (M 128, N 128, K 16) per stage, operands through
//               shared-memory matrix descriptors (K-major, SWIZZLE_128B, stride 1024 B between 8-row groups), accumulator = 128 TMEM columns;
//               tcgen05.commit releases the stage / signals the epilogue
//   warps 2..5  epilogue: warp w reads TMEM lanes 32 (w % 4) .. +32 with tcgen05.ld.32x32b (one output row per thread), adds the bias,
//               converts to fp16 and stores
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gemm_tcgen05 gemm_tcgen05.cu -lcuda ; run: ./gemm_tcgen05 [M N K]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

constexpr int BM = 128, BN = 128, BK = 64, UK = 16, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;   // 16 KB + 16 KB
constexpr int THREADS = 192;
constexpr uint32_t TMEM_COLS = 128;

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major operand, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);            // start address, bits [0,14)
    d |= (uint64_t)0 << 16;                            // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B fp16, both K-major, N = 128, M = 128
__device__ __forceinline__ uint32_t instr_desc() {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {       // arrives on the mbarrier when all MMAs issued so far have completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const __half* __restrict__ bias,
                    __half* __restrict__ out, int M, int N, int K) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int nk = K / BK;
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(s_addr(&full_bar[i]), 1); mbar_init(s_addr(&empty_bar[i]), 1); }
        mbar_init(s_addr(&acc_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {       // one warp allocates the accumulator columns and publishes the TMEM base address
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {                                          // ===== TMA producer =====
            for (int kb = 0; kb < nk; ++kb) {
                const int st = kb % STAGES;
                if (kb >= STAGES) mbar_wait(s_addr(&empty_bar[st]), ((kb / STAGES) - 1) & 1);
                const uint32_t fb = s_addr(&full_bar[st]);
                mbar_expect_tx(fb, STAGE_BYTES);
                tma_load_2d(s_addr(tiles + st * STAGE_BYTES), &map_a, kb * BK, m0, fb);
                tma_load_2d(s_addr(tiles + st * STAGE_BYTES + A_BYTES), &map_w, kb * BK, n0, fb);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                          // ===== MMA issuer =====
            const uint32_t idesc = instr_desc();
            for (int kb = 0; kb < nk; ++kb) {
                const int st = kb % STAGES;
                mbar_wait(s_addr(&full_bar[st]), (kb / STAGES) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da = smem_desc(s_addr(tiles + st * STAGE_BYTES));
                const uint64_t db = smem_desc(s_addr(tiles + st * STAGE_BYTES + A_BYTES));
#pragma unroll
                for (int k = 0; k < BK / UK; ++k)                  // + 32 bytes (2 x 16-byte units) of K per step inside the swizzle atom
                    umma_f16(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                umma_commit(s_addr(&empty_bar[st]));              // stage reusable once these MMAs have read it
            }
            umma_commit(s_addr(&acc_bar));                        // accumulator complete
        }
    } else {                                                      // ===== epilogue (warps 2..5) =====
        const int quad = warp & 3;                                // TMEM lane quadrant this warp may access
        mbar_wait(s_addr(&acc_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + quad * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < M) {
                __align__(16) __half h[16];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int col = n0 + c0 + j;
                    const float b = (bias != nullptr && col < N) ? __half2float(bias[col]) : 0.f;
                    h[j] = __float2half_rn(__uint_as_float(v[j]) + b);
                }
                if (n0 + c0 + 16 <= N) {
                    *reinterpret_cast<uint4*>(out + (size_t)row * N + n0 + c0) = *reinterpret_cast<const uint4*>(h);
                    *reinterpret_cast<uint4*>(out + (size_t)row * N + n0 + c0 + 8) = *reinterpret_cast<const uint4*>(h + 8);
                } else {
                    for (int j = 0; j < 16 && n0 + c0 + j < N; j++) out[(size_t)row * N + n0 + c0 + j] = h[j];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

__global__ void ref_gemm_kernel(const __half* A, const __half* W, const __half* bias, float* out, int M, int N, int K) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
    if (n >= N || m >= M) return;
    float acc = 0.f;
    for (int k = 0; k < K; k++) acc += __half2float(A[(size_t)m * K + k]) * __half2float(W[(size_t)n * K + k]);
    out[(size_t)m * N + n] = acc + (bias ? __half2float(bias[n]) : 0.f);
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeTiled enc, CUtensorMap* map, const __half* base, int rows, int K) {   // [rows][K] fp16, box 64 x 128, 128 B swizzle
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    const cuuint32_t elem[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }
    return 0;
}

int main(int argc, char** argv) {
    const int M = argc > 3 ? atoi(argv[1]) : 2048, N = argc > 3 ? atoi(argv[2]) : 4608, K = argc > 3 ? atoi(argv[3]) : 1536;
    if (K % BK) { printf("K must be a multiple of %d\n", BK); return 1; }
    std::vector<__half> hA((size_t)M * K), hW((size_t)N * K), hb(N);
    srand(1);
    for (auto& x : hA) x = __float2half((rand() % 2001 - 1000) / 1000.0f);
    for (auto& x : hW) x = __float2half((rand() % 2001 - 1000) / 4000.0f);
    for (auto& x : hb) x = __float2half((rand() % 2001 - 1000) / 1000.0f);
    __half *A, *W, *b, *out; float* ref;
    CK(cudaMalloc(&A, hA.size() * 2)); CK(cudaMalloc(&W, hW.size() * 2)); CK(cudaMalloc(&b, N * 2));
    CK(cudaMalloc(&out, (size_t)M * N * 2)); CK(cudaMalloc(&ref, (size_t)M * N * 4));
    CK(cudaMemcpy(A, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(W, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b, hb.data(), N * 2, cudaMemcpyHostToDevice)); CK(cudaMemset(out, 0, (size_t)M * N * 2));
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn) { printf("cuTensorMapEncodeTiled not available\n"); return 1; }
    CUtensorMap ma, mw;
    if (make_map((EncodeTiled)fn, &ma, A, M, K) || make_map((EncodeTiled)fn, &mw, W, N, K)) return 1;
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
    CK(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
    gemm_tcgen05_kernel<<<grid, THREADS, smem>>>(ma, mw, b, out, M, N, K);
    CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    ref_gemm_kernel<<<dim3((N + 127) / 128, M), 128>>>(A, W, b, ref, M, N, K);
    CK(cudaDeviceSynchronize());
    std::vector<__half> ho((size_t)M * N); std::vector<float> hr((size_t)M * N);
    CK(cudaMemcpy(ho.data(), out, ho.size() * 2, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hr.data(), ref, hr.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (size_t i = 0; i < ho.size(); i++) {
        const double e = fabs((double)__half2float(ho[i]) - (double)hr[i]);
        if (e > maxerr) maxerr = e;
        if (fabs(hr[i]) > maxref) maxref = fabs(hr[i]);
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 20;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) gemm_tcgen05_kernel<<<grid, THREADS, smem>>>(ma, mw, b, out, M, N, K);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("M %d N %d K %d: max |err| %.4g (max |ref| %.4g, fp16 output rounding ~ %.3g)  %.3f ms  %.1f TFLOP/s\n", M, N, K, maxerr, maxref,
           maxref / 2048.0, ms / reps, 2.0 * M * N * K / (ms / reps * 1e-3) / 1e12);
    return maxerr <= maxref / 512.0 ? 0 : 2;
}
