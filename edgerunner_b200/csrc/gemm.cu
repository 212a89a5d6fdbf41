// Dense fp16 GEMM for the prefill / encoder / teacher-forced paths:  out = epilogue(A[M,K] @ W[N,K]^T + bias).
//
// Stands in for every nn.Linear that sees more than one row (reference: core/transformer/modeling_opt.py:185-190,
// 232,281,284,497; core/transformer/point.py:64,78-80,202; core/transformer/attention.py:148-153;
// core/models.py:124), with the autocast rounding points fused into the epilogue (fp32 accumulate, +bias, one
// rounding to fp16, then ReLU / residual add in the dtype the reference adds in).
//
// Round-1 implementation: warp-level tensor-core MMA (mma.sync.m16n8k16, fp32 accumulate), 128x128x32 CTA tile,
// 3-stage cp.async pipeline, XOR-swizzled smem + ldmatrix.  (tcgen05/TMEM version: see DESIGN.md "next".)
#include "kernels.h"

#include "common.cuh"

namespace er {

constexpr int BM = 128, BN = 128, BK = 32, STAGES = 3, GEMM_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, const uint32_t* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// tile row r (64 bytes = 4 chunks of 16 B), chunk c -> byte offset with XOR swizzle (conflict-free ldmatrix)
__device__ __forceinline__ int swz(int r, int c) { return r * 64 + ((c ^ ((r >> 1) & 3)) << 4); }

__global__ void __launch_bounds__(GEMM_THREADS) gemm_f16_kernel(GemmArgs g) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;                         // STAGES * BM * 64
    unsigned char* sB = smem + STAGES * BM * 64;      // STAGES * BN * 64
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;          // 2 x 4 warps, warp tile 64 x 32
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int KT = (g.K + BK - 1) / BK;

    auto load_stage = [&](int stage, int kt) {
        const int k0 = kt * BK;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int chunk = tid + i * GEMM_THREADS;   // 0..511
            const int r = chunk >> 2, c = chunk & 3;
            const int k = k0 + c * 8;
            {
                const int row = m0 + r;
                const bool ok = row < g.M && k < g.K;
                const __half* src = g.A + (size_t)(ok ? row : 0) * g.lda + (ok ? k : 0);
                cp_async16(smem_u32(sA + stage * BM * 64 + swz(r, c)), src, ok ? 16 : 0);
            }
            {
                const int row = n0 + r;
                const bool ok = row < g.N && k < g.K;
                const __half* src = g.W + (size_t)(ok ? row : 0) * g.ldw + (ok ? k : 0);
                cp_async16(smem_u32(sB + stage * BN * 64 + swz(r, c)), src, ok ? 16 : 0);
            }
        }
    };

    float acc[4][4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int r = 0; r < 4; r++) acc[i][j][r] = 0.f;

#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < KT) load_stage(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; kt++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) load_stage(nk % STAGES, nk);
            cp_async_commit();
        }
        const unsigned char* a_st = sA + (kt % STAGES) * BM * 64;
        const unsigned char* b_st = sB + (kt % STAGES) * BN * 64;
#pragma unroll
        for (int kk = 0; kk < 2; kk++) {              // two k16 steps per BK = 32
            uint32_t af[4][4], bf[4][2];
#pragma unroll
            for (int mi = 0; mi < 4; mi++) {
                const int r = wm * 64 + mi * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = kk * 2 + (lane >> 4);
                ldmatrix_x4(af[mi][0], af[mi][1], af[mi][2], af[mi][3], smem_u32(a_st + swz(r, c)));
            }
#pragma unroll
            for (int nj = 0; nj < 2; nj++) {
                const int r = wn * 32 + nj * 16 + (lane & 7) + (lane >> 4) * 8;
                const int c = kk * 2 + ((lane >> 3) & 1);
                ldmatrix_x4(bf[nj * 2][0], bf[nj * 2][1], bf[nj * 2 + 1][0], bf[nj * 2 + 1][1], smem_u32(b_st + swz(r, c)));
            }
#pragma unroll
            for (int mi = 0; mi < 4; mi++)
#pragma unroll
                for (int ni = 0; ni < 4; ni++) mma_16816(acc[mi][ni], af[mi], bf[ni]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue ---------------------------------------------------------------------------------------------
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int mi = 0; mi < 4; mi++) {
#pragma unroll
        for (int half_ = 0; half_ < 2; half_++) {
            const int row = m0 + wm * 64 + mi * 16 + gq + half_ * 8;
            if (row >= g.M) continue;
#pragma unroll
            for (int ni = 0; ni < 4; ni++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int col = n0 + wn * 32 + ni * 8 + tq * 2 + e;
                    if (col >= g.N) continue;
                    float v = acc[mi][ni][half_ * 2 + e];
                    if (g.bias) v += __half2float(g.bias[col]);
                    const size_t o = (size_t)row * g.ldo + col;
                    switch (g.mode) {
                        case GEMM_F16: g.out16[o] = __float2half_rn(v); break;
                        case GEMM_F16_RELU: g.out16[o] = __float2half_rn(fmaxf(round_f16(v), 0.f)); break;
                        case GEMM_F16_RES16: g.out16[o] = __float2half_rn(round_f16(v) + __half2float(g.res16[(size_t)row * g.ldr + col])); break;
                        case GEMM_F32_RES32: g.out32[o] = g.res32[(size_t)row * g.ldr + col] + round_f16(v); break;
                        default: g.out32[o] = v; break;
                    }
                }
            }
        }
    }
}

}  // namespace er

cudaError_t er_gemm_tcgen05(const er::GemmArgs& a, cudaStream_t stream);   // gemm_tcgen05.cu: the tensor-memory kernel (production path)

cudaError_t er_gemm(const er::GemmArgs& g, cudaStream_t stream) {
    using namespace er;
    if (g.M <= 0 || g.N <= 0) return cudaSuccess;
    extern int g_er_dense_legacy;
    if (!g_er_dense_legacy) {   // tcgen05 + TMA kernel whenever the operands meet the TMA alignment rules (every shape of the ArAE / tiny presets does); the
        // mma.sync kernel below remains for odd strides
        const cudaError_t e = er_gemm_tcgen05(g, stream);
        if (e != cudaErrorNotSupported) return e;
    }
    if (g.mode == er::GEMM_F16_GEGLU || g.mode == er::GEMM_GATE_RES32) return cudaErrorNotSupported;   // fused DiT epilogues exist in the tcgen05 kernel only
    if ((g.K & 7) || (g.lda & 7) || (g.ldw & 7)) return cudaErrorInvalidValue;
    static bool attr[64] = {};                 // per DEVICE: the attribute belongs to the function on the current device's context
    const int smem = STAGES * (BM + BN) * 64;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM);
    gemm_f16_kernel<<<grid, GEMM_THREADS, smem, stream>>>(g);
    return cudaGetLastError();
}
