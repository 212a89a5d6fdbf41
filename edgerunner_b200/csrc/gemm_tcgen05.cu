// Dense GEMM of the N > 1 paths (prefill, point encoder, teacher-forced forward) on the 5th-generation tensor cores:
//   out[M][N] = epilogue(A[M][K] . W[N][K]^T + bias)      fp16 operands, fp32 accumulation in TENSOR MEMORY
// Replaces the cuBLAS calls behind every nn.Linear of the reference on these paths (modeling_opt.py:185-232, 281-284, 497; point.py:108-126)
// with the autocast rounding points fused into the epilogue (see GemmMode in kernels.h).
//
// Persistent kernel, one CTA per SM, 192 threads, 128 x 256 output tiles walked N-fastest (the A rows of a tile row stay in L2):
//   warp 0      TMA producer: per 64-wide K block two 2-D bulk tensor copies (A tile 128 x 64, W tile 256 x 64, 128-byte swizzle, OOB rows /
//               K tail zero-filled by the copy engine) into a 4-stage shared-memory ring (48 KB per stage), completion by mbarrier tx count
//   warp 1      MMA issuer (one elected lane): 4 x tcgen05.mma.cta_group::1.kind::f16 (M 128, N 256, K 16) per stage from shared-memory
//               matrix descriptors; the accumulator lives in 256 TMEM columns and there are TWO of them (all 512 columns), so the next
//               tile's main loop runs while the epilogue drains the previous one; tcgen05.commit releases smem stages / publishes a tile
//   warps 2..9  epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31 (one output row per thread) and one 128-column half of the tile,
//               tcgen05.ld 32 columns at a time, bias, rounding / ReLU / residual / GEGLU / gated residual of the mode, 32- / 64-byte row
//               segments stored straight from registers.  Eight warps because with K = 1024 a tile's main loop is only ~8k cycles: four
//               warps could not drain an fp32-residual epilogue (16 dependent global round trips per row) in that time
// SASS: UTCHMMA, UTMALDG.2D, LDTM, UTCBAR, SYNCS.  The descriptors follow cute::UMMA (SmemDescriptor / InstrDescriptor of the CUTLASS
// headers vendored under flashinfer/data/cutlass); scripts/microbench/gemm_tcgen05.cu is the single-tile self-checking precursor
// (profiles/r02_microbench_gemm_tcgen05_draft.txt).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace er {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 64, UK = 16, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;   // 16 KB + 32 KB
constexpr int THREADS = 320;
constexpr uint32_t TMEM_COLS = 512;     // two 256-column accumulators

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
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
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // layout type SWIZZLE_128B
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B fp16, both K-major, N = 256, M = 128
__device__ __forceinline__ uint32_t instr_desc() {
    return (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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

struct EpiArgs {
    const __half* bias; int mode;
    __half* out16; float* out32; int ldo;
    const __half* res16; const float* res32; int ldr;
    const __half *gate_tab, *gate_t; long long gate_bs; int n_per;
    int M, N, K;
};

__device__ __forceinline__ void load16h(const __half* p, float* f) {      // 16 fp16 (32-byte aligned run) -> fp32
    const uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + 8);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; j++) { const float2 t = h2f2(w[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ void store16h(__half* p, const float* f) {
    __align__(16) __half h[16];
#pragma unroll
    for (int j = 0; j < 16; j++) h[j] = __float2half_rn(f[j]);
    *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(p + 8) = *reinterpret_cast<const uint4*>(h + 8);
}
// GEMM_F16_GEGLU: columns [col, col+16) are the value half, [col+16, col+32) the gate half of output columns [col/2, col/2+16)
__device__ __forceinline__ void epilogue_geglu32(const EpiArgs& g, const int row, const int col, const uint32_t* v) {
    float ba[16], bg[16], o[16];
    load16h(g.bias + col, ba); load16h(g.bias + col + 16, bg);
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const float a = round_f16(__uint_as_float(v[j]) + ba[j]);
        const float x = round_f16(__uint_as_float(v[16 + j]) + bg[j]);
        o[j] = a * round_f16(0.5f * x * (1.f + erff(x * 0.70710678118654752440f)));
    }
    store16h(g.out16 + (size_t)row * g.ldo + (col >> 1), o);
}
// 16 consecutive columns [col, col+16) of output row `row`: v = fp32 accumulators
__device__ __forceinline__ void epilogue_16(const EpiArgs& g, const int row, const int col, const uint32_t* v) {
    float x[16];
#pragma unroll
    for (int j = 0; j < 16; j++) x[j] = __uint_as_float(v[j]);
    const bool full = col + 16 <= g.N;
    if (g.bias) {
        if (full) {
            const uint4 b0 = *reinterpret_cast<const uint4*>(g.bias + col), b1 = *reinterpret_cast<const uint4*>(g.bias + col + 8);
            const uint32_t bw[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; j++) { const float2 f = h2f2(bw[j]); x[2 * j] += f.x; x[2 * j + 1] += f.y; }
        } else {
            for (int j = 0; j < 16 && col + j < g.N; j++) x[j] += __half2float(g.bias[col + j]);
        }
    }
    const size_t o = (size_t)row * g.ldo + col;
    if (g.mode == GEMM_F16 || g.mode == GEMM_F16_RELU || g.mode == GEMM_F16_RES16) {
        __align__(16) __half h[16];
        if (g.mode == GEMM_F16_RES16) {
            const __half* rp = g.res16 + (size_t)row * g.ldr + col;
            if (full && ((g.ldr & 7) == 0)) {
                const uint4 r0 = *reinterpret_cast<const uint4*>(rp), r1 = *reinterpret_cast<const uint4*>(rp + 8);
                const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int j = 0; j < 8; j++) { const float2 f = h2f2(rw[j]); h[2 * j] = __float2half_rn(round_f16(x[2 * j]) + f.x); h[2 * j + 1] = __float2half_rn(round_f16(x[2 * j + 1]) + f.y); }
            } else {
                for (int j = 0; j < 16; j++) h[j] = __float2half_rn(round_f16(x[j]) + ((col + j < g.N) ? __half2float(rp[j]) : 0.f));
            }
        } else if (g.mode == GEMM_F16_RELU) {
#pragma unroll
            for (int j = 0; j < 16; j++) h[j] = __float2half_rn(fmaxf(round_f16(x[j]), 0.f));
        } else {
#pragma unroll
            for (int j = 0; j < 16; j++) h[j] = __float2half_rn(x[j]);
        }
        if (full && ((g.ldo & 7) == 0)) {
            *reinterpret_cast<uint4*>(g.out16 + o) = *reinterpret_cast<const uint4*>(h);
            *reinterpret_cast<uint4*>(g.out16 + o + 8) = *reinterpret_cast<const uint4*>(h + 8);
        } else {
            for (int j = 0; j < 16 && col + j < g.N; j++) g.out16[o + j] = h[j];
        }
    } else {
        if (g.mode == GEMM_F32_RES32) {
            const float* rp = g.res32 + (size_t)row * g.ldr + col;
            if (full && ((g.ldr & 3) == 0)) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
                    x[j] = r4.x + round_f16(x[j]); x[j + 1] = r4.y + round_f16(x[j + 1]); x[j + 2] = r4.z + round_f16(x[j + 2]); x[j + 3] = r4.w + round_f16(x[j + 3]);
                }
            } else {
                for (int j = 0; j < 16; j++) x[j] = ((col + j < g.N) ? rp[j] : 0.f) + round_f16(x[j]);
            }
        }
        if (full && ((g.ldo & 3) == 0)) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(g.out32 + o + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
        } else {
            for (int j = 0; j < 16 && col + j < g.N; j++) g.out32[o + j] = x[j];
        }
    }
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, const __grid_constant__ EpiArgs g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (g.K + BK - 1) / BK;
    const int tiles_n = (g.N + BN - 1) / BN, tiles_m = (g.M + BM - 1) / BM, ntiles = tiles_n * tiles_m;
    unsigned char* tiles = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    float* scratch = reinterpret_cast<float*>(tiles + STAGES * STAGE_BYTES);      // 8 x 4 KB transpose scratch of the epilogue warps

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; i++) { mbar_init(s_addr(&full_bar[i]), 1); mbar_init(s_addr(&empty_bar[i]), 1); }
        for (int i = 0; i < 2; i++) { mbar_init(s_addr(&acc_full[i]), 1); mbar_init(s_addr(&acc_empty[i]), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {       // one warp allocates all 512 TMEM columns (1 CTA per SM: the shared-memory ring alone guarantees it)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {                                          // ===== TMA producer =====
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int st = it % STAGES;
                    if (it >= (uint32_t)STAGES) mbar_wait(s_addr(&empty_bar[st]), ((it / STAGES) - 1) & 1);
                    const uint32_t fb = s_addr(&full_bar[st]);
                    mbar_expect_tx(fb, STAGE_BYTES);
                    tma_load_2d(s_addr(tiles + st * STAGE_BYTES), &map_a, kb * BK, m0, fb);
                    tma_load_2d(s_addr(tiles + st * STAGE_BYTES + A_BYTES), &map_w, kb * BK, n0, fb);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                          // ===== MMA issuer =====
            const uint32_t idesc = instr_desc();
            uint32_t it = 0, local = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++local) {
                const uint32_t buf = local & 1;
                if (local >= 2) mbar_wait(s_addr(&acc_empty[buf]), ((local >> 1) - 1) & 1);      // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + buf * (uint32_t)BN;
                for (int kb = 0; kb < nk; ++kb, ++it) {
                    const int st = it % STAGES;
                    mbar_wait(s_addr(&full_bar[st]), (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint64_t da = smem_desc(s_addr(tiles + st * STAGE_BYTES));
                    const uint64_t db = smem_desc(s_addr(tiles + st * STAGE_BYTES + A_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k)              // + 32 bytes (2 x 16-byte units) of K per step inside the swizzle atom
                        umma_f16(tmem_d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    umma_commit(s_addr(&empty_bar[st]));          // stage reusable once these MMAs have read it
                }
                umma_commit(s_addr(&acc_full[buf]));              // accumulator of this tile complete
            }
        }
    } else {                                                      // ===== epilogue (warps 2..9) =====
        const int quad = warp & 3;                                // TMEM lane quadrant this warp may access
        const int half = (warp - 2) >> 2;                         // which 128 columns of the tile
        uint32_t local = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++local) {
            const uint32_t buf = local & 1;
            const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN + half * (BN / 2);
            mbar_wait(s_addr(&acc_full[buf]), (local >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = m0 + quad * 32 + lane;
            const int ncols = min(BN / 2, g.N - n0);
#pragma unroll 1
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (uint32_t)BN + (uint32_t)(half * (BN / 2) + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                      "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                      "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (g.mode == GEMM_GATE_RES32 || g.mode == GEMM_F32_RES32) {
                    // fp32 residual stream: one row per thread would make every 128-bit access touch 32 different rows (32 L1 wavefronts
                    // per instruction; ~20k wavefronts per tile against an ~8k-cycle main loop).  Transpose the warp's 32 x 32 block through a
                    // private, XOR-swizzled shared-memory scratch instead, so that each global access is one 128-byte row segment.
                    float* sc = scratch + (warp - 2) * 1024;
                    const int col = n0 + c0;
                    const bool two = col + 32 <= g.N;                 // (N is a multiple of 16 for these modes' callers; ragged N handled per element)
                    float b[32];
                    if (g.bias && two) { load16h(g.bias + col, b); load16h(g.bias + col + 16, b + 16); }
                    else { for (int j = 0; j < 32; j++) b[j] = (g.bias && col + j < g.N) ? __half2float(g.bias[col + j]) : 0.f; }
                    if (g.mode == GEMM_GATE_RES32) {
                        float gt[32], tv[32];
                        const __half* tp = g.gate_t + (size_t)(min(row, g.M - 1) / g.n_per) * g.gate_bs + col;
                        load16h(g.gate_tab + col, gt); load16h(g.gate_tab + col + 16, gt + 16); load16h(tp, tv); load16h(tp + 16, tv + 16);
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            sc[lane * 32 + (j ^ lane)] = round_f16(round_f16(gt[j] + tv[j]) * round_f16(__uint_as_float(v[j]) + b[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++) sc[lane * 32 + (j ^ lane)] = round_f16(__uint_as_float(v[j]) + b[j]);
                    }
                    __syncwarp();
                    const int rbase = m0 + quad * 32;
                    const int nrow = min(32, g.M - rbase);
                    const int cc = col + lane;
                    if (cc < g.N) {
                        const float* rp = g.res32 + (size_t)rbase * g.ldr + cc;
                        float* op = g.out32 + (size_t)rbase * g.ldo + cc;
                        if (nrow == 32) {             // all 32 residual loads in flight before the first use (one memory latency per block, not 32)
                            float r[32];
#pragma unroll
                            for (int rr = 0; rr < 32; rr++) r[rr] = rp[(size_t)rr * g.ldr];
#pragma unroll
                            for (int rr = 0; rr < 32; rr++) {
                                r[rr] += sc[rr * 32 + (lane ^ rr)];
                                op[(size_t)rr * g.ldo] = r[rr];
                            }
                            if (g.out16) {
                                __half* hp = g.out16 + (size_t)rbase * g.ldo + cc;
#pragma unroll
                                for (int rr = 0; rr < 32; rr++) hp[(size_t)rr * g.ldo] = __float2half_rn(r[rr]);
                            }
                        } else {
                            for (int rr = 0; rr < nrow; rr++) {
                                const float o = rp[(size_t)rr * g.ldr] + sc[rr * 32 + (lane ^ rr)];
                                op[(size_t)rr * g.ldo] = o;
                                if (g.out16) g.out16[(size_t)(rbase + rr) * g.ldo + cc] = __float2half_rn(o);
                            }
                        }
                    }
                    __syncwarp();
                } else if (row < g.M) {
                    if (g.mode == GEMM_F16_GEGLU) {
                        epilogue_geglu32(g, row, n0 + c0, v);
                    } else {
                        epilogue_16(g, row, n0 + c0, v);
                        if (c0 + 16 < ncols) epilogue_16(g, row, n0 + c0 + 16, v + 16);
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(s_addr(&acc_empty[buf]));  // 8 arrivals (one per epilogue warp) free the accumulator
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiled encode_fn() {
    static EncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess) fn = (EncodeTiled)p;
    }
    return fn;
}
// [rows][K] fp16 with row pitch ld elements, box 64 x box_rows, 128-byte swizzle, out-of-bounds elements read as zero
static bool make_map(CUtensorMap* map, const __half* base, int rows, int K, int ld, int box_rows) {
    EncodeTiled enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t elem[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tc
}  // namespace er

// -> cudaSuccess, or cudaErrorNotSupported when the operands do not meet the TMA constraints (the caller then uses the mma.sync kernel)
cudaError_t er_gemm_tcgen05(const er::GemmArgs& a, cudaStream_t stream) {
    using namespace er::tc;
    if ((a.lda & 7) || (a.ldw & 7) || ((uintptr_t)a.A & 15) || ((uintptr_t)a.W & 15) || a.K < 8) return cudaErrorNotSupported;
    if (a.bias && ((uintptr_t)a.bias & 15)) return cudaErrorNotSupported;
    if (a.mode == er::GEMM_F16_GEGLU || a.mode == er::GEMM_GATE_RES32) {        // fused DiT epilogues: vector accesses only, no ragged N
        if ((a.N & 31) || !a.bias || (a.ldo & 7)) return cudaErrorInvalidValue;
        if (a.mode == er::GEMM_GATE_RES32 && (!a.res32 || !a.out32 || !a.gate_tab || !a.gate_t || a.n_per <= 0 || (a.ldr & 3) || (a.gate_bs & 7)))
            return cudaErrorInvalidValue;
    }
    CUtensorMap ma, mw;
    if (!make_map(&ma, a.A, a.M, a.K, a.lda, BM) || !make_map(&mw, a.W, a.N, a.K, a.ldw, BN)) return cudaErrorNotSupported;
    EpiArgs g{};
    g.bias = a.bias; g.mode = a.mode; g.out16 = a.out16; g.out32 = a.out32; g.ldo = a.ldo; g.res16 = a.res16; g.res32 = a.res32; g.ldr = a.ldr;
    g.gate_tab = a.gate_tab; g.gate_t = a.gate_t; g.gate_bs = a.gate_bs; g.n_per = a.n_per;
    g.M = a.M; g.N = a.N; g.K = a.K;
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024 + 8 * 4096;
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int ntiles = ((a.M + BM - 1) / BM) * ((a.N + BN - 1) / BN);
    gemm_tcgen05_kernel<<<ntiles < sms ? ntiles : sms, THREADS, smem, stream>>>(ma, mw, g);
    return cudaGetLastError();
}
