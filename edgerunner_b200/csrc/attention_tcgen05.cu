// Multi-row attention (prefill / point encoder / teacher-forced forward) on the 5th-generation tensor cores: softmax(Q K^T / sqrt(D)) V with
// fp16 operands, fp32 scores / statistics / accumulation, probabilities rounded to fp16 before P V — the arithmetic of the flash kernel the
// reference calls (core/transformer/attention.py:44-46).  Replaces the mma.sync kernel of attention.cu where the operands meet the TMA rules.
//
// One CTA = one (batch, head, 128-query tile); key blocks of 64; 320 threads; two CTAs per SM (<= 96 KB of shared memory and 256 TMEM columns
// each).  S is double-buffered in TMEM: Q K^T of block j+1 is issued before the softmax of block j has finished, so the tensor core works
// under the softmax; the second CTA of the SM fills what is left:
//   warp 0      TMA producer: Q once, then per key block K and V as 3-D bulk tensor copies {64 dims, 1 head, 64|128 rows} with 128-byte swizzle.
//               D = 96 is two 64-dim atoms, the second half-filled: the tensor's innermost extent IS the head dim, so the copy engine
//               zero-fills dims 96..127 (K = 128 for Q K^T costs 33 % more MMA time there and nothing anywhere else)
//   warp 1      MMA issuer (one elected lane): S[128 x 64] = Q K^T (tcgen05.mma M 128, N 64, K-major x K-major) into TMEM columns 0..63;
//               O_blk[128 x D] = P V (A = P from shared memory, K-major; B = the V tile AS LOADED, i.e. MN-major: dims contiguous) into
//               TMEM columns 64..64+D; tcgen05.commit publishes S / O_blk and frees the K / V buffers for the next copies
//   warps 2..9  softmax: TWO threads per query row (TMEM lane): warps 2..5 take keys 0..31 and output dims [0, D/2), warps 6..9 keys 32..63
//               and dims [D/2, D) (a warp may touch TMEM lanes 32 (warp % 4) .. +31 only, any columns).  Pass 1 reads S for the row maximum
//               (the two halves meet through shared memory and one named barrier), pass 2 exponentiates the 32 scores still held in
//               registers (96 registers per thread: two CTAs per SM still fit; re-reading TMEM cost a round trip per block), rounds to fp16 and stores P in the K-major swizzled layout the MMA wants.  O ACCUMULATES IN TMEM
//               across key blocks (the MMA's accumulate flag); the reference maximum of a row is only raised when the block maximum exceeds
//               it by more than 2^8 in the exponent (P stays <= 256, exact in the final O / l), and only then are the row's O (tcgen05.ld /
//               .st) and l rescaled — rare after the first blocks.  K(j+1) streams in under softmax(j) / P V(j), V(j+1) under Q K^T(j+1).
// SASS: UTCHMMA, UTMALDG.3D, LDTM, UTCBAR.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace er {
namespace fa {

constexpr int BQ = 128, BKEY = 64, THREADS = 320;
constexpr uint32_t TMEM_COLS = 256;
constexpr int ATOM_Q = BQ * 128;            // bytes of one 64-dim atom of the Q tile (128 rows x 128 B)
constexpr int ATOM_K = BKEY * 128;          // ... of a K / V tile (64 rows x 128 B)

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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor, version 1, SWIZZLE_128B, 8-row groups 1024 B apart)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) {          // rows = M/N index, 64 K-elements (128 B) per row
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr, uint32_t atom_bytes) {   // rows = K index (8-row groups 1024 B apart), 64 N-elements per row,
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((atom_bytes >> 4) & 0x3FFF) << 16) |   // next 64 N-elements `atom_bytes` further (LBO)
           ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// instruction descriptor: D fp32, A/B fp16, A K-major, M = 128
__device__ __forceinline__ uint32_t idesc(int N, int b_mn_major) {
    return (1u << 4) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(BQ >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t id, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(id), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]),
          "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
          "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// exp2 as ONE MUFU.EX2 (ex2.approx.ftz): exp2f() wraps it in range checks and two scalings per call (denormal results), which tripled the
// instruction count of the softmax loop (ncu: 5 400 warp instructions per 128 x 64 block).  Inputs here are <= 8 and results below 2^-126
// may flush to zero: they are probabilities.
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct Args {
    __half* out; long long o_bs; int ldo;
    int B, H, Nq, Nk, causal;
    float scale_log2;        // softmax scale * log2(e)
    float* lse2;             // optional [B][H][Nq]: m * scale_log2 + log2(l) of every row (the training backward's softmax statistic)
};

template <int D>
__global__ void __launch_bounds__(THREADS, 2)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                         const __grid_constant__ Args a) {
    constexpr int NA = (D + 63) / 64;                  // 64-dim atoms per tile row
    constexpr int KS1 = NA * 4;                        // k16 steps of Q K^T (head dim padded to NA * 64 with zeros)
    constexpr int Q_BYTES = NA * ATOM_Q, K_BYTES = NA * ATOM_K;
    constexpr int DH = D / 2;                          // output dims per softmax thread
    constexpr int NB = (D <= 64) ? 2 : 1;              // K and V shared-memory buffers (two where two CTAs still fit an SM: the copy of block
                                                       // j+2 then starts as soon as the MMA of block j has read its buffer, a whole iteration early)
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long q_full, k_full[2], v_full[2], k_empty[2], v_empty[2], s_full[2], p_full, o_full[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ float xmax[2][BQ];                      // row maxima of the two key halves
    __shared__ float xsum[BQ];                         // final: denominator of the upper half
    unsigned char* base = (unsigned char*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    unsigned char* sQ = base;                          // [NA][128 rows][128 B]
    unsigned char* sK = sQ + Q_BYTES;                  // [NA][64 rows][128 B]
    unsigned char* sV = sK + NB * K_BYTES;             // [NA][64 rows][128 B]
    unsigned char* sP = sV + NB * K_BYTES;                  // 2 x [128 rows][128 B]   (64 keys, K-major, swizzled); block j uses buffer j & 1
    constexpr int P_BYTES = BQ * 128;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
    int nblk = (a.Nk + BKEY - 1) / BKEY;
    if (a.causal) nblk = min(nblk, (m0 + BQ + BKEY - 1) / BKEY);

    if (threadIdx.x == 0) {
        mbar_init(s_addr(&q_full), 1);
        for (int i = 0; i < 2; i++) { mbar_init(s_addr(&k_full[i]), 1); mbar_init(s_addr(&v_full[i]), 1); mbar_init(s_addr(&k_empty[i]), 1); mbar_init(s_addr(&v_empty[i]), 1); }
        mbar_init(s_addr(&s_full[0]), 1); mbar_init(s_addr(&s_full[1]), 1); mbar_init(s_addr(&p_full), 8); mbar_init(s_addr(&o_full[0]), 1); mbar_init(s_addr(&o_full[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t tmem_o = tmem_base + 2 * BKEY;      // S0: columns 0..63, S1: 64..127, O_blk: 128..128+D

    if (warp == 0) {
        if (lane == 0) {                                          // ===== TMA producer =====
            mbar_expect_tx(s_addr(&q_full), Q_BYTES);
#pragma unroll
            for (int t = 0; t < NA; t++) tma_load_3d(s_addr(sQ + t * ATOM_Q), &map_q, t * 64, h, b * a.Nq + m0, s_addr(&q_full));
            for (int j = 0; j < nblk; ++j) {                      // block j uses buffer j % NB; its barriers are in phase j / NB
                const int bf = j % NB, ph = j / NB;
                if (j >= NB) mbar_wait(s_addr(&k_empty[bf]), (ph - 1) & 1);   // Q K^T of block j-NB has read this K buffer
                mbar_expect_tx(s_addr(&k_full[bf]), K_BYTES);
#pragma unroll
                for (int t = 0; t < NA; t++) tma_load_3d(s_addr(sK + bf * K_BYTES + t * ATOM_K), &map_k, t * 64, h, b * a.Nk + j * BKEY, s_addr(&k_full[bf]));
                if (j >= NB) mbar_wait(s_addr(&v_empty[bf]), (ph - 1) & 1);   // P V of block j-NB has read this V buffer
                mbar_expect_tx(s_addr(&v_full[bf]), K_BYTES);
#pragma unroll
                for (int t = 0; t < NA; t++) tma_load_3d(s_addr(sV + bf * K_BYTES + t * ATOM_K), &map_v, t * 64, h, b * a.Nk + j * BKEY, s_addr(&v_full[bf]));
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                          // ===== MMA issuer =====
            const uint32_t id1 = idesc(BKEY, 0), id2 = idesc(D, 1);
            auto qk = [&](int j) {                                // S[j & 1] = Q K_j^T
                const int bf = j % NB;
                mbar_wait(s_addr(&k_full[bf]), (j / NB) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_s = tmem_base + (uint32_t)(j & 1) * BKEY;
#pragma unroll
                for (int s = 0; s < KS1; ++s) {
                    const uint64_t da = desc_kmajor(s_addr(sQ + (s >> 2) * ATOM_Q) + (s & 3) * 32);
                    const uint64_t db = desc_kmajor(s_addr(sK + bf * K_BYTES + (s >> 2) * ATOM_K) + (s & 3) * 32);
                    umma_f16(tmem_s, da, db, id1, s != 0);
                }
                umma_commit(s_addr(&k_empty[bf]));
                umma_commit(s_addr(&s_full[j & 1]));
            };
            mbar_wait(s_addr(&q_full), 0);
            qk(0);
            for (int j = 0; j < nblk; ++j) {
                // S[(j+1) & 1] was last read by the softmax of block j-1, which arrived on p_full(j-1) after reading it — waited for below, one
                // iteration ago — so Q K^T of block j+1 can run under the softmax of block j
                if (j + 1 < nblk) qk(j + 1);
                // O_blk = P V once the softmax warps have written P (and have finished reading O_blk of block j-1: program order on their side)
                mbar_wait(s_addr(&p_full), j & 1);
                const int bf = j % NB;
                mbar_wait(s_addr(&v_full[bf]), (j / NB) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int s = 0; s < BKEY / 16; ++s) {
                    const uint64_t da = desc_kmajor(s_addr(sP + (j & 1) * P_BYTES) + s * 32);
                    const uint64_t db = desc_mnmajor(s_addr(sV + bf * K_BYTES) + s * 2048, ATOM_K);      // 16 keys further = two 8-row groups
                    umma_f16(tmem_o, da, db, id2, (j | s) != 0);          // O accumulates in TMEM over all key blocks
                }
                umma_commit(s_addr(&v_empty[bf]));
                umma_commit(s_addr(&o_full[j & 1]));              // P V of blocks j, j+2, ... complete on barrier j & 1
            }
        }
    } else {                                                      // ===== softmax / epilogue: two threads per query row =====
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;                         // 0: keys 0..31, dims [0, DH); 1: keys 32..63, dims [DH, D)
        const int r = quad * 32 + lane;                           // row inside the tile = TMEM lane
        const int qi = m0 + r;
        const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
        float m_ref = -INFINITY, l_run = 0.f;              // m_ref: the (possibly stale) maximum the exponentials of this row are taken against
        for (int j = 0; j < nblk; ++j) {
            const int k0 = j * BKEY + half * 32;                  // first key of this thread's 32
            const bool edge = (a.causal && j * BKEY + BKEY - 1 > m0) || (j * BKEY + BKEY > a.Nk);   // some (row, key) pairs of this block are masked
            const uint32_t tmem_s = tmem_base + (uint32_t)(j & 1) * BKEY + (uint32_t)half * 32;
            mbar_wait(s_addr(&s_full[j & 1]), (j >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[32];
            // pass 1: maximum over this thread's 32 keys, then over the row (other half through shared memory)
            float mx = -INFINITY;
            tmem_ld32(tmem_s + lane_sel, v);
            {   // four independent chains: a 32-long dependent FMNMX chain costs ~150 cycles per block per warp
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                if (edge) {
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        const bool dead = (a.causal && k0 + i > qi) || k0 + i >= a.Nk;
                        m4[i & 3] = fmaxf(m4[i & 3], dead ? -INFINITY : __uint_as_float(v[i]));
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i++) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[i]));
                }
                mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            }
            xmax[half][r] = mx;
            asm volatile("bar.sync 1, 256;" ::: "memory");        // the 8 softmax warps
            mx = fmaxf(mx, xmax[half ^ 1][r]);
            // P is double-buffered: this block's buffer was last read by P V of block j-2, so the softmax of block j runs UNDER P V of block j-1
            // (the exponentials never wait for the tensor core in the steady state).  Two barriers, one per buffer, so that no wait can fall two
            // phases behind (parity waits only tell adjacent phases apart).
            if (j > 1) { mbar_wait(s_addr(&o_full[j & 1]), ((j >> 1) - 1) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
            // raise the reference maximum only when the block exceeds it by more than 2^8 (or it is still -inf)
            const bool raise = mx > m_ref && (m_ref == -INFINITY || (mx - m_ref) * a.scale_log2 > 8.f);
            float factor = 1.f;
            if (raise) {
                factor = (m_ref == -INFINITY) ? 0.f : fast_exp2((m_ref - mx) * a.scale_log2);
                m_ref = mx;
                l_run *= factor;
            }
            if (j > 0 && __any_sync(0xffffffffu, raise)) {        // warp-uniform: tcgen05.ld / .st are warp-wide
                // O is about to be rescaled in place: P V of block j-1 must have finished accumulating into it (rare after the first blocks)
                mbar_wait(s_addr(&o_full[(j - 1) & 1]), ((j - 1) >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c0 = 0; c0 < DH; c0 += 16) {
                    uint32_t w[16];
                    tmem_ld16(tmem_o + lane_sel + (uint32_t)(half * DH + c0), w);
#pragma unroll
                    for (int i = 0; i < 16; i++) w[i] = __float_as_uint(__uint_as_float(w[i]) * factor);
                    tmem_st16(tmem_o + lane_sel + (uint32_t)(half * DH + c0), w);
                }
            }
            const float msc = (m_ref == -INFINITY) ? 0.f : m_ref * a.scale_log2;
            // pass 2: p = exp2(s * scale - m_ref * scale), fp16, into the swizzled K-major P tile (row r: 128 B, 16-byte chunk c at position c ^ (r & 7))
            float lsum = 0.f;
            const uint32_t prow = s_addr(sP + (j & 1) * P_BYTES) + (uint32_t)r * 128;
            // the thread's 32 scores are still in registers from pass 1 (72 -> ~100 registers would still allow two CTAs per SM)
            auto pass2 = [&](const bool masked) {
                float ls[4] = {0.f, 0.f, 0.f, 0.f};               // independent partial sums (a single accumulator is a 32-long FADD chain)
#pragma unroll
                for (int c0 = 0; c0 < 32; c0 += 16) {
                    uint32_t ph[8];
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        float s0 = __uint_as_float(v[c0 + i]), s1 = __uint_as_float(v[c0 + i + 1]);
                        if (masked) {
                            if ((a.causal && k0 + c0 + i > qi) || k0 + c0 + i >= a.Nk) s0 = -INFINITY;
                            if ((a.causal && k0 + c0 + i + 1 > qi) || k0 + c0 + i + 1 >= a.Nk) s1 = -INFINITY;
                        }
                        const float p0 = fast_exp2(s0 * a.scale_log2 - msc), p1 = fast_exp2(s1 * a.scale_log2 - msc);
                        ls[(i >> 1) & 3] += p0 + p1;
                        const __half2 hh = __floats2half2_rn(p0, p1);
                        ph[i >> 1] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
#pragma unroll
                    for (int c = 0; c < 2; c++) {
                        const int chunk = half * 4 + (c0 >> 3) + c;
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(prow + (uint32_t)((chunk ^ (r & 7)) * 16)), "r"(ph[4 * c]), "r"(ph[4 * c + 1]),
                                     "r"(ph[4 * c + 2]), "r"(ph[4 * c + 3]) : "memory");
                    }
                }
                lsum = (ls[0] + ls[1]) + (ls[2] + ls[3]);
            };
            if (edge) pass2(true); else pass2(false);
            l_run += lsum;
            // P is read by the tensor core through the async proxy; S has been fully read; a rescaled O is in place
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(s_addr(&p_full));
        }
        // the last P V (the tensor core completes in order: every earlier one is done too)
        mbar_wait(s_addr(&o_full[(nblk - 1) & 1]), ((nblk - 1) >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // the two halves of a row share the denominator
        if (half == 1) xsum[r] = l_run;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (half == 0) xsum[r] += l_run;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        {   // every lane takes part in the TMEM loads; only rows inside the tile store
            const float l = xsum[r];
            const float inv = l > 0.f ? 1.f / l : 0.f;
            if (a.lse2 && half == 0 && qi < a.Nq) a.lse2[((size_t)b * a.H + h) * a.Nq + qi] = m_ref * a.scale_log2 + log2f(l);
            __half* op = a.out + (size_t)b * a.o_bs + (size_t)qi * a.ldo + (size_t)h * D + half * DH;
#pragma unroll
            for (int c0 = 0; c0 < DH; c0 += 16) {
                uint32_t w[16];
                tmem_ld16(tmem_o + lane_sel + (uint32_t)(half * DH + c0), w);
                __align__(16) __half hh[16];
#pragma unroll
                for (int i = 0; i < 16; i++) hh[i] = __float2half_rn(__uint_as_float(w[i]) * inv);
                if (qi < a.Nq) {
                    *reinterpret_cast<uint4*>(op + c0) = *reinterpret_cast<const uint4*>(hh);
                    *reinterpret_cast<uint4*>(op + c0 + 8) = *reinterpret_cast<const uint4*>(hh + 8);
                }
            }
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
// view [rows][H][D] (row pitch ld elements, head h at h * D): dims {D, H, rows}; box {64, 1, box_rows}; 128-byte swizzle; OOB = zero
static bool make_map(CUtensorMap* map, const __half* base, long long rows, int H, int D, int ld, int box_rows) {
    EncodeTiled enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)H, (cuuint64_t)rows};
    const cuuint64_t strides[2] = {(cuuint64_t)D * 2, (cuuint64_t)ld * 2};
    const cuuint32_t box[3] = {64, 1, (cuuint32_t)box_rows};
    const cuuint32_t elem[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int D>
static cudaError_t launch(const er::AttnArgs& a, cudaStream_t stream) {
    CUtensorMap mq, mk, mv;
    if (!make_map(&mq, a.q, (long long)a.B * a.Nq, a.H, D, a.ldq, BQ) || !make_map(&mk, a.k, (long long)a.B * a.Nk, a.H, D, a.ldk, BKEY) ||
        !make_map(&mv, a.v, (long long)a.B * a.Nk, a.H, D, a.ldv, BKEY))
        return cudaErrorNotSupported;
    Args g{};
    g.out = a.out; g.o_bs = a.o_bs; g.ldo = a.ldo; g.B = a.B; g.H = a.H; g.Nq = a.Nq; g.Nk = a.Nk; g.causal = a.causal;
    g.scale_log2 = rsqrtf((float)D) * 1.4426950408889634f;
    g.lse2 = a.lse2;
    constexpr int NA = (D + 63) / 64;
    constexpr int NB = (D <= 64) ? 2 : 1;
    const size_t smem = (size_t)NA * ATOM_Q + 2 * (size_t)NB * NA * ATOM_K + 2 * (size_t)BQ * 128 + 1024;
    cudaError_t e = cudaFuncSetAttribute(attention_tcgen05_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((a.Nq + BQ - 1) / BQ, a.H, a.B);
    attention_tcgen05_kernel<D><<<grid, THREADS, smem, stream>>>(mq, mk, mv, g);
    return cudaGetLastError();
}

}  // namespace fa
}  // namespace er

// -> cudaErrorNotSupported when the operand views do not meet the TMA rules (the caller then uses the mma.sync kernel)
cudaError_t er_attention_tcgen05(const er::AttnArgs& a, cudaStream_t stream) {
    // batches must be back to back in one row space, rows 16-byte aligned, head h at h * D inside a row
    if ((a.ldq & 7) || (a.ldk & 7) || (a.ldv & 7) || (a.ldo & 7) || ((uintptr_t)a.q & 15) || ((uintptr_t)a.k & 15) || ((uintptr_t)a.v & 15) || ((uintptr_t)a.out & 15))
        return cudaErrorNotSupported;
    if (a.B > 1 && (a.q_bs != (long long)a.Nq * a.ldq || a.k_bs != (long long)a.Nk * a.ldk || a.v_bs != (long long)a.Nk * a.ldv)) return cudaErrorNotSupported;
    if (a.H * a.D > a.ldq || a.H * a.D > a.ldk || a.H * a.D > a.ldv) return cudaErrorNotSupported;
    if (a.D == 96) return er::fa::launch<96>(a, stream);
    if (a.D == 64) return er::fa::launch<64>(a, stream);
    return cudaErrorNotSupported;
}
