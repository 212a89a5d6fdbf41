// Flash-attention backward, second implementation: mma.sync.m16n8k16 with the score / probability / dS blocks held in registers (the C fragment of
// two adjacent n8 tiles is the A fragment of the next k16 step), K / V / Q / dO tiles double-buffered in padded shared memory with cp.async —
// the fragment idioms are those of the forward kernel in attention.cu.  Same mathematics and the same three passes as the wmma version in backward.cu
// (which stays as the A/B reference: er_debug_set(NULL, "attn_bwd_wmma", 1)):
//
//   stats   lse2[r] = log2 sum_j exp2(scale_log2 * q_r . k_j) over the visible keys, dsum[r] = sum_d dO[r][d] O[r][d]
//   dQ      CTA = 64 queries, streams key tiles:   P = exp2(scale_log2 S - lse2), dP = dO V^T, dS = P (dP - dsum) scale, dQ += dS K
//   dK, dV  CTA = 64 keys, streams query tiles on transposed blocks: S^T = K Q^T, dP^T = V dO^T, dV += P^T dO, dK += dS^T Q
//
// Deterministic (every output row is owned by one warp; no atomics).  Reference semantics: torch autograd through flash_attn_func /
// the naive softmax path of core/transformer/attention.py:27-95.
#include "kernels.h"

#include "common.cuh"

namespace er {
namespace bwm {

constexpr int TILE = 64, THREADS = 128;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cpa16(uint32_t dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void ldsm4(uint32_t* r, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t* r, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// 64 rows x D fp16 (row pitch ld) -> shared rows of RS = 2 D + 16 bytes; rows >= nrows are zero-filled
template <int D>
__device__ __forceinline__ void load_rows(unsigned char* dst, const __half* src, int ld, int row0, int nrows, int tid) {
    constexpr int RS = D * 2 + 16, DC = D / 8;
    for (int c = tid; c < TILE * DC; c += THREADS) {
        const int r = c / DC, ch = c % DC;
        const bool ok = row0 + r < nrows;
        cpa16(s_u32(dst + r * RS + ch * 16), src + (size_t)(ok ? row0 + r : 0) * ld + ch * 8, ok ? 16 : 0);
    }
}
// A fragments (16 rows x D) of the warp's 16 rows starting at row r0 of a tile
template <int D>
__device__ __forceinline__ void load_a_frags(uint32_t (*f)[4], const unsigned char* tile, int r0, int lane) {
    constexpr int RS = D * 2 + 16;
#pragma unroll
    for (int kd = 0; kd < D / 16; kd++) {
        const int r = r0 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm4(f[kd], s_u32(tile + r * RS + (kd * 2 + (lane >> 4)) * 16));
    }
}
// c[0..1] (two n8 tiles = 16 columns) += A[16 x D] * T[rows n0 .. n0+15][D]^T      (T row-major [n][d]: the "K" idiom of the forward kernel)
template <int D>
__device__ __forceinline__ void mma_nt16(float (*c)[4], const uint32_t (*a)[4], const unsigned char* tile, int n0, int lane) {
    constexpr int RS = D * 2 + 16;
#pragma unroll
    for (int kd = 0; kd < D / 16; kd++) {
        uint32_t b[4];
        const int r = n0 + (lane & 7) + (lane >> 4) * 8;
        ldsm4(b, s_u32(tile + r * RS + (kd * 2 + ((lane >> 3) & 1)) * 16));
        mma16816(c[0], a[kd], b[0], b[1]);
        mma16816(c[1], a[kd], b[2], b[3]);
    }
}
// acc[16 x D] += A (one k16 step: 16 rows x 16 "k" rows k0 .. k0+15 of the tile) * T[k0 .. k0+15][D]     (the "V" idiom: transposed ldmatrix)
template <int D>
__device__ __forceinline__ void mma_nn16(float (*acc)[4], const uint32_t* a, const unsigned char* tile, int k0, int lane) {
    constexpr int RS = D * 2 + 16;
#pragma unroll
    for (int dj = 0; dj < D / 16; dj++) {
        uint32_t b[4];
        const int r = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
        ldsm4t(b, s_u32(tile + r * RS + (dj * 2 + (lane >> 4)) * 16));
        mma16816(acc[dj * 2], a, b[0], b[1]);
        mma16816(acc[dj * 2 + 1], a, b[2], b[3]);
    }
}
// fp16 store of the warp's [16 x D] accumulator: rows row0 + g and row0 + g + 8
template <int D>
__device__ __forceinline__ void store_acc(const float (*acc)[4], __half* dst, int ld, int row0, int nrows, int lane) {
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int row = row0 + g + r * 8;
        if (row >= nrows) continue;
#pragma unroll
        for (int i = 0; i < D / 8; i++)
            *reinterpret_cast<uint32_t*>(dst + (size_t)row * ld + i * 8 + tq * 2) = pack_h2(acc[i][r * 2], acc[i][r * 2 + 1]);
    }
}

// ---- pass 0: row statistics ------------------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(THREADS) stats_kernel(const AttnBwdArgs p) {
    constexpr int RS = D * 2 + 16, KD = D / 16;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sQ = smem;
    unsigned char* sK = sQ + TILE * RS;              // 2 stages
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
    const __half* qp = p.q + (size_t)b * p.q_bs + (size_t)h * D;
    const __half* kp = p.k + (size_t)b * p.k_bs + (size_t)h * D;
    int nkt = (p.Nk + TILE - 1) / TILE;
    if (p.causal) nkt = min(nkt, blockIdx.x + 1);
    load_rows<D>(sQ, qp, p.ld_qkv_q, q0, p.Nq, tid);
    load_rows<D>(sK, kp, p.ld_qkv_k, 0, p.Nk, tid);
    asm volatile("cp.async.commit_group;");
    uint32_t qf[KD][4];
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const int g = lane >> 2, tq = lane & 3;
    const int qrow0 = q0 + warp * 16 + g;
    for (int kt = 0; kt < nkt; kt++) {
        if (kt + 1 < nkt) load_rows<D>(sK + ((kt + 1) & 1) * TILE * RS, kp, p.ld_qkv_k, (kt + 1) * TILE, p.Nk, tid);
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 1;");
        __syncthreads();
        if (kt == 0) load_a_frags<D>(qf, sQ, warp * 16, lane);
        const unsigned char* kS = sK + (kt & 1) * TILE * RS;
        float s[TILE / 8][4];
#pragma unroll
        for (int i = 0; i < TILE / 8; i++) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int nj = 0; nj < TILE / 16; nj++) mma_nt16<D>(&s[nj * 2], qf, kS, nj * 16, lane);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < TILE / 8; i++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int key = kt * TILE + i * 8 + tq * 2 + (e & 1);
                const int qr = qrow0 + (e >> 1) * 8;
                const bool ok = key < p.Nk && (!p.causal || key <= qr);
                s[i][e] = ok ? s[i][e] * p.scale_log2 : -INFINITY;
                mx[e >> 1] = fmaxf(mx[e >> 1], s[i][e]);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; r++) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float mnew = fmaxf(m_run[r], mx[r]);
            l_run[r] = (m_run[r] == -INFINITY) ? 0.f : l_run[r] * exp2f(m_run[r] - mnew);
            m_run[r] = mnew;
        }
#pragma unroll
        for (int i = 0; i < TILE / 8; i++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float mm = m_run[e >> 1];
                l_run[e >> 1] += (mm == -INFINITY) ? 0.f : exp2f(s[i][e] - mm);
            }
        }
        __syncthreads();
    }
    const __half* op = p.o + (size_t)b * p.o_bs + (size_t)h * D;
    const __half* dop = p.dout + (size_t)b * p.o_bs + (size_t)h * D;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        float l = l_run[r];
        l += __shfl_xor_sync(0xffffffffu, l, 1);
        l += __shfl_xor_sync(0xffffffffu, l, 2);
        const int row = qrow0 + r * 8;
        float ds = 0.f;
        if (row < p.Nq) {
            const __half* o1 = op + (size_t)row * p.ld_o + tq * (D / 4);
            const __half* d1 = dop + (size_t)row * p.ld_o + tq * (D / 4);
            for (int c = 0; c < D / 4; c++) ds += __half2float(o1[c]) * __half2float(d1[c]);
        }
        ds += __shfl_xor_sync(0xffffffffu, ds, 1);
        ds += __shfl_xor_sync(0xffffffffu, ds, 2);
        if (row < p.Nq && tq == 0) {
            const size_t o = ((size_t)b * p.H + h) * p.Nq + row;
            p.lse2[o] = m_run[r] + log2f(l);
            p.dsum[o] = ds;
        }
    }
}

// ---- pass 1: dQ ---------------------------------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(THREADS) dq_kernel(const AttnBwdArgs p) {
    constexpr int RS = D * 2 + 16, KD = D / 16;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sQ = smem;
    unsigned char* sdO = sQ + TILE * RS;
    unsigned char* sK = sdO + TILE * RS;             // 2 stages
    unsigned char* sV = sK + 2 * TILE * RS;          // 2 stages
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
    const __half* qp = p.q + (size_t)b * p.q_bs + (size_t)h * D;
    const __half* kp = p.k + (size_t)b * p.k_bs + (size_t)h * D;
    const __half* vp = p.v + (size_t)b * p.v_bs + (size_t)h * D;
    const __half* dop = p.dout + (size_t)b * p.o_bs + (size_t)h * D;
    int nkt = (p.Nk + TILE - 1) / TILE;
    if (p.causal) nkt = min(nkt, blockIdx.x + 1);
    load_rows<D>(sQ, qp, p.ld_qkv_q, q0, p.Nq, tid);
    load_rows<D>(sdO, dop, p.ld_o, q0, p.Nq, tid);
    load_rows<D>(sK, kp, p.ld_qkv_k, 0, p.Nk, tid);
    load_rows<D>(sV, vp, p.ld_qkv_v, 0, p.Nk, tid);
    asm volatile("cp.async.commit_group;");
    const int g = lane >> 2, tq = lane & 3;
    const int qrow0 = q0 + warp * 16 + g;
    float lse[2], dsm[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int row = qrow0 + r * 8;
        const size_t o = ((size_t)b * p.H + h) * p.Nq + (row < p.Nq ? row : 0);
        lse[r] = row < p.Nq ? p.lse2[o] : 0.f;
        dsm[r] = row < p.Nq ? p.dsum[o] : 0.f;
    }
    uint32_t qf[KD][4], dof[KD][4];
    float dq[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; i++) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
    for (int kt = 0; kt < nkt; kt++) {
        if (kt + 1 < nkt) {
            load_rows<D>(sK + ((kt + 1) & 1) * TILE * RS, kp, p.ld_qkv_k, (kt + 1) * TILE, p.Nk, tid);
            load_rows<D>(sV + ((kt + 1) & 1) * TILE * RS, vp, p.ld_qkv_v, (kt + 1) * TILE, p.Nk, tid);
        }
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 1;");
        __syncthreads();
        if (kt == 0) { load_a_frags<D>(qf, sQ, warp * 16, lane); load_a_frags<D>(dof, sdO, warp * 16, lane); }
        const unsigned char* kS = sK + (kt & 1) * TILE * RS;
        const unsigned char* vS = sV + (kt & 1) * TILE * RS;
        uint32_t dsf[TILE / 16][4];
#pragma unroll
        for (int nj = 0; nj < TILE / 16; nj++) {
            float s[2][4], dp[2][4];
#pragma unroll
            for (int i = 0; i < 2; i++) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f; }
            mma_nt16<D>(s, qf, kS, nj * 16, lane);
            mma_nt16<D>(dp, dof, vS, nj * 16, lane);
#pragma unroll
            for (int i = 0; i < 2; i++) {
                float ds[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int key = kt * TILE + nj * 16 + i * 8 + tq * 2 + (e & 1);
                    const int qr = qrow0 + (e >> 1) * 8;
                    const bool ok = qr < p.Nq && key < p.Nk && (!p.causal || key <= qr);
                    const float pr = ok ? exp2f(s[i][e] * p.scale_log2 - lse[e >> 1]) : 0.f;
                    ds[e] = pr * (dp[i][e] - dsm[e >> 1]) * p.scale;
                }
                dsf[nj][i * 2 + 0] = pack_h2(ds[0], ds[1]);
                dsf[nj][i * 2 + 1] = pack_h2(ds[2], ds[3]);
            }
        }
#pragma unroll
        for (int kk = 0; kk < TILE / 16; kk++) mma_nn16<D>(dq, dsf[kk], kS, kk * 16, lane);
        __syncthreads();
    }
    store_acc<D>(dq, p.dq + (size_t)b * p.dq_bs + (size_t)h * D, p.ld_dq, q0 + warp * 16, p.Nq, lane);
}

// ---- pass 2: dK, dV -------------------------------------------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(THREADS) dkv_kernel(const AttnBwdArgs p) {
    constexpr int RS = D * 2 + 16, KD = D / 16;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sK = smem;
    unsigned char* sV = sK + TILE * RS;
    unsigned char* sQ = sV + TILE * RS;              // 2 stages
    unsigned char* sdO = sQ + 2 * TILE * RS;         // 2 stages
    float* lse_s = reinterpret_cast<float*>(sdO + 2 * TILE * RS);      // [2][64]
    float* dsm_s = lse_s + 2 * TILE;                                    // [2][64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k0 = blockIdx.x * TILE, h = blockIdx.y, b = blockIdx.z;
    const __half* qp = p.q + (size_t)b * p.q_bs + (size_t)h * D;
    const __half* kp = p.k + (size_t)b * p.k_bs + (size_t)h * D;
    const __half* vp = p.v + (size_t)b * p.v_bs + (size_t)h * D;
    const __half* dop = p.dout + (size_t)b * p.o_bs + (size_t)h * D;
    const size_t sbase = ((size_t)b * p.H + h) * p.Nq;
    const int nqt = (p.Nq + TILE - 1) / TILE;
    const int qt0 = p.causal ? blockIdx.x : 0;
    auto load_q = [&](int stage, int qt) {
        load_rows<D>(sQ + stage * TILE * RS, qp, p.ld_qkv_q, qt * TILE, p.Nq, tid);
        load_rows<D>(sdO + stage * TILE * RS, dop, p.ld_o, qt * TILE, p.Nq, tid);
        if (tid < TILE) {
            const int r = qt * TILE + tid;
            lse_s[stage * TILE + tid] = r < p.Nq ? p.lse2[sbase + r] : 0.f;
            dsm_s[stage * TILE + tid] = r < p.Nq ? p.dsum[sbase + r] : 0.f;
        }
    };
    load_rows<D>(sK, kp, p.ld_qkv_k, k0, p.Nk, tid);
    load_rows<D>(sV, vp, p.ld_qkv_v, k0, p.Nk, tid);
    if (qt0 < nqt) load_q(0, qt0);
    asm volatile("cp.async.commit_group;");
    const int g = lane >> 2, tq = lane & 3;
    const int krow0 = k0 + warp * 16 + g;
    uint32_t kf[KD][4], vf[KD][4];
    float dk[D / 8][4], dv[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; i++) { dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f; dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f; }
    for (int qt = qt0; qt < nqt; qt++) {
        const int it = qt - qt0;
        if (qt + 1 < nqt) load_q((it + 1) & 1, qt + 1);
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 1;");
        __syncthreads();
        if (it == 0) { load_a_frags<D>(kf, sK, warp * 16, lane); load_a_frags<D>(vf, sV, warp * 16, lane); }
        const unsigned char* qS = sQ + (it & 1) * TILE * RS;
        const unsigned char* doS = sdO + (it & 1) * TILE * RS;
        const float* lse_t = lse_s + (it & 1) * TILE;
        const float* dsm_t = dsm_s + (it & 1) * TILE;
#pragma unroll
        for (int nj = 0; nj < TILE / 16; nj++) {
            float st[2][4], dpt[2][4];
#pragma unroll
            for (int i = 0; i < 2; i++) { st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f; dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f; }
            mma_nt16<D>(st, kf, qS, nj * 16, lane);
            mma_nt16<D>(dpt, vf, doS, nj * 16, lane);
            uint32_t ptf[4], dstf[4];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                float pr[4], ds[4];
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int ql = nj * 16 + i * 8 + tq * 2 + (e & 1);          // query inside the tile (the block's column)
                    const int qg = qt * TILE + ql;
                    const int key = krow0 + (e >> 1) * 8;
                    const bool ok = key < p.Nk && qg < p.Nq && (!p.causal || key <= qg);
                    pr[e] = ok ? exp2f(st[i][e] * p.scale_log2 - lse_t[ql]) : 0.f;
                    ds[e] = pr[e] * (dpt[i][e] - dsm_t[ql]) * p.scale;
                }
                ptf[i * 2 + 0] = pack_h2(pr[0], pr[1]); ptf[i * 2 + 1] = pack_h2(pr[2], pr[3]);
                dstf[i * 2 + 0] = pack_h2(ds[0], ds[1]); dstf[i * 2 + 1] = pack_h2(ds[2], ds[3]);
            }
            mma_nn16<D>(dv, ptf, doS, nj * 16, lane);
            mma_nn16<D>(dk, dstf, qS, nj * 16, lane);
        }
        __syncthreads();
    }
    store_acc<D>(dk, p.dk + (size_t)b * p.dk_bs + (size_t)h * D, p.ld_dk, k0 + warp * 16, p.Nk, lane);
    store_acc<D>(dv, p.dv + (size_t)b * p.dv_bs + (size_t)h * D, p.ld_dv, k0 + warp * 16, p.Nk, lane);
}

template <int D>
static cudaError_t launch(const AttnBwdArgs& p, cudaStream_t st) {
    constexpr int RS = D * 2 + 16;
    const int smem_stats = 3 * TILE * RS, smem_dq = 6 * TILE * RS, smem_dkv = 6 * TILE * RS + 4 * TILE * 4;
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(stats_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_stats)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dq)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_dkv)) != cudaSuccess) return e;
    const dim3 gq((p.Nq + TILE - 1) / TILE, p.H, p.B), gk((p.Nk + TILE - 1) / TILE, p.H, p.B);
    if (!p.have_lse) stats_kernel<D><<<gq, THREADS, smem_stats, st>>>(p);
    else if ((e = er_attn_rowdot(p, D, st)) != cudaSuccess) return e;
    dq_kernel<D><<<gq, THREADS, smem_dq, st>>>(p);
    dkv_kernel<D><<<gk, THREADS, smem_dkv, st>>>(p);
    return cudaGetLastError();
}

}  // namespace bwm
}  // namespace er

cudaError_t er_attn_bwd_mma(const er::AttnBwdArgs& p, int D, cudaStream_t st) {
    return D == 96 ? er::bwm::launch<96>(p, st) : er::bwm::launch<64>(p, st);
}
