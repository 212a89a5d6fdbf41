// Multi-row attention for the prefill / encoder / teacher-forced paths (flash-style, online softmax).
//
// Replaces core/transformer/attention.py:27-62 (`attention(q,k,v,causal)`, [B,N,H,D] layout) where it is called
// with N > 1: the causal self-attention of the 2050-row prefix (modeling_opt.py:229) and the 2048 x 8192
// cross-attention of the point encoder (attention.py:151).  Same numerics class as the flash-attn kernel the
// reference calls: fp16 operands, fp32 scores / softmax statistics / output accumulation, softmax scale D^-0.5,
// probabilities rounded to fp16 before the P·V product, fp16 output.
//
// Round-1 implementation: mma.sync.m16n8k16 tensor-core MMA, 64-query x 64-key tiles, 4 warps, double-buffered
// cp.async K/V tiles in padded (conflict-free) shared memory.
#include "kernels.h"

#include "common.cuh"

namespace er {

constexpr int AQ = 64, AK = 64, ATT_THREADS = 128;

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

template <int D>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(AttnArgs a) {
    constexpr int RS = D * 2 + 16;               // padded row stride in bytes (conflict-free ldmatrix)
    constexpr int DC = D / 8;                    // 16-byte chunks per row
    constexpr int KD = D / 16;                   // k16 steps over the head dim
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sQ = smem;                    // AQ rows
    unsigned char* sK = sQ + AQ * RS;            // 2 stages x AK rows
    unsigned char* sV = sK + 2 * AK * RS;        // 2 stages x AK rows

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;
    const __half* qp = a.q + (size_t)b * a.q_bs + (size_t)h * D;
    const __half* kp = a.k + (size_t)b * a.k_bs + (size_t)h * D;
    const __half* vp = a.v + (size_t)b * a.v_bs + (size_t)h * D;

    // Q tile
    for (int c = tid; c < AQ * DC; c += ATT_THREADS) {
        const int r = c / DC, ch = c % DC;
        const bool ok = q0 + r < a.Nq;
        cpa16(s_u32(sQ + r * RS + ch * 16), qp + (size_t)(ok ? q0 + r : 0) * a.ldq + ch * 8, ok ? 16 : 0);
    }
    auto load_kv = [&](int stage, int kt) {
        const int k0 = kt * AK;
        for (int c = tid; c < AK * DC; c += ATT_THREADS) {
            const int r = c / DC, ch = c % DC;
            const bool ok = k0 + r < a.Nk;
            const size_t row = ok ? k0 + r : 0;
            cpa16(s_u32(sK + (stage * AK + r) * RS + ch * 16), kp + row * a.ldk + ch * 8, ok ? 16 : 0);
            cpa16(s_u32(sV + (stage * AK + r) * RS + ch * 16), vp + row * a.ldv + ch * 8, ok ? 16 : 0);
        }
    };
    int nkt = (a.Nk + AK - 1) / AK;
    if (a.causal) nkt = min(nkt, (q0 + AQ - 1) / AK + 1);
    load_kv(0, 0);
    asm volatile("cp.async.commit_group;");

    uint32_t qf[KD][4];
    float o[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; i++) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const float sl2 = rsqrtf((float)D) * 1.4426950408889634f;
    const int g = lane >> 2, tq = lane & 3;
    const int qrow0 = q0 + warp * 16 + g;       // this thread's rows: qrow0 and qrow0 + 8

    for (int kt = 0; kt < nkt; kt++) {
        if (kt + 1 < nkt) load_kv((kt + 1) & 1, kt + 1);
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 1;");
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int kd = 0; kd < KD; kd++) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldsm4(qf[kd], s_u32(sQ + r * RS + (kd * 2 + (lane >> 4)) * 16));
            }
        }
        const unsigned char* kS = sK + (kt & 1) * AK * RS;
        const unsigned char* vS = sV + (kt & 1) * AK * RS;
        float s[AK / 8][4];
#pragma unroll
        for (int i = 0; i < AK / 8; i++) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
#pragma unroll
        for (int kd = 0; kd < KD; kd++) {
#pragma unroll
            for (int nj = 0; nj < AK / 16; nj++) {
                uint32_t bfr[4];
                const int r = nj * 16 + (lane & 7) + (lane >> 4) * 8;
                ldsm4(bfr, s_u32(kS + r * RS + (kd * 2 + ((lane >> 3) & 1)) * 16));
                mma16816(s[nj * 2], qf[kd], bfr[0], bfr[1]);
                mma16816(s[nj * 2 + 1], qf[kd], bfr[2], bfr[3]);
            }
        }
        // mask + online softmax (rows g and g+8 of this warp's 16-row slab)
        const int kbase = kt * AK;
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int i = 0; i < AK / 8; i++) {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int key = kbase + i * 8 + tq * 2 + (e & 1);
                const int qr = qrow0 + (e >> 1) * 8;
                const bool ok = key < a.Nk && (!a.causal || key <= qr);
                s[i][e] = ok ? s[i][e] : -INFINITY;
                mx[e >> 1] = fmaxf(mx[e >> 1], s[i][e]);
            }
        }
        float corr[2], mnew[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            mnew[r] = fmaxf(m_run[r], mx[r]);
            corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f((m_run[r] - mnew[r]) * sl2);
            m_run[r] = mnew[r];
        }
        float ls[2] = {0.f, 0.f};
        uint32_t pf[AK / 16][4];
#pragma unroll
        for (int i = 0; i < AK / 8; i++) {
            float pv[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const float mm = mnew[e >> 1];
                pv[e] = (mm == -INFINITY) ? 0.f : exp2f((s[i][e] - mm) * sl2);
                ls[e >> 1] += pv[e];
            }
            // C-fragment of two adjacent n8 tiles == A-fragment of one k16 step
            pf[i >> 1][(i & 1) * 2 + 0] = pack_h2(pv[0], pv[1]);
            pf[i >> 1][(i & 1) * 2 + 1] = pack_h2(pv[2], pv[3]);
        }
#pragma unroll
        for (int r = 0; r < 2; r++) l_run[r] = l_run[r] * corr[r] + ls[r];
#pragma unroll
        for (int i = 0; i < D / 8; i++) { o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1]; }
        // O += P V
#pragma unroll
        for (int kk = 0; kk < AK / 16; kk++) {
#pragma unroll
            for (int dj = 0; dj < D / 16; dj++) {
                uint32_t bfr[4];
                const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                ldsm4t(bfr, s_u32(vS + r * RS + (dj * 2 + (lane >> 4)) * 16));
                mma16816(o[dj * 2], pf[kk], bfr[0], bfr[1]);
                mma16816(o[dj * 2 + 1], pf[kk], bfr[2], bfr[3]);
            }
        }
        __syncthreads();   // everyone done with this stage before it is refilled
    }
    // finalize: quad-reduce the row sums, normalise, store fp16
#pragma unroll
    for (int r = 0; r < 2; r++) {
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
        l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    __half* op = a.out + (size_t)b * a.o_bs + (size_t)h * D;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const int row = qrow0 + r * 8;
        if (row >= a.Nq) continue;
        const float inv = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
        if (a.lse2 && tq == 0) a.lse2[((size_t)b * a.H + h) * a.Nq + row] = m_run[r] * sl2 + log2f(l_run[r]);
#pragma unroll
        for (int i = 0; i < D / 8; i++) {
            const uint32_t v = pack_h2(o[i][r * 2] * inv, o[i][r * 2 + 1] * inv);
            *reinterpret_cast<uint32_t*>(op + (size_t)row * a.ldo + i * 8 + tq * 2) = v;
        }
    }
}

}  // namespace er

cudaError_t er_attention_tcgen05(const er::AttnArgs& a, cudaStream_t stream);   // attention_tcgen05.cu: the tensor-memory kernel (production path)
int g_er_dense_legacy = 0;      // er_debug_set(NULL, "dense_legacy", 1): force the mma.sync GEMM / attention kernels (A/B timing)

cudaError_t er_attention(const er::AttnArgs& a, cudaStream_t stream) {
    using namespace er;
    if (a.Nq <= 0 || a.Nk <= 0) return cudaSuccess;
    if (!g_er_dense_legacy) {
        const cudaError_t e = er_attention_tcgen05(a, stream);
        if (e != cudaErrorNotSupported) return e;
    }
    dim3 grid((a.Nq + AQ - 1) / AQ, a.H, a.B);
    if (a.D == 96) {
        const int smem = (AQ + 4 * AK) * (96 * 2 + 16);
        cudaFuncSetAttribute(attention_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);   // per call: per-device state, microseconds
        attention_kernel<96><<<grid, ATT_THREADS, smem, stream>>>(a);
    } else if (a.D == 64) {
        const int smem = (AQ + 4 * AK) * (64 * 2 + 16);
        cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attention_kernel<64><<<grid, ATT_THREADS, smem, stream>>>(a);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}
