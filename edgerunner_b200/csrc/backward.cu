// Backward-pass kernels of the training step (reference: `accelerator.backward(loss)` in main.py:172 = torch autograd over
// LMM.forward, models.py:147-202 -> modeling_opt.py:253-298, 464-517).  The dgrad / wgrad contractions run on the tcgen05 GEMM of
// gemm_tcgen05.cu (the transposes below put both operands K-major for it); this file holds everything around them:
//
//   transpose_f16_kernel      [R][Cn] -> [Cn][R] (zero-padded to a multiple of 64 columns: the GEMM's K loop then has no tail)
//   add_dropout_kernel        training-forward residual add: out32 = res32 + dropout(y16)          (modeling_opt.py:272-273, 285-287)
//   ln_bwd_kernel             LayerNorm input gradient (+ the dropout-branch gradient in fp16), row statistics
//   ln_param_partial_kernel   LayerNorm weight / bias gradients, column sums over row slabs (fixed order, no float atomics)
//   colsum_f16_partial_kernel bias gradients of the Linear layers
//   reduce_partials_kernel    second stage of both
//   relu_bwd_kernel, ce_bwd_kernel (softmax - onehot on the fp16-rounded logits, shifted labels), embedding gradients
//   attn_bwd_{stats,dq,dkv}_kernel   causal / dense flash-attention backward, warp-level tensor-core MMA (wmma m16n16k16, fp16 in, fp32
//                             accumulate): P is recomputed from the row log-sum-exp, dQ and (dK, dV) are produced by two kernels that each own
//                             their output rows (no atomics, deterministic).  First version: mma.sync class, not tcgen05.
//
// Gradients of activations are fp16 scaled by a static loss scale; weight gradients are fp32 (GEMM_F32 epilogue) and are unscaled when exported.
#include "kernels.h"

#include <mma.h>

#include "common.cuh"

namespace er {
namespace bw {

using namespace nvcuda;

// keep-mask of F.dropout: counter-based (splitmix64 of seed, site and element index), so that the recomputation of a checkpointed layer in the
// backward pass sees the mask of its forward pass.  NOT torch's Philox stream: masks are reproducible per seed, not equal to the reference's.
__device__ __forceinline__ bool drop_keep(unsigned long long seed, unsigned site, unsigned long long idx, unsigned thr) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1ull) + 0xD1B54A32D192ED03ull * (unsigned long long)(site + 1u);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (unsigned)(z >> 32) >= thr;
}

// ---- transpose -------------------------------------------------------------------------------------------------------------------------
// out[c][r] = in[r][c] for r < R, c < Cn; out[c][r] = 0 for R <= r < Rpad (Rpad = R rounded up to 64 <= ld_out).  64 x 64 tiles, block (32, 8).
__global__ void __launch_bounds__(256) transpose_f16_kernel(const __half* __restrict__ in, int R, int Cn, int ld_in, __half* __restrict__ out, int ld_out) {
    __shared__ __half tile[64][66];
    const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
    const int tx = threadIdx.x, ty = threadIdx.y;
    for (int j = ty; j < 64; j += 8) {
        const int r = r0 + j;
        for (int k = 0; k < 2; k++) {
            const int c = c0 + tx + 32 * k;
            tile[j][tx + 32 * k] = (r < R && c < Cn) ? in[(size_t)r * ld_in + c] : __float2half_rn(0.f);
        }
    }
    __syncthreads();
    for (int j = ty; j < 64; j += 8) {
        const int c = c0 + j;
        if (c >= Cn) continue;
        for (int k = 0; k < 2; k++) {
            const int r = r0 + tx + 32 * k;
            if (r < ld_out) out[(size_t)c * ld_out + r] = tile[tx + 32 * k][j];
        }
    }
}

// ---- training forward: residual + dropout(branch) ---------------------------------------------------------------------------------------
// F.dropout on the fp16 Linear output under autocast: kept elements are scaled by 1/(1-p) and rounded to fp16; the add runs in fp32.
__global__ void add_dropout_kernel(const float* __restrict__ res32, const __half* __restrict__ y16, float* __restrict__ out32, size_t n,
                                   unsigned thr, float inv_keep, unsigned long long seed, unsigned site) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float y = __half2float(y16[i]);
    if (thr) y = drop_keep(seed, site, i, thr) ? round_f16(y * inv_keep) : 0.f;
    out32[i] = res32[i] + y;
}

// ---- LayerNorm backward -------------------------------------------------------------------------------------------------------------------
// one warp per row.  y = (s - mean) * rstd * gamma + beta ;  g = dy * gamma ;  ds = rstd * (g - mean(g) - xhat * mean(g * xhat)).
// Writes ds32 (optional), the statistics (mean, rstd) for the parameter-gradient kernel, and (optional) the gradient of the dropout branch that
// was added to form s:  dbr16 = f16( keep ? ds / (1-p) : 0 )  over the dense [M][C] index space (ld_s == C for that use).
__global__ void ln_bwd_kernel(const float* __restrict__ dy32, const float* s32, const __half* s16, int ld_s, const __half* __restrict__ gamma, float* ds32,
                              __half* dbr16, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int C, unsigned thr, float inv_keep,
                              unsigned long long seed, unsigned site) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    auto ld = [&](int i) -> float { return s32 ? s32[(size_t)row * ld_s + i] : __half2float(s16[(size_t)row * ld_s + i]); };
    float a = 0.f;
    for (int i = lane; i < C; i += 32) a += ld(i);
    const float mean = warp_sum(a) / C;
    float q = 0.f;
    for (int i = lane; i < C; i += 32) { const float d = ld(i) - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / C + 1e-5f);
    float sg = 0.f, sgx = 0.f;
    for (int i = lane; i < C; i += 32) {
        const float g = dy32[(size_t)row * C + i] * __half2float(gamma[i]);
        sg += g;
        sgx += g * (ld(i) - mean) * rstd;
    }
    sg = warp_sum(sg) / C;
    sgx = warp_sum(sgx) / C;
    for (int i = lane; i < C; i += 32) {
        const float g = dy32[(size_t)row * C + i] * __half2float(gamma[i]);
        const float xh = (ld(i) - mean) * rstd;
        const float d = rstd * (g - sg - xh * sgx);
        const size_t o = (size_t)row * C + i;
        if (ds32) ds32[o] = d;
        if (dbr16) {
            float b = d;
            if (thr) b = drop_keep(seed, site, o, thr) ? d * inv_keep : 0.f;
            dbr16[o] = __float2half_rn(b);
        }
    }
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// partial[slab][0][c] = sum over the slab's rows of dy * xhat (d gamma), partial[slab][1][c] = sum of dy (d beta).  grid (ceil(C/256), nslab)
__global__ void __launch_bounds__(256) ln_param_partial_kernel(const float* __restrict__ dy32, const float* s32, const __half* s16, int ld_s,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd, int M, int C,
                                                                float* __restrict__ partial) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int nslab = gridDim.y, slab = blockIdx.y;
    const int per = (M + nslab - 1) / nslab;
    const int r0 = slab * per, r1 = min(M, r0 + per);
    if (c >= C) return;
    float dg = 0.f, db = 0.f;
    for (int r = r0; r < r1; r++) {
        const float s = s32 ? s32[(size_t)r * ld_s + c] : __half2float(s16[(size_t)r * ld_s + c]);
        const float d = dy32[(size_t)r * C + c];
        dg += d * (s - mean[r]) * rstd[r];
        db += d;
    }
    partial[((size_t)slab * 2 + 0) * C + c] = dg;
    partial[((size_t)slab * 2 + 1) * C + c] = db;
}

// partial[slab][c] = sum over the slab's rows of x16[r][c].  grid (ceil(ncols/256), nslab)
__global__ void __launch_bounds__(256) colsum_f16_partial_kernel(const __half* __restrict__ x16, int ld, int M, int ncols, float* __restrict__ partial) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int nslab = gridDim.y, slab = blockIdx.y;
    const int per = (M + nslab - 1) / nslab;
    const int r0 = slab * per, r1 = min(M, r0 + per);
    if (c >= ncols) return;
    float a = 0.f;
    for (int r = r0; r < r1; r++) a += __half2float(x16[(size_t)r * ld + c]);
    partial[(size_t)slab * ncols + c] = a;
}

// out[c] (+)= sum over slabs (in slab order) of partial[slab * stride + c]
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int nslab, size_t stride, int ncols, float* __restrict__ out, int accumulate) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncols) return;
    float a = 0.f;
    for (int s = 0; s < nslab; s++) a += partial[(size_t)s * stride + c];
    out[c] = accumulate ? out[c] + a : a;
}

// ---- ReLU backward (in place on the fp16 gradient; 8 elements per thread) -----------------------------------------------------------------
__global__ void relu_bwd_kernel(__half* __restrict__ dh16, const __half* __restrict__ h16, size_t nvec) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nvec) return;
    uint4 d = reinterpret_cast<uint4*>(dh16)[i];
    const uint4 h = reinterpret_cast<const uint4*>(h16)[i];
    __half* dp = reinterpret_cast<__half*>(&d);
    const __half* hp = reinterpret_cast<const __half*>(&h);
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (!(__half2float(hp[k]) > 0.f)) dp[k] = __float2half_rn(0.f);
    reinterpret_cast<uint4*>(dh16)[i] = d;
}

// ---- cross-entropy backward (modeling_opt.py:500-505: row i of a sample predicts label i+1, ignore_index -100, mean over valid rows) -----------
// dl[r][v] = f16( loss_scale / count * (softmax(round_f16(logits[r]))[v] - [v == label]) ), zero rows for ignored / last rows; the pad
// columns [V, ldo) are zero.  One warp per row of the [B * N] rows.
__global__ void ce_bwd_kernel(const float* __restrict__ logits_pre, int ld, const int64_t* __restrict__ labels, int B, int N, int V, const int* __restrict__ count,
                              float loss_scale, __half* __restrict__ dl, int ldo) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= B * N) return;
    const int i = row % N;
    const long long lab = (i + 1 < N) ? labels[(size_t)row + 1] : -100;
    __half* o = dl + (size_t)row * ldo;
    if (lab < 0 || lab >= V) {
        for (int v = lane; v < ldo; v += 32) o[v] = __float2half_rn(0.f);
        return;
    }
    const float* x = logits_pre + (size_t)row * ld;
    float mx = -INFINITY;
    for (int v = lane; v < V; v += 32) mx = fmaxf(mx, round_f16(x[v]));
    mx = warp_max(mx);
    float s = 0.f;
    for (int v = lane; v < V; v += 32) s += expf(round_f16(x[v]) - mx);
    s = warp_sum(s);
    const float sc = loss_scale / (float)max(*count, 1);
    const float inv = 1.f / s;
    for (int v = lane; v < ldo; v += 32) {
        float g = 0.f;
        if (v < V) g = (expf(round_f16(x[v]) - mx) * inv - (v == (int)lab ? 1.f : 0.f)) * sc;
        o[v] = __float2half_rn(g);
    }
}

// ---- embedding gradients (models.py:228-233, modeling_opt.py:355-357) ------------------------------------------------------------------------------
// d embed_positions[n] = sum over samples of dx0[b*N + n]
__global__ void pos_grad_kernel(const float* __restrict__ dx0, int B, int N, int C, float* __restrict__ dpos) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * C) return;
    float a = 0.f;
    for (int b = 0; b < B; b++) a += dx0[(size_t)b * N * C + i];
    dpos[i] = a;
}
// d embd[v] = sum (in token order) of the dx0 rows whose token is v.  grid (V, ceil(C/256)); ids [B][T], token (b, t) sits in row b*N + P + t
__global__ void __launch_bounds__(256) embd_grad_kernel(const float* __restrict__ dx0, const int32_t* __restrict__ ids, int B, int T, int N, int P, int C,
                                                         float* __restrict__ dembd) {
    const int v = blockIdx.x, c = blockIdx.y * 256 + threadIdx.x;
    if (c >= C) return;
    float a = 0.f;
    for (int b = 0; b < B; b++)
        for (int t = 0; t < T; t++)
            if (ids[(size_t)b * T + t] == v) a += dx0[((size_t)b * N + P + t) * C + c];
    dembd[(size_t)v * C + c] = a;
}
// d embed_num_face[k] = sum of the dx0 rows (b*N + row) of the samples whose bucket is k.  grid (10, ceil(C/256))
__global__ void __launch_bounds__(256) numface_grad_kernel(const float* __restrict__ dx0, const int32_t* __restrict__ bucket, int B, int N, int row, int C,
                                                            float* __restrict__ denf) {
    const int k = blockIdx.x, c = blockIdx.y * 256 + threadIdx.x;
    if (c >= C) return;
    float a = 0.f;
    for (int b = 0; b < B; b++)
        if (bucket[b] == k) a += dx0[((size_t)b * N + row) * C + c];
    denf[(size_t)k * C + c] = a;
}

// dst[r][c] = src[r * ld + c] * scale   (gradient export: un-applies the loss scale, drops the row padding)
__global__ void export_f32_kernel(const float* __restrict__ src, int ld, int rows, int cols, float scale, float* __restrict__ dst) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    dst[i] = src[(size_t)r * ld + c] * scale;
}

// ---- GEGLU backward (point.py:69-72: h = [a | g], out = a * gelu_erf(g)) ------------------------------------------------------------------------------
// dh [M][2F] from dout [M][F]: da = dout * gelu(g), dg = dout * a * gelu'(g), gelu'(g) = Phi(g) + g phi(g)
__global__ void geglu_bwd_kernel(const __half* __restrict__ h, const __half* __restrict__ dout, __half* __restrict__ dh, int M, int F) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)M * F) return;
    const size_t r = idx / F, c = idx % F;
    const float a = __half2float(h[r * 2 * F + c]), g = __half2float(h[r * 2 * F + F + c]), d = __half2float(dout[idx]);
    const float cdf = 0.5f * (1.f + erff(g * 0.70710678118654752f));
    const float pdf = 0.3989422804014327f * expf(-0.5f * g * g);
    dh[r * 2 * F + c] = __float2half_rn(d * g * cdf);
    dh[r * 2 * F + F + c] = __float2half_rn(d * a * (cdf + g * pdf));
}
__global__ void add_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}
// dst[r][c] = f16(dst[r][c] + alpha * src[r][c])  (the KL term's gradient, kl_weight * latent, added to the latent gradient)
__global__ void axpy_f16_kernel(__half* __restrict__ dst, int ld_dst, const __half* __restrict__ src, int ld_src, int rows, int cols, float alpha) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * cols) return;
    const int r = (int)(i / cols), c = (int)(i % cols);
    const size_t o = (size_t)r * ld_dst + c;
    dst[o] = __float2half_rn(__half2float(dst[o]) + alpha * __half2float(src[(size_t)r * ld_src + c]));
}

// ---- flash-attention backward ----------------------------------------------------------------------------------------------------------------------------
// Geometry shared by the three kernels: CTA = one 64-row tile of one (batch, head), 4 warps x 16 rows.  Tiles live in shared memory as
// [64][D + 8] fp16 (row pitch 16-byte aligned, wmma ldm multiple of 8); every warp has a private fp32 scratch for the accumulator tiles it has to
// post-process elementwise (wmma fragments have no portable element layout, so S / dP go through shared memory).
constexpr int BT = 64;          // tile rows (queries or keys)
constexpr int SLD = BT + 4;     // fp32 scratch pitch
constexpr int HLD = BT + 8;     // fp16 scratch pitch

template <int D>
__device__ __forceinline__ void load_tile(__half* dst, const __half* src, int ld, int row0, int nrows_total) {
    // 64 rows x D fp16 -> dst [64][D + 8]; rows >= nrows_total are zero.  128 threads, 16-byte vectors.
    constexpr int VPR = D / 8;
    for (int i = threadIdx.x; i < BT * VPR; i += 128) {
        const int r = i / VPR, vv = i % VPR;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (row0 + r < nrows_total) val = *reinterpret_cast<const uint4*>(src + (size_t)(row0 + r) * ld + vv * 8);
        *reinterpret_cast<uint4*>(dst + r * (D + 8) + vv * 8) = val;
    }
}

// acc[16 x 64] (fp32, to scratch with pitch SLD) = A_w[16 x D] * Bt[64 x D]^T ; A_w = the warp's 16 rows of a tile, Bt = a whole tile
template <int D>
__device__ __forceinline__ void mma_rows_by_tile_t(const __half* a_rows, const __half* bt, float* scratch) {
    constexpr int LDT = D + 8;
#pragma unroll
    for (int n = 0; n < BT / 16; n++) {
        wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
        wmma::fill_fragment(acc, 0.f);
#pragma unroll
        for (int kk = 0; kk < D / 16; kk++) {
            wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> a;
            wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::col_major> b;
            wmma::load_matrix_sync(a, a_rows + kk * 16, LDT);
            wmma::load_matrix_sync(b, bt + n * 16 * LDT + kk * 16, LDT);      // element (k, n) at bt[n * LDT + k]
            wmma::mma_sync(acc, a, b, acc);
        }
        wmma::store_matrix_sync(scratch + n * 16, acc, SLD, wmma::mem_row_major);
    }
}

// acc[n] (16 x 16 blocks of a [16 x D] accumulator) += P_w[16 x 64] (fp16 scratch, pitch HLD) * Bm[64 x D] (a whole tile, row-major)
template <int D>
__device__ __forceinline__ void mma_acc_rows(wmma::fragment<wmma::accumulator, 16, 16, 16, float>* acc, const __half* p_rows, const __half* bm) {
    constexpr int LDT = D + 8;
#pragma unroll
    for (int kk = 0; kk < BT / 16; kk++) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> a;
        wmma::load_matrix_sync(a, p_rows + kk * 16, HLD);
#pragma unroll
        for (int n = 0; n < D / 16; n++) {
            wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> b;
            wmma::load_matrix_sync(b, bm + kk * 16 * LDT + n * 16, LDT);
            wmma::mma_sync(acc[n], a, b, acc[n]);
        }
    }
}

// the warp's [16 x D] accumulator -> fp16 rows of the output (through the warp's fp32 scratch, pitch D + 4; needs 16 * (D + 4) floats)
template <int D>
__device__ __forceinline__ void store_rows(wmma::fragment<wmma::accumulator, 16, 16, 16, float>* acc, float* scratch, __half* dst, int ld, int row0,
                                           int nrows_total) {
    constexpr int OLD = D + 4;
    const int lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int n = 0; n < D / 16; n++) wmma::store_matrix_sync(scratch + n * 16, acc[n], OLD, wmma::mem_row_major);
    __syncwarp();
    for (int i = lane; i < 16 * (D / 2); i += 32) {
        const int r = i / (D / 2), c = (i % (D / 2)) * 2;
        if (row0 + r < nrows_total)
            *reinterpret_cast<__half2*>(dst + (size_t)(row0 + r) * ld + c) = __floats2half2_rn(scratch[r * OLD + c], scratch[r * OLD + c + 1]);
    }
    __syncwarp();
}

// per-warp scratch: S (fp32 16 x SLD) | dP (fp32 16 x SLD) | two fp16 [16 x HLD] blocks.  2 * 16 * SLD floats >= 16 * (D + 4) for D <= 128.
constexpr int WARP_SCRATCH_BYTES = 2 * 16 * SLD * 4 + 2 * 16 * HLD * 2;

template <int D>
constexpr int attn_bwd_smem() { return 4 * BT * (D + 8) * 2 + 4 * WARP_SCRATCH_BYTES + 2 * BT * 4 + 128; }

// ---- pass 0: row statistics.  lse2[r] = log2 sum_j exp2(scale_log2 * q_r . k_j) over the visible keys; dsum[r] = sum_d dO[r][d] * O[r][d] ------------------
template <int D>
__global__ void __launch_bounds__(128) attn_bwd_stats_kernel(const AttnBwdArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int LDT = D + 8;
    __half* Qs = reinterpret_cast<__half*>(smem_raw);
    __half* Ks = Qs + BT * LDT;
    float* scr = reinterpret_cast<float*>(smem_raw + 4 * BT * LDT * 2) + (threadIdx.x >> 5) * (WARP_SCRATCH_BYTES / 4);
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __half* qb = p.q + (size_t)b * p.q_bs + h * D;
    const __half* kb = p.k + (size_t)b * p.k_bs + h * D;
    load_tile<D>(Qs, qb, p.ld_qkv_q, qt * BT, p.Nq);
    const int row = lane >> 1, half = lane & 1;
    const int row_g = qt * BT + warp * 16 + row;
    float m = -INFINITY, l = 0.f;
    const int nkt = p.causal ? min((p.Nk + BT - 1) / BT, qt + 1) : (p.Nk + BT - 1) / BT;
    for (int kt = 0; kt < nkt; kt++) {
        __syncthreads();
        load_tile<D>(Ks, kb, p.ld_qkv_k, kt * BT, p.Nk);
        __syncthreads();
        mma_rows_by_tile_t<D>(Qs + warp * 16 * LDT, Ks, scr);
        __syncwarp();
        float sv[32];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; c++) {
            const int cc = 2 * c + half, col_g = kt * BT + cc;     // the two lanes of a row interleave over the columns (adjacent banks)
            const bool ok = col_g < p.Nk && (!p.causal || col_g <= row_g);
            sv[c] = ok ? scr[row * SLD + cc] * p.scale_log2 : -INFINITY;
            mx = fmaxf(mx, sv[c]);
        }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        const float mn = fmaxf(m, mx);
        float s = 0.f;
        if (mn > -INFINITY) {
#pragma unroll
            for (int c = 0; c < 32; c++) s += exp2f(sv[c] - mn);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        l = (m > -INFINITY ? l * exp2f(m - mn) : 0.f) + s;
        m = mn;
        __syncwarp();
    }
    // rowsum(dO * O): the two lanes of a row take half of the head dimension each (no divergence around the shuffle)
    const bool row_ok = row_g < p.Nq;
    float ds = 0.f;
    if (row_ok) {
        const __half* op = p.o + (size_t)b * p.o_bs + (size_t)row_g * p.ld_o + h * D + half * (D / 2);
        const __half* dp = p.dout + (size_t)b * p.o_bs + (size_t)row_g * p.ld_o + h * D + half * (D / 2);
        for (int c = 0; c < D / 2; c++) ds += __half2float(op[c]) * __half2float(dp[c]);
    }
    ds += __shfl_xor_sync(0xffffffffu, ds, 1);
    if (row_ok && half == 0) {
        const size_t o = ((size_t)b * p.H + h) * p.Nq + row_g;
        p.lse2[o] = m + log2f(l);
        p.dsum[o] = ds;
    }
}

// ---- pass 1: dQ.  dS = P * (dP - dsum) * scale with P = exp2(scale_log2 * S - lse2), dP = dO V^T ;  dQ = dS K ---------------------------------------
template <int D>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const AttnBwdArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int LDT = D + 8;
    __half* Qs = reinterpret_cast<__half*>(smem_raw);
    __half* dOs = Qs + BT * LDT;
    __half* Ks = dOs + BT * LDT;
    __half* Vs = Ks + BT * LDT;
    unsigned char* wbase = smem_raw + 4 * BT * LDT * 2 + (threadIdx.x >> 5) * WARP_SCRATCH_BYTES;
    float* Sf = reinterpret_cast<float*>(wbase);
    float* dPf = Sf + 16 * SLD;
    __half* dSh = reinterpret_cast<__half*>(wbase + 2 * 16 * SLD * 4);
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __half* qb = p.q + (size_t)b * p.q_bs + h * D;
    const __half* kb = p.k + (size_t)b * p.k_bs + h * D;
    const __half* vb = p.v + (size_t)b * p.v_bs + h * D;
    const __half* dob = p.dout + (size_t)b * p.o_bs + h * D;
    load_tile<D>(Qs, qb, p.ld_qkv_q, qt * BT, p.Nq);
    load_tile<D>(dOs, dob, p.ld_o, qt * BT, p.Nq);
    const int row = lane >> 1, half = lane & 1;
    const int row_g = qt * BT + warp * 16 + row;
    const bool row_ok = row_g < p.Nq;
    const size_t so = ((size_t)b * p.H + h) * p.Nq + (row_ok ? row_g : 0);
    const float lse = row_ok ? p.lse2[so] : 0.f, dsm = row_ok ? p.dsum[so] : 0.f;
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[D / 16];
#pragma unroll
    for (int n = 0; n < D / 16; n++) wmma::fill_fragment(acc[n], 0.f);
    const int nkt = p.causal ? min((p.Nk + BT - 1) / BT, qt + 1) : (p.Nk + BT - 1) / BT;
    for (int kt = 0; kt < nkt; kt++) {
        __syncthreads();
        load_tile<D>(Ks, kb, p.ld_qkv_k, kt * BT, p.Nk);
        load_tile<D>(Vs, vb, p.ld_qkv_v, kt * BT, p.Nk);
        __syncthreads();
        mma_rows_by_tile_t<D>(Qs + warp * 16 * LDT, Ks, Sf);
        mma_rows_by_tile_t<D>(dOs + warp * 16 * LDT, Vs, dPf);
        __syncwarp();
#pragma unroll 8
        for (int c = 0; c < 32; c++) {
            const int cc = 2 * c + half, col_g = kt * BT + cc;
            const bool ok = row_ok && col_g < p.Nk && (!p.causal || col_g <= row_g);
            float ds = 0.f;
            if (ok) {
                const float pr = exp2f(Sf[row * SLD + cc] * p.scale_log2 - lse);
                ds = pr * (dPf[row * SLD + cc] - dsm) * p.scale;
            }
            dSh[row * HLD + cc] = __float2half_rn(ds);
        }
        __syncwarp();
        mma_acc_rows<D>(acc, dSh, Ks);
    }
    store_rows<D>(acc, Sf, p.dq + (size_t)b * p.dq_bs + h * D, p.ld_dq, qt * BT + warp * 16, p.Nq);
}

// ---- pass 2: dK, dV.  CTA = one key tile; streams the query tiles that see it.  Works on transposed blocks: S^T = K Q^T, dP^T = V dO^T -----------------
template <int D>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const AttnBwdArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int LDT = D + 8;
    __half* Ks = reinterpret_cast<__half*>(smem_raw);
    __half* Vs = Ks + BT * LDT;
    __half* Qs = Vs + BT * LDT;
    __half* dOs = Qs + BT * LDT;
    unsigned char* wbase = smem_raw + 4 * BT * LDT * 2 + (threadIdx.x >> 5) * WARP_SCRATCH_BYTES;
    float* STf = reinterpret_cast<float*>(wbase);
    float* dPTf = STf + 16 * SLD;
    __half* PTh = reinterpret_cast<__half*>(wbase + 2 * 16 * SLD * 4);
    __half* dSTh = PTh + 16 * HLD;
    float* lse_s = reinterpret_cast<float*>(smem_raw + 4 * BT * LDT * 2 + 4 * WARP_SCRATCH_BYTES);
    float* dsm_s = lse_s + BT;
    const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __half* qb = p.q + (size_t)b * p.q_bs + h * D;
    const __half* kb = p.k + (size_t)b * p.k_bs + h * D;
    const __half* vb = p.v + (size_t)b * p.v_bs + h * D;
    const __half* dob = p.dout + (size_t)b * p.o_bs + h * D;
    load_tile<D>(Ks, kb, p.ld_qkv_k, kt * BT, p.Nk);
    load_tile<D>(Vs, vb, p.ld_qkv_v, kt * BT, p.Nk);
    const int krow = lane >> 1, half = lane & 1;
    const int key_g = kt * BT + warp * 16 + krow;
    const bool key_ok = key_g < p.Nk;
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> dk[D / 16], dv[D / 16];
#pragma unroll
    for (int n = 0; n < D / 16; n++) { wmma::fill_fragment(dk[n], 0.f); wmma::fill_fragment(dv[n], 0.f); }
    const int nqt = (p.Nq + BT - 1) / BT;
    for (int qt = p.causal ? kt : 0; qt < nqt; qt++) {
        __syncthreads();
        load_tile<D>(Qs, qb, p.ld_qkv_q, qt * BT, p.Nq);
        load_tile<D>(dOs, dob, p.ld_o, qt * BT, p.Nq);
        if (threadIdx.x < BT) {
            const int r = qt * BT + threadIdx.x;
            const size_t so = ((size_t)b * p.H + h) * p.Nq + (r < p.Nq ? r : 0);
            lse_s[threadIdx.x] = r < p.Nq ? p.lse2[so] : 0.f;
            dsm_s[threadIdx.x] = r < p.Nq ? p.dsum[so] : 0.f;
        }
        __syncthreads();
        mma_rows_by_tile_t<D>(Ks + warp * 16 * LDT, Qs, STf);
        mma_rows_by_tile_t<D>(Vs + warp * 16 * LDT, dOs, dPTf);
        __syncwarp();
#pragma unroll 8
        for (int c = 0; c < 32; c++) {
            const int cc = 2 * c + half, q_g = qt * BT + cc;
            const bool ok = key_ok && q_g < p.Nq && (!p.causal || key_g <= q_g);
            float pr = 0.f, ds = 0.f;
            if (ok) {
                pr = exp2f(STf[krow * SLD + cc] * p.scale_log2 - lse_s[cc]);
                ds = pr * (dPTf[krow * SLD + cc] - dsm_s[cc]) * p.scale;
            }
            PTh[krow * HLD + cc] = __float2half_rn(pr);
            dSTh[krow * HLD + cc] = __float2half_rn(ds);
        }
        __syncwarp();
        mma_acc_rows<D>(dv, PTh, dOs);
        mma_acc_rows<D>(dk, dSTh, Qs);
    }
    store_rows<D>(dk, STf, p.dk + (size_t)b * p.dk_bs + h * D, p.ld_dk, kt * BT + warp * 16, p.Nk);
    store_rows<D>(dv, STf, p.dv + (size_t)b * p.dv_bs + h * D, p.ld_dv, kt * BT + warp * 16, p.Nk);
}

// dsum[b][h][r] = sum_d dO[r][h*D + d] * O[r][h*D + d]: one warp per (row, head)
template <int D>
__global__ void attn_rowdot_kernel(const AttnBwdArgs p) {
    const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const size_t total = (size_t)p.B * p.Nq * p.H;
    if (w >= total) return;
    const int h = (int)(w % p.H);
    const size_t br = w / p.H;
    const int r = (int)(br % p.Nq), b = (int)(br / p.Nq);
    const __half* op = p.o + (size_t)b * p.o_bs + (size_t)r * p.ld_o + h * D;
    const __half* dp = p.dout + (size_t)b * p.o_bs + (size_t)r * p.ld_o + h * D;
    float a = 0.f;
    for (int c = lane * 2; c < D; c += 64) {
        const float2 x = __half22float2(*reinterpret_cast<const __half2*>(op + c)), y = __half22float2(*reinterpret_cast<const __half2*>(dp + c));
        a += x.x * y.x + x.y * y.y;
    }
    a = warp_sum(a);
    if (lane == 0) p.dsum[((size_t)b * p.H + h) * p.Nq + r] = a;
}

template <int D>
static cudaError_t launch_attn_bwd(const AttnBwdArgs& p, cudaStream_t st) {
    const int smem = attn_bwd_smem<D>();
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(attn_bwd_stats_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(attn_bwd_dq_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)) != cudaSuccess) return e;
    const dim3 gq((p.Nq + BT - 1) / BT, p.H, p.B), gk((p.Nk + BT - 1) / BT, p.H, p.B);
    if (!p.have_lse) attn_bwd_stats_kernel<D><<<gq, 128, smem, st>>>(p);
    else attn_rowdot_kernel<D><<<(unsigned)(((size_t)p.B * p.Nq * p.H + 7) / 8), 256, 0, st>>>(p);
    attn_bwd_dq_kernel<D><<<gq, 128, smem, st>>>(p);
    attn_bwd_dkv_kernel<D><<<gk, 128, smem, st>>>(p);
    return cudaGetLastError();
}

}  // namespace bw
}  // namespace er

using namespace er;
using namespace er::bw;

cudaError_t er_transpose_f16(const __half* in, int R, int Cn, int ld_in, __half* out, int ld_out, cudaStream_t st) {
    if (R <= 0 || Cn <= 0) return cudaSuccess;
    if (ld_out < (R + 63) / 64 * 64) return cudaErrorInvalidValue;
    const dim3 grid((Cn + 63) / 64, (R + 63) / 64), block(32, 8);
    transpose_f16_kernel<<<grid, block, 0, st>>>(in, R, Cn, ld_in, out, ld_out);
    return cudaGetLastError();
}
cudaError_t er_add_dropout(const float* res32, const __half* y16, float* out32, size_t n, float p, unsigned long long seed, unsigned site, cudaStream_t st) {
    if (!n) return cudaSuccess;
    const unsigned thr = p > 0.f ? (unsigned)fmin((double)p * 4294967296.0, 4294967295.0) : 0u;
    add_dropout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(res32, y16, out32, n, thr, p > 0.f ? 1.f / (1.f - p) : 1.f, seed, site);
    return cudaGetLastError();
}
cudaError_t er_ln_bwd(const float* dy32, const float* s32, const __half* s16, int ld_s, const __half* gamma, float* ds32, __half* dbr16, float* mean,
                      float* rstd, int M, int C, float p, unsigned long long seed, unsigned site, cudaStream_t st) {
    if (M <= 0) return cudaSuccess;
    const unsigned thr = p > 0.f ? (unsigned)fmin((double)p * 4294967296.0, 4294967295.0) : 0u;
    ln_bwd_kernel<<<(M + 7) / 8, 256, 0, st>>>(dy32, s32, s16, ld_s, gamma, ds32, dbr16, mean, rstd, M, C, thr, p > 0.f ? 1.f / (1.f - p) : 1.f, seed, site);
    return cudaGetLastError();
}
cudaError_t er_ln_param_grad(const float* dy32, const float* s32, const __half* s16, int ld_s, const float* mean, const float* rstd, int M, int C,
                             float* partial /*[ER_BW_SLABS][2][C]*/, float* dgamma, float* dbeta, cudaStream_t st, int accumulate) {
    if (M <= 0) return cudaSuccess;
    const dim3 grid((C + 255) / 256, ER_BW_SLABS);
    ln_param_partial_kernel<<<grid, 256, 0, st>>>(dy32, s32, s16, ld_s, mean, rstd, M, C, partial);
    reduce_partials_kernel<<<(C + 255) / 256, 256, 0, st>>>(partial, ER_BW_SLABS, (size_t)2 * C, C, dgamma, accumulate);
    reduce_partials_kernel<<<(C + 255) / 256, 256, 0, st>>>(partial + C, ER_BW_SLABS, (size_t)2 * C, C, dbeta, accumulate);
    return cudaGetLastError();
}
cudaError_t er_colsum_f16(const __half* x16, int ld, int M, int ncols, float* partial /*[ER_BW_SLABS][ncols]*/, float* out, cudaStream_t st, int accumulate) {
    if (M <= 0 || ncols <= 0) return cudaSuccess;
    const dim3 grid((ncols + 255) / 256, ER_BW_SLABS);
    colsum_f16_partial_kernel<<<grid, 256, 0, st>>>(x16, ld, M, ncols, partial);
    reduce_partials_kernel<<<(ncols + 255) / 256, 256, 0, st>>>(partial, ER_BW_SLABS, (size_t)ncols, ncols, out, accumulate);
    return cudaGetLastError();
}
cudaError_t er_geglu_bwd(const __half* h, const __half* dout, __half* dh, int M, int F, cudaStream_t st) {
    const size_t n = (size_t)M * F;
    if (!n) return cudaSuccess;
    geglu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h, dout, dh, M, F);
    return cudaGetLastError();
}
cudaError_t er_add_f32(float* dst, const float* src, size_t n, cudaStream_t st) {
    if (!n) return cudaSuccess;
    add_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, src, n);
    return cudaGetLastError();
}
cudaError_t er_axpy_f16(__half* dst, int ld_dst, const __half* src, int ld_src, int rows, int cols, float alpha, cudaStream_t st) {
    const size_t n = (size_t)rows * cols;
    if (!n) return cudaSuccess;
    axpy_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, ld_dst, src, ld_src, rows, cols, alpha);
    return cudaGetLastError();
}
cudaError_t er_relu_bwd(__half* dh16, const __half* h16, size_t n, cudaStream_t st) {
    if (n & 7) return cudaErrorInvalidValue;
    const size_t nvec = n >> 3;
    if (!nvec) return cudaSuccess;
    relu_bwd_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, st>>>(dh16, h16, nvec);
    return cudaGetLastError();
}
cudaError_t er_ce_bwd(const float* logits_pre, int ld, const int64_t* labels, int B, int N, int V, const int* count_dev, float loss_scale, __half* dl, int ldo,
                      cudaStream_t st) {
    const int M = B * N;
    if (M <= 0) return cudaSuccess;
    ce_bwd_kernel<<<(M + 7) / 8, 256, 0, st>>>(logits_pre, ld, labels, B, N, V, count_dev, loss_scale, dl, ldo);
    return cudaGetLastError();
}
cudaError_t er_embed_bwd(const float* dx0, const int32_t* ids, const int32_t* bucket_dev, int B, int T, int N, int P, int C, int V, int numface_row,
                         float* dpos, float* dembd, float* denf, cudaStream_t st) {
    const size_t n = (size_t)N * C;
    pos_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dx0, B, N, C, dpos);
    embd_grad_kernel<<<dim3(V, (C + 255) / 256), 256, 0, st>>>(dx0, ids, B, T, N, P, C, dembd);
    if (denf) numface_grad_kernel<<<dim3(10, (C + 255) / 256), 256, 0, st>>>(dx0, bucket_dev, B, N, numface_row, C, denf);
    return cudaGetLastError();
}
cudaError_t er_export_f32(const float* src, int ld, int rows, int cols, float scale, float* dst, cudaStream_t st) {
    const size_t n = (size_t)rows * cols;
    if (!n) return cudaSuccess;
    export_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, ld, rows, cols, scale, dst);
    return cudaGetLastError();
}

// wmma version (first implementation; kept as the A/B reference of the register-resident kernels in attention_bwd_mma.cu)
cudaError_t er_attn_bwd_wmma(const er::AttnBwdArgs& p, int D, cudaStream_t st) { return D == 96 ? launch_attn_bwd<96>(p, st) : launch_attn_bwd<64>(p, st); }

int g_er_attn_bwd_wmma = 0;     // er_debug_set(NULL, "attn_bwd_wmma", 1): use the wmma kernels of this file

cudaError_t er_attn_rowdot(const er::AttnBwdArgs& p, int D, cudaStream_t st) {
    const unsigned grid = (unsigned)(((size_t)p.B * p.Nq * p.H + 7) / 8);
    if (D == 96) attn_rowdot_kernel<96><<<grid, 256, 0, st>>>(p); else attn_rowdot_kernel<64><<<grid, 256, 0, st>>>(p);
    return cudaGetLastError();
}

cudaError_t er_attention_bwd(const er::AttnArgs& a, const __half* dout, __half* dq, __half* dk, __half* dv, int ld_dq, int ld_dk, int ld_dv, long long dq_bs,
                             long long dk_bs, long long dv_bs, float* lse2, float* dsum, cudaStream_t st) {
    if (a.B <= 0 || a.Nq <= 0 || a.Nk <= 0) return cudaSuccess;
    if (a.D != 64 && a.D != 96) return cudaErrorInvalidValue;
    if ((a.ldq | a.ldk | a.ldv | a.ldo | ld_dq | ld_dk | ld_dv) & 7) return cudaErrorInvalidValue;     // 16-byte vector loads, half2 stores
    if (a.causal && a.Nq != a.Nk) return cudaErrorInvalidValue;
    AttnBwdArgs p{};
    p.q = a.q; p.k = a.k; p.v = a.v; p.o = a.out; p.dout = dout; p.dq = dq; p.dk = dk; p.dv = dv; p.lse2 = lse2; p.dsum = dsum;
    p.q_bs = a.q_bs; p.k_bs = a.k_bs; p.v_bs = a.v_bs; p.o_bs = a.o_bs; p.dq_bs = dq_bs; p.dk_bs = dk_bs; p.dv_bs = dv_bs;
    p.ld_qkv_q = a.ldq; p.ld_qkv_k = a.ldk; p.ld_qkv_v = a.ldv; p.ld_o = a.ldo; p.ld_dq = ld_dq; p.ld_dk = ld_dk; p.ld_dv = ld_dv;
    p.B = a.B; p.H = a.H; p.Nq = a.Nq; p.Nk = a.Nk; p.causal = a.causal;
    p.scale = 1.f / sqrtf((float)a.D);
    p.scale_log2 = p.scale * 1.4426950408889634f;
    p.have_lse = a.lse2 != nullptr && a.lse2 == lse2;      // the forward described by `a` already wrote the row statistic into the same buffer
    return g_er_attn_bwd_wmma ? er_attn_bwd_wmma(p, a.D, st) : er_attn_bwd_mma(p, a.D, st);
}
