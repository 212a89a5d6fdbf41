// Host-callable launchers of the dense (multi-row) kernels: prefill, point encoder, teacher-forced forward.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace er {

enum GemmMode {
    GEMM_F16 = 0,        // out16 = f16(acc + bias)
    GEMM_F16_RELU = 1,   // out16 = relu(f16(acc + bias))
    GEMM_F16_RES16 = 2,  // out16 = f16( f16(acc + bias) + res16 )        (fp16 + fp16 residual, encoder)
    GEMM_F32_RES32 = 3,  // out32 = res32 + f16(acc + bias)               (fp32 + fp16 residual, decoder)
    GEMM_F32 = 4,        // out32 = acc (+ bias)                          (logits before the fp16 store)
    // tcgen05 kernel only (DiT denoiser, dit.cu); N must be a multiple of 32:
    GEMM_F16_GEGLU = 5,  // W rows interleaved in groups of 16 (16 value rows, then their 16 gate rows): out16[row][n / 2 ..] =
                         // f16( f16(acc_a + bias) * f16(gelu_erf(f16(acc_g + bias))) ), ldo = row pitch of the N / 2 wide output
    GEMM_GATE_RES32 = 6, // y = f16(acc + bias); gate = f16(gate_tab[n] + gate_t[(row / n_per) * gate_bs + n]);
                         // out32 = res32 + f16(gate * y); out16 (optional) = f16(out32)
};

struct GemmArgs {
    const __half* A; int lda;       // [M][K]
    const __half* W; int ldw;       // [N][K]  (nn.Linear weight)
    const __half* bias;             // [N] or null
    int M, N, K;
    int mode;
    __half* out16; float* out32; int ldo;
    const __half* res16; const float* res32; int ldr;
    const __half *gate_tab, *gate_t; long long gate_bs; int n_per;     // GEMM_GATE_RES32
};

struct AttnArgs {   // softmax(q k^T / sqrt(D)) v ; q [B][Nq][H][D], k/v [B][Nk][H][D] with element strides
    const __half *q, *k, *v; __half* out;
    long long q_bs, k_bs, v_bs, o_bs;   // batch strides (elements)
    int ldq, ldk, ldv, ldo;             // row strides (elements); head h starts at h*D inside a row
    int B, H, Nq, Nk, D, causal;
    float* lse2;                        // optional output [B][H][Nq]: log2-domain log-sum-exp of the scaled scores (training: saves the backward's statistics pass)
};

struct AttnBwdArgs {
    const __half *q, *k, *v, *o, *dout;      // q/k/v with row pitch ld_qkv, o / dout with row pitch ld_o
    __half *dq, *dk, *dv;                    // row pitch ld_dqkv
    float *lse2, *dsum;                      // [B][H][Nq]: log2-domain log-sum-exp of the scaled scores, rowsum(dO * O)
    long long q_bs, k_bs, v_bs, o_bs, dq_bs, dk_bs, dv_bs;
    int ld_qkv_q, ld_qkv_k, ld_qkv_v, ld_o, ld_dq, ld_dk, ld_dv;
    int B, H, Nq, Nk, causal;
    float scale, scale_log2;                 // 1/sqrt(D), scale * log2(e)
    int have_lse;                            // lse2 was written by the forward kernel: skip the statistics pass, compute dsum only
};

}  // namespace er

// sets the thread-local message behind er_last_error() and returns `code` (engine.cu)
int er_set_error(int code, const char* fmt, ...);

cudaError_t er_gemm(const er::GemmArgs& g, cudaStream_t stream);
cudaError_t er_attention(const er::AttnArgs& a, cudaStream_t stream);

// ---- elementwise / row kernels (elementwise.cu) ----------------------------------------------------------------------
// LayerNorm rows (eps 1e-5, fp16 affine params, fp32 math): in32 or in16 -> out32 (optional) and out16 (optional)
cudaError_t er_layernorm(const float* in32, const __half* in16, int ld_in, const __half* gamma, const __half* beta,
                         float* out32, __half* out16, int ld_out, int M, int C, cudaStream_t stream);
// Fourier point embedding (point.py:54-63): xyz fp32 [n][3] -> A16 [n][ldo]: sin(24) | cos(24) | xyz | zero pad
cudaError_t er_point_embed(const float* xyz, const __half* basis /*[3][24]*/, __half* out, int ldo, int n, cudaStream_t stream);
// GEGLU (point.py:69-72): h [M][2F] -> out [M][F] = f16( a * f16(gelu_erf(g)) )
cudaError_t er_geglu(const __half* h, __half* out, int M, int F, cudaStream_t stream);
// Prefix embeddings (models.py:228-233 + modeling_opt.py:355-357): rows [0,P) from cond32, rows [P,P+n) = embd[ids];
// + pos[row0 + i]; writes x32 and its fp16 copy.
cudaError_t er_embed_prefix(const float* cond32, int P, const int32_t* ids_dev, int n_ids, const __half* embd, const __half* pos,
                            int C, float* x32, __half* x16, cudaStream_t stream);
// f16 -> f32 row copy (num-face embedding row appended to cond_embeds)
cudaError_t er_f16_to_f32(const __half* src, float* dst, int n, cudaStream_t stream);
cudaError_t er_f32_to_f16(const float* src, __half* dst, size_t n, cudaStream_t stream);
// Scatter k,v rows of qkv16 [N][3C] (positions pos0..pos0+N) into the decode kernel's KV cache layouts
cudaError_t er_kv_store(const __half* qkv16, int N, int C, int H, int layer, int pos0, int Lmax, int nkb, __half* kc, __half* vc,
                        cudaStream_t stream);
// cross-entropy over rows with label != -100 on fp16-rounded logits (modeling_opt.py:500-505): per-row losses -> row_loss / row_valid
// (scratch, M entries), then *loss_sum += sum, *count += valid rows, in a fixed order (no float atomics)
cudaError_t er_cross_entropy(const float* logits_pre, int ld, const int64_t* labels, int M, int V, float* row_loss, unsigned char* row_valid,
                             double* loss_sum, int* count, cudaStream_t stream);
// *out += sum of squares of x (partial296: scratch of 296 floats), fixed order
cudaError_t er_sum_squares(const __half* x, size_t n, float* partial296, double* out, cudaStream_t stream);
// rows of a16 [M][C] whose mask byte is 0 become zero (flash_attn pad_input)
cudaError_t er_zero_masked_rows(__half* a16, const unsigned char* mask, int M, int C, cudaStream_t stream);

// ---- training step: backward-pass kernels (backward.cu) ------------------------------------------------------------------------------
#define ER_BW_SLABS 64      // row slabs of the two-stage column reductions (bias / LayerNorm parameter gradients)
// out [Cn][ld_out] = in [R][ld_in]^T, columns [R, round_up(R, 64)) of out zeroed; ld_out >= round_up(R, 64)
cudaError_t er_transpose_f16(const __half* in, int R, int Cn, int ld_in, __half* out, int ld_out, cudaStream_t st);
// out32 = res32 + dropout_p(y16) over n elements; the keep mask is a counter-based function of (seed, site, element index); p == 0: plain add
cudaError_t er_add_dropout(const float* res32, const __half* y16, float* out32, size_t n, float p, unsigned long long seed, unsigned site, cudaStream_t st);
// LayerNorm backward (eps 1e-5): dy32 [M][C] (gradient of the output), s32 | s16 [M][ld_s] (the input) -> ds32 [M][C] (optional), dbr16 [M][C]
// (optional: f16 of ds through the dropout mask of (p, seed, site)), mean / rstd [M]
cudaError_t er_ln_bwd(const float* dy32, const float* s32, const __half* s16, int ld_s, const __half* gamma, float* ds32, __half* dbr16, float* mean,
                      float* rstd, int M, int C, float p, unsigned long long seed, unsigned site, cudaStream_t st);
// dgamma [C] = sum_r dy * xhat, dbeta [C] = sum_r dy; partial: scratch of ER_BW_SLABS * 2 * C floats
cudaError_t er_ln_param_grad(const float* dy32, const float* s32, const __half* s16, int ld_s, const float* mean, const float* rstd, int M, int C,
                             float* partial, float* dgamma, float* dbeta, cudaStream_t st, int accumulate = 0);
// out [ncols] = column sums of x16 [M][ld]; partial: scratch of ER_BW_SLABS * ncols floats
cudaError_t er_colsum_f16(const __half* x16, int ld, int M, int ncols, float* partial, float* out, cudaStream_t st, int accumulate = 0);
// GEGLU backward: h [M][2F] (the forward input), dout [M][F] -> dh [M][2F]
cudaError_t er_geglu_bwd(const __half* h, const __half* dout, __half* dh, int M, int F, cudaStream_t st);
cudaError_t er_add_f32(float* dst, const float* src, size_t n, cudaStream_t st);                       // dst += src
cudaError_t er_axpy_f16(__half* dst, int ld_dst, const __half* src, int ld_src, int rows, int cols, float alpha, cudaStream_t st);   // dst = f16(dst + alpha src)
// dh16 = h16 > 0 ? dh16 : 0 (in place), n multiple of 8
cudaError_t er_relu_bwd(__half* dh16, const __half* h16, size_t n, cudaStream_t st);
// dl [B*N][ldo] = f16(loss_scale / *count * (softmax(round_f16(logits)) - onehot(label of the NEXT row))), zero for ignored rows and pad columns
cudaError_t er_ce_bwd(const float* logits_pre, int ld, const int64_t* labels, int B, int N, int V, const int* count_dev, float loss_scale, __half* dl, int ldo,
                      cudaStream_t st);
// gradients of embed_positions rows [0, N), embd [V][C] and (optional) embed_num_face [10][C] from dx0 [B*N][C]
cudaError_t er_embed_bwd(const float* dx0, const int32_t* ids, const int32_t* bucket_dev, int B, int T, int N, int P, int C, int V, int numface_row,
                         float* dpos, float* dembd, float* denf, cudaStream_t st);
cudaError_t er_export_f32(const float* src, int ld, int rows, int cols, float scale, float* dst, cudaStream_t st);
// flash-attention backward for the forward described by `a` (a.out = the forward output): dq / dk / dv with their own pitches and batch strides;
// lse2, dsum: scratch of B * H * Nq floats each
cudaError_t er_attention_bwd(const er::AttnArgs& a, const __half* dout, __half* dq, __half* dk, __half* dv, int ld_dq, int ld_dk, int ld_dv, long long dq_bs,
                             long long dk_bs, long long dv_bs, float* lse2, float* dsum, cudaStream_t st);
cudaError_t er_attn_bwd_wmma(const er::AttnBwdArgs& p, int D, cudaStream_t st);     // backward.cu: wmma through shared memory
cudaError_t er_attn_bwd_mma(const er::AttnBwdArgs& p, int D, cudaStream_t st);      // attention_bwd_mma.cu: mma.sync, scores / dS in registers
// dsum [B][H][Nq] = sum_d dO * O per (row, head): the only statistic the attention backward still needs when the forward wrote lse2
cudaError_t er_attn_rowdot(const er::AttnBwdArgs& p, int D, cudaStream_t st);
