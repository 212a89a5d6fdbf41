// Row / elementwise kernels of the prefill, point-encoder and teacher-forced paths.  All are HBM-bound
// streaming kernels: 128-bit accesses where the layout allows, one warp (or one CTA) per row.
#include "kernels.h"

#include "common.cuh"

namespace er {

// ---- LayerNorm (nn.LayerNorm, eps 1e-5; autocast runs it in fp32: SURVEY.md Appendix B) -----------------------------
// one warp per row; reference call sites: modeling_opt.py:274,288 ; point.py:193,122-123 ; models.py:124
__global__ void layernorm_kernel(const float* in32, const __half* in16, int ld_in, const __half* __restrict__ gamma,
                                 const __half* __restrict__ beta, float* out32, __half* out16, int ld_out, int M, int C) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    auto ld = [&](int i) -> float {
        return in32 ? in32[(size_t)row * ld_in + i] : __half2float(in16[(size_t)row * ld_in + i]);
    };
    float s = 0.f;
    for (int i = lane; i < C; i += 32) s += ld(i);
    const float mean = warp_sum(s) / C;
    float q = 0.f;
    for (int i = lane; i < C; i += 32) { const float d = ld(i) - mean; q += d * d; }
    const float rstd = rsqrtf(warp_sum(q) / C + 1e-5f);
    for (int i = lane; i < C; i += 32) {
        const float y = (ld(i) - mean) * rstd * __half2float(gamma[i]) + __half2float(beta[i]);
        if (out32) out32[(size_t)row * ld_out + i] = y;
        if (out16) out16[(size_t)row * ld_out + i] = __float2half_rn(y);
    }
}

// ---- Fourier point embedding (point.py:54-63) ------------------------------------------------------------------------
// autocast: einsum runs in fp16 (xyz and basis rounded to fp16, product rounded to fp16), sin/cos evaluated on the
// fp16 value and rounded to fp16; the trailing xyz is cast to fp16 by the following Linear.
__global__ void point_embed_kernel(const float* xyz, const __half* __restrict__ basis, __half* out, int ldo, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __half* o = out + (size_t)i * ldo;
    for (int d = 0; d < 3; d++) {
        const float x = xyz[(size_t)i * 3 + d];
        const float xh = round_f16(x);
        for (int j = 0; j < 8; j++) {
            const float pr = round_f16(xh * __half2float(basis[d * 24 + d * 8 + j]));
            o[d * 8 + j] = __float2half_rn(sinf(pr));
            o[24 + d * 8 + j] = __float2half_rn(cosf(pr));
        }
        o[48 + d] = __float2half_rn(x);
    }
    for (int c = 51; c < ldo; c++) o[c] = __float2half_rn(0.f);
}

// ---- GEGLU (point.py:69-72): fp16 in/out, erf-GELU evaluated in fp32 and rounded, then the product rounded ------------
__global__ void geglu_kernel(const __half* h, __half* out, int M, int F) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)M * F) return;
    const size_t r = idx / F, c = idx % F;
    const float a = __half2float(h[r * 2 * F + c]);
    const float g = __half2float(h[r * 2 * F + F + c]);
    const float ge = round_f16(0.5f * g * (1.f + erff(g * 0.70710678118654752f)));
    out[idx] = __float2half_rn(a * ge);
}

// ---- prefix embeddings: cat(cond_embeds fp32, embd[ids] fp16) + pos (models.py:228-233, modeling_opt.py:355-357) -----
__global__ void embed_prefix_kernel(const float* cond32, int P, const int32_t* ids, int n_ids, const __half* __restrict__ embd,
                                    const __half* __restrict__ pos, int C, float* x32, __half* x16) {
    const int row = blockIdx.x;
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        float v = row < P ? cond32[(size_t)row * C + i] : __half2float(embd[(size_t)ids[row - P] * C + i]);
        v += __half2float(pos[(size_t)row * C + i]);
        x32[(size_t)row * C + i] = v;
        x16[(size_t)row * C + i] = __float2half_rn(v);
    }
}

__global__ void f16_to_f32_kernel(const __half* s, float* d, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = __half2float(s[i]);
}
__global__ void f32_to_f16_kernel(const float* s, __half* d, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = __float2half_rn(s[i]);
}

// ---- KV cache store: qkv16 [N][3C] -> K and V, both blocked [layer][h][key/32][d/8][key%32][8] ------------------------------
__global__ void kv_store_kernel(const __half* qkv16, int N, int C, int H, int layer, int pos0, int Lmax, int nkb, __half* kc, __half* vc) {
    const int D = C / H, DV = D / 8;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one 16-byte vector each
    const size_t total = (size_t)N * H * DV;
    if (idx >= total) return;
    const int vec = idx % DV;
    const int h = (idx / DV) % H;
    const int n = idx / ((size_t)DV * H);
    const int key = pos0 + n;
    const uint4 kv = *reinterpret_cast<const uint4*>(qkv16 + (size_t)n * 3 * C + C + h * D + vec * 8);
    const uint4 vv = *reinterpret_cast<const uint4*>(qkv16 + (size_t)n * 3 * C + 2 * C + h * D + vec * 8);
    const size_t kidx = ((((size_t)layer * H + h) * nkb + (key >> 5)) * DV + vec) * 256 + (size_t)(key & 31) * 8;
    *reinterpret_cast<uint4*>(kc + kidx) = kv;
    *reinterpret_cast<uint4*>(vc + kidx) = vv;
}

// ---- cross entropy on fp16-rounded logits, ignore_index = -100 (modeling_opt.py:500-505) ------------------------------
// row r of logits predicts labels[r] (the caller passes already-shifted views); one warp per row.  Per-row losses go to row_loss
// (0 and valid = 0 for ignored rows); the sums are formed by reduce_rows_kernel in a FIXED order (no float atomics: the loss is
// bit-reproducible run to run and independent of how rows are spread over ranks up to the final sum).
__global__ void cross_entropy_kernel(const float* logits_pre, int ld, const int64_t* labels, int M, int V, float* row_loss, unsigned char* row_valid) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= M) return;
    const long long lab = labels[row];
    if (lab < 0) { if (lane == 0) { row_loss[row] = 0.f; row_valid[row] = 0; } return; }
    const float* x = logits_pre + (size_t)row * ld;
    float mx = -INFINITY;
    for (int i = lane; i < V; i += 32) mx = fmaxf(mx, round_f16(x[i]));
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < V; i += 32) s += expf(round_f16(x[i]) - mx);
    s = warp_sum(s);
    if (lane == 0) { row_loss[row] = logf(s) + mx - round_f16(x[lab]); row_valid[row] = 1; }
}
// one block: sum[0] += sum of vals (double accumulation, fixed strided order + fixed tree), count[0] += number of valid rows
__global__ void reduce_rows_kernel(const float* vals, const unsigned char* valid, int M, double* sum, int* count) {
    __shared__ double sh[1024];
    __shared__ int shc[1024];
    double s = 0.0; int c = 0;
    for (int i = threadIdx.x; i < M; i += blockDim.x) { s += (double)vals[i]; c += valid ? valid[i] : 0; }
    sh[threadIdx.x] = s; shc[threadIdx.x] = c;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) { sh[threadIdx.x] += sh[threadIdx.x + o]; shc[threadIdx.x] += shc[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { *sum += sh[0]; if (count) *count += shc[0]; }
}

// per-block partial sums of squares (fixed order inside a block), reduced by reduce_rows_kernel
__global__ void sum_squares_kernel(const __half* x, size_t n, float* partial) {
    __shared__ float sh[256];
    float s = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = __half2float(x[i]);
        s += v * v;
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// flash_attn's pad_input (attention.py:88-93): rows of the attention output whose mask is false are zero
__global__ void zero_masked_rows_kernel(__half* a16, const unsigned char* mask, int M, int C) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one 16-byte vector each
    const size_t total = (size_t)M * (C >> 3);
    if (i >= total) return;
    const int row = (int)(i / (C >> 3));
    if (!mask[row]) reinterpret_cast<uint4*>(a16)[i] = make_uint4(0, 0, 0, 0);
}

}  // namespace er

using namespace er;

cudaError_t er_layernorm(const float* in32, const __half* in16, int ld_in, const __half* gamma, const __half* beta, float* out32,
                         __half* out16, int ld_out, int M, int C, cudaStream_t stream) {
    if (M <= 0) return cudaSuccess;
    layernorm_kernel<<<(M + 7) / 8, 256, 0, stream>>>(in32, in16, ld_in, gamma, beta, out32, out16, ld_out, M, C);
    return cudaGetLastError();
}
cudaError_t er_point_embed(const float* xyz, const __half* basis, __half* out, int ldo, int n, cudaStream_t stream) {
    point_embed_kernel<<<(n + 127) / 128, 128, 0, stream>>>(xyz, basis, out, ldo, n);
    return cudaGetLastError();
}
cudaError_t er_geglu(const __half* h, __half* out, int M, int F, cudaStream_t stream) {
    const size_t n = (size_t)M * F;
    geglu_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(h, out, M, F);
    return cudaGetLastError();
}
cudaError_t er_embed_prefix(const float* cond32, int P, const int32_t* ids_dev, int n_ids, const __half* embd, const __half* pos, int C,
                            float* x32, __half* x16, cudaStream_t stream) {
    embed_prefix_kernel<<<P + n_ids, 256, 0, stream>>>(cond32, P, ids_dev, n_ids, embd, pos, C, x32, x16);
    return cudaGetLastError();
}
cudaError_t er_f16_to_f32(const __half* src, float* dst, int n, cudaStream_t stream) {
    f16_to_f32_kernel<<<(n + 255) / 256, 256, 0, stream>>>(src, dst, n);
    return cudaGetLastError();
}
cudaError_t er_f32_to_f16(const float* src, __half* dst, size_t n, cudaStream_t stream) {
    f32_to_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, dst, n);
    return cudaGetLastError();
}
cudaError_t er_kv_store(const __half* qkv16, int N, int C, int H, int layer, int pos0, int Lmax, int nkb, __half* kc, __half* vc,
                        cudaStream_t stream) {
    const size_t total = (size_t)N * H * (C / H / 8);
    kv_store_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(qkv16, N, C, H, layer, pos0, Lmax, nkb, kc, vc);
    return cudaGetLastError();
}
cudaError_t er_cross_entropy(const float* logits_pre, int ld, const int64_t* labels, int M, int V, float* row_loss, unsigned char* row_valid,
                             double* loss_sum, int* count, cudaStream_t stream) {
    cross_entropy_kernel<<<(M + 7) / 8, 256, 0, stream>>>(logits_pre, ld, labels, M, V, row_loss, row_valid);
    reduce_rows_kernel<<<1, 1024, 0, stream>>>(row_loss, row_valid, M, loss_sum, count);
    return cudaGetLastError();
}
cudaError_t er_sum_squares(const __half* x, size_t n, float* partial296, double* out, cudaStream_t stream) {
    sum_squares_kernel<<<296, 256, 0, stream>>>(x, n, partial296);
    reduce_rows_kernel<<<1, 1024, 0, stream>>>(partial296, nullptr, 296, out, nullptr);
    return cudaGetLastError();
}
cudaError_t er_zero_masked_rows(__half* a16, const unsigned char* mask, int M, int C, cudaStream_t stream) {
    const size_t total = (size_t)M * (C >> 3);
    zero_masked_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a16, mask, M, C);
    return cudaGetLastError();
}
