// DiT denoiser engine: the image-conditioned path's latent generator (BASELINE configs[4]).
//
// Stands in for, on the inference path (infer_dit.py:104-113):
//   DiT.forward / DiTLayer._forward      core/transformer/dit.py:100-196     (PixArt-alpha style adaLN-single, 24 x [self-attn, cross-attn, GEGLU FF])
//   Timesteps / TimestepEmbedding        core/transformer/dit.py:45-97
//   MDiT.get_cond (after the CLIP tower) core/models_dit.py:106-118          (proj_cond + norm_cond)
//   MDiT.run                             core/models_dit.py:184-229          (classifier-free guidance + DDIM update, 100 steps)
// The CLIP vision tower itself is a third-party library model in the reference (transformers.CLIPVisionModel) and stays one here.
//
// B200-first choices (not in the reference): the whole sampling loop runs on the device — one CUDA graph per step, replayed; everything that
// does not depend on the latents is hoisted out of the loop: the timestep MLP + adaLN vectors of all steps (one batched GEMM), and the
// cross-attention K / V of the condition for every layer (the reference recomputes both every step); guidance + scheduler update are one
// fused kernel; GEMMs / attention are the tcgen05 kernels of this directory.
//
// Rounding points follow the reference under `torch.autocast(fp16)` with `.half()` weights (infer_dit.py:70,106): Linear outputs fp16,
// LayerNorm outputs fp32, the residual stream fp32 from the first norm on (the reference re-binds x to the NORMALISED, modulated tensor
// before each residual add — dit.py:127-136 — and so does this), `1 + scale`, `table + t_adaln` and `gate * y` rounded to fp16.
#include "../../include/edgerunner_b200.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "kernels.h"

#define CK(call)                                                                                                          \
    do {                                                                                                                  \
        cudaError_t _e = (call);                                                                                          \
        if (_e != cudaSuccess) return er_set_error(ER_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)
#define CKL(e, call) do { (e)->launches++; CK(call); } while (0)

namespace {

__device__ __forceinline__ float h2f(__half h) { return __half2float(h); }
__device__ __forceinline__ float rh(float x) { return __half2float(__float2half_rn(x)); }   // round through fp16

// fp16 table row + fp16 per-sample vector, rounded to fp16 (`scale_shift_table[None] + t_adaln`, dit.py:125)
__device__ __forceinline__ float mod_val(const __half* tab, const __half* tv, int c) { return rh(h2f(tab[c]) + h2f(tv[c])); }

// LayerNorm(eps 1e-6, no affine) in fp32, then x * (1 + scale) + shift (dit.py:127-128,134-135,191-192).
// in: in32 [M][C], or in16 [M][C] (+ add16 [(row % n_per)][C]: the learned positions, summed in fp16 first — dit.py:176).
// shift = f16(tab_shift + t_shift[b]), scale = f16(tab_scale + t_scale[b]), b = row / n_per.  One block (128 threads) per row.
// Rows below `split` read alt32 + cadd instead of in32 (the unconditional half of a guided step, see denoiser()).
__global__ void __launch_bounds__(128) dit_ln_mod_kernel(const float* __restrict__ in32, const __half* __restrict__ in16, const __half* __restrict__ add16,
                                                         const __half* tab_shift, const __half* tab_scale, const __half* t_shift, const __half* t_scale,
                                                         long long t_bs, int n_per, float* __restrict__ out32, __half* __restrict__ out16, int C,
                                                         const float* __restrict__ alt32, const __half* __restrict__ cadd, int split) {
    const int row = blockIdx.x, tid = threadIdx.x;
    const int b = row / n_per;
    __shared__ float red[4];
    __shared__ float srow[2048];
    const int Cs = C <= 2048 ? C : 0;      // rows up to 2048 wide are staged in shared memory; wider ones are re-read
    auto load = [&](int c) -> float {
        if (in32) return row < split ? __fadd_rn(alt32[(size_t)row * C + c], h2f(cadd[c])) : in32[(size_t)row * C + c];
        float v = h2f(in16[(size_t)row * C + c]);
        if (add16) v = rh(v + h2f(add16[(size_t)(row % n_per) * C + c]));
        return v;
    };
    float s = 0.f;
    for (int c = tid; c < C; c += 128) { const float v = load(c); if (Cs) srow[c] = v; s += v; }
    auto block_sum = [&](float v) -> float {
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        return red[0] + red[1] + red[2] + red[3];
    };
    const float mean = block_sum(s) / C;
    float q = 0.f;
    for (int c = tid; c < C; c += 128) { const float d = (Cs ? srow[c] : load(c)) - mean; q += d * d; }
    const float rstd = rsqrtf(block_sum(q) / C + 1e-6f);
    const __half* ts = t_shift + (size_t)b * t_bs; const __half* tc = t_scale + (size_t)b * t_bs;
    for (int c = tid; c < C; c += 128) {
        const float xn = ((Cs ? srow[c] : load(c)) - mean) * rstd;
        const float sc = rh(1.f + mod_val(tab_scale, tc, c));
        const float y = __fadd_rn(__fmul_rn(xn, sc), mod_val(tab_shift, ts, c));
        if (out32) out32[(size_t)row * C + c] = y;
        out16[(size_t)row * C + c] = __float2half_rn(y);
    }
}

// The same with ONE WARP per row and the row in registers (C = 128 * NCH): one global read, 16-byte accesses, no block barrier.  At the preset
// width this kernel moves 160 MB per call (fp32 in, fp32 + fp16 out) and is bound by HBM.
template <int NCH>
__global__ void __launch_bounds__(128) dit_ln_mod_warp_kernel(const float* __restrict__ in32, const __half* __restrict__ in16, const __half* __restrict__ add16,
                                                              const __half* tab_shift, const __half* tab_scale, const __half* t_shift, const __half* t_scale,
                                                              long long t_bs, int n_per, float* __restrict__ out32, __half* __restrict__ out16, int M,
                                                              const float* __restrict__ alt32, const __half* __restrict__ cadd, int split) {
    constexpr int C = 128 * NCH;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const int b = row / n_per;
    float x[NCH][4];
#pragma unroll
    for (int k = 0; k < NCH; k++) {
        const int c = k * 128 + lane * 4;
        if (in32) {
            const float4 v = *reinterpret_cast<const float4*>((row < split ? alt32 : in32) + (size_t)row * C + c);
            x[k][0] = v.x; x[k][1] = v.y; x[k][2] = v.z; x[k][3] = v.w;
            if (row < split) {
                const uint2 cw = *reinterpret_cast<const uint2*>(cadd + c);
                const __half* ch = reinterpret_cast<const __half*>(&cw);
#pragma unroll
                for (int i = 0; i < 4; i++) x[k][i] = __fadd_rn(x[k][i], h2f(ch[i]));
            }
        } else {
            const uint2 w = *reinterpret_cast<const uint2*>(in16 + (size_t)row * C + c);
            const __half* hw = reinterpret_cast<const __half*>(&w);
#pragma unroll
            for (int i = 0; i < 4; i++) x[k][i] = h2f(hw[i]);
            if (add16) {
                const uint2 aw = *reinterpret_cast<const uint2*>(add16 + (size_t)(row % n_per) * C + c);
                const __half* ah = reinterpret_cast<const __half*>(&aw);
#pragma unroll
                for (int i = 0; i < 4; i++) x[k][i] = rh(x[k][i] + h2f(ah[i]));
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; k++) s += (x[k][0] + x[k][1]) + (x[k][2] + x[k][3]);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; k++)
#pragma unroll
        for (int i = 0; i < 4; i++) { const float d = x[k][i] - mean; q += d * d; }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + 1e-6f);
    const __half* ts = t_shift + (size_t)b * t_bs; const __half* tc = t_scale + (size_t)b * t_bs;
#pragma unroll
    for (int k = 0; k < NCH; k++) {
        const int c = k * 128 + lane * 4;
        float y[4];
        __align__(8) __half yh[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float sc = rh(1.f + mod_val(tab_scale, tc, c + i));
            y[i] = __fadd_rn(__fmul_rn((x[k][i] - mean) * rstd, sc), mod_val(tab_shift, ts, c + i));
            yh[i] = __float2half_rn(y[i]);
        }
        if (out32) *reinterpret_cast<float4*>(out32 + (size_t)row * C + c) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<uint2*>(out16 + (size_t)row * C + c) = *reinterpret_cast<const uint2*>(yh);
    }
}

// x = x + gate * y (dit.py:129,136): gate = f16(tab + t[b]); g = f16(gate * y16); out32 = x32 + g; out16 = f16(out32)
__global__ void dit_gate_res_kernel(const float* __restrict__ x32, const __half* __restrict__ y16, const __half* tab, const __half* tv, long long t_bs,
                                    int n_per, float* __restrict__ out32, __half* __restrict__ out16, size_t M, int C) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= M * C) return;
    const size_t row = i / C; const int c = (int)(i % C);
    const __half* t = tv + (row / n_per) * t_bs;
    const float4 x = *reinterpret_cast<const float4*>(x32 + i);
    const uint2 yy = *reinterpret_cast<const uint2*>(y16 + i);
    const __half* y = reinterpret_cast<const __half*>(&yy);
    float o[4] = {x.x, x.y, x.z, x.w};
    __half oh[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        o[k] = o[k] + rh(mod_val(tab, t, c + k) * h2f(y[k]));
        oh[k] = __float2half_rn(o[k]);
    }
    *reinterpret_cast<float4*>(out32 + i) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint2*>(out16 + i) = *reinterpret_cast<uint2*>(oh);
}

// Timesteps(num_channels = 256) (dit.py:45-76): [sin(t * w_i) | cos(t * w_i)], w_i = exp(-ln(10000) * i / 128), fp32, stored as the fp16 the
// first Linear sees under autocast.  One block per timestep.
__global__ void dit_timestep_kernel(const float* __restrict__ t, __half* __restrict__ out) {
    const int s = blockIdx.x, i = threadIdx.x;       // 128 threads
    const float w = expf((-9.210340371976184f * (float)i) / 128.f);
    const float a = t[s] * w;
    out[(size_t)s * 256 + i] = __float2half_rn(sinf(a));
    out[(size_t)s * 256 + 128 + i] = __float2half_rn(cosf(a));
}
// F.silu on fp16 (dit.py:92,180)
__global__ void dit_silu_kernel(const __half* __restrict__ in, __half* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = h2f(in[i]);
    out[i] = __float2half_rn(x / (1.f + expf(-x)));
}
// GEGLU (dit.py:26-29): h [M][2F] -> out [M][F] = f16(a * f16(gelu_erf(g))), 8 outputs per thread
__global__ void dit_geglu_kernel(const __half* __restrict__ h, __half* __restrict__ out, size_t M, int F) {
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= M * F) return;
    const size_t row = i / F; const int c = (int)(i % F);
    const uint4 av = *reinterpret_cast<const uint4*>(h + row * 2 * F + c);
    const uint4 gv = *reinterpret_cast<const uint4*>(h + row * 2 * F + F + c);
    const __half* a = reinterpret_cast<const __half*>(&av); const __half* g = reinterpret_cast<const __half*>(&gv);
    __half o[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const float x = h2f(g[k]);
        o[k] = __float2half_rn(h2f(a[k]) * rh(0.5f * x * (1.f + erff(x * 0.70710678118654752440f))));
    }
    *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<uint4*>(o);
}
// torch.cat([latents] * 2) in the fp16 the first Linear sees (models_dit.py:212 + autocast)
__global__ void dit_dup_f16_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const __half v = __float2half_rn(in[i]);
    out[i] = v; out[n + i] = v;
}
// cond = cat([zeros_like(cond), cond]) as fp16 (models_dit.py:209 + the autocast cast in k_proj / v_proj)
__global__ void dit_cfg_cond_kernel(const float* __restrict__ in, __half* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = __float2half_rn(0.f); out[n + i] = __float2half_rn(in[i]);
}

struct StepCoef { float sa, sb, sap, dir; };   // sqrt(alpha_t), sqrt(1 - alpha_t), sqrt(alpha_prev), sqrt(1 - alpha_prev - sigma^2)

// First node of a step's graph: picks step *counter's adaLN vectors / t_emb / scheduler coefficients into fixed buffers, so that every
// step replays the SAME graph.
__global__ void dit_select_step_kernel(const int* __restrict__ counter, const __half* __restrict__ ada_all, const __half* __restrict__ temb_all,
                                       const StepCoef* __restrict__ coef_all, __half* __restrict__ ada_cur, __half* __restrict__ temb_cur,
                                       StepCoef* __restrict__ coef_cur, int C) {
    const int s = *counter;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < 6 * C; c += gridDim.x * blockDim.x) {
        ada_cur[c] = ada_all[(size_t)s * 6 * C + c];
        if (c < C) temb_cur[c] = temb_all[(size_t)s * C + c];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *coef_cur = coef_all[s];
}
// Last node: guidance (models_dit.py:223-224, fp16 arithmetic) + DDIMScheduler.step (diffusers 0.30.2 scheduling_ddim.py, eta = 0, no clipping):
//   v_prediction: x0 = sa * x - f16(sb * m);   eps = f16(sa * m) + sb * x
//   epsilon:      x0 = (x - f16(sb * m)) / sa; eps = m
//   x_prev = sap * x0 + dir * eps            (0-dim fp32 coefficients times an fp16 tensor give fp16 in torch: those products are rounded)
// `guided` = 0: pred holds one prediction per sample (no guidance).  Thread 0 advances the step counter.
__global__ void dit_guide_step_kernel(const __half* __restrict__ pred, float* __restrict__ lat, size_t n, float gscale, int guided, int v_pred,
                                      const StepCoef* __restrict__ coef, int* counter) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *counter += 1;          // safe: the next reader is a later kernel
    if (i >= n) return;
    const StepCoef k = *coef;
    float m;
    if (guided) {
        const float u = h2f(pred[i]), c = h2f(pred[n + i]);
        m = rh(u + rh(gscale * rh(c - u)));
    } else {
        m = h2f(pred[i]);
    }
    const float x = lat[i];
    float x0, prev;
    if (v_pred) {
        x0 = __fsub_rn(__fmul_rn(k.sa, x), rh(k.sb * m));
        const float eps = __fadd_rn(rh(k.sa * m), __fmul_rn(k.sb, x));
        prev = __fadd_rn(__fmul_rn(k.sap, x0), __fmul_rn(k.dir, eps));
    } else {
        x0 = __fdiv_rn(__fsub_rn(x, rh(k.sb * m)), k.sa);
        prev = __fadd_rn(__fmul_rn(k.sap, x0), rh(k.dir * m));
    }
    lat[i] = prev;
}
__global__ void dit_convert_kernel(const void* src, int dtype, __half* dst, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[i] = dtype == ER_DTYPE_F16 ? ((const __half*)src)[i] : __float2half_rn(((const float*)src)[i]);
}
__global__ void dit_set_int_kernel(int* p, int v) { *p = v; }
// ff.net.0 re-ordered for the fused GEGLU epilogue: physical row p of [2F][K] = value row 16 (p / 32) + p % 32 when p % 32 < 16, else gate row
// F + 16 (p / 32) + p % 32 - 16; bias likewise (K = 1 call)
__global__ void dit_interleave_kernel(const __half* __restrict__ src, __half* __restrict__ dst, int F, int K) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)2 * F * K) return;
    const int p = (int)(i / K), k = (int)(i % K);
    const int grp = p >> 5, w = p & 31;
    const int srow = w < 16 ? grp * 16 + w : F + grp * 16 + (w - 16);
    dst[i] = src[(size_t)srow * K + k];
}

struct Slot { __half* dst; size_t n; };

}  // namespace

struct er_dit {
    er_dit_config cfg;
    int C, H, D, NL, N, DL, M, CD;
    long long launches = 0;
    std::vector<void*> allocs;
    std::map<std::string, Slot> slots;
    std::set<std::string> loaded;
    bool finalized = false;
    struct Layer { __half *qkv_w, *qkv_b, *o_w, *o_b, *q_w, *q_b, *kv_w, *kv_b, *o2_w, *o2_b, *ff1_w, *ff1_b, *ff2_w, *ff2_b, *table, *ff1_wi, *ff1_bi; };
    std::vector<Layer> L;
    __half *pin_w, *pin_b, *pos, *t1_w, *t1_b, *t2_w, *t2_b, *ada_w, *ada_b, *table2, *pout_w, *pout_b, *pc_w, *pc_b, *nc_g, *nc_b;
    // workspace for `rows` = batch * N token rows and `crows` = batch * M condition rows
    int max_batch = 0;
    float *xa32 = nullptr, *xb32 = nullptr;
    __half *x16 = nullptr, *qkv16 = nullptr, *a16 = nullptr, *y16 = nullptr, *h16 = nullptr, *g16 = nullptr, *in16 = nullptr, *pred16 = nullptr;
    __half *c16 = nullptr, *kv16 = nullptr;        // condition fp16 [batch][M][C]; per layer K|V [NL][batch * M][2C]
    __half *pc16 = nullptr;                        // proj_cond output
    __half *cconst16 = nullptr;                    // [NL][C] cross-attention branch of a zero-condition sample
    int uncond_shortcut = 1;
    // timestep path: capacity max_steps rows
    int max_steps = 0;
    float* t_dev = nullptr; __half *te256 = nullptr, *te1 = nullptr, *te1s = nullptr, *temb = nullptr, *tembs = nullptr, *ada = nullptr;
    StepCoef* coef_dev = nullptr;
    __half *ada_cur = nullptr, *temb_cur = nullptr; StepCoef* coef_cur = nullptr; int* counter = nullptr;
    float* lat_dev = nullptr; float* cond_dev = nullptr;     // for er_dit_run_host
    // cached step graph
    cudaGraphExec_t graph = nullptr; int graph_batch = 0, graph_guided = 0, graph_vpred = 0; float graph_gscale = 0.f; float* graph_lat = nullptr;
    int use_graph = 1, fuse = 1, graph_fuse = -1; long long graph_kernels = 0;
};

static cudaError_t ln_mod(const float* in32, const __half* in16, const __half* add16, const __half* tab_shift, const __half* tab_scale, const __half* t_shift,
                          const __half* t_scale, long long t_bs, int n_per, float* out32, __half* out16, int M, int C, const float* alt32, const __half* cadd,
                          int split, cudaStream_t st) {
    if (C == 1024)
        dit_ln_mod_warp_kernel<8><<<(M + 3) / 4, 128, 0, st>>>(in32, in16, add16, tab_shift, tab_scale, t_shift, t_scale, t_bs, n_per, out32, out16, M, alt32, cadd, split);
    else if (C == 128)
        dit_ln_mod_warp_kernel<1><<<(M + 3) / 4, 128, 0, st>>>(in32, in16, add16, tab_shift, tab_scale, t_shift, t_scale, t_bs, n_per, out32, out16, M, alt32, cadd, split);
    else
        dit_ln_mod_kernel<<<M, 128, 0, st>>>(in32, in16, add16, tab_shift, tab_scale, t_shift, t_scale, t_bs, n_per, out32, out16, C, alt32, cadd, split);
    return cudaGetLastError();
}

template <typename T>
static int dalloc(er_dit* e, T** p, size_t n) {
    void* q = nullptr;
    CK(cudaMalloc(&q, n * sizeof(T) + 256));
    e->allocs.push_back(q);
    *p = (T*)q;
    return ER_OK;
}
template <typename T>
static void dfree(er_dit* e, T** p) {
    if (!*p) return;
    for (auto it = e->allocs.begin(); it != e->allocs.end(); ++it)
        if (*it == (void*)*p) { e->allocs.erase(it); break; }
    cudaFree(*p);
    *p = nullptr;
}
#define ALLOC(ptr, n) do { int _r = dalloc(e, &(ptr), (size_t)(n)); if (_r) return _r; } while (0)

static void add_slot(er_dit* e, const std::string& name, __half** dst, size_t n, int* rc) {
    if (*rc) return;
    *rc = dalloc(e, dst, n);
    if (!*rc) e->slots[name] = Slot{*dst, n};
}

extern "C" int er_dit_create(const er_dit_config* cfg, er_dit** out) {
    if (!cfg || !out) return er_set_error(ER_ERR_INVALID, "null argument");
    if (cfg->hidden_dim <= 0 || cfg->num_heads <= 0 || cfg->hidden_dim % cfg->num_heads || cfg->num_layers <= 0 || cfg->latent_size <= 0 ||
        cfg->latent_dim <= 0 || cfg->cond_tokens <= 0 || cfg->cond_dim <= 0)
        return er_set_error(ER_ERR_INVALID, "bad DiT dimensions");
    const int D = cfg->hidden_dim / cfg->num_heads;
    if (D != 64 && D != 96) return er_set_error(ER_ERR_INVALID, "DiT head_dim %d not supported (64 or 96)", D);
    if ((cfg->hidden_dim & 7) || (cfg->latent_dim & 7) || (cfg->cond_dim & 7)) return er_set_error(ER_ERR_INVALID, "dimensions must be multiples of 8");
    CK(cudaSetDevice(cfg->device));
    er_dit* e = new er_dit();
    e->cfg = *cfg;
    e->C = cfg->hidden_dim; e->H = cfg->num_heads; e->D = D; e->NL = cfg->num_layers; e->N = cfg->latent_size; e->DL = cfg->latent_dim;
    e->M = cfg->cond_tokens; e->CD = cfg->cond_dim;
    const size_t C = e->C;
    int rc = ER_OK;
    e->L.resize(e->NL);
    // state-dict keys of MDiT (core/models_dit.py:33-76): `dit.*`, `proj_cond.*`, `norm_cond.*`; k_proj / v_proj share one [2C][C] buffer
    add_slot(e, "dit.proj_in.weight", &e->pin_w, C * e->DL, &rc); add_slot(e, "dit.proj_in.bias", &e->pin_b, C, &rc);
    add_slot(e, "dit.pos_embed", &e->pos, (size_t)e->N * C, &rc);
    add_slot(e, "dit.timestep_proj.linear_1.weight", &e->t1_w, C * 256, &rc); add_slot(e, "dit.timestep_proj.linear_1.bias", &e->t1_b, C, &rc);
    add_slot(e, "dit.timestep_proj.linear_2.weight", &e->t2_w, C * C, &rc); add_slot(e, "dit.timestep_proj.linear_2.bias", &e->t2_b, C, &rc);
    add_slot(e, "dit.adaln_linear.weight", &e->ada_w, 6 * C * C, &rc); add_slot(e, "dit.adaln_linear.bias", &e->ada_b, 6 * C, &rc);
    add_slot(e, "dit.scale_shift_table", &e->table2, 2 * C, &rc);
    add_slot(e, "dit.proj_out.weight", &e->pout_w, (size_t)e->DL * C, &rc); add_slot(e, "dit.proj_out.bias", &e->pout_b, e->DL, &rc);
    add_slot(e, "proj_cond.weight", &e->pc_w, C * e->CD, &rc); add_slot(e, "proj_cond.bias", &e->pc_b, C, &rc);
    add_slot(e, "norm_cond.weight", &e->nc_g, C, &rc); add_slot(e, "norm_cond.bias", &e->nc_b, C, &rc);
    for (int l = 0; l < e->NL && !rc; l++) {
        er_dit::Layer& y = e->L[l];
        const std::string p = "dit.layers." + std::to_string(l) + ".";
        add_slot(e, p + "attn1.qkv_proj.weight", &y.qkv_w, 3 * C * C, &rc); add_slot(e, p + "attn1.qkv_proj.bias", &y.qkv_b, 3 * C, &rc);
        add_slot(e, p + "attn1.out_proj.weight", &y.o_w, C * C, &rc); add_slot(e, p + "attn1.out_proj.bias", &y.o_b, C, &rc);
        add_slot(e, p + "attn2.q_proj.weight", &y.q_w, C * C, &rc); add_slot(e, p + "attn2.q_proj.bias", &y.q_b, C, &rc);
        if (!rc) rc = dalloc(e, &y.kv_w, 2 * C * C);
        if (!rc) rc = dalloc(e, &y.kv_b, 2 * C);
        if (!rc) {
            e->slots[p + "attn2.k_proj.weight"] = Slot{y.kv_w, C * C}; e->slots[p + "attn2.v_proj.weight"] = Slot{y.kv_w + C * C, C * C};
            e->slots[p + "attn2.k_proj.bias"] = Slot{y.kv_b, C}; e->slots[p + "attn2.v_proj.bias"] = Slot{y.kv_b + C, C};
        }
        add_slot(e, p + "attn2.out_proj.weight", &y.o2_w, C * C, &rc); add_slot(e, p + "attn2.out_proj.bias", &y.o2_b, C, &rc);
        add_slot(e, p + "ff.net.0.weight", &y.ff1_w, 8 * C * C, &rc); add_slot(e, p + "ff.net.0.bias", &y.ff1_b, 8 * C, &rc);
        add_slot(e, p + "ff.net.2.weight", &y.ff2_w, 4 * C * C, &rc); add_slot(e, p + "ff.net.2.bias", &y.ff2_b, C, &rc);
        add_slot(e, p + "scale_shift_table", &y.table, 6 * C, &rc);
        if (!rc) rc = dalloc(e, &y.ff1_wi, 8 * C * C);
        if (!rc) rc = dalloc(e, &y.ff1_bi, 8 * C);
    }
    if (!rc) rc = dalloc(e, &e->ada_cur, 6 * C);
    if (!rc) rc = dalloc(e, &e->temb_cur, C);
    if (!rc) rc = dalloc(e, &e->coef_cur, 1);
    if (!rc) rc = dalloc(e, &e->counter, 1);
    if (!rc) rc = dalloc(e, &e->cconst16, (size_t)e->NL * C + 64);
    if (rc) { er_dit_destroy(e); return rc; }
    *out = e;
    return ER_OK;
}

extern "C" void er_dit_destroy(er_dit* e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
    if (e->graph) cudaGraphExecDestroy(e->graph);
    for (void* p : e->allocs) cudaFree(p);
    delete e;
}

extern "C" int er_dit_load_weight(er_dit* e, const char* name, const void* data_dev, int32_t dtype, int64_t numel, void* stream) {
    if (!e || !name || !data_dev) return er_set_error(ER_ERR_INVALID, "null argument");
    auto it = e->slots.find(name);
    if (it == e->slots.end()) return er_set_error(ER_ERR_INVALID, "unknown DiT tensor '%s'", name);
    if ((size_t)numel != it->second.n) return er_set_error(ER_ERR_INVALID, "tensor '%s': %lld elements, expected %zu", name, (long long)numel, it->second.n);
    if (dtype != ER_DTYPE_F16 && dtype != ER_DTYPE_F32) return er_set_error(ER_ERR_INVALID, "dtype");
    CKL(e, (dit_convert_kernel<<<(unsigned)((numel + 255) / 256), 256, 0, (cudaStream_t)stream>>>(data_dev, dtype, it->second.dst, (size_t)numel), cudaGetLastError()));
    e->loaded.insert(name);
    e->finalized = false;
    return ER_OK;
}

extern "C" int er_dit_finalize_weights(er_dit* e, void* stream) {
    if (!e) return er_set_error(ER_ERR_INVALID, "null engine");
    for (auto& kv : e->slots)
        if (!e->loaded.count(kv.first)) return er_set_error(ER_ERR_STATE, "DiT tensor '%s' was not loaded", kv.first.c_str());
    for (int l = 0; l < e->NL; l++) {          // value / gate rows of ff.net.0 interleaved in groups of 16 for the fused GEGLU epilogue
        const int C = e->C;
        CKL(e, (dit_interleave_kernel<<<(unsigned)(((size_t)8 * C * C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(e->L[l].ff1_w, e->L[l].ff1_wi, 4 * C, C), cudaGetLastError()));
        CKL(e, (dit_interleave_kernel<<<(unsigned)((8 * C + 255) / 256), 256, 0, (cudaStream_t)stream>>>(e->L[l].ff1_b, e->L[l].ff1_bi, 4 * C, 1), cudaGetLastError()));
    }
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    e->finalized = true;
    return ER_OK;
}

static int ensure_batch(er_dit* e, int batch) {
    if (batch <= e->max_batch) return ER_OK;
    CK(cudaDeviceSynchronize());
    if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
    dfree(e, &e->xa32); dfree(e, &e->xb32); dfree(e, &e->x16); dfree(e, &e->qkv16); dfree(e, &e->a16); dfree(e, &e->y16); dfree(e, &e->h16);
    dfree(e, &e->g16); dfree(e, &e->in16); dfree(e, &e->pred16); dfree(e, &e->c16); dfree(e, &e->kv16); dfree(e, &e->pc16); dfree(e, &e->lat_dev);
    dfree(e, &e->cond_dev);
    e->max_batch = 0;
    const size_t rows = (size_t)batch * e->N + 128, crows = (size_t)batch * e->M + 128, C = e->C;      // + 128: TMA boxes past the last row stay inside the allocation
    ALLOC(e->xa32, rows * C); ALLOC(e->xb32, rows * C); ALLOC(e->x16, rows * C); ALLOC(e->qkv16, rows * 3 * C); ALLOC(e->a16, rows * C);
    ALLOC(e->y16, rows * C); ALLOC(e->h16, rows * 8 * C); ALLOC(e->g16, rows * 4 * C); ALLOC(e->in16, rows * e->DL); ALLOC(e->pred16, rows * e->DL);
    ALLOC(e->c16, crows * C); ALLOC(e->kv16, (size_t)e->NL * crows * 2 * C); ALLOC(e->pc16, crows * C);
    ALLOC(e->lat_dev, rows * e->DL); ALLOC(e->cond_dev, crows * C);
    e->max_batch = batch;
    return ER_OK;
}
static int ensure_steps(er_dit* e, int steps) {
    if (steps <= e->max_steps) return ER_OK;
    CK(cudaDeviceSynchronize());
    if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
    dfree(e, &e->t_dev); dfree(e, &e->te256); dfree(e, &e->te1); dfree(e, &e->te1s); dfree(e, &e->temb); dfree(e, &e->tembs); dfree(e, &e->ada);
    dfree(e, &e->coef_dev);
    e->max_steps = 0;
    const size_t S = steps + 128, C = e->C;
    ALLOC(e->t_dev, S); ALLOC(e->te256, S * 256); ALLOC(e->te1, S * C); ALLOC(e->te1s, S * C); ALLOC(e->temb, S * C); ALLOC(e->tembs, S * C);
    ALLOC(e->ada, S * 6 * C); ALLOC(e->coef_dev, S);
    e->max_steps = steps;
    return ER_OK;
}

static cudaError_t gemm(const __half* A, int lda, const __half* W, const __half* bias, int M, int N, int K, int mode, __half* out16, float* out32,
                        int ldo, const float* res32, cudaStream_t st) {
    er::GemmArgs g{};
    g.A = A; g.lda = lda; g.W = W; g.ldw = K; g.bias = bias; g.M = M; g.N = N; g.K = K; g.mode = mode;
    g.out16 = out16; g.out32 = out32; g.ldo = ldo; g.res32 = res32; g.ldr = ldo;
    return er_gemm(g, st);
}
// out32 = res32 + f16(gate * f16(A W^T + bias)), gate = f16(tab + t[b]); out16 (optional) = f16(out32): `x = x + gate * f(x)` in the GEMM epilogue
static cudaError_t gemm_gate(const __half* A, int lda, const __half* W, const __half* bias, int M, int N, int K, const float* res32, float* out32,
                             __half* out16, const __half* tab, const __half* t, long long t_bs, int n_per, cudaStream_t st) {
    er::GemmArgs g{};
    g.A = A; g.lda = lda; g.W = W; g.ldw = K; g.bias = bias; g.M = M; g.N = N; g.K = K; g.mode = er::GEMM_GATE_RES32;
    g.out16 = out16; g.out32 = out32; g.ldo = N; g.res32 = res32; g.ldr = N; g.gate_tab = tab; g.gate_t = t; g.gate_bs = t_bs; g.n_per = n_per;
    return er_gemm(g, st);
}

// timestep MLP + adaLN vectors for n timesteps already in e->t_dev: temb [n][C], ada [n][6C]   (dit.py:178-180)
static int timestep_path(er_dit* e, int n, cudaStream_t st) {
    const int C = e->C;
    CKL(e, (dit_timestep_kernel<<<n, 128, 0, st>>>(e->t_dev, e->te256), cudaGetLastError()));
    CKL(e, gemm(e->te256, 256, e->t1_w, e->t1_b, n, C, 256, er::GEMM_F16, e->te1, nullptr, C, nullptr, st));
    CKL(e, (dit_silu_kernel<<<(unsigned)(((size_t)n * C + 255) / 256), 256, 0, st>>>(e->te1, e->te1s, (size_t)n * C), cudaGetLastError()));
    CKL(e, gemm(e->te1s, C, e->t2_w, e->t2_b, n, C, C, er::GEMM_F16, e->temb, nullptr, C, nullptr, st));
    CKL(e, (dit_silu_kernel<<<(unsigned)(((size_t)n * C + 255) / 256), 256, 0, st>>>(e->temb, e->tembs, (size_t)n * C), cudaGetLastError()));
    CKL(e, gemm(e->tembs, C, e->ada_w, e->ada_b, n, 6 * C, C, er::GEMM_F16, e->ada, nullptr, 6 * C, nullptr, st));
    return ER_OK;
}
// K | V of the condition rows for every layer (CrossAttention.k_proj / v_proj, attention.py:147-148): kv16 [NL][batch * M][2C]
// Samples [0, uncond) have an all-zero condition: no K / V rows are needed for them (see denoiser()); their cross-attention branch is the
// constant row cconst16[l] = out_proj(value bias).
static int cond_kv(er_dit* e, int batch, int uncond, cudaStream_t st) {
    const int C = e->C, rows = batch * e->M, skip = uncond * e->M;
    for (int l = 0; l < e->NL; l++) {
        CKL(e, gemm(e->c16 + (size_t)skip * C, C, e->L[l].kv_w, e->L[l].kv_b, rows - skip, 2 * C, C, er::GEMM_F16,
                    e->kv16 + ((size_t)l * rows + skip) * 2 * C, nullptr, 2 * C, nullptr, st));
        if (uncond) CKL(e, gemm(e->L[l].kv_b + C, C, e->L[l].o2_w, e->L[l].o2_b, 1, C, C, er::GEMM_F16, e->cconst16 + (size_t)l * C, nullptr, C, nullptr, st));
    }
    return ER_OK;
}

// One denoiser forward over `batch` samples: in16 [batch][N][DL] -> pred16 [batch][N][DL].  adaLN vectors at ada / temb with sample stride
// (ada_bs, temb_bs) (0: all samples share one timestep).  Needs cond_kv() for this batch.
// uncond > 0 (a guided step: samples [0, uncond) carry the all-zero condition): their cross-attention is skipped.  With a zero condition every
// key / value row of such a sample equals the projection bias, so softmax weights are uniform and the attention output is the value bias
// exactly (k * v_fp16 is exact in fp32 for k <= 257); out_proj of identical rows is one row, computed once per run by the same GEMM kernel
// (cconst16, bit-identical to any row of the full product).  The branch then is `x += const` and is folded into the next LayerNorm's load.
static int denoiser(er_dit* e, int batch, const __half* ada, long long ada_bs, const __half* temb, long long temb_bs, int uncond, cudaStream_t st) {
    const int C = e->C, N = e->N, H = e->H, D = e->D, rows = batch * N, crows = batch * e->M;
    const int urows = uncond * N, ucrows = uncond * e->M;
    const unsigned ew4 = (unsigned)(((size_t)rows * C / 4 + 255) / 256);
    const bool fuse = e->fuse && !(C & 31);      // GEGLU and the two gated residuals in the GEMM epilogues (3 kernels and ~0.9 GB of traffic less per layer)
    CKL(e, gemm(e->in16, e->DL, e->pin_w, e->pin_b, rows, C, e->DL, er::GEMM_F16, e->x16, nullptr, C, nullptr, st));
    float* xin32 = nullptr;            // layer 0 normalises the fp16 (proj_in + pos_embed) tensor, later layers the fp32 stream
    for (int l = 0; l < e->NL; l++) {
        const er_dit::Layer& y = e->L[l];
        // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = rows 0..5 (dit.py:125)
        CKL(e, ln_mod(xin32, xin32 ? nullptr : e->x16, xin32 ? nullptr : e->pos, y.table, y.table + C, ada, ada + C, ada_bs, N, e->xa32, e->x16, rows, C, nullptr,
                      nullptr, 0, st));
        CKL(e, gemm(e->x16, C, y.qkv_w, y.qkv_b, rows, 3 * C, C, er::GEMM_F16, e->qkv16, nullptr, 3 * C, nullptr, st));
        er::AttnArgs a{};
        a.q = e->qkv16; a.k = e->qkv16 + C; a.v = e->qkv16 + 2 * C; a.out = e->a16;
        a.ldq = a.ldk = a.ldv = 3 * C; a.ldo = C; a.q_bs = a.k_bs = a.v_bs = (long long)N * 3 * C; a.o_bs = (long long)N * C;
        a.B = batch; a.H = H; a.Nq = N; a.Nk = N; a.D = D; a.causal = 0;
        CKL(e, er_attention(a, st));
        if (fuse) {
            CKL(e, gemm_gate(e->a16, C, y.o_w, y.o_b, rows, C, C, e->xa32, e->xb32, e->x16, y.table + 2 * C, ada + 2 * C, ada_bs, N, st));
        } else {
            CKL(e, gemm(e->a16, C, y.o_w, y.o_b, rows, C, C, er::GEMM_F16, e->y16, nullptr, C, nullptr, st));
            CKL(e, (dit_gate_res_kernel<<<ew4, 256, 0, st>>>(e->xa32, e->y16, y.table + 2 * C, ada + 2 * C, ada_bs, N, e->xb32, e->x16, (size_t)rows, C), cudaGetLastError()));
        }
        // cross-attention to the condition (dit.py:131): x = x + attn2(x, c)
        const size_t uo = (size_t)urows * C;
        CKL(e, gemm(e->x16 + uo, C, y.q_w, y.q_b, rows - urows, C, C, er::GEMM_F16, e->qkv16 + uo, nullptr, C, nullptr, st));
        const __half* kv = e->kv16 + ((size_t)l * crows + ucrows) * 2 * C;
        er::AttnArgs x{};
        x.q = e->qkv16 + uo; x.k = kv; x.v = kv + C; x.out = e->a16 + uo;
        x.ldq = C; x.ldk = x.ldv = 2 * C; x.ldo = C; x.q_bs = x.o_bs = (long long)N * C; x.k_bs = x.v_bs = (long long)e->M * 2 * C;
        x.B = batch - uncond; x.H = H; x.Nq = N; x.Nk = e->M; x.D = D; x.causal = 0;
        CKL(e, er_attention(x, st));
        CKL(e, gemm(e->a16 + uo, C, y.o2_w, y.o2_b, rows - urows, C, C, er::GEMM_F32_RES32, nullptr, e->xa32 + uo, C, e->xb32 + uo, st));
        // feed-forward (dit.py:133-136)
        CKL(e, ln_mod(e->xa32, nullptr, nullptr, y.table + 3 * C, y.table + 4 * C, ada + 3 * C, ada + 4 * C, ada_bs, N, e->xb32, e->x16, rows, C, e->xb32,
                      e->cconst16 + (size_t)l * C, urows, st));
        if (fuse) {
            CKL(e, gemm(e->x16, C, y.ff1_wi, y.ff1_bi, rows, 8 * C, C, er::GEMM_F16_GEGLU, e->g16, nullptr, 4 * C, nullptr, st));
            CKL(e, gemm_gate(e->g16, 4 * C, y.ff2_w, y.ff2_b, rows, C, 4 * C, e->xb32, e->xa32, nullptr, y.table + 5 * C, ada + 5 * C, ada_bs, N, st));
        } else {
            CKL(e, gemm(e->x16, C, y.ff1_w, y.ff1_b, rows, 8 * C, C, er::GEMM_F16, e->h16, nullptr, 8 * C, nullptr, st));
            CKL(e, (dit_geglu_kernel<<<(unsigned)(((size_t)rows * 4 * C / 8 + 255) / 256), 256, 0, st>>>(e->h16, e->g16, (size_t)rows, 4 * C), cudaGetLastError()));
            CKL(e, gemm(e->g16, 4 * C, y.ff2_w, y.ff2_b, rows, C, 4 * C, er::GEMM_F16, e->y16, nullptr, C, nullptr, st));
            CKL(e, (dit_gate_res_kernel<<<ew4, 256, 0, st>>>(e->xb32, e->y16, y.table + 5 * C, ada + 5 * C, ada_bs, N, e->xa32, e->x16, (size_t)rows, C), cudaGetLastError()));
        }
        xin32 = e->xa32;
    }
    // shift, scale = (table2 + t_emb).chunk(2) ; norm_out ; modulate ; proj_out  (dit.py:189-194)
    CKL(e, ln_mod(xin32, nullptr, nullptr, e->table2, e->table2 + C, temb, temb, temb_bs, N, nullptr, e->x16, rows, C, nullptr, nullptr, 0, st));
    CKL(e, gemm(e->x16, C, e->pout_w, e->pout_b, rows, e->DL, C, er::GEMM_F16, e->pred16, nullptr, e->DL, nullptr, st));
    return ER_OK;
}

extern "C" int er_dit_cond(er_dit* e, const void* clip_hidden_dev, int32_t B, float* cond_out_dev, void* stream) {
    if (!e || !clip_hidden_dev || !cond_out_dev || B <= 0) return er_set_error(ER_ERR_INVALID, "bad argument");
    if (!e->finalized) return er_set_error(ER_ERR_STATE, "DiT weights are not finalized");
    CK(cudaSetDevice(e->cfg.device));
    int rc = ensure_batch(e, B);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = B * e->M, C = e->C;
    // norm_cond(proj_cond(h)) (models_dit.py:116): Linear -> fp16, LayerNorm(eps 1e-5, affine) -> fp32
    CKL(e, gemm((const __half*)clip_hidden_dev, e->CD, e->pc_w, e->pc_b, rows, C, e->CD, er::GEMM_F16, e->pc16, nullptr, C, nullptr, st));
    CKL(e, er_layernorm(nullptr, e->pc16, C, e->nc_g, e->nc_b, cond_out_dev, nullptr, C, rows, C, st));
    return ER_OK;
}

extern "C" int er_dit_forward(er_dit* e, const float* x_dev, const float* cond_dev, const float* t_dev, int32_t B, void* out_dev, void* stream) {
    if (!e || !x_dev || !cond_dev || !t_dev || !out_dev || B <= 0) return er_set_error(ER_ERR_INVALID, "bad argument");
    if (!e->finalized) return er_set_error(ER_ERR_STATE, "DiT weights are not finalized");
    CK(cudaSetDevice(e->cfg.device));
    int rc = ensure_batch(e, B);
    if (!rc) rc = ensure_steps(e, B);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nx = (size_t)B * e->N * e->DL, nc = (size_t)B * e->M * e->C;
    CKL(e, er_f32_to_f16(x_dev, e->in16, nx, st));
    CKL(e, er_f32_to_f16(cond_dev, e->c16, nc, st));
    CK(cudaMemcpyAsync(e->t_dev, t_dev, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
    if ((rc = timestep_path(e, B, st))) return rc;
    if ((rc = cond_kv(e, B, 0, st))) return rc;
    if ((rc = denoiser(e, B, e->ada, 6LL * e->C, e->temb, e->C, 0, st))) return rc;
    CK(cudaMemcpyAsync(out_dev, e->pred16, nx * sizeof(__half), cudaMemcpyDeviceToDevice, st));
    return ER_OK;
}

// the per-step work: select step constants -> [latents] * 2 -> denoiser -> guidance + scheduler update
static int one_step(er_dit* e, float* lat, int R, int guided, float gscale, int v_pred, cudaStream_t st) {
    const int batch = guided ? 2 * R : R;
    const size_t n = (size_t)R * e->N * e->DL;
    CKL(e, (dit_select_step_kernel<<<8, 256, 0, st>>>(e->counter, e->ada, e->temb, e->coef_dev, e->ada_cur, e->temb_cur, e->coef_cur, e->C), cudaGetLastError()));
    if (guided) CKL(e, (dit_dup_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(lat, e->in16, n), cudaGetLastError()));
    else CKL(e, er_f32_to_f16(lat, e->in16, n, st));
    int rc = denoiser(e, batch, e->ada_cur, 0, e->temb_cur, 0, guided && e->uncond_shortcut ? R : 0, st);
    if (rc) return rc;
    CKL(e, (dit_guide_step_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(e->pred16, lat, n, gscale, guided, v_pred, e->coef_cur, e->counter), cudaGetLastError()));
    return ER_OK;
}

extern "C" int er_dit_run(er_dit* e, const float* cond_dev, float* latents_dev, int32_t R, int32_t n_steps, const float* timesteps_host,
                          const float* coef_host, float guidance_scale, int32_t guided, int32_t prediction_type, void* stream) {
    if (!e || !cond_dev || !latents_dev || R <= 0 || n_steps < 0 || (n_steps && (!timesteps_host || !coef_host)))
        return er_set_error(ER_ERR_INVALID, "bad argument");
    if (prediction_type != ER_DIT_PRED_EPSILON && prediction_type != ER_DIT_PRED_V) return er_set_error(ER_ERR_INVALID, "prediction_type");
    if (!e->finalized) return er_set_error(ER_ERR_STATE, "DiT weights are not finalized");
    CK(cudaSetDevice(e->cfg.device));
    const int batch = guided ? 2 * R : R;
    int rc = ensure_batch(e, batch);
    if (!rc) rc = ensure_steps(e, n_steps);
    if (rc) return rc;
    if (n_steps == 0) return ER_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nc = (size_t)R * e->M * e->C;
    // hoisted out of the loop: condition K / V of every layer, timestep MLP + adaLN vectors of every step, scheduler coefficients
    if (guided) CKL(e, (dit_cfg_cond_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(cond_dev, e->c16, nc), cudaGetLastError()));
    else CKL(e, er_f32_to_f16(cond_dev, e->c16, nc, st));
    if ((rc = cond_kv(e, batch, guided && e->uncond_shortcut ? R : 0, st))) return rc;
    CK(cudaMemcpyAsync(e->t_dev, timesteps_host, sizeof(float) * n_steps, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->coef_dev, coef_host, sizeof(StepCoef) * n_steps, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));        // the host arrays may be pageable / reused by the caller
    if ((rc = timestep_path(e, n_steps, st))) return rc;
    CKL(e, (dit_set_int_kernel<<<1, 1, 0, st>>>(e->counter, 0), cudaGetLastError()));
    const int v_pred = prediction_type == ER_DIT_PRED_V;
    if (!e->use_graph) {
        for (int s = 0; s < n_steps; s++)
            if ((rc = one_step(e, latents_dev, R, guided, guidance_scale, v_pred, st))) return rc;
        return ER_OK;
    }
    if (!e->graph || e->graph_batch != batch || e->graph_guided != guided || e->graph_vpred != v_pred || e->graph_gscale != guidance_scale ||
        e->graph_lat != latents_dev || e->graph_fuse != e->fuse) {
        if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; }
        cudaStream_t cs;
        CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaError_t ce = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
        if (ce != cudaSuccess) { cudaStreamDestroy(cs); return er_set_error(ER_ERR_CUDA, "graph capture: %s", cudaGetErrorString(ce)); }
        const long long before = e->launches;
        rc = one_step(e, latents_dev, R, guided, guidance_scale, v_pred, cs);
        e->graph_kernels = e->launches - before;      // kernels one replay launches (3 + 10 per layer + 3; 13 per layer with the epilogue fusions off)
        e->launches = before;                         // capturing launched nothing
        cudaGraph_t g = nullptr;
        ce = cudaStreamEndCapture(cs, &g);
        cudaStreamDestroy(cs);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) return er_set_error(ER_ERR_CUDA, "graph capture: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&e->graph, g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) { e->graph = nullptr; return er_set_error(ER_ERR_CUDA, "graph instantiate: %s", cudaGetErrorString(ce)); }
        e->graph_batch = batch; e->graph_guided = guided; e->graph_vpred = v_pred; e->graph_gscale = guidance_scale; e->graph_lat = latents_dev; e->graph_fuse = e->fuse;
    }
    for (int s = 0; s < n_steps; s++) { CK(cudaGraphLaunch(e->graph, st)); e->launches += e->graph_kernels; }
    return ER_OK;
}

extern "C" int er_dit_run_host(er_dit* e, const float* cond_host, float* latents_host, int32_t R, int32_t n_steps, const float* timesteps_host,
                               const float* coef_host, float guidance_scale, int32_t guided, int32_t prediction_type) {
    if (!e || !cond_host || !latents_host || R <= 0) return er_set_error(ER_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(e->cfg.device));
    int rc = ensure_batch(e, guided ? 2 * R : R);
    if (rc) return rc;
    const size_t nl = (size_t)R * e->N * e->DL, nc = (size_t)R * e->M * e->C;
    CK(cudaMemcpyAsync(e->lat_dev, latents_host, nl * sizeof(float), cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(e->cond_dev, cond_host, nc * sizeof(float), cudaMemcpyHostToDevice, 0));
    if ((rc = er_dit_run(e, e->cond_dev, e->lat_dev, R, n_steps, timesteps_host, coef_host, guidance_scale, guided, prediction_type, nullptr))) return rc;
    CK(cudaMemcpyAsync(latents_host, e->lat_dev, nl * sizeof(float), cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return ER_OK;
}

extern "C" int64_t er_dit_kernel_launches(const er_dit* e) { return e ? e->launches : 0; }
extern "C" int er_dit_debug_set(er_dit* e, const char* key, int64_t value) {
    if (!e || !key) return er_set_error(ER_ERR_INVALID, "null argument");
    if (!strcmp(key, "graph")) { e->use_graph = value != 0; return ER_OK; }
    if (!strcmp(key, "fuse")) { e->fuse = value != 0; return ER_OK; }
    if (!strcmp(key, "uncond_shortcut")) { e->uncond_shortcut = value != 0; if (e->graph) { cudaGraphExecDestroy(e->graph); e->graph = nullptr; } return ER_OK; }
    return er_set_error(ER_ERR_INVALID, "unknown key '%s'", key);
}
// algorithmic FLOPs of one denoiser forward over `batch` samples (GEMMs + attention; bench.py roofline)
extern "C" double er_dit_flops_per_forward(const er_dit* e, int32_t batch) {
    if (!e) return 0;
    const double C = e->C, N = e->N, M = e->M, rows = batch * N, crows = batch * M;
    double per_layer = 2 * rows * C * (3 * C + C + C + C + 8 * C + 4 * C) + 4.0 * batch * N * N * C + 4.0 * batch * N * M * C;
    return e->NL * per_layer + 2 * rows * C * 2 * e->DL + 0 * crows;
}
