// Persistent auto-regressive decode kernel (sm_100a).
//
// Replaces, for every generated token, the whole per-step stack of the reference:
//   HF _sample step (logits -> constraint mask -> argmax | top-k sample)            third-party, restated
//   constraint FSM                         /root/reference/core/models.py:245-271
//   ShapeOPTDecoder.forward (N == 1)       /root/reference/core/transformer/modeling_opt.py:340-357
//   24 x OPTDecoderLayer.forward           modeling_opt.py:264-288   (post-LN, ReLU MLP)
//   OptFlashAttention2.forward             modeling_opt.py:185-232   (q/k/v GEMV, KV append, 1 x L attention)
//   lm_head                                modeling_opt.py:497
//
// One cooperative launch generates up to `steps` tokens: one CTA per SM, all CTAs walk the same phase
// list and meet at a grid barrier between dependent phases (5 per layer).  The token loop, the FSM and the
// sampler live on the device, so there is no host round trip per token (the reference has >= 5).
//
// HBM-bound by construction (B = 1 GEMV + single-query attention, ~1 flop/byte): every weight byte and every
// cached K/V byte is read exactly once per token with 128-bit coalesced loads; the new K/V row is appended
// in place (the reference re-allocates and copies the whole cache per layer per step, modeling_opt.py:191).
// The next phase's weight slice is bulk-prefetched into L2 (TMA unit, cp.async.bulk.prefetch.L2) while the
// current phase runs, so HBM stays busy across the grid barriers.
//
// dtype ledger (mirrored by oracle/er_oracle.py mode='ledger'): fp16 weights and KV,
// fp32 accumulation everywhere, activations rounded to fp16 exactly where model.half()+autocast(fp16) does.
#include "decode_kernel.h"

#include "common.cuh"

namespace er {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int HD = 96;          // decoder head_dim (ArAE: 1536 / 16)
constexpr int HV = HD / 8;      // 16-byte vectors per head row (12)

// ---- grid barrier ----------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += 1;
        const unsigned target = epoch * gridDim.x;
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

// ---- optional phase timeline (debug / profiles): CTA `prof_cta`, thread 0 stamps %globaltimer ------------------------
__device__ __forceinline__ void prof_stamp(const DecodeParams& p, int& slot, bool on) {
    if (on && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.prof[slot] = t;
    }
    slot++;
}

// ---- block reductions (512 threads) ------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kWarps; i++) t += red[i];
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = -INFINITY;
#pragma unroll
    for (int i = 0; i < kWarps; i++) t = fmaxf(t, red[i]);
    return t;
}

// LayerNorm over xres[0..C) in place (fp32, eps 1e-5, affine fp16 params), also emits the fp16 copy.
__device__ __noinline__ void layer_norm_inplace(float* xres, __half* x16, const __half* __restrict__ g, const __half* __restrict__ b,
                                   int C, float* red) {
    float s = 0.f;
    for (int i = threadIdx.x; i < C; i += kThreads) s += xres[i];
    const float mean = block_sum(s, red) / C;
    float q = 0.f;
    for (int i = threadIdx.x; i < C; i += kThreads) { float d = xres[i] - mean; q += d * d; }
    const float var = block_sum(q, red) / C;
    const float rstd = rsqrtf(var + 1e-5f);
    for (int i = threadIdx.x; i < C; i += kThreads) {
        float y = (xres[i] - mean) * rstd * __half2float(g[i]) + __half2float(b[i]);
        xres[i] = y;
        x16[i] = __float2half_rn(y);
    }
    __syncthreads();
}

// ---- GEMV phase --------------------------------------------------------------------------------------------
// Rows [0,R) of W[R][K] are split into contiguous per-CTA ranges; inside the CTA a unit = (row, k-slice of KU
// elements) and warps take units round-robin, two at a time (12 independent 16-byte loads in flight per lane).
struct RowRange { int r0, r1; };
__device__ __forceinline__ RowRange cta_rows(int R) {
    RowRange rr;
    rr.r0 = (int)(((long long)R * blockIdx.x) / gridDim.x);
    rr.r1 = (int)(((long long)R * (blockIdx.x + 1)) / gridDim.x);
    return rr;
}

constexpr int kMaxUnits = 256;

// partial dot products of this CTA's units -> red_units[unit]
__device__ __noinline__ void gemv_units(const __half* __restrict__ W, int K, int KU, RowRange rr, const __half* xin, float* red_units) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ks = K / KU;                       // k-slices per row
    const int nunits = (rr.r1 - rr.r0) * ks;
    const int nvec = KU >> 3;                    // 16-byte vectors per unit
    const uint4* xv = reinterpret_cast<const uint4*>(xin);
    for (int u = warp; u < nunits; u += 2 * kWarps) {
        const int u2 = u + kWarps;
        const bool has2 = u2 < nunits;
        const int rowA = rr.r0 + u / ks, kA = (u % ks) * nvec;
        const int rowB = has2 ? rr.r0 + u2 / ks : rowA, kB = has2 ? (u2 % ks) * nvec : kA;
        const uint4* wA = reinterpret_cast<const uint4*>(W + (size_t)rowA * K) + kA;
        const uint4* wB = reinterpret_cast<const uint4*>(W + (size_t)rowB * K) + kB;
        float accA = 0.f, accB = 0.f;
        for (int base = 0; base < nvec; base += 32 * 6) {
            uint4 a[6], b[6];
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const int v = base + j * 32 + lane;
                if (v < nvec) { a[j] = ldg_stream(wA + v); if (has2) b[j] = ldg_stream(wB + v); }
            }
#pragma unroll
            for (int j = 0; j < 6; j++) {
                const int v = base + j * 32 + lane;
                if (v < nvec) {
                    accA = dot8(a[j], xv[kA + v], accA);
                    if (has2) accB = dot8(b[j], xv[kB + v], accB);
                }
            }
        }
        accA = warp_sum(accA);
        accB = warp_sum(accB);
        if (lane == 0) {
            red_units[u] = accA;
            if (has2) red_units[u2] = accB;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ float unit_row_sum(const float* red_units, int local_row, int ks) {
    float s = 0.f;
    for (int j = 0; j < ks; j++) s += red_units[local_row * ks + j];
    return s;
}

// ---- sampler (HF _sample step + constraint FSM), executed redundantly by warp 0 of every CTA ----------------
__device__ __forceinline__ uint32_t mix32(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (uint32_t)(x >> 32);
}

// FSM transition on the last fed token (core/models.py:254-260); executed by every thread so all agree.
__device__ __forceinline__ void fsm_update(int& counter, int last_tok) {
    if (last_tok == 5) counter = 9;
    else if (last_tok == 3 || last_tok == 4) counter = 3;
    else if (last_tok >= 6) counter -= 1;
}

__device__ __noinline__ int sample_warp(const float* __restrict__ logits_pre, float* sc, int V, int step, int counter,
                           const DecodeParams& p) {
    const int lane = threadIdx.x & 31;
    int lo = 0, hi = 0;    // allowed = [lo, hi) plus the specials below
    bool specials = false; // {3,4,5,eos}
    bool only_bom = false;
    bool eos_extra = false;
    if (p.use_fsm) {       // core/models.py:252,264-268
        if (step == 0) only_bom = true;
        else if (counter > 0) { lo = 6; hi = V; }
        else specials = true;
    } else {               // core/models.py:237-242
        lo = 3; hi = V; eos_extra = (step % 9 == 1);
    }
    // scores = float(fp16(logit)) + mask
    for (int i = lane; i < V; i += 32) {
        bool ok;
        if (only_bom) ok = (i == 5);
        else if (specials) ok = (i == 3 || i == 4 || i == 5 || i == p.eos);
        else ok = (i >= lo && i < hi) || (eos_extra && i == p.eos);
        sc[i] = ok ? round_f16(ldg_cg_f32(logits_pre + i)) : -INFINITY;
    }
    __syncwarp();
    // argmax, lowest index wins ties (torch.argmax on CUDA)
    float bv = -INFINITY; int bi = 0x7fffffff;
    for (int i = lane; i < V; i += 32) { float v = sc[i]; if (v > bv) { bv = v; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (p.mode == 0) return bi;
    // ---- sample: TopKLogitsWarper(top_k) keeps ties with the k-th value; softmax; inverse-CDF draw ----------
    const float vmax = bv;
    float thr = vmax;
    {
        // k-th largest with multiplicity: walk down distinct values counting occurrences
        int taken = 0; float cur = INFINITY;
        while (taken < p.top_k) {
            float nb = -INFINITY; int cnt = 0;
            for (int i = lane; i < V; i += 32) { float v = sc[i]; if (v < cur && v > nb) nb = v; }
            nb = warp_max(nb);
            for (int i = lane; i < V; i += 32) cnt += (sc[i] == nb);
            cnt = (int)warp_sum((float)cnt);
            thr = nb; taken += cnt; cur = nb;
            if (nb == -INFINITY) break;
        }
    }
    // per-lane contiguous segment sums of exp(score - max) over kept entries
    const int seg = (V + 31) / 32;
    const int s0 = lane * seg, s1 = min(V, s0 + seg);
    float local = 0.f;
    for (int i = s0; i < s1; i++) { float v = sc[i]; if (v >= thr && v > -INFINITY) local += __expf(v - vmax); }
    float incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    const float total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t r = mix32(p.seed * 0x100000001B3ull + (uint64_t)step);
    const float target = ((r >> 8) + 0.5f) * (1.0f / 16777216.0f) * total;
    const float excl = incl - local;
    int pick = -1;
    if (target >= excl && target < incl) {
        float run = excl;
        for (int i = s0; i < s1; i++) {
            float v = sc[i];
            if (v >= thr && v > -INFINITY) { run += __expf(v - vmax); pick = i; if (run > target) break; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pick = max(pick, __shfl_xor_sync(0xffffffffu, pick, o));
    return pick < 0 ? bi : pick;
}

// ---- P2: single-query attention over the cached keys 0..L, one (head, KV split) per CTA ------------------------
// K pass: lane = key inside a 32-key block (12 coalesced 512-byte loads per block, no cross-lane reduction);
// V pass: thread = (key group, 16-byte vector of the head row).  Writes (o[96], max, sum) of the split.
__device__ __noinline__ void attention_phase(const DecodeParams& p, int layer, int L, float* qs, float* sc, float* vred,
                                             float* red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = p.H;
    const size_t nkb = (size_t)p.nkb;
    const float sm_scale_log2 = rsqrtf((float)HD) * 1.4426950408889634f;
    const int nkeys = L + 1;
    const int nblk = (nkeys + 31) >> 5;
    const int bps = (nblk + p.S - 1) / p.S;
    if ((int)blockIdx.x < H * p.S) {
        const int h = blockIdx.x / p.S, s = blockIdx.x % p.S;
        const int b0 = min(s * bps, nblk), b1 = min(b0 + bps, nblk);
        const int k0 = b0 * 32, k1 = min(b1 * 32, nkeys);
        float* outp = p.part + ((size_t)h * p.S + s) * 100;
        if (k1 <= k0) {
            for (int i = tid; i < 100; i += kThreads) outp[i] = (i == HD) ? -INFINITY : 0.f;
        } else {
            if (tid < HD) qs[tid] = __half2float(__ushort_as_half(ldg_cg_u16(p.q16 + h * HD + tid)));
            __syncthreads();
            // K pass: lane = key inside a 32-key block, 12 coalesced 512-byte loads per block
            const __half* kbase = p.kc + (((size_t)layer * H + h) * nkb) * (HV * 256);
            float lmax = -INFINITY;
            for (int b = b0 + warp; b < b1; b += kWarps) {
                const uint4* kb = reinterpret_cast<const uint4*>(kbase + (size_t)b * (HV * 256)) + lane;
                uint4 kv[HV];
#pragma unroll
                for (int j = 0; j < HV; j++) kv[j] = ldg_cg(kb + j * 32);
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < HV; j++) {
                    const float4 qa = *reinterpret_cast<const float4*>(qs + j * 8);
                    const float4 qb = *reinterpret_cast<const float4*>(qs + j * 8 + 4);
                    float2 f;
                    f = h2f2(kv[j].x); acc = fmaf(f.x, qa.x, acc); acc = fmaf(f.y, qa.y, acc);
                    f = h2f2(kv[j].y); acc = fmaf(f.x, qa.z, acc); acc = fmaf(f.y, qa.w, acc);
                    f = h2f2(kv[j].z); acc = fmaf(f.x, qb.x, acc); acc = fmaf(f.y, qb.y, acc);
                    f = h2f2(kv[j].w); acc = fmaf(f.x, qb.z, acc); acc = fmaf(f.y, qb.w, acc);
                }
                const int key = b * 32 + lane;
                const float sv = (key < nkeys) ? acc : -INFINITY;
                sc[key - k0] = sv;
                lmax = fmaxf(lmax, sv);
            }
            const float m = block_max(lmax, red);         // raw dot units (unscaled)
            float lsum = 0.f;
            for (int i = tid; i < k1 - k0; i += kThreads) {
                const float pr = exp2f((sc[i] - m) * sm_scale_log2);
                sc[i] = pr;
                lsum += pr;
            }
            const float l = block_sum(lsum, red);          // (also orders the sc[] writes before the V pass)
            // V pass: thread = (key group kg, 16-byte vector of the head row), coalesced 512-byte warp loads
            const int vec = tid % HV, kg = tid / HV;       // kg in [0,42] ; 504 active threads
            float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (kg < 42) {
                const __half* vbase = p.vc + (((size_t)layer * H + h) * p.Lmax) * HD + vec * 8;
                int key = k0 + kg;
                for (; key + 3 * 42 < k1; key += 4 * 42) {
                    uint4 v0 = ldg_cg(reinterpret_cast<const uint4*>(vbase + (size_t)key * HD));
                    uint4 v1 = ldg_cg(reinterpret_cast<const uint4*>(vbase + (size_t)(key + 42) * HD));
                    uint4 v2 = ldg_cg(reinterpret_cast<const uint4*>(vbase + (size_t)(key + 84) * HD));
                    uint4 v3 = ldg_cg(reinterpret_cast<const uint4*>(vbase + (size_t)(key + 126) * HD));
                    const float p0 = sc[key - k0], p1 = sc[key + 42 - k0], p2 = sc[key + 84 - k0], p3 = sc[key + 126 - k0];
                    float2 f;
#define ER_ACC(vv, pp)                                                                            \
    f = h2f2(vv.x); o[0] = fmaf(pp, f.x, o[0]); o[1] = fmaf(pp, f.y, o[1]);                       \
    f = h2f2(vv.y); o[2] = fmaf(pp, f.x, o[2]); o[3] = fmaf(pp, f.y, o[3]);                       \
    f = h2f2(vv.z); o[4] = fmaf(pp, f.x, o[4]); o[5] = fmaf(pp, f.y, o[5]);                       \
    f = h2f2(vv.w); o[6] = fmaf(pp, f.x, o[6]); o[7] = fmaf(pp, f.y, o[7]);
                    ER_ACC(v0, p0) ER_ACC(v1, p1) ER_ACC(v2, p2) ER_ACC(v3, p3)
                }
                for (; key < k1; key += 42) {
                    uint4 v0 = ldg_cg(reinterpret_cast<const uint4*>(vbase + (size_t)key * HD));
                    const float p0 = sc[key - k0];
                    float2 f;
                    ER_ACC(v0, p0)
                }
#undef ER_ACC
#pragma unroll
                for (int e = 0; e < 8; e++) vred[kg * HD + vec * 8 + e] = o[e];
            }
            __syncthreads();
            if (tid < HD) {
                float a = 0.f;
                for (int g = 0; g < 42; g++) a += vred[g * HD + tid];
                outp[tid] = a;
            } else if (tid == HD) {
                outp[HD] = m * rsqrtf((float)HD);      // max in softmax (scaled) units
                outp[HD + 1] = l;
            }
        }
    }
}

// ---- the kernel ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) decode_persistent_kernel(const DecodeParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = p.C, F = p.F, H = p.H, V = p.V;
    // smem carve-up
    float* xres = reinterpret_cast<float*>(smem_raw);                 // [C]   residual stream (fp32)
    __half* xin = reinterpret_cast<__half*>(xres + C);                // [max(C,F)] GEMV input (fp16)
    float* red_units = reinterpret_cast<float*>(xin + (F > C ? F : C)); // [kMaxUnits]
    float* red = red_units + kMaxUnits;                               // [32]
    float* qs = red + 32;                                             // [96] query of this CTA's head (fp32)
    float* wsplit = qs + HD;                                          // [H*S] combine weights
    float* vred = wsplit + H * p.S;                                   // [42*96] V-pass cross-thread reduction
    float* sc = vred + 42 * HD;                                       // [max(sc_keys, V)] scores / sampler scratch
    __shared__ int s_tok;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned epoch = 0;

    int t = p.st->t, L = p.st->L, counter = p.st->counter, last_tok = p.st->last_tok;
    const int done0 = p.st->done;
    if (done0) return;

    const size_t nkb = (size_t)p.nkb;                      // key blocks (32 keys) per head in the K cache

    for (int it = 0; it < p.steps; ++it, ++t) {
        const bool prof_on = p.prof != nullptr && t == p.prof_token && (int)blockIdx.x == p.prof_cta;
        int pslot = 0;
        prof_stamp(p, pslot, prof_on);
        // ================= sample token t from the current logits =================================================
        if (p.use_fsm && t > 0) fsm_update(counter, last_tok);
        if (warp == 0) {
            const int tok = sample_warp(p.logits, sc, V, t, counter, p);
            if (lane == 0) s_tok = tok;
        }
        __syncthreads();
        const int chosen = s_tok;
        const int fed = p.forced ? p.forced[t] : chosen;
        if (blockIdx.x == 0) {
            if (tid == 0) p.out_ids[t] = chosen;
            if (p.out_logits)
                for (int i = tid; i < V; i += kThreads) p.out_logits[(size_t)t * V + i] = ldg_cg_f32(p.logits + i);
        }
        last_tok = fed;
        const bool finished = (fed == p.eos) || (t + 1 >= p.max_new);
        if (finished) {
            if (blockIdx.x == 0 && tid == 0) { p.st->done = 1; p.st->t = t + 1; p.st->L = L; p.st->counter = counter; p.st->last_tok = last_tok; }
            return;
        }

        // ================= embed: x = fp16(embd[tok] + pos[L]) =====================================================
        for (int i = tid; i < C; i += kThreads) {
            float v = __half2float(p.embd[(size_t)fed * C + i]) + __half2float(p.pos[(size_t)L * C + i]);
            __half h = __float2half_rn(v);
            xin[i] = h;
            xres[i] = __half2float(h);
        }
        __syncthreads();

        for (int layer = 0; layer < p.layers; ++layer) {
            const __half* wqkv = p.wqkv + (size_t)layer * 3 * C * C;
            const __half* wo = p.wo + (size_t)layer * C * C;
            const __half* w1 = p.w1 + (size_t)layer * F * C;
            const __half* w2 = p.w2 + (size_t)layer * C * F;
            // ---------------- P1: q,k,v = x16 @ Wqkv^T + b ; KV append in place ----------------------------------
            {
                RowRange rr = cta_rows(3 * C);
                if (tid == 0) {   // next phase: out_proj slice (weights) — KV is demand-streamed
                    RowRange nx = cta_rows(C);
                    prefetch_l2_range(wo + (size_t)nx.r0 * C, (size_t)(nx.r1 - nx.r0) * C * 2);
                }
                gemv_units(wqkv, C, C, rr, xin, red_units);
                const __half* bq = p.bqkv + (size_t)layer * 3 * C;
                for (int i = tid; i < rr.r1 - rr.r0; i += kThreads) {
                    const int r = rr.r0 + i;
                    const __half hv = __float2half_rn(red_units[i] + __half2float(bq[r]));
                    if (r < C) {
                        p.q16[r] = hv;
                    } else if (r < 2 * C) {
                        const int c = r - C, h = c / HD, d = c % HD;
                        const size_t idx = ((((size_t)layer * H + h) * nkb + (L >> 5)) * HV + (d >> 3)) * 256 + (size_t)(L & 31) * 8 + (d & 7);
                        p.kc[idx] = hv;
                    } else {
                        const int c = r - 2 * C, h = c / HD, d = c % HD;
                        p.vc[(((size_t)layer * H + h) * p.Lmax + L) * HD + d] = hv;
                    }
                }
            }
            prof_stamp(p, pslot, prof_on);
            grid_barrier(p.bar, epoch);
            prof_stamp(p, pslot, prof_on);
            // ---------------- P2: single-query attention, one (head, split) per CTA ------------------------------
            {
                attention_phase(p, layer, L, qs, sc, vred, red);
            }
            prof_stamp(p, pslot, prof_on);
            grid_barrier(p.bar, epoch);
            prof_stamp(p, pslot, prof_on);
            // ---------------- P3: combine splits -> attn16 ; out_proj ---------------------------------------------
            {
                // per (head, split) weights exp(m_s - M) / sum_s exp(m_s - M) l_s
                if (tid < H) {
                    float M = -INFINITY;
                    for (int s = 0; s < p.S; s++) M = fmaxf(M, ldg_cg_f32(p.part + ((size_t)tid * p.S + s) * 100 + HD));
                    float den = 0.f;
                    for (int s = 0; s < p.S; s++) {
                        const float ms = ldg_cg_f32(p.part + ((size_t)tid * p.S + s) * 100 + HD);
                        const float ls = ldg_cg_f32(p.part + ((size_t)tid * p.S + s) * 100 + HD + 1);
                        const float w = (ms == -INFINITY) ? 0.f : __expf(ms - M);
                        wsplit[tid * p.S + s] = w;
                        den += w * ls;
                    }
                    const float inv = 1.f / den;
                    for (int s = 0; s < p.S; s++) wsplit[tid * p.S + s] *= inv;
                }
                __syncthreads();
                for (int i = tid; i < C; i += kThreads) {
                    const int h = i / HD, d = i % HD;
                    float a = 0.f;
                    for (int s = 0; s < p.S; s++) {
                        const float w = wsplit[h * p.S + s];
                        if (w != 0.f) a = fmaf(w, ldg_cg_f32(p.part + ((size_t)h * p.S + s) * 100 + d), a);
                    }
                    xin[i] = __float2half_rn(a);
                }
                __syncthreads();
                RowRange rr = cta_rows(C);
                const int ku = C / p.ks_out;
                if (tid == 0) {   // next weight phase: fc1 slice
                    RowRange nx = cta_rows(F);
                    prefetch_l2_range(w1 + (size_t)nx.r0 * C, (size_t)(nx.r1 - nx.r0) * C * 2);
                }
                gemv_units(wo, C, ku, rr, xin, red_units);
                const __half* bo = p.bo + (size_t)layer * C;
                for (int i = tid; i < rr.r1 - rr.r0; i += kThreads)
                    p.y1[rr.r0 + i] = __float2half_rn(unit_row_sum(red_units, i, C / ku) + __half2float(bo[rr.r0 + i]));
            }
            prof_stamp(p, pslot, prof_on);
            grid_barrier(p.bar, epoch);
            prof_stamp(p, pslot, prof_on);
            // ---------------- P4: x = LN1(x + y1) ; h1 = relu(fc1(x)) ----------------------------------------------
            {
                for (int i = tid; i < C; i += kThreads) {
                    float v = xres[i] + __half2float(__ushort_as_half(ldg_cg_u16(p.y1 + i)));
                    if (layer == 0) v = round_f16(v);     // fp16 + fp16 residual add on the first layer of a decode step
                    xres[i] = v;
                }
                __syncthreads();
                layer_norm_inplace(xres, xin, p.ln1_w + (size_t)layer * C, p.ln1_b + (size_t)layer * C, C, red);
                RowRange rr = cta_rows(F);
                if (tid == 0) {   // next weight phase: fc2 slice
                    RowRange nx = cta_rows(C);
                    prefetch_l2_range(w2 + (size_t)nx.r0 * F, (size_t)(nx.r1 - nx.r0) * F * 2);
                }
                gemv_units(w1, C, C, rr, xin, red_units);
                const __half* b1 = p.b1 + (size_t)layer * F;
                for (int i = tid; i < rr.r1 - rr.r0; i += kThreads) {
                    const float v = round_f16(red_units[i] + __half2float(b1[rr.r0 + i]));
                    p.h1[rr.r0 + i] = __float2half_rn(fmaxf(v, 0.f));
                }
            }
            prof_stamp(p, pslot, prof_on);
            grid_barrier(p.bar, epoch);
            prof_stamp(p, pslot, prof_on);
            // ---------------- P5: y2 = fc2(h1) -------------------------------------------------------------------------
            {
                if (tid == 0) {   // next layer's qkv slice (or lm_head after the last layer)
                    if (layer + 1 < p.layers) {
                        RowRange nx = cta_rows(3 * C);
                        prefetch_l2_range(wqkv + (size_t)3 * C * C + (size_t)nx.r0 * C, (size_t)(nx.r1 - nx.r0) * C * 2);
                    } else {
                        RowRange nx = cta_rows(V);
                        prefetch_l2_range(p.lm_head + (size_t)nx.r0 * C, (size_t)(nx.r1 - nx.r0) * C * 2);
                    }
                }
                for (int i = tid; i < F / 8; i += kThreads)
                    reinterpret_cast<uint4*>(xin)[i] = ldg_cg(reinterpret_cast<const uint4*>(p.h1) + i);
                __syncthreads();
                RowRange rr = cta_rows(C);
                const int ku = F / p.ks_fc2;
                gemv_units(w2, F, ku, rr, xin, red_units);
                const __half* b2 = p.b2 + (size_t)layer * C;
                for (int i = tid; i < rr.r1 - rr.r0; i += kThreads)
                    p.y2[rr.r0 + i] = __float2half_rn(unit_row_sum(red_units, i, F / ku) + __half2float(b2[rr.r0 + i]));
            }
            prof_stamp(p, pslot, prof_on);
            grid_barrier(p.bar, epoch);
            prof_stamp(p, pslot, prof_on);
            // ---------------- x = LN2(x + y2) ------------------------------------------------------------------------------
            for (int i = tid; i < C; i += kThreads)
                xres[i] += __half2float(__ushort_as_half(ldg_cg_u16(p.y2 + i)));
            __syncthreads();
            layer_norm_inplace(xres, xin, p.ln2_w + (size_t)layer * C, p.ln2_b + (size_t)layer * C, C, red);
        }
        // ================= lm_head: logits_pre = fp16(x) @ W^T (fp32 value before the fp16 store) ================
        {
            RowRange rr = cta_rows(V);
            if (tid == 0) {   // next token, layer 0 qkv slice
                RowRange nx = cta_rows(3 * C);
                prefetch_l2_range(p.wqkv + (size_t)nx.r0 * C, (size_t)(nx.r1 - nx.r0) * C * 2);
            }
            const int ku = C / p.ks_lm;
            gemv_units(p.lm_head, C, ku, rr, xin, red_units);
            for (int i = tid; i < rr.r1 - rr.r0; i += kThreads) p.logits[rr.r0 + i] = unit_row_sum(red_units, i, C / ku);
        }
        L += 1;
        prof_stamp(p, pslot, prof_on);
        grid_barrier(p.bar, epoch);
        prof_stamp(p, pslot, prof_on);
    }
    if (blockIdx.x == 0 && tid == 0) { p.st->t = t; p.st->L = L; p.st->counter = counter; p.st->last_tok = last_tok; }
}

}  // namespace er

// ---- host launcher ----------------------------------------------------------------------------------------------
size_t er_decode_smem_bytes(const er::DecodeParams& p, int sc_keys) {
    const int C = p.C, F = p.F;
    size_t fl = (size_t)C /*xres*/ + er::kMaxUnits + 32 + er::HD + (size_t)p.H * p.S + 42 * er::HD +
                (size_t)(sc_keys > p.V ? sc_keys : p.V);
    return fl * 4 + (size_t)(F > C ? F : C) * 2 + 64;
}

cudaError_t er_decode_launch(const er::DecodeParams& p, int grid, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(er::decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&p};
    return cudaLaunchCooperativeKernel((const void*)er::decode_persistent_kernel, dim3(grid), dim3(er::kThreads), args, smem, stream);
}

int er_decode_max_grid(size_t smem) {
    int dev = 0, sms = 0, per = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(er::decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, er::decode_persistent_kernel, er::kThreads, smem);
    return per > 0 ? sms : 0;
}
