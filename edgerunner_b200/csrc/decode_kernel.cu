// Persistent auto-regressive decode kernel (sm_100a), v1: TMA-fed shared-memory ring.
//
// Replaces, for every generated token, the whole per-step stack of the reference:
//   HF _sample step (logits -> constraint mask -> argmax | top-k sample)            third-party, restated
//   constraint FSM                         /root/reference/core/models.py:245-271
//   ShapeOPTDecoder.forward (N == 1)       /root/reference/core/transformer/modeling_opt.py:340-357
//   24 x OPTDecoderLayer.forward           modeling_opt.py:264-288   (post-LN, ReLU MLP)
//   OptFlashAttention2.forward             modeling_opt.py:185-232   (q/k/v GEMV, KV append, 1 x L attention)
//   lm_head                                modeling_opt.py:497
//
// One cooperative launch generates up to `steps` tokens: one CTA per SM (16 consumer warps + 1 producer warp).  All
// CTAs walk the same phase list and meet at a grid barrier between dependent phases (5 per layer).  The token loop, the
// FSM and the sampler live on the device, so there is no host round trip per token (the reference has >= 5).
//
// HBM-bound by construction (B = 1 GEMV + single-query attention, ~1 flop/byte).  Every weight byte and every cached
// K/V byte is read exactly once per token.  The producer warp streams this CTA's slice of every phase — weight rows,
// K blocks, V rows, in consumption order — from HBM into a shared-memory ring with TMA bulk copies
// (cp.async.bulk.shared.global + mbarrier complete_tx; SASS UBLKCP).  Weights and old K/V rows do not depend on the
// activations, so the stream runs AHEAD of the consumers across grid barriers: HBM stays busy while the consumers
// wait for each other, and a phase starts computing out of shared memory the moment its input vector arrives.
// The new K/V row is appended in place with plain stores (the reference re-allocates and copies the whole cache per
// layer per step, modeling_opt.py:191-192) and is the only key read straight from global memory.
//
// dtype ledger (mirrored by oracle/er_oracle.py mode='ledger'): fp16 weights and KV,
// fp32 accumulation everywhere, activations rounded to fp16 exactly where model.half()+autocast(fp16) does.
#include "decode_kernel.h"

#include <cstdlib>

#include "common.cuh"

namespace er {

constexpr int kConsumerWarps = 14;               // 7 owner pairs, one per ring stage; 15 warps total -> 128 registers/thread
constexpr int kConsumers = kConsumerWarps * 32;   // 512 compute threads
constexpr int kThreads = kConsumers + 32;         // + one producer warp
constexpr int HD = 96;                            // decoder head_dim (ArAE: 1536 / 16)
constexpr int HV = HD / 8;                        // 16-byte vectors per head row (12)
constexpr int kStageBytes = 24576;                // 8 weight rows of 1536 fp16 = 4 K blocks = 128 V rows
constexpr int kMaxStages = 7;                    // == kConsumerWarps / 2: stage s is processed by warp pair s
constexpr int kKBlockBytes = HV * 32 * 16;        // 6144: one 32-key block of the blocked K cache
constexpr int kMaxUnits = 256;

// ---- mbarrier / TMA bulk copy -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// (all shared-memory operands are 32-bit shared-space addresses so that the compiler emits LDS / SYNCS, never generic LD)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
    return r;
}

// ---- consumer-only block barrier (the producer warp never joins it) ---------------------------------------------------------
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }

// ---- grid barrier over the consumers of all CTAs -----------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch) {
    cbar();
    if (threadIdx.x == 0) {
        epoch += 1;
        const unsigned target = epoch * gridDim.x;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    cbar();
}

// ---- optional phase timeline (profiles/): one CTA, thread 0 stamps %globaltimer ------------------------------------------------
// slot map: [0] token start; per layer l, base = 1 + 16*l: +0 residual+LN2 done, +1 qkv gemv done, +2 P1 epilogue done, +3 B1,
// +4 attention done, +5 B2, +6 combine done, +7 out_proj done, +8 B3, +9 LN1 done, +10 fc1 done, +11 B4, +12 h1 loaded,
// +13 fc2 done, +14 B5, +15 unused; after the layers: +0 lm_head done, +1 final barrier.
__device__ volatile uint32_t* g_prod_it_ptr;
__device__ __forceinline__ void prof_stamp(const DecodeParams& p, int slot, bool on, uint32_t cons_it = 0, volatile uint32_t* prod_it = nullptr) {
    if (on && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.prof[slot] = t;
        if (prod_it) p.prof[2048 + slot] = (unsigned long long)(*prod_it - cons_it);
    }
}

// ---- block reductions over the 512 consumers -----------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    cbar();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    cbar();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; i++) t += red[i];
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    cbar();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    cbar();
    float t = -INFINITY;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; i++) t = fmaxf(t, red[i]);
    return t;
}

// x = LayerNorm(xres + y) in place (fp32 statistics, eps 1e-5, fp16 affine params); also emits the fp16 copy used as the next
// GEMV input.  y is the fp16 phase output other CTAs just published (read at L2); round_first: the residual add of the first
// layer of a decode step is an fp16 + fp16 add.  One block reduction (sum and sum of squares), 2 barriers.
struct LnParams { float g[4], b[4]; };                      // this thread's affine parameters (elements tid + j*512)
__device__ __forceinline__ LnParams ln_load(const __half* __restrict__ g, const __half* __restrict__ b, int C) {
    LnParams lp;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int i = threadIdx.x + j * kConsumers;
        lp.g[j] = (i < C) ? __half2float(g[i]) : 0.f;
        lp.b[j] = (i < C) ? __half2float(b[i]) : 0.f;
    }
    return lp;
}
__device__ __noinline__ void residual_layer_norm(float* xres, __half* x16, const __half* y, bool round_first, const LnParams lp, int C, float* red) {
    constexpr int PER = 4;                                  // supports C <= 4 * 512
    float v[PER];
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const int i = threadIdx.x + j * kConsumers;
        v[j] = 0.f;
        if (i < C) {
            float t = xres[i] + __half2float(__ushort_as_half(ldg_cg_u16(y + i)));
            if (round_first) t = round_f16(t);
            v[j] = t; s += t; q += t * t;
        }
    }
    s = warp_sum(s); q = warp_sum(q);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = s; red[16 + (threadIdx.x >> 5)] = q; }   // kConsumerWarps <= 16
    cbar();
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; i++) { ts += red[i]; tq += red[16 + i]; }
    const float mean = ts / C;
    const float var = fmaxf(tq / C - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const int i = threadIdx.x + j * kConsumers;
        if (i < C) {
            const float yv = (v[j] - mean) * rstd * lp.g[j] + lp.b[j];
            xres[i] = yv;
            x16[i] = __float2half_rn(yv);
        }
    }
    cbar();
}

// ---- work partition (identical on the producer and the consumer side) ------------------------------------------------------------
struct RowRange { int r0, r1; };
__device__ __forceinline__ RowRange cta_rows(int R) {
    RowRange rr;
    rr.r0 = (int)(((long long)R * blockIdx.x) / gridDim.x);
    rr.r1 = (int)(((long long)R * (blockIdx.x + 1)) / gridDim.x);
    return rr;
}
struct AttnRange { int h, b0, b1, k0, k1, is_new; };   // old keys [k0,k1) in K blocks [b0,b1); is_new: this CTA also owns key L
__device__ __forceinline__ bool attn_range(const DecodeParams& p, int L, AttnRange& a) {
    if ((int)blockIdx.x >= p.H * p.S) return false;
    a.h = blockIdx.x / p.S;
    const int s = blockIdx.x % p.S;
    const int nblk = (L + 31) >> 5;                  // blocks holding old keys 0..L-1
    const int bps = (nblk + p.S - 1) / p.S;
    a.b0 = min(s * bps, nblk);
    a.b1 = min(a.b0 + bps, nblk);
    a.k0 = a.b0 * 32;
    a.k1 = max(a.k0, min(a.b1 * 32, L));   // empty splits: b0 == b1 == nblk, k0 may exceed L
    a.is_new = (s == p.S - 1);
    return true;
}

struct Ring {          // passed by value (registers): shared-space addresses of the stage data and the mbarrier arrays
    uint32_t data, full, empty;
    int nstage;
    __device__ __forceinline__ uint32_t stage(uint32_t s) const { return data + s * kStageBytes; }
    __device__ __forceinline__ uint32_t fullb(uint32_t s) const { return full + s * 8; }
    __device__ __forceinline__ uint32_t emptyb(uint32_t s) const { return empty + s * 8; }
};

// Consumer-side ring cursor.  Every consumer warp advances it identically over every stage of every job, but a stage is
// PROCESSED by one warp pair only: ring stage s belongs to warp pair s (nstage <= kConsumerWarps / 2).  Different pairs therefore work on different
// stages at the same time (up to `nstage` of them), which overlaps the per-stage dependency chains (barrier wait, smem
// latency, shuffle reduction) instead of running all 16 warps in lockstep over one stage.
struct Cursor {
    uint32_t it, stage, parity;
    __device__ __forceinline__ void advance(int nstage) {
        ++it;
        if (++stage == (uint32_t)nstage) { stage = 0; parity ^= 1; }
    }
    __device__ __forceinline__ bool mine(int warp) const { return stage == (uint32_t)(warp >> 1); }
};
constexpr int kOwnersPerStage = 2;   // warps arriving on a stage's empty barrier

// ---- producer: stream one contiguous byte range through the ring ---------------------------------------------------------------------
// returns false if the consumers raised `stop` (EOS) while we were waiting for a free stage
__device__ __forceinline__ bool produce(const Ring r, uint32_t& it, const void* base, size_t bytes, volatile int* stop, volatile uint32_t* prod_it) {
    const char* src = reinterpret_cast<const char*>(base);
    for (size_t off = 0; off < bytes; off += kStageBytes, ++it) {
        const uint32_t s = it % r.nstage, par = (it / r.nstage) & 1;
        while (!mbar_try_wait(r.emptyb(s), par ^ 1)) {
            if (*stop) return false;
        }
        if (*stop) return false;
        const uint32_t n = (uint32_t)(bytes - off < (size_t)kStageBytes ? bytes - off : (size_t)kStageBytes);
        mbar_arrive_expect_tx(r.fullb(s), n);
        bulk_g2s(r.stage(s), src + off, n, r.fullb(s));
        *prod_it = it + 1;
    }
    return true;
}

__device__ __noinline__ void producer_loop(const DecodeParams& p, const Ring r, volatile int* stop, volatile uint32_t* cons_it, volatile uint32_t* prod_it) {
    const int C = p.C, F = p.F, H = p.H;
    int t = p.st->t, L = p.st->L;
    if (p.st->done) return;
    // forward passes of this launch: token tt is followed by a pass iff tt + 1 < max_new (EOS is handled through `stop`)
    const int n_fwd = min(p.steps, p.max_new - 1 - t);
    uint32_t it = 0;
    bool ok = true;
    for (int pass = 0; pass < n_fwd && ok; ++pass, ++L) {
        for (int layer = 0; layer < p.layers && ok; ++layer) {
            RowRange rr = cta_rows(3 * C);
            ok = produce(r, it, p.wqkv + ((size_t)layer * 3 * C + rr.r0) * C, (size_t)(rr.r1 - rr.r0) * C * 2, stop, prod_it);
            AttnRange a;
            if (ok && attn_range(p, L, a)) {
                const __half* kbase = p.kc + (((size_t)layer * H + a.h) * p.nkb + a.b0) * (size_t)(HV * 256);
                ok = produce(r, it, kbase, (size_t)(a.b1 - a.b0) * kKBlockBytes, stop, prod_it);
                const __half* vbase = p.vc + (((size_t)layer * H + a.h) * p.Lmax + a.k0) * HD;
                if (ok) ok = produce(r, it, vbase, (size_t)(a.k1 - a.k0) * HD * 2, stop, prod_it);
            }
            rr = cta_rows(C);
            if (ok) ok = produce(r, it, p.wo + ((size_t)layer * C + rr.r0) * C, (size_t)(rr.r1 - rr.r0) * C * 2, stop, prod_it);
            RowRange rf = cta_rows(F);
            if (ok) ok = produce(r, it, p.w1 + ((size_t)layer * F + rf.r0) * C, (size_t)(rf.r1 - rf.r0) * C * 2, stop, prod_it);
            if (ok) ok = produce(r, it, p.w2 + ((size_t)layer * C + rr.r0) * F, (size_t)(rr.r1 - rr.r0) * F * 2, stop, prod_it);
        }
        RowRange rv = cta_rows(p.V);
        if (ok) ok = produce(r, it, p.lm_head + (size_t)rv.r0 * C, (size_t)(rv.r1 - rv.r0) * C * 2, stop, prod_it);
    }
    if (!ok) {
        // EOS: the consumers stopped at stage *cons_it; every copy we issued beyond it must land before the CTA may exit
        for (uint32_t j = *cons_it; j < it; ++j) mbar_wait(r.fullb(j % r.nstage), (j / r.nstage) & 1);
    }
}

// ---- consumer: GEMV over one streamed weight slice ------------------------------------------------------------------------------------
// The slice is n_units units of C fp16 (a row of K = nu_row * C elements is nu_row consecutive units).  The stage's owner
// pair splits every unit in two halves (warp parity = half); each warp walks the units of the stage two at a time (independent
// accumulator chains).  XREG: nu_row == 1, the x half this warp needs is the same for every unit -> registers.
// NVL = 16-byte vectors per lane per half-unit (3 for C = 1536).  Partial sums go to red_units[unit*2 + half].
template <int NVL>
__device__ __forceinline__ float half_unit_dot(uint32_t wb, const float (*xr)[8], uint32_t xb, bool xreg, int lane, int hv) {
    uint4 wv[NVL], xv[NVL];
#pragma unroll
    for (int j = 0; j < NVL; j++) {
        const int v = lane + 32 * j;
        wv[j] = (v < hv) ? lds128(wb + v * 16) : make_uint4(0, 0, 0, 0);
        if (!xreg) xv[j] = (v < hv) ? lds128(xb + v * 16) : make_uint4(0, 0, 0, 0);
    }
    float tot = 0.f;
#pragma unroll
    for (int j = 0; j < NVL; j++) {
        float x[8];
        if (xreg) {
#pragma unroll
            for (int e = 0; e < 8; e++) x[e] = xr[j][e];
        } else {
            float2 f;
            f = h2f2(xv[j].x); x[0] = f.x; x[1] = f.y;
            f = h2f2(xv[j].y); x[2] = f.x; x[3] = f.y;
            f = h2f2(xv[j].z); x[4] = f.x; x[5] = f.y;
            f = h2f2(xv[j].w); x[6] = f.x; x[7] = f.y;
        }
        float2 f; float a0, a1;
        f = h2f2(wv[j].x); a0 = f.x * x[0];           a1 = f.y * x[1];
        f = h2f2(wv[j].y); a0 = fmaf(f.x, x[2], a0);  a1 = fmaf(f.y, x[3], a1);
        f = h2f2(wv[j].z); a0 = fmaf(f.x, x[4], a0);  a1 = fmaf(f.y, x[5], a1);
        f = h2f2(wv[j].w); a0 = fmaf(f.x, x[6], a0);  a1 = fmaf(f.y, x[7], a1);
        tot += a0 + a1;
    }
    return tot;
}
template <int NVL>
__device__ __forceinline__ void gemv_job_t(const Ring r, Cursor& cur, int n_units, int nu_row, int C, uint32_t xin_s, float* red_units) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, half = warp & 1;
    const int upstage = kStageBytes / (2 * C);
    const int hv = C >> 4;                                   // 16-byte vectors per half-unit
    const int nch = (n_units + upstage - 1) / upstage;
    const bool xreg = (nu_row == 1);
    float xr[NVL][8];
    if (xreg) {
        const uint32_t xb = xin_s + (uint32_t)half * C;      // bytes: half * (C/2 elements) * 2
#pragma unroll
        for (int j = 0; j < NVL; j++) {
            const int v = lane + 32 * j;
            const uint4 xv = (v < hv) ? lds128(xb + v * 16) : make_uint4(0, 0, 0, 0);
            float2 f;
            f = h2f2(xv.x); xr[j][0] = f.x; xr[j][1] = f.y;
            f = h2f2(xv.y); xr[j][2] = f.x; xr[j][3] = f.y;
            f = h2f2(xv.z); xr[j][4] = f.x; xr[j][5] = f.y;
            f = h2f2(xv.w); xr[j][6] = f.x; xr[j][7] = f.y;
        }
    }
    for (int c = 0; c < nch; ++c, cur.advance(r.nstage)) {
        if (!cur.mine(warp)) continue;
        mbar_wait(r.fullb(cur.stage), cur.parity);
        const int here = min(upstage, n_units - c * upstage);
        const uint32_t st = r.stage(cur.stage) + (uint32_t)half * C;
        for (int u = 0; u < here; u += 2) {
            const int gu = c * upstage + u;
            const bool two = (u + 1 < here);
            const uint32_t xb0 = xin_s + (uint32_t)((gu % nu_row) * C) * 2 + (uint32_t)half * C;
            const uint32_t xb1 = xin_s + (uint32_t)(((gu + 1) % nu_row) * C) * 2 + (uint32_t)half * C;
            float t0 = half_unit_dot<NVL>(st + (uint32_t)u * 2 * C, xr, xb0, xreg, lane, hv);
            float t1 = two ? half_unit_dot<NVL>(st + (uint32_t)(u + 1) * 2 * C, xr, xb1, xreg, lane, hv) : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                t0 += __shfl_xor_sync(0xffffffffu, t0, o);
                t1 += __shfl_xor_sync(0xffffffffu, t1, o);
            }
            if (lane == 0) {
                red_units[gu * 2 + half] = t0;
                if (two) red_units[(gu + 1) * 2 + half] = t1;
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(r.emptyb(cur.stage));
    }
    cbar();
}
__device__ __noinline__ void gemv_job(const Ring r, Cursor& cur, int n_units, int nu_row, int C, uint32_t xin_s, float* red_units) {
    const int nvl = ((C >> 4) + 31) >> 5;
    if (nvl == 3) gemv_job_t<3>(r, cur, n_units, nu_row, C, xin_s, red_units);
    else if (nvl == 2) gemv_job_t<2>(r, cur, n_units, nu_row, C, xin_s, red_units);
    else gemv_job_t<1>(r, cur, n_units, nu_row, C, xin_s, red_units);
}
__device__ __forceinline__ float row_sum(const float* red_units, int local_row, int nu_row) {
    float s = 0.f;
    for (int j = 0; j < 2 * nu_row; j++) s += red_units[local_row * 2 * nu_row + j];
    return s;
}

// ---- sampler (HF _sample step + constraint FSM), executed redundantly by warp 0 of every CTA --------------------------------
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
    // scores = float(fp16(logit)) + mask   (core/utils.py:143-158)
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

// ---- P2: single-query attention over the cached keys, one (head, KV split) per CTA, K/V streamed through the ring -------------
// K pass: lane = key inside a 32-key block (12 conflict-free 16-byte smem reads, one complete q.k per lane, no cross-lane
// reduction); V pass: warp = 2 rows x 12 vectors.  The new key L (appended by P1 of this step) is read from global by the
// CTA of the last split.  Each split publishes (o[96], max, sum); the LAST split of a head to finish (atomic ticket) merges
// the S partials of that head and publishes attn16[h*96 ..] — so the out_proj phase of every CTA starts from one 3 KB vector.
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ float dot_k8(const uint4 kv, const float4 qa, const float4 qb, float acc) {
    float2 f;
    f = h2f2(kv.x); acc = fmaf(f.x, qa.x, acc); acc = fmaf(f.y, qa.y, acc);
    f = h2f2(kv.y); acc = fmaf(f.x, qa.z, acc); acc = fmaf(f.y, qa.w, acc);
    f = h2f2(kv.z); acc = fmaf(f.x, qb.x, acc); acc = fmaf(f.y, qb.y, acc);
    f = h2f2(kv.w); acc = fmaf(f.x, qb.z, acc); acc = fmaf(f.y, qb.w, acc);
    return acc;
}
__device__ __noinline__ void attention_phase(const DecodeParams& p, const Ring r, Cursor& cur, int layer, int L, float* qs,
                                             float* sc, float* vred, float* red, int* s_flag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    AttnRange a;
    if (!attn_range(p, L, a)) return;
    const int H = p.H;
    const float sm_scale_log2 = rsqrtf((float)HD) * 1.4426950408889634f;
    float* outp = p.part + ((size_t)a.h * p.S + (blockIdx.x % p.S)) * 100;
    const int nold = a.k1 - a.k0;
    const int nk = nold + a.is_new;
    const int new_slot = (a.b1 - a.b0) * 32;               // score slot just past the old blocks
    float m = -INFINITY, l = 0.f;
    const uint32_t qs_s = s_addr(qs), sc_s = s_addr(sc);
    if (nk > 0) {
        if (tid < HD) qs[tid] = __half2float(__ushort_as_half(ldg_cg_u16(p.q16 + a.h * HD + tid)));
        cbar();
        // ---- K pass ----
        float lmax = -INFINITY;
        {
            const int nb = a.b1 - a.b0;
            const int nch = (nb + 3) >> 2;
            for (int c = 0; c < nch; ++c, cur.advance(r.nstage)) {
                if (!cur.mine(warp)) continue;
                mbar_wait(r.fullb(cur.stage), cur.parity);
                const int here = min(4, nb - c * 4);
                for (int bb = (warp & 1); bb < here; bb += 2) {            // the owner pair splits the stage's blocks
                    const int blk = c * 4 + bb;
                    const uint32_t kb = r.stage(cur.stage) + (uint32_t)bb * kKBlockBytes + lane * 16;
                    uint4 kv[HV];
#pragma unroll
                    for (int j = 0; j < HV; j++) kv[j] = lds128(kb + j * 512);
                    float a0 = 0.f, a1 = 0.f;
#pragma unroll
                    for (int j = 0; j < HV; j += 2) {
                        a0 = dot_k8(kv[j], lds_f4(qs_s + j * 32), lds_f4(qs_s + j * 32 + 16), a0);
                        a1 = dot_k8(kv[j + 1], lds_f4(qs_s + j * 32 + 32), lds_f4(qs_s + j * 32 + 48), a1);
                    }
                    const int key = (a.b0 + blk) * 32 + lane;
                    const float sv = (key < a.k1) ? a0 + a1 : -INFINITY;      // slots >= L hold stale / unwritten data
                    sc[blk * 32 + lane] = sv;
                    lmax = fmaxf(lmax, sv);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(r.emptyb(cur.stage));
            }
        }
        // ---- the new key: q . k_L straight from the cache line P1 just wrote (visible after the grid barrier) ----
        if (a.is_new && warp == 0) {
            float acc = 0.f;
            if (lane < HV) {
                const __half* kp = p.kc + ((((size_t)layer * H + a.h) * p.nkb + (L >> 5)) * HV + lane) * 256 + (size_t)(L & 31) * 8;
                acc = dot_k8(ldg_cg(reinterpret_cast<const uint4*>(kp)), lds_f4(qs_s + lane * 32), lds_f4(qs_s + lane * 32 + 16), 0.f);
            }
            acc = warp_sum(acc);
            if (lane == 0) sc[new_slot] = acc;
            lmax = fmaxf(lmax, acc);
        }
        m = block_max(lmax, red);                              // raw dot units (unscaled); also orders the sc[] writes
        float lsum = 0.f;
        const int nsc = new_slot + a.is_new;                   // score slots (old blocks, then the new key)
        for (int i = tid; i < nsc; i += kConsumers) {
            const float pr = exp2f((sc[i] - m) * sm_scale_log2);     // masked slots: exp2(-inf) = 0
            sc[i] = pr;
            lsum += pr;
        }
        l = block_sum(lsum, red);
        // ---- V pass ----
        float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int sub = lane / HV, vec = lane % HV;            // lanes 0..23: 2 rows x 12 vectors; lanes 24..31 idle
        {
            const int nch = (nold + 127) >> 7;
            for (int c = 0; c < nch; ++c, cur.advance(r.nstage)) {
                if (!cur.mine(warp)) continue;
                mbar_wait(r.fullb(cur.stage), cur.parity);
                const int here = min(128, nold - c * 128);
                if (lane < 2 * HV) {
                    const uint32_t st = r.stage(cur.stage) + vec * 16;
                    const uint32_t pc = sc_s + c * 512;           // old key k0 + c*128 + row sits at slot c*128 + row (k0 is block aligned)
                    int row = (warp & 1) * 2 + sub;                // the owner pair interleaves row pairs: rows 4i + 2*(warp&1) + sub
                    for (; row + 4 < here; row += 8) {             // two rows in flight
                        const uint4 v0 = lds128(st + (uint32_t)row * (HD * 2));
                        const uint4 v1 = lds128(st + (uint32_t)(row + 4) * (HD * 2));
                        const float p0 = lds32(pc + row * 4), p1 = lds32(pc + row * 4 + 16);
                        float2 f;
                        f = h2f2(v0.x); o[0] = fmaf(p0, f.x, o[0]); o[1] = fmaf(p0, f.y, o[1]);
                        f = h2f2(v0.y); o[2] = fmaf(p0, f.x, o[2]); o[3] = fmaf(p0, f.y, o[3]);
                        f = h2f2(v0.z); o[4] = fmaf(p0, f.x, o[4]); o[5] = fmaf(p0, f.y, o[5]);
                        f = h2f2(v0.w); o[6] = fmaf(p0, f.x, o[6]); o[7] = fmaf(p0, f.y, o[7]);
                        f = h2f2(v1.x); o[0] = fmaf(p1, f.x, o[0]); o[1] = fmaf(p1, f.y, o[1]);
                        f = h2f2(v1.y); o[2] = fmaf(p1, f.x, o[2]); o[3] = fmaf(p1, f.y, o[3]);
                        f = h2f2(v1.z); o[4] = fmaf(p1, f.x, o[4]); o[5] = fmaf(p1, f.y, o[5]);
                        f = h2f2(v1.w); o[6] = fmaf(p1, f.x, o[6]); o[7] = fmaf(p1, f.y, o[7]);
                    }
                    for (; row < here; row += 4) {
                        const uint4 v0 = lds128(st + (uint32_t)row * (HD * 2));
                        const float p0 = lds32(pc + row * 4);
                        float2 f;
                        f = h2f2(v0.x); o[0] = fmaf(p0, f.x, o[0]); o[1] = fmaf(p0, f.y, o[1]);
                        f = h2f2(v0.y); o[2] = fmaf(p0, f.x, o[2]); o[3] = fmaf(p0, f.y, o[3]);
                        f = h2f2(v0.z); o[4] = fmaf(p0, f.x, o[4]); o[5] = fmaf(p0, f.y, o[5]);
                        f = h2f2(v0.w); o[6] = fmaf(p0, f.x, o[6]); o[7] = fmaf(p0, f.y, o[7]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(r.emptyb(cur.stage));
            }
        }
        if (lane < 2 * HV) {
#pragma unroll
            for (int e = 0; e < 8; e++) vred[(warp * 2 + sub) * HD + vec * 8 + e] = o[e];
        }
        cbar();
    }
    // ---- publish the split partial ----
    if (tid < HD) {
        float acc = 0.f;
        if (nk > 0) {
            for (int g = 0; g < 2 * kConsumerWarps; g++) acc += vred[g * HD + tid];
            if (a.is_new) {
                const float vn = __half2float(__ushort_as_half(ldg_cg_u16(p.vc + (((size_t)layer * H + a.h) * p.Lmax + L) * HD + tid)));
                acc = fmaf(sc[new_slot], vn, acc);
            }
        }
        outp[tid] = acc;
    } else if (tid == HD) {
        outp[HD] = (nk > 0) ? m * rsqrtf((float)HD) : -INFINITY;      // max in softmax (scaled) units
        outp[HD + 1] = l;
    }
    // ---- last split of this head to finish merges the S partials (release/acquire ticket on a monotonic counter) ----
    cbar();
    if (tid == 0) {
        unsigned ticket;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(p.head_cnt + a.h) : "memory");
        *s_flag = ((ticket % (unsigned)p.S) == (unsigned)p.S - 1) ? 1 : 0;
    }
    cbar();
    if (*s_flag && tid < HD) {
        const float* pp = p.part + (size_t)a.h * p.S * 100;
        float ms[16], ls[16], ov[16];
#pragma unroll
        for (int s = 0; s < 16; s++) {
            ms[s] = -INFINITY; ls[s] = 0.f; ov[s] = 0.f;
            if (s < p.S) {
                ms[s] = ldg_cg_f32(pp + (size_t)s * 100 + HD);
                ls[s] = ldg_cg_f32(pp + (size_t)s * 100 + HD + 1);
                ov[s] = ldg_cg_f32(pp + (size_t)s * 100 + tid);
            }
        }
        float M = -INFINITY;
#pragma unroll
        for (int s = 0; s < 16; s++) M = fmaxf(M, ms[s]);
        float den = 0.f, num = 0.f;
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const float w = (ms[s] == -INFINITY) ? 0.f : __expf(ms[s] - M);
            den = fmaf(w, ls[s], den);
            num = fmaf(w, ov[s], num);
        }
        p.attn16[a.h * HD + tid] = __float2half_rn(num / den);
    }
}

// ---- the kernel ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) decode_persistent_kernel(const DecodeParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = p.C, F = p.F, H = p.H, V = p.V;
    // smem carve-up: ring first (128-byte aligned stages), then the small arrays
    Ring ring;
    ring.data = s_addr(smem_raw);
    ring.nstage = p.nstage;
    unsigned char* q = smem_raw + (size_t)p.nstage * kStageBytes;
    ring.full = s_addr(q);
    ring.empty = ring.full + 8 * 8;
    float* xres = reinterpret_cast<float*>(q + 2 * 8 * 8);   // [C]   residual stream (fp32)
    float* red_units = xres + C;                                      // [kMaxUnits]
    float* red = red_units + kMaxUnits;                               // [32]
    float* qs = red + 32;                                             // [96] query of this CTA's head (fp32)
    float* vred = qs + HD;                                            // [32*96] V-pass cross-warp reduction
    float* sc = vred + 2 * kConsumerWarps * HD;                       // [sc_len] scores / sampler scratch
    __half* xin = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(sc + p.sc_len) + 15) & ~uintptr_t(15));   // [max(C,F)] GEMV input (fp16)
    __shared__ int s_tok;
    __shared__ int s_stop;
    __shared__ uint32_t s_cons_it;
    __shared__ int s_flag;
    __shared__ uint32_t s_prod_it;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < p.nstage; i++) { mbar_init(ring.fullb(i), 1); mbar_init(ring.emptyb(i), kOwnersPerStage); }
        s_stop = 0;
        s_cons_it = 0;
        s_prod_it = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ===== producer warp: one elected lane streams this CTA's byte ranges through the ring =====
        if (lane == 0) producer_loop(p, ring, &s_stop, &s_cons_it, &s_prod_it);
    } else {
        // ===== consumers =====
        unsigned epoch = 0;
        Cursor cur{0u, 0u, 0u};
        const uint32_t xin_s = s_addr(xin);
        int t = p.st->t, L = p.st->L, counter = p.st->counter, last_tok = p.st->last_tok;
        const bool done0 = p.st->done != 0;
        bool state_written = done0;
        const size_t nkb = (size_t)p.nkb;
        for (int iter = 0; iter < p.steps && !done0; ++iter, ++t) {
            const bool prof_on = p.prof != nullptr && t == p.prof_token && (int)blockIdx.x == p.prof_cta;
            prof_stamp(p, 0, prof_on);
            // ================= sample token t from the current logits =============================================================
            if (p.use_fsm && t > 0) fsm_update(counter, last_tok);
            if (warp == 0) {
                const int tok = sample_warp(p.logits, sc, V, t, counter, p);
                if (lane == 0) s_tok = tok;
            }
            cbar();
            const int chosen = s_tok;
            const int fed = p.forced ? p.forced[t] : chosen;
            if (blockIdx.x == 0) {
                if (tid == 0) p.out_ids[t] = chosen;
                if (p.out_logits)
                    for (int i = tid; i < V; i += kConsumers) p.out_logits[(size_t)t * V + i] = ldg_cg_f32(p.logits + i);
            }
            last_tok = fed;
            if ((fed == p.eos) || (t + 1 >= p.max_new)) {
                if (blockIdx.x == 0 && tid == 0) { p.st->done = 1; p.st->t = t + 1; p.st->L = L; p.st->counter = counter; p.st->last_tok = last_tok; }
                if (tid == 0) { s_cons_it = cur.it; __threadfence_block(); s_stop = 1; }
                state_written = true;
                break;
            }

            // ================= embed: x = fp16(embd[tok] + pos[L]) =================================================================
            for (int i = tid; i < C; i += kConsumers) {
                float v = __half2float(p.embd[(size_t)fed * C + i]) + __half2float(p.pos[(size_t)L * C + i]);
                __half h = __float2half_rn(v);
                xin[i] = h;
                xres[i] = __half2float(h);
            }
            cbar();

            for (int layer = 0; layer < p.layers; ++layer) {
                const int pb = 1 + 16 * layer;
                // ---------------- P1: q,k,v = x16 @ Wqkv^T + b ; KV append in place ----------------------------------------------
                {
                    RowRange rr = cta_rows(3 * C);
                    prof_stamp(p, pb + 0, prof_on, cur.it, &s_prod_it);
                    const float bias = (tid < rr.r1 - rr.r0) ? __half2float(p.bqkv[(size_t)layer * 3 * C + rr.r0 + tid]) : 0.f;   // in flight during the GEMV
                    gemv_job(ring, cur, rr.r1 - rr.r0, 1, C, xin_s, red_units);
                    prof_stamp(p, pb + 1, prof_on, cur.it, &s_prod_it);
                    for (int i = tid; i < rr.r1 - rr.r0; i += kConsumers) {
                        const int r = rr.r0 + i;
                        const __half hv = __float2half_rn(row_sum(red_units, i, 1) + bias);
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
                prof_stamp(p, pb + 2, prof_on, cur.it, &s_prod_it);
                grid_barrier(p.bar, epoch);
                prof_stamp(p, pb + 3, prof_on, cur.it, &s_prod_it);
                // ---------------- P2: attention -----------------------------------------------------------------------------------------
                attention_phase(p, ring, cur, layer, L, qs, sc, vred, red, &s_flag);
                prof_stamp(p, pb + 4, prof_on, cur.it, &s_prod_it);
                grid_barrier(p.bar, epoch);
                prof_stamp(p, pb + 5, prof_on, cur.it, &s_prod_it);
                // ---------------- P3: combine splits -> attn16 ; out_proj ---------------------------------------------------------
                {
                    for (int i = tid; i < C / 8; i += kConsumers)
                        reinterpret_cast<uint4*>(xin)[i] = ldg_cg(reinterpret_cast<const uint4*>(p.attn16) + i);
                    cbar();
                    RowRange rr = cta_rows(C);
                    prof_stamp(p, pb + 6, prof_on, cur.it, &s_prod_it);
                    const float bias = (tid < rr.r1 - rr.r0) ? __half2float(p.bo[(size_t)layer * C + rr.r0 + tid]) : 0.f;
                    gemv_job(ring, cur, rr.r1 - rr.r0, 1, C, xin_s, red_units);
                    for (int i = tid; i < rr.r1 - rr.r0; i += kConsumers)
                        p.y1[rr.r0 + i] = __float2half_rn(row_sum(red_units, i, 1) + bias);
                }
                const LnParams lp1 = ln_load(p.ln1_w + (size_t)layer * C, p.ln1_b + (size_t)layer * C, C);   // lands while we wait at the barrier
                prof_stamp(p, pb + 7, prof_on, cur.it, &s_prod_it);
                grid_barrier(p.bar, epoch);
                prof_stamp(p, pb + 8, prof_on, cur.it, &s_prod_it);
                // ---------------- P4: x = LN1(x + y1) ; h1 = relu(fc1(x)) -----------------------------------------------------------
                {
                    residual_layer_norm(xres, xin, p.y1, layer == 0, lp1, C, red);
                    RowRange rr = cta_rows(F);
                    prof_stamp(p, pb + 9, prof_on, cur.it, &s_prod_it);
                    const float bias = (tid < rr.r1 - rr.r0) ? __half2float(p.b1[(size_t)layer * F + rr.r0 + tid]) : 0.f;
                    gemv_job(ring, cur, rr.r1 - rr.r0, 1, C, xin_s, red_units);
                    for (int i = tid; i < rr.r1 - rr.r0; i += kConsumers) {
                        const float v = round_f16(row_sum(red_units, i, 1) + bias);
                        p.h1[rr.r0 + i] = __float2half_rn(fmaxf(v, 0.f));
                    }
                }
                prof_stamp(p, pb + 10, prof_on, cur.it, &s_prod_it);
                grid_barrier(p.bar, epoch);
                prof_stamp(p, pb + 11, prof_on, cur.it, &s_prod_it);
                // ---------------- P5: y2 = fc2(h1) -----------------------------------------------------------------------------------------
                {
                    for (int i = tid; i < F / 8; i += kConsumers)
                        reinterpret_cast<uint4*>(xin)[i] = ldg_cg(reinterpret_cast<const uint4*>(p.h1) + i);
                    cbar();
                    RowRange rr = cta_rows(C);
                    const int nu_row = F / C;
                    prof_stamp(p, pb + 12, prof_on, cur.it, &s_prod_it);
                    const float bias = (tid < rr.r1 - rr.r0) ? __half2float(p.b2[(size_t)layer * C + rr.r0 + tid]) : 0.f;
                    gemv_job(ring, cur, (rr.r1 - rr.r0) * nu_row, nu_row, C, xin_s, red_units);
                    for (int i = tid; i < rr.r1 - rr.r0; i += kConsumers)
                        p.y2[rr.r0 + i] = __float2half_rn(row_sum(red_units, i, nu_row) + bias);
                }
                const LnParams lp2 = ln_load(p.ln2_w + (size_t)layer * C, p.ln2_b + (size_t)layer * C, C);
                prof_stamp(p, pb + 13, prof_on, cur.it, &s_prod_it);
                grid_barrier(p.bar, epoch);
                prof_stamp(p, pb + 14, prof_on, cur.it, &s_prod_it);
                // ---------------- x = LN2(x + y2) ------------------------------------------------------------------------------------------
                residual_layer_norm(xres, xin, p.y2, false, lp2, C, red);
            }
            // ================= lm_head: logits_pre = fp16(x) @ W^T (fp32 value before the fp16 store) ================
            {
                RowRange rr = cta_rows(V);
                gemv_job(ring, cur, rr.r1 - rr.r0, 1, C, xin_s, red_units);
                for (int i = tid; i < rr.r1 - rr.r0; i += kConsumers) p.logits[rr.r0 + i] = row_sum(red_units, i, 1);
            }
            L += 1;
            prof_stamp(p, 1 + 16 * p.layers, prof_on);
            grid_barrier(p.bar, epoch);
            prof_stamp(p, 2 + 16 * p.layers, prof_on);
        }
        if (!state_written && blockIdx.x == 0 && tid == 0) { p.st->t = t; p.st->L = L; p.st->counter = counter; p.st->last_tok = last_tok; }
    }
    __syncthreads();   // nobody leaves while bulk copies into this CTA's shared memory may still be in flight
}

}  // namespace er

// ---- host launcher -------------------------------------------------------------------------------------------------------------------------
size_t er_decode_small_smem_bytes(const er::DecodeParams& p) {
    const int C = p.C, F = p.F;
    size_t fl = (size_t)C + er::kMaxUnits + 32 + er::HD + 2 * er::kConsumerWarps * er::HD + (size_t)p.sc_len;
    return 2 * 8 * 8 + fl * 4 + (size_t)(F > C ? F : C) * 2 + 128;
}
size_t er_decode_smem_bytes(const er::DecodeParams& p) { return (size_t)p.nstage * er::kStageBytes + er_decode_small_smem_bytes(p); }
int er_decode_pick_stages(const er::DecodeParams& p, size_t smem_limit) {
    const size_t small = er_decode_small_smem_bytes(p);
    if (smem_limit < small + 2 * (size_t)er::kStageBytes) return 0;
    int n = (int)((smem_limit - small) / er::kStageBytes);
    return n > er::kMaxStages ? er::kMaxStages : n;
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
    if (cudaFuncSetAttribute(er::decode_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, er::decode_persistent_kernel, er::kThreads, smem);
    return per > 0 ? sms : 0;
}
