// Persistent auto-regressive decode kernel (sm_100a): TMA-fed shared-memory ring, latency-minimised consumers.
//
// Replaces, for every generated token, the whole per-step stack of the reference:
//   HF _sample step (logits -> constraint mask -> argmax | top-k sample)            third-party, restated
//   constraint FSM                         /root/reference/core/models.py:245-271
//   ShapeOPTDecoder.forward (N == 1)       /root/reference/core/transformer/modeling_opt.py:340-357
//   24 x OPTDecoderLayer.forward           modeling_opt.py:264-288   (post-LN, ReLU MLP)
//   OptFlashAttention2.forward             modeling_opt.py:185-232   (q/k/v GEMV, KV append, 1 x L attention)
//   lm_head                                modeling_opt.py:497
//
// One cooperative launch generates up to `steps` tokens: one CTA per SM (8 consumer warps + producer warp + L2 run-ahead warp +
// accumulator janitor warp).  The token loop, the FSM and the sampler live on the device, so there is no host round trip per token
// (the reference has >= 5).  Two layer variants (template FUSE): the tensor-parallel layer (default; two head-local flagged-word
// exchanges + two counting-atomic all-reduces, zero grid barriers — see the block comment above fix_add_cnt) and the five-exchange layer
// of round 1 (a grid barrier between dependent phases).  One grid barrier per token remains in both (after lm_head).
//
// HBM-bound by construction (B = 1 GEMV + single-query attention, ~1 flop/byte).  Every weight byte and every cached
// K/V byte is read exactly once per token.  The producer warp streams this CTA's slice of every phase — weight units,
// K blocks, V blocks, in consumption order — from HBM into a shared-memory ring with TMA bulk copies
// (cp.async.bulk.shared.global + mbarrier complete_tx; SASS UBLKCP); the run-ahead warp walks the same ranges further ahead with
// cp.async.bulk.prefetch.L2 (UBLKPF).  Weights and old K/V rows do not depend on the activations, so both streams run AHEAD of the
// consumers across every wait (7 x 24 KB of ring + 128 KB of L2 per SM).
// What bounds a token is not bandwidth but the chain of dependent steps, so the consumer side keeps that chain short: GEMV, q.K and P.V
// on mma.sync with the operands in registers / ldmatrix, no shuffle reductions inside a phase, fused residual + LayerNorm with one block
// reduction, per-warp softmax maxima merged with one barrier, biases / LayerNorm parameters fetched before the wait they follow.
// The new K/V row is appended in place with plain stores (the reference re-allocates and copies the whole cache per
// layer per step, modeling_opt.py:191-192).
//
// dtype ledger (mirrored by oracle/er_oracle.py mode='ledger'): fp16 weights and KV,
// fp32 accumulation everywhere, activations rounded to fp16 exactly where model.half()+autocast(fp16) does.
#include "decode_kernel.h"
#include "decode_partition.h"

#include <cstdio>

#include "common.cuh"

namespace er {

#ifndef ER_CONSUMER_WARPS
#define ER_CONSUMER_WARPS 8
#endif
constexpr int kConsumerWarps = ER_CONSUMER_WARPS;   // 8 or 16
constexpr int kUnitDiv = 1;                       // a weight unit is one K-slice of C fp16
static_assert(kConsumerWarps == 8, "the GEMV consumers assume 8 warps (one K-eighth / one unit per warp)");
constexpr int kConsumers = kConsumerWarps * 32;   // 256 compute threads
constexpr int kThreads = kConsumers + 96;         // + producer warp + L2 run-ahead warp + accumulator janitor warp
constexpr int HD = 96;                            // decoder head_dim (ArAE: 1536 / 16)
constexpr int HV = HD / 8;                        // 16-byte vectors per head row (12)
constexpr int kStageBytes = 24704;                // 8 padded weight units of (1536 + 8) fp16; also holds 4 K blocks / 128 V rows
constexpr int kKVChunk = 24576;                   // bytes per stage of the K / V jobs (4 K blocks = 128 V rows)
constexpr int kMaxStages = 8;
constexpr int kKBlockBytes = HV * 32 * 16;        // 6144: one 32-key block of the blocked K cache
constexpr int kHoStride = HD + 8;                 // fused phases: fp16 per (head, row) unit of the per-head out_proj copy (208 B)
constexpr int kHoUnitsPerStage = 112;            // (head, row) units per stage: 14 groups of 8 (23,296 B <= kStageBytes)
constexpr int kMaxUnits = 64 * kUnitDiv;          // weight units a CTA owns in one phase
constexpr int kPartStride = kMaxUnits + 1;        // lane-partial matrix [32][kPartStride] (odd stride: conflict-free both ways)

// ---- shared-memory / mbarrier / TMA bulk copy primitives (32-bit shared-space addresses -> LDS / SYNCS / UBLKCP) ----
__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
// three / four 16-byte loads issued back to back in ONE asm statement: the compiler cannot serialise load -> use -> load
__device__ __forceinline__ void lds128x3(uint32_t addr, uint4& a, uint4& b, uint4& c) {   // addr, addr + 512, addr + 1024
    asm volatile(
        "ld.shared.v4.u32 {%0,%1,%2,%3}, [%12];\n\t"
        "ld.shared.v4.u32 {%4,%5,%6,%7}, [%12+512];\n\t"
        "ld.shared.v4.u32 {%8,%9,%10,%11}, [%12+1024];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w), "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w)
        : "r"(addr));
}
__device__ __forceinline__ void lds128x4(uint32_t addr, uint4& a, uint4& b, uint4& c, uint4& d) {   // stride 512 bytes
    asm volatile(
        "ld.shared.v4.u32 {%0,%1,%2,%3}, [%16];\n\t"
        "ld.shared.v4.u32 {%4,%5,%6,%7}, [%16+512];\n\t"
        "ld.shared.v4.u32 {%8,%9,%10,%11}, [%16+1024];\n\t"
        "ld.shared.v4.u32 {%12,%13,%14,%15}, [%16+1536];"
        : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w), "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w),
          "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w)
        : "r"(addr));
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
    float r;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ void sts32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- barriers ---------------------------------------------------------------------------------------------------------------
// consumer-only block barrier (the producer warp never joins it)
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory"); }
// warps 0..3 only (the threads that publish a split partial)
__device__ __forceinline__ void pbar() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// grid barrier over the consumers of all CTAs (release/acquire on one monotonically increasing counter)
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch, const bool nosync) {
    cbar();
    if (nosync) return;                 // diagnostics only (DecodeParams::dbg_nosync)
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

// ---- flagged exchange between CTAs ("LL" words) -----------------------------------------------------------------------------------
// A grid barrier costs ~2.5 us here (drain the CTA's stores for the release, one atomic, one polled acquire) and the data it
// guards is then fetched with one more L2 round trip.  The small vectors that cross the S CTAs of a head (q / new k / new v, the
// split partials) are instead published as 8-byte words {payload, flag} with single-copy-atomic 64-bit stores; a reader polls the
// words it needs until each carries the flag of the current (token, layer).  Arrival of the data IS the synchronisation: one store +
// one load on the critical path, no fences.  (Round 1 published EVERY exchanged vector this way, all 148 CTAs polling 24 KB of fc1
// output: a polling storm, slower than barriers.  With S = 9 pollers per word there is none.)  Reuse is safe without handshakes: a
// word is rewritten for layer l+1 only by a CTA that has seen an all-reduce of layer l complete, i.e. after every reader of layer l is
// done.  Flags are unique per request (t * layers + layer + 1); the host zeroes the words before each launch.
constexpr int kSpinLimit = 1 << 22;   // ~ seconds; a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void ll_store(unsigned long long* ptr, uint32_t data, uint32_t flag) {
    const unsigned long long v = ((unsigned long long)flag << 32) | data;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
}
__device__ __forceinline__ uint2 ll_load(const unsigned long long* ptr) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return make_uint2((uint32_t)v, (uint32_t)(v >> 32));
}
__device__ __forceinline__ uint4 ll_load2(const unsigned long long* ptr) {   // two consecutive words (16-byte aligned)
    unsigned long long a, b;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(ptr) : "memory");
    return make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
}
// Poll up to N words per thread (bit j of `mask`: word j wanted).  Must be called by ALL lanes of a warp (mask 0 = nothing wanted).
// All loads of a round are in flight together.  After `rounds` unsuccessful rounds the warp stops hammering the L2 with 32 x N
// loads per round: one lane that still misses a word spins on that word alone and the others wait for it at a warp barrier.
template <int N, typename AddrFn>
__device__ __forceinline__ void ll_poll(uint32_t (&out)[N], uint32_t mask, uint32_t flag, int rounds, AddrFn addr) {
    int spins = 0;
    for (;;) {
        uint2 v[N];
#pragma unroll
        for (int j = 0; j < N; j++) if ((mask >> j) & 1u) v[j] = ll_load(addr(j));
#pragma unroll
        for (int j = 0; j < N; j++) if (((mask >> j) & 1u) && v[j].y == flag) { out[j] = v[j].x; mask &= ~(1u << j); }
        const uint32_t missing = __ballot_sync(0xffffffffu, mask != 0u);
        if (!missing) break;
        if (++spins > kSpinLimit) asm volatile("trap;");
        if (spins >= rounds) {
            if ((int)(threadIdx.x & 31) == __ffs(missing) - 1) {
                const unsigned long long* w = addr(__ffs(mask) - 1);
                int sp2 = 0;
                while (ll_load(w).y != flag) { if (++sp2 > kSpinLimit) asm volatile("trap;"); }
            }
            __syncwarp();
        }
    }
}
// same with 16-byte loads: N double-words, out[2j], out[2j+1]
template <int N, typename AddrFn>
__device__ __forceinline__ void ll_poll2(uint32_t (&out)[2 * N], uint32_t mask, uint32_t flag, int rounds, AddrFn addr) {
    int spins = 0;
    for (;;) {
        uint4 v[N];
#pragma unroll
        for (int j = 0; j < N; j++) if ((mask >> j) & 1u) v[j] = ll_load2(addr(j));
#pragma unroll
        for (int j = 0; j < N; j++)
            if (((mask >> j) & 1u) && v[j].y == flag && v[j].w == flag) { out[2 * j] = v[j].x; out[2 * j + 1] = v[j].z; mask &= ~(1u << j); }
        const uint32_t missing = __ballot_sync(0xffffffffu, mask != 0u);
        if (!missing) break;
        if (++spins > kSpinLimit) asm volatile("trap;");
        if (spins >= rounds) {
            if ((int)(threadIdx.x & 31) == __ffs(missing) - 1) {
                const unsigned long long* w = addr(__ffs(mask) - 1);
                int sp2 = 0;
                for (;;) {
                    const uint4 t4 = ll_load2(w);
                    if (t4.y == flag && t4.w == flag) break;
                    if (++sp2 > kSpinLimit) asm volatile("trap;");
                }
            }
            __syncwarp();
        }
    }
}
// a consumer thread holding the fp16 output of row (r0 + tid / nu_row) on threads with tid % nu_row == 0 publishes row pairs;
// every lane of a publishing warp must call it
__device__ __forceinline__ void ll_publish_rows(unsigned long long* dst, int r0, int nu, int nu_row, __half hv, uint32_t flag) {
    const uint32_t mine = __half_as_ushort(hv);
    const uint32_t other = __shfl_down_sync(0xffffffffu, mine, nu_row);
    const int tid = threadIdx.x;
    const bool wr = tid < nu && (tid % (2 * nu_row)) == 0;
    if (wr) ll_store(dst + ((r0 + tid / nu_row) >> 1), mine | (other << 16), flag);
}

// the same barrier in two halves: work placed between them overlaps the counter round trip
__device__ __forceinline__ void grid_arrive(unsigned* counter, unsigned& epoch, const bool nosync) {
    cbar();
    if (threadIdx.x == 0 && !nosync) {
        epoch += 1;
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    }
}
__device__ __forceinline__ void grid_wait(unsigned* counter, unsigned epoch, const bool nosync) {
    if (threadIdx.x == 0 && !nosync) {
        const unsigned target = epoch * gridDim.x;
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
    }
    cbar();
}

// ---- optional phase timeline (profiles/): one CTA, thread 0 stamps %globaltimer ------------------------------------------------
// slot map: [0] token start; per layer l, base = 1 + 16*l: +0 residual+LN2 done, +1 qkv gemv done, +2 P1 epilogue done, +3 B1,
// +4 attention done, +5 B2, +6 attn16 loaded, +7 out_proj done, +8 B3, +9 LN1 done, +10 fc1 done, +11 B4, +12 h1 loaded,
// +13 fc2 done, +14 B5; after the layers: +0 lm_head done, +1 final barrier.
__device__ __forceinline__ void prof_stamp(const DecodeParams& p, int slot, bool on) {
    if (on && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.prof[slot] = t;
    }
}
// the same 15 per-layer stamps for EVERY CTA, for one layer (5) of the profiled token: prof[4096 + 16 * cta + k]
__device__ __forceinline__ void prof_all(const DecodeParams& p, int k, bool on) {
    if (on && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        p.prof[4096 + 16 * blockIdx.x + k] = t;
    }
}

// ---- work partition (identical on the producer and the consumer side) ------------------------------------------------------------
__device__ __forceinline__ RowRange cta_rows(int R) { return cta_rows_of(R, blockIdx.x, gridDim.x); }
__device__ __forceinline__ bool attn_range(int H, int S, int split_handicap, int L, AttnRange& a) {
    return attn_range_of(H, S, split_handicap, L, blockIdx.x, a);
}

// P1 rows of this CTA in the 3C-row qkv matrix.  Five-exchange layer: an even 148-way split.  Tensor-parallel layer: CTA (head h, split s)
// computes 288 / S rows of ITS head (q96 | k96 | v96; the host picks S in {6, 9, 12}: a range never straddles two matrices, <= 48 rows).
template <bool FUSE>
__device__ __forceinline__ RowRange p1_rows(const DecodeParams& p) {
    if (!FUSE) return cta_rows(3 * p.C);
    RowRange rr{0, 0};
    const int S = p.S, b = (int)blockIdx.x;
    if (b >= p.H * S) return rr;
    const int h = b / S, s = b % S, per = 3 * HD / S;
    const int j0 = per * s, m = j0 / HD;
    rr.r0 = m * p.C + h * HD + (j0 - m * HD);
    rr.r1 = rr.r0 + per;
    return rr;
}

struct Ring {          // passed by value (registers): shared-space addresses of the stage data and the mbarrier arrays
    uint32_t data, full, empty;
    int nstage;
    __device__ __forceinline__ uint32_t stage(uint32_t s) const { return data + s * kStageBytes; }
    __device__ __forceinline__ uint32_t fullb(uint32_t s) const { return full + s * 8; }
    __device__ __forceinline__ uint32_t emptyb(uint32_t s) const { return empty + s * 8; }
};
struct Cursor {        // ring position; producer and every consumer warp advance it identically (one step per stage)
    uint32_t it, stage, parity;
    __device__ __forceinline__ void advance(int nstage) {
        ++it;
        if (++stage == (uint32_t)nstage) { stage = 0; parity ^= 1; }
    }
};

// ---- producer: stream one contiguous byte range through the ring ---------------------------------------------------------------------
// returns false if the consumers raised `stop` (EOS) while we were waiting for a free stage
__device__ __forceinline__ bool produce(const Ring r, Cursor& cur, const void* base, size_t bytes, uint32_t chunk, volatile int* stop) {
    const char* src = reinterpret_cast<const char*>(base);
    for (size_t off = 0; off < bytes; off += chunk, cur.advance(r.nstage)) {
        while (!mbar_try_wait(r.emptyb(cur.stage), cur.parity ^ 1)) {
            if (*stop) return false;
        }
        if (*stop) return false;
        const uint32_t n = (uint32_t)(bytes - off < (size_t)chunk ? bytes - off : (size_t)chunk);
        mbar_arrive_expect_tx(r.fullb(cur.stage), n);
        bulk_g2s(r.stage(cur.stage), src + off, n, r.fullb(cur.stage));
    }
    return true;
}

// ---- the byte ranges a CTA streams, in consumption order (shared by the ring producer and the L2 run-ahead warp) --------------------
// The kernel parameters are copied into registers up front: the consumers' acquire loads at the grid barrier invalidate the L1, and a
// parameter fetched through a reference (local / generic memory) right after that costs an L2 round trip per field.
// job(base, bytes, chunk, kv_pass): stream `bytes` from `base` in pieces of `chunk`; kv_pass > 0 marks the first K/V range of forward
// pass kv_pass (the rows of the previous token must be complete before a bulk COPY may read them); returns false to stop the walk.
struct SnapState { int t, L, counter, last_tok, done; };     // one snapshot of DecodeState per CTA (shared memory)
template <bool FUSE, typename Job>
__device__ __forceinline__ void walk_jobs(const DecodeParams& p, const int t0, int L, Job job) {
    const int C = p.C, F = p.F, H = p.H, S = p.S, V = p.V, layers = p.layers, ustride = p.ustride, handicap = p.split_handicap;
    const int nkb = p.nkb;
    const __half* const wdec = p.wdec;
    const __half* const wfuse = FUSE ? p.wfuse : nullptr;
    const size_t fuse_layer = FUSE ? (size_t)H * C * kHoStride + (size_t)F * ustride : 0;   // fp16 per layer of the fused-phase weight copies
    const __half* const kc = p.kc;
    const __half* const vc = p.vc;
    // forward passes of this launch: token tt is followed by a pass iff tt + 1 < max_new (EOS is handled through `stop`)
    const int n_fwd = min(p.steps, p.max_new - 1 - t0);
    bool ok = true;
    // decode weights live in `wdec` as units of C fp16 padded to `ustride` (bank-conflict-free ldmatrix rows); per layer:
    // [3C qkv rows][C out_proj rows][F fc1 rows][C fc2 rows x F/C units]; after the layers: V lm_head rows.
    const size_t ub = (size_t)ustride * 2;
    const uint32_t wchunk = (uint32_t)(p.upstage * ub);
    const int nuf = F / C;
    const size_t UL = (size_t)4 * C + 2 * (size_t)F;
    const RowRange rq = p1_rows<FUSE>(p), rc = cta_rows(C), rf = cta_rows(F), rv = cta_rows(V);
    for (int pass = 0; pass < n_fwd && ok; ++pass, ++L) {
        AttnRange a;
        const bool has_attn = attn_range(H, S, handicap, L, a);
        for (int layer = 0; layer < layers && ok; ++layer) {
            const __half* wl = wdec + (size_t)layer * UL * ustride;
            if (rq.r1 > rq.r0) ok = job(wl + (size_t)rq.r0 * ustride, (size_t)(rq.r1 - rq.r0) * ub, wchunk, 0);
            if (ok && has_attn) {
                const __half* kbase = kc + (((size_t)layer * H + a.h) * nkb + a.b0) * (size_t)(HV * 256);
                ok = job(kbase, (size_t)(a.b1 - a.b0) * kKBlockBytes, kKVChunk, layer == 0 ? pass : 0);
                const __half* vbase = vc + (((size_t)layer * H + a.h) * nkb + a.b0) * (size_t)(HV * 256);
                if (ok) ok = job(vbase, (size_t)(a.b1 - a.b0) * kKBlockBytes, kKVChunk, 0);
                if (FUSE && ok) {   // this split's rows of the head's out_proj columns
                    const RowRange rs = cta_rows_of(C, blockIdx.x % (unsigned)S, (unsigned)S);
                    ok = job(wfuse + (size_t)layer * fuse_layer + ((size_t)a.h * C + rs.r0) * kHoStride, (size_t)(rs.r1 - rs.r0) * kHoStride * 2,
                             (uint32_t)(kHoUnitsPerStage * kHoStride * 2), 0);
                }
            }
            if (!FUSE && ok) ok = job(wl + ((size_t)3 * C + rc.r0) * ustride, (size_t)(rc.r1 - rc.r0) * ub, wchunk, 0);
            if (ok) ok = job(wl + ((size_t)4 * C + rf.r0) * ustride, (size_t)(rf.r1 - rf.r0) * ub, wchunk, 0);
            if (!FUSE && ok) ok = job(wl + ((size_t)4 * C + F + (size_t)rc.r0 * nuf) * ustride, (size_t)(rc.r1 - rc.r0) * nuf * ub, wchunk, 0);
            if (FUSE && ok)     // the fc2 columns matching this CTA's fc1 rows (one transposed unit per column)
                ok = job(wfuse + (size_t)layer * fuse_layer + (size_t)H * C * kHoStride + (size_t)rf.r0 * ustride, (size_t)(rf.r1 - rf.r0) * ub, wchunk, 0);
        }
        if (ok) ok = job(wdec + ((size_t)layers * UL + rv.r0) * ustride, (size_t)(rv.r1 - rv.r0) * ub, wchunk, 0);
    }
}

// Ring producer.  Run-ahead bound: weights never change, but the K block / V rows that hold key L-1 were written by the PREVIOUS
// token's P1.  The ring alone does not bound the producer by tokens (a CTA that owns few rows of a small model needs fewer stages per
// token than the ring has slots), so K/V copies of pass j are only issued once the consumers have left the token-end grid barrier of
// pass j-1 (`tok_done`).  `issued` publishes the running byte count to the L2 run-ahead warp.
template <bool FUSE>
__device__ __noinline__ void producer_loop(const DecodeParams& p, const Ring r, const SnapState* snap, volatile int* stop, volatile uint32_t* cons_it,
                                           volatile int* tok_done, volatile unsigned long long* issued) {
    if (snap->done) return;
    Cursor cur{0u, 0u, 0u};
    bool ok = true;
    unsigned long long total = 0;
    walk_jobs<FUSE>(p, snap->t, snap->L, [&](const void* base, size_t bytes, uint32_t chunk, int kv_pass) -> bool {
        if (kv_pass > 0) {
            while (*tok_done < kv_pass) { if (*stop) { ok = false; return false; } }
            asm volatile("fence.proxy.async.global;" ::: "memory");   // rows written through the generic proxy, read by the bulk copy below
        }
        const char* src = reinterpret_cast<const char*>(base);
        for (size_t off = 0; off < bytes; off += chunk, cur.advance(r.nstage)) {
            while (!mbar_try_wait(r.emptyb(cur.stage), cur.parity ^ 1)) {
                if (*stop) { ok = false; return false; }
            }
            if (*stop) { ok = false; return false; }
            const uint32_t n = (uint32_t)(bytes - off < (size_t)chunk ? bytes - off : (size_t)chunk);
            mbar_arrive_expect_tx(r.fullb(cur.stage), n);
            bulk_g2s(r.stage(cur.stage), src + off, n, r.fullb(cur.stage));
            total += n;
            *issued = total;
        }
        return true;
    });
    if (!ok) {
        // EOS: the consumers stopped at stage *cons_it; every copy we issued beyond it must land before the CTA may exit
        for (uint32_t j = *cons_it; j < cur.it; ++j) mbar_wait(r.fullb(j % r.nstage), (j / r.nstage) & 1);
    }
}

// L2 run-ahead warp (one lane): walks the same byte ranges and asks the TMA unit to pull them into L2 (cp.async.bulk.prefetch.L2,
// fire and forget), staying at most pf_dist bytes ahead of what the ring producer has requested.  The ring (7 x 24 KB) covers ~4 us
// of HBM streaming; a layer has ~15 us of exchanges.  With the run-ahead HBM keeps streaming through them and the ring refills from
// L2 afterwards.  Pieces the producer has already requested are skipped (no duplicate DRAM traffic).  Prefetching a K/V line that
// the previous token is still writing is harmless: L2 is the point of coherence for those stores.
template <bool FUSE>
__device__ __noinline__ void prefetch_loop(const DecodeParams& p, const SnapState* snap, volatile int* stop, volatile unsigned long long* issued) {
    if (snap->done) return;
    const unsigned long long dist = (unsigned long long)p.pf_dist;
    unsigned long long total = 0;
    walk_jobs<FUSE>(p, snap->t, snap->L, [&](const void* base, size_t bytes, uint32_t chunk, int) -> bool {
        const char* src = reinterpret_cast<const char*>(base);
        for (size_t off = 0; off < bytes; off += chunk) {
            const uint32_t n = (uint32_t)(bytes - off < (size_t)chunk ? bytes - off : (size_t)chunk);
            unsigned long long have;
            while (total + n > (have = *issued) + dist) {
                if (*stop) return false;
                __nanosleep(100);
            }
            if (total >= have) prefetch_l2_bulk(src + off, n);
            total += n;
        }
        return !*stop;
    });
}

// ---- consumer: GEMV over one streamed weight slice ------------------------------------------------------------------------------------
// The slice is n_units units of C fp16 (a row of K = nu_row * C elements is nu_row consecutive units).  Warp w takes units
// w, w+8, ... of every stage; 8 and the units-per-stage are multiples of nu_row, so the x slice a warp needs is the same for every
// unit it ever touches in this job and lives in registers.  No cross-lane reduction here: lane partials go to part[lane][unit]
// and are summed per output row by reduce_rows() after the job.  NVL = 16-byte vectors per lane per unit (6 for C = 1536).
template <int NVL>
__device__ __forceinline__ Cursor gemv_job_t(const Ring r, Cursor cur, int n_units, int nu_row, int C, int ustride, int upstage, uint32_t xin_s, uint32_t part_s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ubytes = ustride * 2;                          // bytes per (padded) unit
    const int nvec = C >> 3;                                 // 16-byte vectors of payload per unit
    const int nch = (n_units + upstage - 1) / upstage;
    float xr[NVL][8];
    {
        const uint32_t xb = xin_s + (uint32_t)((warp % nu_row) * C) * 2;
#pragma unroll
        for (int j = 0; j < NVL; j++) {
            const int v = lane + 32 * j;
            const uint4 xv = (v < nvec) ? lds128(xb + v * 16) : make_uint4(0, 0, 0, 0);
            float2 f;
            f = h2f2(xv.x); xr[j][0] = f.x; xr[j][1] = f.y;
            f = h2f2(xv.y); xr[j][2] = f.x; xr[j][3] = f.y;
            f = h2f2(xv.z); xr[j][4] = f.x; xr[j][5] = f.y;
            f = h2f2(xv.w); xr[j][6] = f.x; xr[j][7] = f.y;
        }
    }
    const uint32_t pl = part_s + (uint32_t)lane * kPartStride * 4;
    for (int c = 0; c < nch; ++c, cur.advance(r.nstage)) {
        mbar_wait(r.fullb(cur.stage), cur.parity);
        const int here = min(upstage, n_units - c * upstage);
        const uint32_t st = r.stage(cur.stage);
        for (int u = warp; u < here; u += kConsumerWarps) {
            const uint32_t wb = st + (uint32_t)u * ubytes;
            uint4 wv[NVL];
#pragma unroll
            for (int j = 0; j < NVL; j++) {
                const int v = lane + 32 * j;
                wv[j] = (v < nvec) ? lds128(wb + v * 16) : make_uint4(0, 0, 0, 0);
            }
            float tot = 0.f;
#pragma unroll
            for (int j = 0; j < NVL; j++) {
                float2 f; float a0, a1;
                f = h2f2(wv[j].x); a0 = f.x * xr[j][0];           a1 = f.y * xr[j][1];
                f = h2f2(wv[j].y); a0 = fmaf(f.x, xr[j][2], a0);  a1 = fmaf(f.y, xr[j][3], a1);
                f = h2f2(wv[j].z); a0 = fmaf(f.x, xr[j][4], a0);  a1 = fmaf(f.y, xr[j][5], a1);
                f = h2f2(wv[j].w); a0 = fmaf(f.x, xr[j][6], a0);  a1 = fmaf(f.y, xr[j][7], a1);
                tot += a0 + a1;
            }
            sts32(pl + (uint32_t)(c * upstage + u) * 4, tot);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(r.emptyb(cur.stage));
    }
    cbar();
    return cur;
}

// ---- tensor-core GEMV (C % 256 == 0): one stage = 8 weight units = the 8 columns of an m16n8k16 B operand -------------------------
// The CUDA-core loop above needs ~17 instructions per 16 weights per lane (fp16->fp32 converts + FMAs) and is issue/latency
// bound at 8 warps.  Here each warp owns 1/8 of the K range of ALL 8 units of a stage: ldmatrix.x4 pulls two k16 steps of the
// 8 unit rows (rows are padded to C+8 fp16 in HBM, so the 8 row addresses hit distinct banks), mma.sync accumulates
// x . W^T in fp32 (fp16 products are exact; same numerics class as the cuBLAS GEMV the reference calls).  The A operand
// carries x: row m holds the x slice of unit column m % nu_row (a K = nu_row * C row spans nu_row consecutive units), kept in
// registers for the whole phase.  The wanted outputs D[n % nu_row][n] are written as per-warp partials to part[warp][unit].
// KS = k16 steps per warp (C / 128).
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float* d, uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a0), "r"(a2), "r"(a2), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
// six / three / ... ldmatrix.x4 issued back to back in one asm statement (two k16 steps each, 64 bytes apart)
template <int N>
__device__ __forceinline__ void ldmatrix_batch(uint32_t addr, uint32_t (*b)[4]) {
#pragma unroll
    for (int i = 0; i < N; i++) ldmatrix_x4(addr + i * 64, b[i][0], b[i][1], b[i][2], b[i][3]);
}
template <>
__device__ __forceinline__ void ldmatrix_batch<6>(uint32_t addr, uint32_t (*b)[4]) {
    asm volatile(
        "ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%24];\n\t"
        "ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%4,%5,%6,%7}, [%24+64];\n\t"
        "ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%8,%9,%10,%11}, [%24+128];\n\t"
        "ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%12,%13,%14,%15}, [%24+192];\n\t"
        "ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%16,%17,%18,%19}, [%24+256];\n\t"
        "ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%20,%21,%22,%23}, [%24+320];"
        : "=r"(b[0][0]), "=r"(b[0][1]), "=r"(b[0][2]), "=r"(b[0][3]), "=r"(b[1][0]), "=r"(b[1][1]), "=r"(b[1][2]), "=r"(b[1][3]),
          "=r"(b[2][0]), "=r"(b[2][1]), "=r"(b[2][2]), "=r"(b[2][3]), "=r"(b[3][0]), "=r"(b[3][1]), "=r"(b[3][2]), "=r"(b[3][3]),
          "=r"(b[4][0]), "=r"(b[4][1]), "=r"(b[4][2]), "=r"(b[4][3]), "=r"(b[5][0]), "=r"(b[5][1]), "=r"(b[5][2]), "=r"(b[5][3])
        : "r"(addr));
}
template <int KS>
__device__ __forceinline__ Cursor gemv_job_mma(const Ring r, Cursor cur, int n_units, int nu_row, int C, int ustride, uint32_t xin_s, uint32_t part_s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ubytes = ustride * 2;
    const int nch = (n_units + 7) >> 3;
    const int k0 = warp * (C >> 3);                          // this warp's K range: [k0, k0 + C/8)
    uint32_t xa[KS][2];                                      // A fragments (rows g and g+8 carry the same x slice)
    {
        const uint32_t xb = xin_s + (uint32_t)((g % nu_row) * C + k0 + 2 * t) * 2;
#pragma unroll
        for (int s = 0; s < KS; s++) { xa[s][0] = lds_u32(xb + s * 32); xa[s][1] = lds_u32(xb + s * 32 + 16); }
    }
    // D[row g][cols 2t, 2t+1]: column n is wanted from row n % nu_row — decided once, not per stage
    const bool w0 = g == ((2 * t) % nu_row), w1 = g == ((2 * t + 1) % nu_row);
    const uint32_t row_off = (uint32_t)(lane & 7) * ubytes + (uint32_t)(k0 + (lane >> 3) * 8) * 2;
    const uint32_t pw = part_s + (uint32_t)warp * kPartStride * 4 + (uint32_t)(2 * t) * 4;
    int left = n_units;
    for (int c = 0; c < nch; ++c, left -= 8, cur.advance(r.nstage)) {
        mbar_wait(r.fullb(cur.stage), cur.parity);
        uint32_t b[KS / 2][4];
        ldmatrix_batch<KS / 2>(r.stage(cur.stage) + row_off, b);
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};      // two independent accumulator chains
#pragma unroll
        for (int s2 = 0; s2 < KS / 2; s2++) {
            mma_16816(d0, xa[2 * s2][0], xa[2 * s2][1], b[s2][0], b[s2][1]);
            mma_16816(d1, xa[2 * s2 + 1][0], xa[2 * s2 + 1][1], b[s2][2], b[s2][3]);
        }
        if (w0 && 2 * t < left) sts32(pw + (uint32_t)c * 32, d0[0] + d1[0]);
        if (w1 && 2 * t + 1 < left) sts32(pw + (uint32_t)c * 32 + 4, d0[1] + d1[1]);
        __syncwarp();
        if (lane == 0) mbar_arrive(r.emptyb(cur.stage));
    }
    cbar();
    return cur;
}
// (the cursor travels by value so that it stays in registers across the call)
struct GemvCfg { int C, ustride, upstage, use_mma; };     // by value: stays in registers (a reference to the kernel params would be LDL traffic)
__device__ __noinline__ Cursor gemv_job(const Ring r, Cursor cur, int n_units, int nu_row, const GemvCfg gc, uint32_t xin_s, uint32_t part_s) {
    const int C = gc.C;
    if (gc.use_mma) {
        if (C == 1536) return gemv_job_mma<12>(r, cur, n_units, nu_row, C, gc.ustride, xin_s, part_s);
        if (C == 768) return gemv_job_mma<6>(r, cur, n_units, nu_row, C, gc.ustride, xin_s, part_s);
        if (C == 1024) return gemv_job_mma<8>(r, cur, n_units, nu_row, C, gc.ustride, xin_s, part_s);
        if (C == 512) return gemv_job_mma<4>(r, cur, n_units, nu_row, C, gc.ustride, xin_s, part_s);
        return gemv_job_mma<2>(r, cur, n_units, nu_row, C, gc.ustride, xin_s, part_s);   // C == 256
    }
    const int nvl = ((C >> 3) + 31) >> 5;
    if (nvl <= 1) return gemv_job_t<1>(r, cur, n_units, nu_row, C, gc.ustride, gc.upstage, xin_s, part_s);
    if (nvl == 2) return gemv_job_t<2>(r, cur, n_units, nu_row, C, gc.ustride, gc.upstage, xin_s, part_s);
    if (nvl <= 4) return gemv_job_t<4>(r, cur, n_units, nu_row, C, gc.ustride, gc.upstage, xin_s, part_s);
    return gemv_job_t<6>(r, cur, n_units, nu_row, C, gc.ustride, gc.upstage, xin_s, part_s);
}
// Sum the partials of each unit (32 lanes on the CUDA-core path, 8 warps on the tensor-core path) and the nu_row (1, 2, 4 or 8) units of each output row; thread t owns unit t.
// Returns the row sum on threads with t % nu_row == 0 (row t / nu_row); every thread of a participating warp must call it.
__device__ __forceinline__ float reduce_rows(uint32_t part_s, int n_units, int nu_row, int nparts) {
    const int t = threadIdx.x;
    float s = 0.f;
    if (t < n_units) {
        const uint32_t pb = part_s + (uint32_t)t * 4;
#pragma unroll 8
        for (int l = 0; l < nparts; l++) s += lds32(pb + (uint32_t)l * kPartStride * 4);
    }
    for (int o = 1; o < nu_row; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// ---- tensor-parallel layer (FUSE = true): K-split GEMVs whose partial sums meet in L2, arrival counted INSIDE the data words ----
// Measured on the B200 (profiles/r02_diag_runahead_nosync_fuse.json): with the grid barriers switched off the default kernel runs a
// layer in 16 us at L = 2050, with them in 28 us — five exchanges of ~2.4 us each (drain stores for the release, atomic, polled
// acquire, then fetch the vector: four dependent L2 trips), none of which can overlap anything because the layer is one dependency
// chain.  This variant cuts the layer the way tensor-parallel training cuts it, so that a vector only crosses CTAs where it must:
//   * q/k/v rows of head h are computed BY the S CTAs that attend for head h (288 rows / S each): q, the new k and the new v travel
//     as flagged words among S CTAs (no storm: 9 pollers per word), no grid barrier;
//   * out_proj is row-parallel per head: after an S-way flagged-word merge every CTA of the head multiplies the head's output by ITS
//     rows of the head's 96 out_proj columns (tensor cores) and adds the partial rows into an accumulator in L2;
//   * fc2 is row-parallel over fc1's column split: a CTA multiplies ITS 41-42 fc1 outputs by the matching columns of W2 (stored
//     transposed, one unit per column; ldmatrix.trans + mma.m16n8k8) and adds the 1536 partial sums into a second accumulator.
// The accumulators are u64 words: bits 0..55 hold a biased fixed-point sum (2^-32; integer addition is associative, so the result
// is independent of arrival order and runs stay bit-reproducible), bits 56..63 COUNT the addends.  A reader polls the words it needs
// until the count is complete: the data is its own flag — no release fence, no barrier, no separate fetch.  Two head-local
// exchanges + two L2 all-reduces per layer, zero grid barriers (one per token remains, after lm_head).
// Accumulator reuse: reduction k uses copy k & 3, and EVERY CTA adds to every word of every reduction (a CTA with nothing to add adds
// zero).  A CTA that has seen reduction k complete knows every CTA has finished reading reduction k-2 (everybody reads k-2 before adding
// to k-1, and k-1 completed before k could), so its janitor warp zeroes this CTA's slice of copy (k+2) & 3 and fences; the CTA adds to
// reduction k+1 only after its janitor has caught up.  Whoever adds to reduction k+2 has seen k+1 complete, i.e. every CTA added to
// k+1, i.e. every slice of copy (k+2) & 3 was zeroed at L2 before.
constexpr float kFixScale = 4294967296.0f;         // 2^32
constexpr float kFixInv = 1.0f / 4294967296.0f;
constexpr unsigned long long kFixBias = 1ull << 47;        // every addend is non-negative and < 2^48: 255 of them cannot carry into the count
constexpr unsigned long long kFixOne = 1ull << 56;
constexpr unsigned long long kFixMask = (1ull << 56) - 1;
// Element u lives at word u * kAccStride.  Measured: a fc2 reduction (148 x 1536 atomics) is observed complete ~5.5 us after a CTA has
// issued its own adds whether the accumulators are packed (stride 1) or spread one per 32-byte sector (stride 4: 7.9 us — more sectors per
// warp-wide RED), and ~3.5 us later with 4x fewer atomics (groups of four, below): the cost is the all-to-all hop itself.
constexpr int kAccStride = 1;
__device__ __forceinline__ void fix_add_cnt(unsigned long long* acc, float v) {
    v = fminf(fmaxf(v, -32000.f), 32000.f);                 // also maps NaN to a number: the count must always arrive
    const long long q = __float2ll_rn(v * kFixScale);       // exact for |v| >= 2^-9, rounded to 2^-32 below
    const unsigned long long w = kFixOne + kFixBias + (unsigned long long)q;
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(acc), "l"(w) : "memory");
}
// thread polls its 8 accumulators (elements 8t .. 8t+7, `acc` points at the first) until each has `target` addends; -> the 8 sums.  While
// waiting only ONE word per thread is polled (all elements complete within a fraction of a microsecond of each other; a light poll does not
// slow the atomics it is waiting for); then all eight are read and checked.  All lanes of a warp must call it (act = false: nothing
// wanted); the exit is warp-uniform.  nosync (diagnostics): take whatever is there.
__device__ __forceinline__ void acc_poll8(const unsigned long long* acc, const int target, float* out, const bool act, const bool nosync) {
    const unsigned long long bias = (unsigned long long)target * kFixBias;
    int spins = 0;
    if (!nosync) {
        for (;;) {
            bool ok = true;
            if (act) {
                unsigned long long w0;
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w0) : "l"(acc + 7 * kAccStride) : "memory");
                ok = (int)(w0 >> 56) == target;
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (++spins > kSpinLimit) asm volatile("trap;");
        }
    }
    for (;;) {
        bool ok = true;
        if (act) {
            unsigned long long w[8];
#pragma unroll
            for (int j = 0; j < 8; j++)
                asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w[j]) : "l"(acc + j * kAccStride) : "memory");
#pragma unroll
            for (int e = 0; e < 8; e++) ok = ok && ((int)(w[e] >> 56) == target);
            if (ok || nosync) {
#pragma unroll
                for (int e = 0; e < 8; e++) out[e] = __ll2float_rn((long long)((w[e] & kFixMask) - bias)) * kFixInv;
            }
        }
        if (__all_sync(0xffffffffu, ok) || nosync) break;
        if (++spins > kSpinLimit) asm volatile("trap;");
    }
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_1688(float* d, uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}

// out_proj, row-parallel per head: this CTA's n rows of the head's 96 out_proj columns (units of HD fp16 padded to kHoStride) times the
// head's merged attention output o16[96].  A group of 8 units is the B operand of 6 k16 steps; the full K lives in one warp, so there
// is no cross-warp reduction: lanes 0..3 of the warp that owns a group hold its 8 outputs.  Warp w takes groups w, w+8, ... of a stage.
__device__ __forceinline__ Cursor outproj_job_mma(const Ring r, Cursor cur, const int n, const uint32_t o16_s, const uint32_t out_s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 3;
    uint32_t xa[HD / 16][2];
#pragma unroll
    for (int s = 0; s < HD / 16; s++) { xa[s][0] = lds_u32(o16_s + (uint32_t)(16 * s + 2 * t) * 2); xa[s][1] = lds_u32(o16_s + (uint32_t)(16 * s + 8 + 2 * t) * 2); }
    const uint32_t row_off = (uint32_t)(lane & 7) * (kHoStride * 2) + (uint32_t)(lane >> 3) * 16;
    const int nch = (n + kHoUnitsPerStage - 1) / kHoUnitsPerStage;
    for (int c = 0; c < nch; ++c, cur.advance(r.nstage)) {
        mbar_wait(r.fullb(cur.stage), cur.parity);
        const int here = min(kHoUnitsPerStage, n - c * kHoUnitsPerStage);
        const int ngrp = (here + 7) >> 3;
        for (int gi = warp; gi < ngrp; gi += kConsumerWarps) {
            const uint32_t base = r.stage(cur.stage) + (uint32_t)gi * (8 * kHoStride * 2) + row_off;
            uint32_t b[HD / 32][4];
#pragma unroll
            for (int i = 0; i < HD / 32; i++) ldmatrix_x4(base + i * 64, b[i][0], b[i][1], b[i][2], b[i][3]);
            float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int i = 0; i < HD / 32; i++) {
                mma_16816(d0, xa[2 * i][0], xa[2 * i][1], b[i][0], b[i][1]);
                mma_16816(d1, xa[2 * i + 1][0], xa[2 * i + 1][1], b[i][2], b[i][3]);
            }
            if (lane < 4) {     // row 0 of D (all 16 rows carry the same x): columns 2t, 2t+1 = units 8 gi + 2t (+1)
                const int u = c * kHoUnitsPerStage + gi * 8 + 2 * t;
                if (u < n) sts32(out_s + (uint32_t)u * 4, d0[0] + d1[0]);
                if (u + 1 < n) sts32(out_s + (uint32_t)(u + 1) * 4, d0[1] + d1[1]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(r.emptyb(cur.stage));
    }
    return cur;
}

// fc2, row-parallel over fc1's column split: y2_partial[C] = sum over this CTA's nr fc1 outputs h[j] of column j of W2.  The columns
// arrive as transposed units (one column = C fp16 + pad, 8 per stage).  ldmatrix.trans turns an 8-column x 8-row patch into the A
// fragment of mma.m16n8k8 (output rows x columns-as-k); the B operand carries h in its column 0.  Warp w owns output rows
// [w C/8, (w+1) C/8): MT = C/128 m-tiles of 16 rows, accumulated in registers over the whole phase.
template <int MT>
__device__ __forceinline__ Cursor fc2_job_mma(const Ring r, Cursor cur, const int nr, const int C, const int ustride, const uint32_t h_s, const uint32_t out_s) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int ubytes = ustride * 2;
    const int ib = warp * (C >> 3);
    float acc[MT][4];
#pragma unroll
    for (int m = 0; m < MT; m++) { acc[m][0] = 0.f; acc[m][1] = 0.f; acc[m][2] = 0.f; acc[m][3] = 0.f; }
    const uint32_t lane_off = (uint32_t)(lane & 7) * ubytes + (uint32_t)(ib + (lane >> 3) * 8) * 2;
    const int nch = (nr + 7) >> 3;
    for (int c = 0; c < nch; ++c, cur.advance(r.nstage)) {
        mbar_wait(r.fullb(cur.stage), cur.parity);
        const uint32_t b0 = (g == 0) ? lds_u32(h_s + (uint32_t)(8 * c + 2 * t) * 2) : 0u;      // h is zero-padded to a multiple of 8
        const uint32_t base = r.stage(cur.stage) + lane_off;
#pragma unroll
        for (int m2 = 0; m2 < MT / 2; m2++) {
            uint32_t a0, a1, a2, a3;
            ldmatrix_x4_trans(base + m2 * 64, a0, a1, a2, a3);
            mma_1688(acc[2 * m2], a0, a1, b0);
            mma_1688(acc[2 * m2 + 1], a2, a3, b0);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(r.emptyb(cur.stage));
    }
    if (t == 0) {          // column 0 of D
#pragma unroll
        for (int m = 0; m < MT; m++) {
            sts32(out_s + (uint32_t)(ib + 16 * m + g) * 4, acc[m][0]);
            sts32(out_s + (uint32_t)(ib + 16 * m + g + 8) * 4, acc[m][2]);
        }
    }
    return cur;
}
__device__ __noinline__ Cursor fc2_job(const Ring r, Cursor cur, const int nr, const int C, const int ustride, const uint32_t h_s, const uint32_t out_s) {
    if (C == 1536) return fc2_job_mma<12>(r, cur, nr, C, ustride, h_s, out_s);
    if (C == 1024) return fc2_job_mma<8>(r, cur, nr, C, ustride, h_s, out_s);
    if (C == 768) return fc2_job_mma<6>(r, cur, nr, C, ustride, h_s, out_s);
    if (C == 512) return fc2_job_mma<4>(r, cur, nr, C, ustride, h_s, out_s);
    return fc2_job_mma<2>(r, cur, nr, C, ustride, h_s, out_s);      // C == 256
}

// ---- fused residual + LayerNorm --------------------------------------------------------------------------------------------------------------
// x = LayerNorm(xres + y) in place (fp32 statistics, eps 1e-5, fp16 affine params), plus the fp16 copy used as the next GEMV
// input.  y is the fp16 phase output other CTAs just published (read at L2).  Thread t < C/8 owns elements 8t..8t+7;
// its affine parameters were fetched into registers BEFORE the grid barrier (LnParams).  round_first: the residual add of
// the first layer of a decode step is an fp16 + fp16 add.  One block reduction, two barriers.
struct LnParams { uint4 g, b; };
__device__ __forceinline__ LnParams ln_load(const __half* __restrict__ g, const __half* __restrict__ b, int C) {
    LnParams lp;
    lp.g = make_uint4(0, 0, 0, 0); lp.b = lp.g;
    if ((int)threadIdx.x < (C >> 3)) {
        lp.g = *reinterpret_cast<const uint4*>(g + threadIdx.x * 8);
        lp.b = *reinterpret_cast<const uint4*>(b + threadIdx.x * 8);
    }
    return lp;
}
__device__ __forceinline__ void unpack8(const uint4 u, float* f) {
    float2 t;
    t = h2f2(u.x); f[0] = t.x; f[1] = t.y;
    t = h2f2(u.y); f[2] = t.x; f[3] = t.y;
    t = h2f2(u.z); f[4] = t.x; f[5] = t.y;
    t = h2f2(u.w); f[6] = t.x; f[7] = t.y;
}
// yacc != nullptr (tensor-parallel layer): y = fp16(all-reduced sum + bias); the sum is polled from the counting accumulator.
__device__ __forceinline__ void residual_layer_norm(uint32_t xres_s, uint32_t x16_s, const __half* y, bool round_first, const LnParams lp, int C, float inv_c,
                                                 float* red, const unsigned long long* yacc = nullptr, const int target = 0,
                                                 const __half* ybias = nullptr, const bool nosync = false) {
    const int t = threadIdx.x;
    const bool act = t < (C >> 3);
    float v[8];
    float s = 0.f, q = 0.f;
    float yv[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (yacc) {
        uint4 bvv = make_uint4(0, 0, 0, 0);
        if (act) bvv = *reinterpret_cast<const uint4*>(ybias + 8 * t);     // in flight during the poll
        acc_poll8(yacc + (size_t)8 * t * kAccStride, target, yv, act, nosync);
        float bv[8];
        unpack8(bvv, bv);
#pragma unroll
        for (int e = 0; e < 8; e++) yv[e] = round_f16(yv[e] + bv[e]);
    } else if (act) {
        unpack8(ldg_cg(reinterpret_cast<const uint4*>(y) + t), yv);
    }
    if (act) {
        const float4 a = lds_f4(xres_s + t * 32), b = lds_f4(xres_s + t * 32 + 16);
        v[0] = a.x + yv[0]; v[1] = a.y + yv[1]; v[2] = a.z + yv[2]; v[3] = a.w + yv[3];
        v[4] = b.x + yv[4]; v[5] = b.y + yv[5]; v[6] = b.z + yv[6]; v[7] = b.w + yv[7];
#pragma unroll
        for (int e = 0; e < 8; e++) {
            if (round_first) v[e] = round_f16(v[e]);
            s += v[e]; q += v[e] * v[e];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if ((t & 31) == 0) { red[t >> 5] = s; red[16 + (t >> 5)] = q; }
    cbar();
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int i = 0; i < kConsumerWarps; i++) { ts += red[i]; tq += red[16 + i]; }
    const float mean = ts * inv_c;
    const float var = fmaxf(tq * inv_c - mean * mean, 0.f);
    const float rstd = rsqrtf(var + 1e-5f);
    if (act) {
        float g[8], b[8], o[8];
        unpack8(lp.g, g); unpack8(lp.b, b);
#pragma unroll
        for (int e = 0; e < 8; e++) o[e] = (v[e] - mean) * rstd * g[e] + b[e];
        sts_f4(xres_s + t * 32, make_float4(o[0], o[1], o[2], o[3]));
        sts_f4(xres_s + t * 32 + 16, make_float4(o[4], o[5], o[6], o[7]));
        uint4 h;
        __half2 h0 = __floats2half2_rn(o[0], o[1]), h1 = __floats2half2_rn(o[2], o[3]), h2 = __floats2half2_rn(o[4], o[5]), h3 = __floats2half2_rn(o[6], o[7]);
        h.x = *reinterpret_cast<uint32_t*>(&h0); h.y = *reinterpret_cast<uint32_t*>(&h1);
        h.z = *reinterpret_cast<uint32_t*>(&h2); h.w = *reinterpret_cast<uint32_t*>(&h3);
        sts128(x16_s + t * 16, h);
    }
    cbar();
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
// K pass: warp = 32-key block (blocks round-robin over the 8 warps), lane = key: 12 conflict-free 16-byte smem reads give one
// complete q.k per lane, no cross-lane reduction.  Each warp normalises ITS keys with its own running maximum; the 8 warp
// maxima are merged with ONE barrier and folded into the V pass as a per-block scale, the softmax denominator is accumulated in
// the V pass too.  V pass: warp = 2 rows x 12 vectors per step.  The new key L (appended by P1 of this step) is read from
// global by the CTA of the last split.  Each split publishes (o[96], max, sum); the LAST split of a head to finish (atomic
// ticket) merges the S partials of that head and publishes attn16[h*96 ..].
__device__ __forceinline__ float dot_k8(const uint4 kv, const float4 qa, const float4 qb, float acc) {
    float2 f;
    f = h2f2(kv.x); acc = fmaf(f.x, qa.x, acc); acc = fmaf(f.y, qa.y, acc);
    f = h2f2(kv.y); acc = fmaf(f.x, qa.z, acc); acc = fmaf(f.y, qa.w, acc);
    f = h2f2(kv.z); acc = fmaf(f.x, qb.x, acc); acc = fmaf(f.y, qb.y, acc);
    f = h2f2(kv.w); acc = fmaf(f.x, qb.z, acc); acc = fmaf(f.y, qb.w, acc);
    return acc;
}
#ifdef ER_ATTN_NOINLINE
#define ER_ATTN_INLINE __noinline__
#else
#define ER_ATTN_INLINE __forceinline__
#endif
template <bool FUSE>
__device__ ER_ATTN_INLINE Cursor attention_phase(const DecodeParams& p, const Ring r, Cursor cur, int layer, int L, float* qs,
                                               float* sc, float* vred, float* red, int* s_flag, float* mg, uint32_t flag,
                                               unsigned long long* acc_y1, volatile int* zeroed, const int red_idx, const bool nosync,
                                               const bool prof_on, const int pb) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    AttnRange a;
    if (!attn_range(p.H, p.S, p.split_handicap, L, a)) return cur;
    const float cl2 = rsqrtf((float)HD) * 1.4426950408889634f;   // softmax scale in log2 units
    float* outp = p.part + ((size_t)a.h * p.S + (blockIdx.x % p.S)) * 100;
    const int nold = a.k1 - a.k0;
    const int nk = nold + a.is_new;
    const int nb = a.b1 - a.b0;
    const int new_slot = nb * 32;                            // score slot just past the old blocks
    const uint32_t qs_s = s_addr(qs), sc_s = s_addr(sc);
    __half* const q16s = reinterpret_cast<__half*>(red + 80);   // [96] the query as the fp16 tensor it is (A operand of the K pass)
    float M = -INFINITY;
    // the new key / value row (written by P1 of this step) are fetched NOW and used at the end of the K pass / at publish time, so
    // their L2 round trips hide behind the streamed passes
    uint4 knew = make_uint4(0, 0, 0, 0);
    unsigned short vnew = 0;
    if (FUSE) {
        // q / new k / new v of this head arrive as flagged words from the CTAs of the SAME head that own those qkv rows: threads 128..175
        // take the 48 q words, (last split only) lanes 0..11 of warp 0 the new key (4 words each) and threads 0..95 the new value (word tid / 2)
        if (nk > 0) {
            const bool qrole = tid >= 128 && tid < 128 + HD / 2;
            const unsigned long long* kq = p.ll_q + ((p.C + a.h * HD) >> 1) + tid * 4;
            const unsigned long long* vq = p.ll_q + ((2 * p.C + a.h * HD) >> 1) + (tid >> 1);
            const unsigned long long* qq = p.ll_q + ((a.h * HD) >> 1) + (tid - 128);
            uint32_t w[5] = {0u, 0u, 0u, 0u, 0u};
            const uint32_t mask = ((a.is_new && tid < HV) ? 0xFu : 0u) | (((a.is_new && tid < HD) || qrole) ? 0x10u : 0u);
            ll_poll<5>(w, nosync ? 0u : mask, flag, p.poll_rounds, [&](int j) { return j < 4 ? kq + j : (qrole ? qq : vq); });
            if (a.is_new && tid < HV) knew = make_uint4(w[0], w[1], w[2], w[3]);
            if (a.is_new && tid < HD) vnew = (unsigned short)((tid & 1) ? (w[4] >> 16) : (w[4] & 0xffffu));
            if (qrole) {
                const float2 f = h2f2(w[4]);
                qs[2 * (tid - 128)] = f.x; qs[2 * (tid - 128) + 1] = f.y;
                reinterpret_cast<uint32_t*>(q16s)[tid - 128] = w[4];
            }
        }
    } else if (a.is_new) {
        if (warp == 0 && lane < HV) knew = ldg_cg(reinterpret_cast<const uint4*>(p.q16 + p.C + a.h * HD) + lane);
        if (tid < HD) vnew = ldg_cg_u16(p.q16 + 2 * p.C + a.h * HD + tid);
    }
    if (FUSE) prof_stamp(p, pb + 5, prof_on);          // q (and the new k / v) polled
    if (nk > 0) {
        if (!FUSE && tid < HD) {
            const unsigned short qh = ldg_cg_u16(p.q16 + a.h * HD + tid);
            qs[tid] = __half2float(__ushort_as_half(qh));
            q16s[tid] = __ushort_as_half(qh);
        }
        cbar();
        // Both passes run on the tensor cores (mma.sync m16n8k16, fp16 operands, fp32 accumulate — the arithmetic of the flash kernel the
        // reference calls: S = Q K^T in fp32, P rounded to fp16 for P V).  K and V share one blocked cache layout
        // [key/32][d/8][key%32][8]; a stage holds 4 blocks = 128 keys and warp w owns keys 16 w .. 16 w + 15 of EVERY stage in both passes,
        // so the per-warp softmax maximum it normalises its scores with is the one it rescales its own P V partial with.
        const int t4 = lane & 3, g8 = lane >> 2;
        const uint32_t blk_off = (uint32_t)(warp >> 1) * kKBlockBytes;
        // ---- K pass: scores[key] = q . k.  B operand = two key groups x two 8-dim chunks per ldmatrix.x4 (rows = keys: conflict-free) ----
        float lmax = -INFINITY;
        {
            uint32_t qa[HD / 16][2];
            const uint32_t q16_s = s_addr(q16s);
#pragma unroll
            for (int s = 0; s < HD / 16; s++) { qa[s][0] = lds_u32(q16_s + (uint32_t)(16 * s + 2 * t4) * 2); qa[s][1] = lds_u32(q16_s + (uint32_t)(16 * s + 8 + 2 * t4) * 2); }
            const uint32_t k_off = blk_off + (uint32_t)((lane >> 3) & 1) * 512 + (uint32_t)((warp & 1) * 16 + (lane >> 4) * 8 + (lane & 7)) * 16;
            const int nch = (nb + 3) >> 2;
            for (int c = 0; c < nch; ++c, cur.advance(r.nstage)) {
                mbar_wait(r.fullb(cur.stage), cur.parity);
                if ((warp >> 1) < nb - c * 4) {
                    uint32_t b[HD / 16][4];
                    const uint32_t base = r.stage(cur.stage) + k_off;
#pragma unroll
                    for (int s = 0; s < HD / 16; s++) ldmatrix_x4(base + s * 1024, b[s][0], b[s][1], b[s][2], b[s][3]);
                    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};      // key group 0 / 1 of this warp's 16 keys
#pragma unroll
                    for (int s = 0; s < HD / 16; s++) {
                        mma_16816(d0, qa[s][0], qa[s][1], b[s][0], b[s][1]);
                        mma_16816(d1, qa[s][0], qa[s][1], b[s][2], b[s][3]);
                    }
                    // every row of D carries the same q: row 0 (lanes 0..3) holds keys 2 t4, 2 t4 + 1 of each group
                    const int sl = c * 128 + 16 * warp + 2 * t4;                 // slot of d0[0]; key = a.k0 + slot (k0 is block aligned)
                    const float s00 = (a.k0 + sl < a.k1) ? d0[0] : -INFINITY, s01 = (a.k0 + sl + 1 < a.k1) ? d0[1] : -INFINITY;   // slots >= L: stale data
                    const float s10 = (a.k0 + sl + 8 < a.k1) ? d1[0] : -INFINITY, s11 = (a.k0 + sl + 9 < a.k1) ? d1[1] : -INFINITY;
                    if (g8 == 0) {
                        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(sc_s + (uint32_t)sl * 4), "f"(s00), "f"(s01) : "memory");
                        asm volatile("st.shared.v2.f32 [%0], {%1,%2};" ::"r"(sc_s + (uint32_t)(sl + 8) * 4), "f"(s10), "f"(s11) : "memory");
                    }
                    lmax = fmaxf(fmaxf(lmax, fmaxf(s00, s01)), fmaxf(s10, s11));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(r.emptyb(cur.stage));
            }
        }
        // the new key: q . k_L; owned by warp 0
        if (a.is_new && warp == 0) {
            float acc = 0.f;
            if (lane < HV) acc = dot_k8(knew, lds_f4(qs_s + lane * 32), lds_f4(qs_s + lane * 32 + 16), 0.f);
            acc = warp_sum(acc);
            if (lane == 0) sts32(sc_s + (uint32_t)new_slot * 4, acc);
            lmax = fmaxf(lmax, acc);
        }
        // per-warp maximum, and this warp's keys normalised by it: p = exp2((s - m_w) * cl2)
        const float mw = warp_max(lmax);
        __syncwarp();
        for (int sl = (lane >> 4) * 128 + 16 * warp + (lane & 15); sl < nb * 32; sl += 256) {
            const uint32_t ad = sc_s + (uint32_t)sl * 4;
            sts32(ad, exp2f((lds32(ad) - mw) * cl2));                            // masked slots: exp2(-inf) = 0
        }
        if (a.is_new && warp == 0 && lane == 0) sts32(sc_s + (uint32_t)new_slot * 4, exp2f((lds32(sc_s + (uint32_t)new_slot * 4) - mw) * cl2));
        if (lane == 0) red[warp] = mw;
        cbar();
        // merge the warp maxima; fw rescales keys normalised by this warp (red[32] = warp 0's factor, used for the new key at publish time)
#pragma unroll
        for (int w = 0; w < kConsumerWarps; w++) M = fmaxf(M, red[w]);
        const float fw = (mw == -INFINITY) ? 0.f : exp2f((mw - M) * cl2);
        if (tid == 0) red[32] = fw;
        // ---- V pass: o[96] += p[key] * v[key][:]; A operand = p (fp16), B operand = ldmatrix.trans of (8 keys x 8 dims) patches ----
        float o[HV][4];
#pragma unroll
        for (int j = 0; j < HV; j++) { o[j][0] = 0.f; o[j][1] = 0.f; o[j][2] = 0.f; o[j][3] = 0.f; }
        float lsum = 0.f;
        {
            const uint32_t v_off = blk_off + (uint32_t)(lane >> 4) * 512 + (uint32_t)((warp & 1) * 16 + ((lane >> 3) & 1) * 8 + (lane & 7)) * 16;
            const int nch = (nb + 3) >> 2;
            for (int c = 0; c < nch; ++c, cur.advance(r.nstage)) {
                mbar_wait(r.fullb(cur.stage), cur.parity);
                const int sl0 = c * 128 + 16 * warp;
                if (sl0 < nold) {
                    const uint32_t pa = sc_s + (uint32_t)(sl0 + 2 * t4) * 4;
                    float p0, p1, p2, p3;
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(p0), "=f"(p1) : "r"(pa));
                    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(p2), "=f"(p3) : "r"(pa + 32));
                    p0 *= fw; p1 *= fw; p2 *= fw; p3 *= fw;                      // slots in [nold, block end) hold exp2(-inf) = 0
                    if (g8 == 0) lsum += (p0 + p1) + (p2 + p3);
                    const __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
                    const uint32_t a_lo = *reinterpret_cast<const uint32_t*>(&h01), a_hi = *reinterpret_cast<const uint32_t*>(&h23);
                    const uint32_t base = r.stage(cur.stage) + v_off;
                    const int lim = nold - sl0;                                  // keys of this group that exist (>= 16: all)
#pragma unroll
                    for (int j2 = 0; j2 < HV / 2; j2++) {
                        uint32_t b0, b1, b2, b3;
                        ldmatrix_x4_trans(base + j2 * 1024, b0, b1, b2, b3);
                        if (lim < 16) {
                            // rows of the cache beyond the last key were never written (they may hold NaN bit patterns): 0 x NaN must not reach o
                            const uint32_t m_lo = (2 * t4 < lim ? 0x0000ffffu : 0u) | (2 * t4 + 1 < lim ? 0xffff0000u : 0u);
                            const uint32_t m_hi = (2 * t4 + 8 < lim ? 0x0000ffffu : 0u) | (2 * t4 + 9 < lim ? 0xffff0000u : 0u);
                            b0 &= m_lo; b2 &= m_lo; b1 &= m_hi; b3 &= m_hi;
                        }
                        mma_16816(o[2 * j2], a_lo, a_hi, b0, b1);
                        mma_16816(o[2 * j2 + 1], a_lo, a_hi, b2, b3);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(r.emptyb(cur.stage));
            }
        }
        if (g8 == 0) {
#pragma unroll
            for (int j = 0; j < HV; j++) { vred[warp * HD + 8 * j + 2 * t4] = o[j][0]; vred[warp * HD + 8 * j + 2 * t4 + 1] = o[j][1]; }
        }
        lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
        lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
        if (lane == 0) red[64 + warp] = lsum;
        cbar();
    }
    if (FUSE) prof_stamp(p, pb + 6, prof_on);          // K / V passes done
    if (warp >= 4 && !FUSE) return cur;           // warps 0..3 publish; the others go on to the grid barrier
    // ---- publish the split partial ----
    unsigned long long* outl = p.ll_part + ((size_t)a.h * p.S + (blockIdx.x % p.S)) * 100;
    if (tid < HD) {
        float acc = 0.f;
        if (nk > 0) {
#pragma unroll
            for (int g = 0; g < kConsumerWarps; g++) acc += vred[g * HD + tid];
            if (a.is_new) acc = fmaf(sc[new_slot] * red[32], __half2float(__ushort_as_half(vnew)), acc);   // new key: normalised by warp 0
        }
        if (FUSE) ll_store(outl + tid, __float_as_uint(acc), flag); else outp[tid] = acc;
    } else if (tid == HD) {
        float l = 0.f;
        if (nk > 0) {
            for (int g = 0; g < kConsumerWarps; g++) l += red[64 + g];
            if (a.is_new) l += sc[new_slot] * red[32];
        }
        const float mval = (nk > 0) ? M * rsqrtf((float)HD) : -INFINITY;        // max in softmax (scaled) units
        if (FUSE) { ll_store(outl + HD, __float_as_uint(mval), flag); ll_store(outl + HD + 1, __float_as_uint(l), flag); }
        else { outp[HD] = mval; outp[HD + 1] = l; }
    }
    if (FUSE) {
        // ---- tensor-parallel tail: every CTA of the head merges the S partials itself (flagged words, S pollers per word: no storm),
        // then multiplies the head's output by ITS rows of the head's out_proj columns and adds the partial rows to the y1 accumulator ----
        const int S = p.S, s = blockIdx.x % S;
        float* pm = mg;                                   // [S][100] merged-partial staging (aliases the GEMV partials)
        cbar();
        {
            const unsigned long long* src = p.ll_part + (size_t)a.h * S * 100;
            const int nw = S * (HD + 2);
            uint32_t w[7] = {0u, 0u, 0u, 0u, 0u, 0u, 0u};
            uint32_t mask = 0;
#pragma unroll
            for (int j = 0; j < 7; j++) mask |= (tid + kConsumers * j < nw) ? (1u << j) : 0u;
            const uint32_t want = mask;
            ll_poll<7>(w, nosync ? 0u : mask, flag, p.poll_rounds, [&](int j) { const int i = tid + kConsumers * j; return src + (i / (HD + 2)) * 100 + i % (HD + 2); });
#pragma unroll
            for (int j = 0; j < 7; j++)
                if ((want >> j) & 1u) { const int i = tid + kConsumers * j; pm[(i / (HD + 2)) * 100 + i % (HD + 2)] = __uint_as_float(w[j]); }
        }
        cbar();
        __half* o16 = reinterpret_cast<__half*>(qs);      // the head's attention output, fp16 as in the reference; q is no longer needed
        if (tid < HD) {                                   // same fold as the single-merger path, split order
            float Mx = -INFINITY;
            for (int q = 0; q < S; q++) Mx = fmaxf(Mx, pm[q * 100 + HD]);
            float den = 0.f, num = 0.f;
            for (int q = 0; q < S; q++) {
                const float ms = pm[q * 100 + HD];
                const float w8 = (ms == -INFINITY) ? 0.f : __expf(ms - Mx);
                den = fmaf(w8, pm[q * 100 + HD + 1], den);
                num = fmaf(w8, pm[q * 100 + tid], num);
            }
            o16[tid] = __float2half_rn(num / den);
        }
        prof_stamp(p, pb + 15, prof_on);                  // partials of the S splits polled and merged
        if (tid == 0) {                                   // the janitor has recycled every copy this CTA has been told about (see fix_add_cnt)
            int spins = 0;
            while (*zeroed < red_idx) { if (++spins > kSpinLimit) asm volatile("trap;"); }
        }
        cbar();
        const RowRange rs = cta_rows_of(p.C, (unsigned)s, (unsigned)S);
        const int n = rs.r1 - rs.r0;
        const uint32_t out_s = s_addr(vred);              // n <= C / 3 <= 512 floats
        cur = outproj_job_mma(r, cur, n, s_addr(o16), out_s);
        prof_stamp(p, pb + 12, prof_on);                  // out_proj rows computed (atomics follow)
        cbar();
        for (int u = tid; u < n; u += kConsumers) fix_add_cnt(acc_y1 + (size_t)(rs.r0 + u) * kAccStride, vred[u]);
        return cur;
    }
    // ---- last split of this head to finish merges the S partials (release/acquire ticket on a monotonic counter) ----
    pbar();
    if (tid == 0) {
        unsigned ticket;
        asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(p.head_cnt + a.h) : "memory");
        *s_flag = ((ticket % (unsigned)p.S) == (unsigned)p.S - 1) ? 1 : 0;
    }
    pbar();
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
        float Mx = -INFINITY;
#pragma unroll
        for (int s = 0; s < 16; s++) Mx = fmaxf(Mx, ms[s]);
        float den = 0.f, num = 0.f;
#pragma unroll
        for (int s = 0; s < 16; s++) {
            const float w = (ms[s] == -INFINITY) ? 0.f : __expf(ms[s] - Mx);
            den = fmaf(w, ls[s], den);
            num = fmaf(w, ov[s], num);
        }
        const __half av = __float2half_rn(num / den);
        for (int c = 0; c < p.xrep; c++) p.attn16[(size_t)c * p.C + a.h * HD + tid] = av;
    }
    return cur;
}

// ---- the kernel ---------------------------------------------------------------------------------------------------------------------
// PROF: the timeline instrumentation is a separate instantiation so that the production kernel carries neither its registers nor its branches
// FUSE: the tensor-parallel layer (see above); false = the five-exchange layer (grid barrier between dependent phases)
template <bool PROF, bool FUSE>
__global__ void __launch_bounds__(kThreads, 1) decode_persistent_kernel(const __grid_constant__ DecodeParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int C = p.C, F = p.F, H = p.H, V = p.V;
    // smem carve-up: ring first (128-byte aligned stages), then the small arrays
    Ring ring;
    ring.data = s_addr(smem_raw);
    ring.nstage = p.nstage;
    unsigned char* q = smem_raw + (size_t)p.nstage * kStageBytes;
    ring.full = s_addr(q);
    ring.empty = ring.full + kMaxStages * 8;
    float* xres = reinterpret_cast<float*>(q + 2 * kMaxStages * 8);   // [C]   residual stream (fp32)
    float* part = xres + C;                                           // [32][kPartStride] GEMV lane partials / merge staging / fc2 partial sums
    float* red = part + 32 * kPartStride;                             // [128] block-reduction scratch
    float* qs = red + 128;                                            // [96] query of this CTA's head (fp32); later the head's output (fp16) / the fc1 slice
    float* vred = qs + HD;                                            // [16*96] V-pass cross-warp reduction / out_proj partial rows
    float* sc = vred + 2 * kConsumerWarps * HD;                       // [sc_len] scores / sampler scratch
    __half* xin = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(sc + p.sc_len) + 15) & ~uintptr_t(15));   // [max(C,F)] GEMV input (fp16)
    __shared__ int s_tok;
    __shared__ int s_stop;
    __shared__ uint32_t s_cons_it;
    __shared__ int s_flag;
    __shared__ int s_tok_done;   // tokens of this launch whose token-end grid barrier the consumers have passed
    __shared__ int s_rows[10];  // this CTA's row ranges of the qkv / C / F / V phases (+ out_proj slice); computed once (registers are scarce)
    __shared__ SnapState s_snap; // ONE snapshot of the device state per CTA: producer, run-ahead warp and consumers must agree on `done`
    __shared__ unsigned long long s_issued;   // bytes the ring producer has requested so far (read by the L2 run-ahead warp)
    __shared__ int s_red_done;   // FUSE: L2 reductions this CTA has seen complete
    __shared__ int s_zeroed;     // FUSE: ... and whose recycled accumulator slice the janitor has zeroed (and fenced)
    __shared__ int s_exit;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool nosync = p.dbg_nosync != 0;
    if (tid == 0) {
        for (int i = 0; i < p.nstage; i++) { mbar_init(ring.fullb(i), 1); mbar_init(ring.emptyb(i), kConsumerWarps); }
        s_stop = 0;
        s_cons_it = 0;
        s_tok_done = 0;
        s_issued = 0ull;
        s_red_done = 0; s_zeroed = 0; s_exit = 0;
        s_snap.t = p.st->t; s_snap.L = p.st->L; s_snap.counter = p.st->counter; s_snap.last_tok = p.st->last_tok; s_snap.done = p.st->done;
        const RowRange r3 = p1_rows<FUSE>(p), r1 = cta_rows(C), rf = cta_rows(F), rv = cta_rows(V);
        s_rows[0] = r3.r0; s_rows[1] = r3.r1; s_rows[2] = r1.r0; s_rows[3] = r1.r1; s_rows[4] = rf.r0; s_rows[5] = rf.r1; s_rows[6] = rv.r0; s_rows[7] = rv.r1;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (FUSE) {
        // the fc2 consumer reads whole 8-column stages through ldmatrix: a partially filled last stage must not hold NaN bit patterns left
        // behind by an earlier kernel (0 x NaN = NaN); stale finite weights are harmless (their h is zero)
        for (uint32_t o = tid * 16; o < (uint32_t)p.nstage * kStageBytes; o += kThreads * 16) sts128(ring.data + o, make_uint4(0, 0, 0, 0));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == kConsumerWarps) {
        // ===== producer warp: one elected lane streams this CTA's byte ranges through the ring =====
        if (lane == 0) producer_loop<FUSE>(p, ring, &s_snap, &s_stop, &s_cons_it, &s_tok_done, &s_issued);
    } else if (warp == kConsumerWarps + 1) {
        // ===== L2 run-ahead warp =====
        if (lane == 0 && p.pf_dist > 0) prefetch_loop<FUSE>(p, &s_snap, &s_stop, &s_issued);
    } else if (warp == kConsumerWarps + 2) {
        // ===== accumulator janitor (FUSE): zero this CTA's slice of the copy that reduction k + 2 will use once reduction k is complete =====
        if (FUSE && !s_snap.done) {
            const RowRange rz = cta_rows(C);
            const int zn = rz.r1 - rz.r0;                      // <= 32 for every supported shape (host-checked)
            int seen = 0;
            for (;;) {
                const int k = *(volatile int*)&s_red_done;
                if (k > seen) {
                    for (int j = seen; j < k; ++j)
                        if (lane < zn) p.acc[((size_t)((j + 2) & 3) * C + rz.r0 + lane) * kAccStride] = 0ull;
                    __threadfence();                           // the zeros are performed at L2 before anybody is told
                    __syncwarp();
                    if (lane == 0) *(volatile int*)&s_zeroed = k;
                    seen = k;
                } else {
                    if (*(volatile int*)&s_exit) break;
                    __nanosleep(200);
                }
            }
        }
    } else {
        // ===== consumers =====
        unsigned epoch = 0;
        Cursor cur{0u, 0u, 0u};
        const uint32_t xin_s = s_addr(xin), xres_s = s_addr(xres), part_s = s_addr(part);
        const float inv_c = 1.0f / (float)C;
        int t = s_snap.t, L = s_snap.L, counter = s_snap.counter, last_tok = s_snap.last_tok;
        const bool done0 = s_snap.done != 0;
        bool state_written = done0;
        const int nu_fc2 = (F / C) * kUnitDiv;      // units per fc2 row
        const int nu1 = kUnitDiv;                   // units per row of the C-wide phases
        const int nparts = p.use_mma ? kConsumerWarps : 32;
        const GemvCfg gc{p.C, p.ustride, p.upstage, p.use_mma};
        const int xcopy = (int)(blockIdx.x % (unsigned)p.xrep);   // which replica of the exchanged vectors this CTA reads
        for (int iter = 0; iter < p.steps && !done0; ++iter, ++t) {
            const bool prof_on = PROF && p.prof != nullptr && t == p.prof_token && (int)blockIdx.x == p.prof_cta;
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
                const uint32_t flag = (uint32_t)(t * p.layers + layer + 1);   // exchange-word flag of this (token, layer)
                const bool all_on = PROF && p.prof != nullptr && t == p.prof_token && layer == 5;
                const int gl = iter * p.layers + layer;                          // layers run by this launch so far
                // FUSE: reduction 2 gl (y1) uses accumulator copy (2 gl) & 3, reduction 2 gl + 1 (y2) copy (2 gl + 1) & 3
                unsigned long long* const acc_y1 = FUSE ? p.acc + (size_t)((2 * gl) & 3) * C * kAccStride : nullptr;
                unsigned long long* const acc_y2 = FUSE ? p.acc + (size_t)((2 * gl + 1) & 3) * C * kAccStride : nullptr;
                // ---------------- P1: q,k,v = x16 @ Wqkv^T + b ; KV append in place ----------------------------------------------
                {
                    const RowRange rr{s_rows[0], s_rows[1]};
                    const int nr = rr.r1 - rr.r0;
                    prof_stamp(p, pb + 0, prof_on); prof_all(p, 0, all_on);
                    const int nu = nr * nu1;
                    const bool own = tid < nu && (tid % nu1) == 0;
                    const float bias = own ? __half2float(p.bqkv[(size_t)layer * 3 * C + rr.r0 + tid / nu1]) : 0.f;   // in flight during the GEMV
                    if (nu > 0 || !FUSE) cur = gemv_job(ring, cur, nu, nu1, gc, xin_s, part_s);
                    prof_stamp(p, pb + 1, prof_on); prof_all(p, 1, all_on);
                    // q, new k, new v go to a small dense vector the attention CTAs read (q16[3C]) — or, FUSE, as flagged words to the CTAs of the
                    // same head; the K/V CACHE rows (scattered 2-byte stores into pages only this token touches: TLB misses that hold the storing
                    // warp for ~1 us) are only needed by later tokens and are written after the values are on their way
                    __half hv = __float2half_rn(0.f);
                    const int r = rr.r0 + tid / nu1;
                    if (warp * 32 < nu) {
                        const float sum = reduce_rows(part_s, nu, nu1, nparts);
                        hv = __float2half_rn(sum + bias);
                        if (FUSE) ll_publish_rows(p.ll_q, rr.r0, nu, nu1, hv, flag);
                        else if (own) p.q16[r] = hv;
                    }
                    prof_stamp(p, pb + 2, prof_on); prof_all(p, 2, all_on);
                    if (!FUSE) grid_arrive(p.bar, epoch, nosync);
                    if (own && r >= C) {
                        const int c = r < 2 * C ? r - C : r - 2 * C, h = c / HD, d = c % HD;     // K and V share the blocked layout
                        const size_t idx = ((((size_t)layer * H + h) * (size_t)p.nkb + (L >> 5)) * HV + (d >> 3)) * 256 + (size_t)(L & 31) * 8 + (d & 7);
                        (r < 2 * C ? p.kc : p.vc)[idx] = hv;
                    }
                    if (!FUSE) grid_wait(p.bar, epoch, nosync);
                }
                prof_stamp(p, pb + 3, prof_on); prof_all(p, 3, all_on);
                // ---------------- P2: attention (+ FUSE: head merge, row-parallel out_proj into the y1 accumulator) ------------------------
                cur = attention_phase<FUSE>(p, ring, cur, layer, L, qs, sc, vred, red, &s_flag, part, flag, acc_y1, &s_zeroed, 2 * gl, nosync, prof_on, pb);
                if (FUSE && (int)blockIdx.x >= H * p.S) {
                    // CTAs without an attention role still ADD (zero) to every y1 word: every reduction then has an addend from every CTA, which
                    // is what makes "reduction k+1 complete" imply "every CTA has recycled its slice for reduction k+2"
                    if (tid == 0) {
                        int spins = 0;
                        while (*(volatile int*)&s_zeroed < 2 * gl) { if (++spins > kSpinLimit) asm volatile("trap;"); }
                    }
                    cbar();
                    for (int u = tid; u < C; u += kConsumers) fix_add_cnt(acc_y1 + (size_t)u * kAccStride, 0.f);
                }
                prof_stamp(p, pb + 4, prof_on); prof_all(p, 4, all_on);
                if (!FUSE) grid_barrier(p.bar, epoch, nosync);
                if (!FUSE) prof_stamp(p, pb + 5, prof_on);
                prof_all(p, 5, all_on);
                // ---------------- P3: out_proj on the merged attention output -----------------------------------------------------------
                if (!FUSE) {
                    for (int i = tid; i < C / 8; i += kConsumers)
                        reinterpret_cast<uint4*>(xin)[i] = ldg_cg(reinterpret_cast<const uint4*>(p.attn16 + (size_t)xcopy * C) + i);
                    cbar();
                    const RowRange rr{s_rows[2], s_rows[3]};
                    const int nr = rr.r1 - rr.r0;
                    prof_stamp(p, pb + 6, prof_on); prof_all(p, 6, all_on);
                    const int nu = nr * nu1;
                    const bool own = tid < nu && (tid % nu1) == 0;
                    const float bias = own ? __half2float(p.bo[(size_t)layer * C + rr.r0 + tid / nu1]) : 0.f;
                    cur = gemv_job(ring, cur, nu, nu1, gc, xin_s, part_s);
                    if (warp * 32 < nu) {
                        const float sum = reduce_rows(part_s, nu, nu1, nparts);
                        if (own) { const __half yv = __float2half_rn(sum + bias); for (int c = 0; c < p.xrep; c++) p.y1[(size_t)c * C + rr.r0 + tid / nu1] = yv; }
                    }
                }
                const LnParams lp1 = ln_load(p.ln1_w + (size_t)layer * C, p.ln1_b + (size_t)layer * C, C);   // lands while we wait
                prof_stamp(p, pb + 7, prof_on); prof_all(p, 7, all_on);
                if (!FUSE) grid_barrier(p.bar, epoch, nosync);
                prof_stamp(p, pb + 8, prof_on); prof_all(p, 8, all_on);
                // ---------------- P4: x = LN1(x + y1) ; h1 = relu(fc1(x)) -----------------------------------------------------------
                {
                    residual_layer_norm(xres_s, xin_s, p.y1 + (size_t)xcopy * C, layer == 0, lp1, C, inv_c, red, acc_y1, H + (int)gridDim.x - H * p.S, FUSE ? p.bo + (size_t)layer * C : nullptr, nosync);
                    if (FUSE && tid == 0) *(volatile int*)&s_red_done = 2 * gl + 1;
                    const RowRange rr{s_rows[4], s_rows[5]};
                    const int nr = rr.r1 - rr.r0;
                    prof_stamp(p, pb + 9, prof_on); prof_all(p, 9, all_on);
                    const int nu = nr * nu1;
                    const bool own = tid < nu && (tid % nu1) == 0;
                    const float bias = own ? __half2float(p.b1[(size_t)layer * F + rr.r0 + tid / nu1]) : 0.f;
                    cur = gemv_job(ring, cur, nu, nu1, gc, xin_s, part_s);
                    __half* const h1s = reinterpret_cast<__half*>(qs);      // FUSE: this CTA's slice of h1 stays on chip (<= 64 fp16, zero padded)
                    if (FUSE && tid >= nu && tid < kMaxUnits) h1s[tid] = __float2half_rn(0.f);
                    if (warp * 32 < nu) {
                        const float sum = reduce_rows(part_s, nu, nu1, nparts);
                        const __half hv = __float2half_rn(fmaxf(round_f16(sum + bias), 0.f));
                        if (FUSE) { if (own) h1s[tid] = hv; }
                        else if (own) for (int c = 0; c < p.xrep; c++) p.h1[(size_t)c * F + rr.r0 + tid / nu1] = hv;
                    }
                    if (FUSE) {
                        prof_stamp(p, pb + 11, prof_on);      // fc1 done
                        // fc2, row-parallel: y2 += W2[:, j] * h1[j] over this CTA's columns j, then into the counting accumulator
                        if (tid == 0) {
                            int spins = 0;
                            while (*(volatile int*)&s_zeroed < 2 * gl + 1) { if (++spins > kSpinLimit) asm volatile("trap;"); }
                        }
                        cbar();
                        cur = fc2_job(ring, cur, nr, C, p.ustride, s_addr(h1s), part_s);
                        prof_stamp(p, pb + 13, prof_on);      // fc2 partial sums computed (atomics follow)
                        cbar();
                        if (p.red_group == 4) {
                            // (experiment, er_debug_set red_group4 = 1; off by default) pre-reduce among 4 CTAs through flagged words, then add: each CTA
                            // sends the three quarters it does not own to their owners, sums the three it receives with its own (fixed order) and issues
                            // C/4 atomics: 4x fewer atomics per element (37 addends).  Measured: the extra hop costs 3 us and the wait for the completed
                            // reduction shrinks by only 2 us — what a reduction costs is the all-to-all hop (skew of 148 CTAs + atomics in flight + poll),
                            // not the atomic count (profiles/r02_phase_timeline_tensor_parallel_v4_group4.json)
                            const int me = (int)(blockIdx.x & 3u), g4 = (int)(blockIdx.x >> 2), Q = C >> 2;
                            for (int i = tid; i < 3 * Q; i += kConsumers) {
                                const int q = (me + 1 + i / Q) & 3, e = i % Q;
                                ll_store(p.xq + ((size_t)(4 * g4 + q) * 4 + me) * Q + e, __float_as_uint(part[q * Q + e]), flag);
                            }
                            for (int e0 = 0; e0 < Q; e0 += kConsumers) {
                                const int e = e0 + tid;
                                uint32_t w[3] = {0u, 0u, 0u};
                                ll_poll<3>(w, (e < Q && !nosync) ? 7u : 0u, flag, p.poll_rounds,
                                           [&](int j) { return p.xq + ((size_t)blockIdx.x * 4 + ((me + 1 + j) & 3)) * Q + e; });
                                if (e < Q) fix_add_cnt(acc_y2 + (size_t)(me * Q + e) * kAccStride, ((part[me * Q + e] + __uint_as_float(w[0])) + __uint_as_float(w[1])) + __uint_as_float(w[2]));
                            }
                        } else {
                            for (int u = tid; u < C; u += kConsumers) fix_add_cnt(acc_y2 + (size_t)u * kAccStride, part[u]);
                        }
                    }
                }
                prof_stamp(p, pb + 10, prof_on); prof_all(p, 10, all_on);
                if (!FUSE) grid_barrier(p.bar, epoch, nosync);
                if (!FUSE) prof_stamp(p, pb + 11, prof_on);
                prof_all(p, 11, all_on);
                // ---------------- P5: y2 = fc2(h1) -----------------------------------------------------------------------------------------
                if (!FUSE) {
                    for (int i = tid; i < F / 8; i += kConsumers)
                        reinterpret_cast<uint4*>(xin)[i] = ldg_cg(reinterpret_cast<const uint4*>(p.h1 + (size_t)xcopy * F) + i);
                    cbar();
                    const RowRange rr{s_rows[2], s_rows[3]};
                    const int nr = rr.r1 - rr.r0, nu = nr * nu_fc2;
                    prof_stamp(p, pb + 12, prof_on); prof_all(p, 12, all_on);
                    const float bias = (tid < nu && (tid % nu_fc2) == 0) ? __half2float(p.b2[(size_t)layer * C + rr.r0 + tid / nu_fc2]) : 0.f;
                    cur = gemv_job(ring, cur, nu, nu_fc2, gc, xin_s, part_s);
                    if (warp * 32 < nu) {
                        const float sum = reduce_rows(part_s, nu, nu_fc2, nparts);
                        if (tid < nu && (tid % nu_fc2) == 0) { const __half yv = __float2half_rn(sum + bias); for (int c = 0; c < p.xrep; c++) p.y2[(size_t)c * C + rr.r0 + tid / nu_fc2] = yv; }
                    }
                }
                const LnParams lp2 = ln_load(p.ln2_w + (size_t)layer * C, p.ln2_b + (size_t)layer * C, C);
                if (!FUSE) prof_stamp(p, pb + 13, prof_on);
                prof_all(p, 13, all_on);
                if (!FUSE) grid_barrier(p.bar, epoch, nosync);
                prof_stamp(p, pb + 14, prof_on); prof_all(p, 14, all_on);
                // ---------------- x = LN2(x + y2) ------------------------------------------------------------------------------------------
                residual_layer_norm(xres_s, xin_s, p.y2 + (size_t)xcopy * C, false, lp2, C, inv_c, red, acc_y2, (int)gridDim.x / p.red_group, FUSE ? p.b2 + (size_t)layer * C : nullptr, nosync);
                if (FUSE && tid == 0) *(volatile int*)&s_red_done = 2 * gl + 2;
            }
            // ================= lm_head: logits_pre = fp16(x) @ W^T (fp32 value before the fp16 store) ================
            {
                const RowRange rr{s_rows[6], s_rows[7]};
                const int nr = rr.r1 - rr.r0;
                const int nu = nr * nu1;
                cur = gemv_job(ring, cur, nu, nu1, gc, xin_s, part_s);
                if (warp * 32 < nu) {
                    const float sum = reduce_rows(part_s, nu, nu1, nparts);
                    if (tid < nu && (tid % nu1) == 0) p.logits[rr.r0 + tid / nu1] = sum;
                }
            }
            L += 1;
            prof_stamp(p, 1 + 16 * p.layers, prof_on);
            grid_barrier(p.bar, epoch, nosync);
            if (tid == 0) { __threadfence_block(); *(volatile int*)&s_tok_done = iter + 1; }
            prof_stamp(p, 2 + 16 * p.layers, prof_on);
        }
        if (!state_written && blockIdx.x == 0 && tid == 0) { p.st->t = t; p.st->L = L; p.st->counter = counter; p.st->last_tok = last_tok; }
        if (tid == 0) *(volatile int*)&s_exit = 1;
    }
    __syncthreads();   // nobody leaves while bulk copies into this CTA's shared memory may still be in flight
}

}  // namespace er

// ---- host launcher -------------------------------------------------------------------------------------------------------------------------
size_t er_decode_small_smem_bytes(const er::DecodeParams& p) {
    const int C = p.C, F = p.F;
    size_t fl = (size_t)C + 32 * er::kPartStride + 128 + er::HD + 2 * er::kConsumerWarps * er::HD + (size_t)p.sc_len;
    return 2 * er::kMaxStages * 8 + fl * 4 + (size_t)(F > C ? F : C) * 2 + 128;
}
size_t er_decode_smem_bytes(const er::DecodeParams& p) { return (size_t)p.nstage * er::kStageBytes + er_decode_small_smem_bytes(p); }
int er_decode_pick_stages(const er::DecodeParams& p, size_t smem_limit) {
    const size_t small = er_decode_small_smem_bytes(p);
    if (smem_limit < small + 2 * (size_t)er::kStageBytes) return 0;
    int n = (int)((smem_limit - small) / er::kStageBytes);
    return n > er::kMaxStages ? er::kMaxStages : n;
}
int er_decode_max_units() { return er::kMaxUnits; }
int er_decode_stage_bytes() { return er::kStageBytes; }

static const void* er_decode_kernel_fn(bool prof, bool fuse) {
    if (fuse) return prof ? (const void*)er::decode_persistent_kernel<true, true> : (const void*)er::decode_persistent_kernel<false, true>;
    return prof ? (const void*)er::decode_persistent_kernel<true, false> : (const void*)er::decode_persistent_kernel<false, false>;
}
cudaError_t er_decode_launch(const er::DecodeParams& p, int grid, size_t smem, cudaStream_t stream) {
    const void* fn = er_decode_kernel_fn(p.prof != nullptr, p.use_fuse != 0);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&p};
    return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(er::kThreads), args, smem, stream);
}

// diagnostics: resource usage / occupancy of the four instantiations as the driver sees them
extern "C" int er_debug_decode_report(char* buf, int n, unsigned long long smem) {
    int off = 0;
    for (int v = 0; v < 4; ++v) {
        const void* fn = er_decode_kernel_fn((v & 1) != 0, (v & 2) != 0);
        cudaFuncAttributes a{};
        const cudaError_t e0 = cudaFuncGetAttributes(&a, fn);
        const cudaError_t e1 = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int per = -1;
        const cudaError_t e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fn, er::kThreads, smem);
        off += snprintf(buf + off, n - off, "prof=%d fuse=%d: regs %d static_smem %zu local %zu maxThreads %d | getattr %d setattr %d occ %d -> blocks/SM %d (threads %d, dyn smem %llu)\n",
                        v & 1, (v >> 1) & 1, a.numRegs, a.sharedSizeBytes, a.localSizeBytes, a.maxThreadsPerBlock, (int)e0, (int)e1, (int)e2, per, er::kThreads, smem);
        cudaGetLastError();
    }
    return off;
}

int er_decode_max_grid(size_t smem) {
    int dev = 0, sms = 0, per = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int ok = 1;
    for (int v = 0; v < 4; ++v) {
        const void* fn = er_decode_kernel_fn((v & 1) != 0, (v & 2) != 0);
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, fn, er::kThreads, smem) != cudaSuccess || per < 1) ok = 0;
    }
    return ok ? sms : 0;
}
