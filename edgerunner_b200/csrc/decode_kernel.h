// Parameter block of the persistent decode kernel (see decode_kernel.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace er {

struct DecodeState {   // lives in device memory; carries the token loop across launches
    int t;             // tokens generated so far (== FSM idx)
    int L;             // rows in the KV cache (position of the next token)
    int counter;       // FSM: coordinate tokens still owed
    int last_tok;      // last token fed to the model
    int done;          // EOS seen or max_new reached
    int pad[3];
};

struct DecodeParams {
    // dimensions
    int C, H, F, V, layers;
    int S;            // KV splits per head in the attention phase (H*S <= grid); tensor-parallel layer: S in {12, 9, 6}, also the qkv row split of a head
    int Lmax;         // V-cache rows per head
    int nkb;          // K-cache 32-key blocks per head (ceil(Lmax/32))
    int nstage;       // shared-memory ring stages (24 KB each)
    int sc_len;       // floats of score scratch (>= max keys per split + 33, >= V)
    // decoder weights, fp16, nn.Linear layout [out][in] (q,k,v rows concatenated in that order)
    const __half *wqkv, *bqkv, *wo, *bo, *ln1_w, *ln1_b, *w1, *b1, *w2, *b2, *ln2_w, *ln2_b;
    const __half *lm_head, *embd, *pos;
    // the same decoder weights re-packed for the decode stream: units of C fp16 padded to `ustride` (see decode_kernel.cu)
    const __half *wdec; int ustride, upstage, use_mma;
    int split_handicap;   // K blocks the last KV split gives up (it also owns the new key and usually merges the head)
    // KV cache: K and V both blocked [layer][head][key/32][d/8][key%32][8]
    __half *kc, *vc;
    // cross-CTA scratch (global, read back with ld.cg)
    int xrep;   // replicas of attn16 / y1 / h1 / y2 (writers store all, CTA b reads replica b % xrep): spreads 148 readers over L2 slices
    __half *q16, *y1, *h1, *y2, *attn16;   // q16: [3C] = q | new k | new v of the current token
    unsigned *head_cnt;   // [H] monotonic split-completion tickets (zeroed with the barrier counter)
    float *part;      // [H][S][100]: o[96], m, l
    float *logits;    // [V] fp32 lm_head output before the fp16 rounding
    // flagged exchange (tensor-parallel layer): 8-byte {payload, flag} words that the S CTAs of a head poll instead of meeting at a
    // barrier; ll_q: [3C/2] two fp16 each (q | new k | new v), ll_part: [H][S][100] one fp32 each (split partials).  Zeroed by the
    // host before each launch; flags are unique per (token, layer).
    unsigned long long *ll_q, *ll_part;
    int poll_rounds;   // all-thread polling rounds before a warp falls back to one spinning lane
    DecodeState *st;
    unsigned *bar;    // grid barrier counter (zeroed by the host before each launch)
    // io
    int32_t *out_ids;        // [max_new]
    float *out_logits;       // optional [max_new][V]
    const int32_t *forced;   // optional teacher-forcing stream [max_new]
    int max_new, steps, mode /*0 greedy, 1 sample*/, top_k, use_fsm, eos;
    unsigned long long seed;
    // optional phase timeline: slot 0 = token start, then (end of phase, end of barrier) x 5 per layer, + lm_head pair
    unsigned long long *prof; int prof_token, prof_cta;
    // tensor-parallel layer (use_fuse): per layer [H][C][HD+8] per-head out_proj units, then [F][ustride] transposed fc2 units;
    // acc: four copies of [C] u64 counting fixed-point accumulators (reduction k uses copy k & 3), zeroed by the host before each launch
    const __half *wfuse; unsigned long long *acc; int use_fuse;
    // fc2 reduction: CTAs pre-reduce in groups of red_group (4 or 1) through the flagged words xq [grid][4][C/4] before the atomics
    unsigned long long *xq; int red_group;
    // L2 run-ahead: a second streaming warp issues cp.async.bulk.prefetch.L2 for this CTA's future ring bytes, staying at most
    // pf_dist bytes ahead of the ring producer, so that HBM keeps streaming while the consumers sit in an exchange (0 = off)
    int pf_dist;
    // diagnostics only (er_debug_set): skip the grid barriers (results are garbage; shows the streaming / consumption rate alone)
    int dbg_nosync;
};

}  // namespace er

size_t er_decode_smem_bytes(const er::DecodeParams& p);
int er_decode_pick_stages(const er::DecodeParams& p, size_t smem_limit);
cudaError_t er_decode_launch(const er::DecodeParams& p, int grid, size_t smem, cudaStream_t stream);
int er_decode_max_grid(size_t smem);
int er_decode_stage_bytes();
