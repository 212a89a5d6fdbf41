// Training step of LMM: forward in training mode + backward, on the engine's own fp16 weights.
//
// Reference: main.py:160-172 (`out = model(data); accelerator.backward(out['loss'])`) over models.py:147-202 (LMM.forward), modeling_opt.py:253-298
// (post-LN decoder layer with F.dropout(p = config.dropout) on both branches), :464-517 (lm_head + shifted cross-entropy), point.py:186-206 (the point
// encoder, always trained in cond_mode 'point': models.py:54), with `opt.checkpointing = True` (options.py:126: every layer is re-run in the backward
// pass).  The reference autocasts to bf16; this engine computes in fp16 (weights are the engine's fp16 copies, activations fp16, residual stream /
// LayerNorm / softmax / loss fp32) with a static loss scale on the fp16 activation gradients; weight gradients are accumulated in fp32 by the
// tensor-core GEMM and exported unscaled.
//
// Structure (B200-first, not autograd):
//   * what a layer's backward needs (LayerAct below) is KEPT from the forward pass — 1.64 GB per layer at 4 x 10 243 rows, 39 GB for 24 layers, next to
//     the fp32 layer inputs; when device memory is short (or on er_debug_set("train_recompute", 1)) only the fp32 layer inputs are kept and each layer
//     is re-run in the backward pass as the reference does.  Both modes are bit-identical.
//   * per Linear one dgrad GEMM (weight transposed on the fly: 2..19 MB) and one wgrad GEMM (both activations transposed to K-major, zero-padded to a
//     multiple of 64 rows: er_transpose_f16) on the tcgen05 kernel of gemm_tcgen05.cu, plus the row kernels of backward.cu and the flash-attention
//     backward of attention_bwd_mma.cu (its softmax statistic is written by the forward attention kernel: AttnArgs.lse2).
//   * everything is deterministic: no float atomics anywhere, fixed-order two-stage column sums, the dropout mask a counter-based function of
//     (seed, site, element).
//
// Gradients are produced for: every decoder layer, lm_head, embd, embed_positions, proj_cond, norm_cond, embed_num_face and — with train_encoder
// (opt.freeze_encoder = False: the ArAE preset and the reference's only setting in 'point' mode) — the point encoder and the KL term (encoder_bwd
// below); with the encoder frozen er_grad_get on its keys fails.
#include "engine_internal.h"

#include <algorithm>

struct GradSlot { float* ptr; int rows, cols, ld; };
// what the backward of one decoder layer reads from its forward: q|k|v, the attention output (+ its row log-sum-exp), the two pre-LayerNorm sums,
// LN1's fp16 output, the ReLU output.  Either one set per layer (kept from the forward pass: 1.6 GB per layer at 4 x 10 243 rows — 180 GB of HBM
// make the reference's per-layer recomputation unnecessary) or ONE set refilled by re-running the layer (opt.checkpointing, when memory is short).
struct LayerAct { __half *qkv16, *a16, *x1_16, *h16; float *s1, *s2, *lse2; };

struct er_train {
    int M_cap = 0, Mp = 0, B_cap = 0;
    float *ckpt = nullptr, *s1 = nullptr, *s2 = nullptr, *x1_32 = nullptr, *g32a = nullptr, *g32b = nullptr;
    __half *x1_16 = nullptr, *o16 = nullptr, *dbr16 = nullptr, *dwide16 = nullptr, *da16 = nullptr, *tA = nullptr, *tB = nullptr, *dl16 = nullptr;
    float *lse2 = nullptr, *dsum = nullptr, *mean = nullptr, *rstd = nullptr;
    __half *pc16_all = nullptr, *dpc16 = nullptr;
    float* dcond32 = nullptr;
    // independent of the row count
    __half* wT = nullptr;
    float* partial = nullptr;
    int32_t* bucket_dev = nullptr;
    float* gflat = nullptr; size_t gtotal = 0;
    float *gwqkv, *gbqkv, *gwo, *gbo, *gln1w, *gln1b, *gw1, *gb1, *gw2, *gb2, *gln2w, *gln2b, *glm, *gembd, *gpos, *gpcw, *gpcb, *gncw, *gncb, *genf;
    std::map<std::string, GradSlot> gslots;
    float loss_scale = 1.f;
    int Vp = 0, wide = 0;
    bool have_grads = false;
    std::vector<void*> row_allocs;      // the buffers sized by M (re-allocated when a larger batch arrives)
    bool store = false;                 // activations of every layer kept from the forward pass (else: recomputed per layer in the backward pass)
    std::vector<LayerAct> acts;         // NL sets (store) or one
    // point-encoder backward (opt.freeze_encoder = False): gradient arrays, one cloud's scratch
    float *gqe = nullptr, *gmlpw, *gmlpb, *glnw, *glnb, *gcl1w, *gcl1b, *gcl2w, *gcl2b, *gcqw, *gcqb, *gckvw, *gckvb, *gcow, *gcob, *gff0w, *gff0b, *gff2w, *gff2b,
          *glinw, *glinb;
    __half *dlat16 = nullptr, *ed16a = nullptr, *ed16b = nullptr, *ed16c = nullptr; float *ef32a = nullptr, *ef32b = nullptr, *eacc32 = nullptr, *gtmp = nullptr;
    bool encoder_grads = false;         // the last er_train_step trained the point encoder
    __half *qkv_all = nullptr, *a_all = nullptr, *x1_all = nullptr, *h_all = nullptr; float *s1_all = nullptr, *s2_all = nullptr, *lse_all = nullptr;
};

int g_er_train_fwd_lse = 1;     // er_debug_set(NULL, "train_fwd_lse", 0): statistics pass in the backward instead of: the recomputed forward attention writes the row log-sum-exp, the backward skips its statistics pass

int g_er_train_recompute = 0;   // er_debug_set(NULL, "train_recompute", 1): per-layer recomputation in the backward pass (the reference's opt.checkpointing) even when memory allows keeping the activations

static inline int round64(int x) { return (x + 63) / 64 * 64; }

void er_train_destroy(er_engine* e) {
    if (!e->train) return;
    delete e->train;      // device memory is owned by e->allocs (freed by er_destroy)
    e->train = nullptr;
}

namespace {

__global__ void train_losses_kernel(const double* loss_sum, const int* count, const double* sq_sum, float kl_weight, int has_kl, float* losses, double* sums) {
    const double ce = *loss_sum / (double)max(*count, 1);
    const double kl = has_kl ? 0.5 * *sq_sum : 0.0;
    losses[1] = (float)ce; losses[2] = (float)kl; losses[0] = (float)(ce + (has_kl ? (double)kl_weight * kl : 0.0));
    if (sums) { sums[0] = *loss_sum; sums[1] = (double)*count; sums[2] = kl; }
}

}  // namespace

template <typename T>
static int row_alloc(er_engine* e, er_train* t, T** p, size_t n) {
    int r = dev_alloc(e, p, n);
    if (r) return r;
    t->row_allocs.push_back((void*)*p);
    return ER_OK;
}
#define RALLOC(ptr, n) do { int _r = row_alloc(e, t, &(ptr), (size_t)(n)); if (_r) return _r; } while (0)

// one-time state: gradient buffers mirroring the engine's weight arrays, slot views by state-dict key
static int create_train_impl(er_engine* e, er_train* t);
static int create_train(er_engine* e) {
    if (e->C % 64 || e->F % 64 || e->LDP % 8) return set_err(ER_ERR_INVALID, "training needs hidden_dim and ffn_dim multiples of 64 (got %d, %d)", e->C, e->F);
    er_train* t = new er_train();
    const int r = create_train_impl(e, t);
    if (r) { delete t; return r; }        // device buffers already recorded in e->allocs are released by er_destroy
    e->train = t;
    return ER_OK;
}
static int create_train_impl(er_engine* e, er_train* t) {
    const size_t C = e->C, F = e->F, V = e->V, NL = e->NL, LDP = e->LDP;
    t->Vp = round64(e->V);
    t->wide = (int)std::max(F, 3 * C);
    struct Arr { const __half* w; size_t n; float** g; };
    std::vector<Arr> arrs = {
        {e->wqkv, NL * 3 * C * C, &t->gwqkv}, {e->bqkv, NL * 3 * C, &t->gbqkv}, {e->wo, NL * C * C, &t->gwo}, {e->bo, NL * C, &t->gbo},
        {e->ln1w, NL * C, &t->gln1w}, {e->ln1b, NL * C, &t->gln1b}, {e->w1, NL * F * C, &t->gw1}, {e->b1, NL * F, &t->gb1},
        {e->w2, NL * C * F, &t->gw2}, {e->b2, NL * C, &t->gb2}, {e->ln2w, NL * C, &t->gln2w}, {e->ln2b, NL * C, &t->gln2b},
        {e->lm_head, V * C, &t->glm}, {e->embd, V * C, &t->gembd}, {e->pos, (size_t)e->cfg.max_positions * C, &t->gpos},
        {e->pc_w, C * LDP, &t->gpcw}, {e->pc_b, C, &t->gpcb}, {e->ncw, C, &t->gncw}, {e->ncb, C, &t->gncb},
        {e->cfg.use_num_face_cond ? e->enf : nullptr, e->cfg.use_num_face_cond ? 10 * C : 0, &t->genf},
    };
    if (e->cfg.has_point_encoder) {
        const size_t E = e->E, LQ = e->LQ;
        const Arr enc[] = {
            {e->qe, LQ * E, &t->gqe}, {e->mlp_w, E * 64, &t->gmlpw}, {e->mlp_b, E, &t->gmlpb}, {e->ln_w, E, &t->glnw}, {e->ln_b, E, &t->glnb},
            {e->cl1w, E, &t->gcl1w}, {e->cl1b, E, &t->gcl1b}, {e->cl2w, E, &t->gcl2w}, {e->cl2b, E, &t->gcl2b}, {e->cq_w, E * E, &t->gcqw}, {e->cq_b, E, &t->gcqb},
            {e->ckv_w, 2 * E * E, &t->gckvw}, {e->ckv_b, 2 * E, &t->gckvb}, {e->co_w, E * E, &t->gcow}, {e->co_b, E, &t->gcob},
            {e->ff0w, 8 * E * E, &t->gff0w}, {e->ff0b, 8 * E, &t->gff0b}, {e->ff2w, 4 * E * E, &t->gff2w}, {e->ff2b, E, &t->gff2b},
            {e->lin_w, LDP * E, &t->glinw}, {e->lin_b, LDP, &t->glinb},
        };
        for (const Arr& a : enc) arrs.push_back(a);
    }
    size_t total = 0;
    std::vector<size_t> offs;
    for (const Arr& a : arrs) { offs.push_back(total); total += (a.n + 63) / 64 * 64; }
    t->gtotal = total;
    ALLOC(t->gflat, total);
    CK(cudaMemset(t->gflat, 0, total * 4));
    for (size_t i = 0; i < arrs.size(); i++) *arrs[i].g = arrs[i].n ? t->gflat + offs[i] : nullptr;
    for (const auto& kv : e->slots) {
        const Slot& s = kv.second;
        for (size_t i = 0; i < arrs.size(); i++) {
            const Arr& a = arrs[i];
            if (a.w && s.dst >= a.w && s.dst < a.w + a.n) {
                t->gslots[kv.first] = GradSlot{*a.g + (s.dst - a.w), s.rows, s.cols, s.dst_ld};
                break;
            }
        }
    }
    const size_t E8 = e->cfg.has_point_encoder ? (size_t)8 * e->E : 0;
    ALLOC(t->wT, std::max(std::max(std::max(F * C, 3 * C * C), C * (size_t)t->Vp), E8 * e->E) + 64);
    ALLOC(t->partial, (size_t)ER_BW_SLABS * 2 * std::max((size_t)t->wide, E8));
    if (e->cfg.has_point_encoder) ALLOC(t->gtmp, E8 * e->E);
    ALLOC(t->bucket_dev, 4096);
    return ER_OK;
}

static int ensure_train(er_engine* e, int M, int B) {
    if (!e->train) { int r = create_train(e); if (r) return r; }
    er_train* t = e->train;
    if (B > 4096) return set_err(ER_ERR_CAPACITY, "batch too large");
    const size_t per_layer = (size_t)M * ((size_t)e->C * (3 + 1 + 1) * 2 + (size_t)e->F * 2 + (size_t)e->C * 2 * 4 + (size_t)e->H * 4);
    bool want_store = !g_er_train_recompute;
    if (want_store && !(M <= t->M_cap && B <= t->B_cap && t->store)) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        size_t mine = 0;                      // what the current row buffers would give back
        if (t->M_cap) mine = (size_t)t->M_cap * e->C * 4 * (e->NL + 8) + (t->store ? (size_t)e->NL * per_layer / M * t->M_cap : 0);
        const size_t need = (size_t)e->NL * per_layer + (size_t)M * e->C * 4 * (e->NL + 12) + (size_t)3 * round64(M) * t->wide * 2 + ((size_t)3 << 30);
        if (free_b + mine < need) want_store = false;
    }
    if (M <= t->M_cap && B <= t->B_cap && want_store == t->store) return ER_OK;
    CK(cudaDeviceSynchronize());
    for (void* p : t->row_allocs) {
        for (auto it = e->allocs.begin(); it != e->allocs.end(); ++it)
            if (*it == p) { e->allocs.erase(it); break; }
        cudaFree(p);
    }
    t->row_allocs.clear();
    t->M_cap = 0;
    const size_t C = e->C, F = e->F, NL = e->NL, H = e->H;
    const size_t Mc = std::max(M, t->M_cap), Bc = std::max(B, t->B_cap);
    const size_t Mp = round64((int)Mc);
    const size_t RL = Bc * e->LQ;
    RALLOC(t->ckpt, (NL + 1) * Mc * C);
    RALLOC(t->s1, Mc * C); RALLOC(t->s2, Mc * C); RALLOC(t->x1_32, Mc * C); RALLOC(t->g32a, Mc * C); RALLOC(t->g32b, Mc * C);
    RALLOC(t->x1_16, Mc * C); RALLOC(t->o16, Mc * C); RALLOC(t->dbr16, Mc * C); RALLOC(t->da16, Mc * C);
    RALLOC(t->dwide16, Mc * t->wide);
    // one cloud of the point encoder (its backward runs per sample): np points, LQ queries, width E
    const size_t E = e->cfg.has_point_encoder ? e->E : 0, np = e->cfg.has_point_encoder ? e->cfg.max_points : 0, LQ = e->LQ;
    const size_t enc_tA = std::max(8 * E * round64((int)LQ), 2 * E * round64((int)np)), enc_tB = std::max(4 * E * round64((int)LQ), std::max(E, (size_t)64) * round64((int)np));
    RALLOC(t->tA, std::max((size_t)std::max<size_t>(t->wide, e->V) * Mp, enc_tA)); RALLOC(t->tB, std::max(std::max<size_t>(std::max(F, C), e->LDP) * Mp, enc_tB));
    RALLOC(t->dl16, Mc * t->Vp);
    const size_t enc_stats = e->cfg.has_point_encoder ? (size_t)e->EH * LQ : 0;       // attention statistics of one cloud
    RALLOC(t->lse2, std::max(Mc * H, enc_stats)); RALLOC(t->dsum, std::max(Mc * H, enc_stats)); RALLOC(t->mean, std::max(Mc, np)); RALLOC(t->rstd, std::max(Mc, np));
    RALLOC(t->dlat16, RL * 64 * ((e->LDP + 63) / 64));
    CK(cudaMemset(t->dlat16, 0, RL * 64 * ((e->LDP + 63) / 64) * 2));            // pad columns [LDP, round64) stay zero: they are K of the lin dgrad GEMM
    if (e->cfg.has_point_encoder) {
        const size_t big = std::max(LQ * 8 * E, np * 2 * E), rows = std::max(LQ, np);
        RALLOC(t->ed16a, big); RALLOC(t->ed16b, big); RALLOC(t->ed16c, big); RALLOC(t->ef32a, rows * E); RALLOC(t->ef32b, rows * E); RALLOC(t->eacc32, LQ * E);
    }
    RALLOC(t->pc16_all, RL * C); RALLOC(t->dpc16, RL * C); RALLOC(t->dcond32, RL * C);
    t->store = want_store;
    t->acts.assign(want_store ? NL : 1, LayerAct{});
    if (want_store) {
        RALLOC(t->qkv_all, NL * Mc * 3 * C); RALLOC(t->a_all, NL * Mc * C); RALLOC(t->x1_all, NL * Mc * C); RALLOC(t->h_all, NL * Mc * F);
        RALLOC(t->s1_all, NL * Mc * C); RALLOC(t->s2_all, NL * Mc * C); RALLOC(t->lse_all, NL * Mc * H);
        for (size_t l = 0; l < NL; l++)
            t->acts[l] = LayerAct{t->qkv_all + l * Mc * 3 * C, t->a_all + l * Mc * C, t->x1_all + l * Mc * C, t->h_all + l * Mc * F,
                                  t->s1_all + l * Mc * C, t->s2_all + l * Mc * C, t->lse_all + l * Mc * H};
    }
    t->M_cap = (int)Mc; t->Mp = (int)Mp; t->B_cap = (int)Bc;
    return ER_OK;
}

// one OPTDecoderLayer in training mode (modeling_opt.py:264-288): in32/in16 -> out32/out16 (skipped when out32 is null: backward recomputation);
// what the backward needs goes to `act`.  want_lse: the attention kernel also writes the rows' log-sum-exp (act.lse2).
static int layer_fwd(er_engine* e, int l, const LayerAct& act, const float* in32, const __half* in16, float* out32, __half* out16, int B, int N,
                     const unsigned char* row_mask, float p, unsigned long long seed, bool want_lse, cudaStream_t st) {
    er_train* t = e->train;
    const int C = e->C, F = e->F, H = e->H, M = B * N;
    er::GemmArgs g = mk_gemm(in16, C, e->wqkv + (size_t)l * 3 * C * C, C, e->bqkv + (size_t)l * 3 * C, M, 3 * C, C, er::GEMM_F16);
    g.out16 = act.qkv16; g.ldo = 3 * C; CKL(e, er_gemm(g, st));
    er::AttnArgs a{};
    a.q = act.qkv16; a.k = act.qkv16 + C; a.v = act.qkv16 + 2 * C; a.out = act.a16;
    a.ldq = a.ldk = a.ldv = 3 * C; a.ldo = C; a.q_bs = a.k_bs = a.v_bs = (long long)N * 3 * C; a.o_bs = (long long)N * C;
    a.B = B; a.H = H; a.Nq = N; a.Nk = N; a.D = 96; a.causal = 1;
    if (want_lse) a.lse2 = act.lse2;
    CKL(e, er_attention(a, st));
    if (row_mask) CKL(e, er_zero_masked_rows(act.a16, row_mask, M, C, st));
    g = mk_gemm(act.a16, C, e->wo + (size_t)l * C * C, C, e->bo + (size_t)l * C, M, C, C, er::GEMM_F16);
    g.out16 = t->o16; g.ldo = C; CKL(e, er_gemm(g, st));
    CKL(e, er_add_dropout(in32, t->o16, act.s1, (size_t)M * C, p, seed, 2 * l, st));
    CKL(e, er_layernorm(act.s1, nullptr, C, e->ln1w + (size_t)l * C, e->ln1b + (size_t)l * C, t->x1_32, act.x1_16, C, M, C, st));
    g = mk_gemm(act.x1_16, C, e->w1 + (size_t)l * F * C, C, e->b1 + (size_t)l * F, M, F, C, er::GEMM_F16_RELU);
    g.out16 = act.h16; g.ldo = F; CKL(e, er_gemm(g, st));
    g = mk_gemm(act.h16, F, e->w2 + (size_t)l * C * F, F, e->b2 + (size_t)l * C, M, C, F, er::GEMM_F16);
    g.out16 = t->o16; g.ldo = C; CKL(e, er_gemm(g, st));
    CKL(e, er_add_dropout(t->x1_32, t->o16, act.s2, (size_t)M * C, p, seed, 2 * l + 1, st));
    if (out32) CKL(e, er_layernorm(act.s2, nullptr, C, e->ln2w + (size_t)l * C, e->ln2b + (size_t)l * C, out32, out16, C, M, C, st));
    return ER_OK;
}

// dW [Nout][Kin] (fp32) = dY^T X for dY [M][Nout] (pitch ldy), X [M][Kin] (pitch ldx): both transposed to K-major, then one GEMM with K = Mp
static int wgrad(er_engine* e, const __half* dy, int ldy, int Nout, const __half* x, int ldx, int Kin, int M, float* dW, int ldw, cudaStream_t st) {
    er_train* t = e->train;
    const int Mp = round64(M);
    CKL(e, er_transpose_f16(dy, M, Nout, ldy, t->tA, Mp, st));
    CKL(e, er_transpose_f16(x, M, Kin, ldx, t->tB, Mp, st));
    er::GemmArgs g = mk_gemm(t->tA, Mp, t->tB, Mp, nullptr, Nout, Kin, Mp, er::GEMM_F32);
    g.out32 = dW; g.ldo = ldw; CKL(e, er_gemm(g, st));
    return ER_OK;
}
// dX = dY W for dY [M][Nout] (pitch ldy), W [Nout][Kin] row-major: W is transposed into wT [Kin][Nout]; mode GEMM_F16 -> out16, GEMM_F32 -> out32,
// GEMM_F32_RES32 -> out32 = res32 + f16(dY W)
static int dgrad(er_engine* e, const __half* dy, int ldy, int Nout, const __half* W, int Kin, int M, int mode, __half* out16, float* out32, const float* res32,
                 cudaStream_t st, int ldo = 0) {
    er_train* t = e->train;
    const int Np = round64(Nout);
    CKL(e, er_transpose_f16(W, Nout, Kin, Kin, t->wT, Np, st));
    er::GemmArgs g = mk_gemm(dy, ldy, t->wT, Np, nullptr, M, Kin, Np, mode);
    g.out16 = out16; g.out32 = out32; g.ldo = ldo ? ldo : Kin; g.res32 = res32; g.ldr = Kin; CKL(e, er_gemm(g, st));
    return ER_OK;
}

static int layer_bwd(er_engine* e, int l, int B, int N, const unsigned char* row_mask, float p, unsigned long long seed, cudaStream_t st) {
    er_train* t = e->train;
    const int C = e->C, F = e->F, H = e->H, M = B * N;
    const float* in32 = t->ckpt + (size_t)l * M * C;
    const LayerAct& act = t->store ? t->acts[l] : t->acts[0];
    // x16 = the fp16 copy of the layer input the forward pass used (operand of the q/k/v weight gradient)
    CKL(e, er_f32_to_f16(in32, e->x16, (size_t)M * C, st));
    if (!t->store) {        // opt.checkpointing: re-run the layer from its checkpointed input
        int r = layer_fwd(e, l, act, in32, e->x16, nullptr, nullptr, B, N, row_mask, p, seed, g_er_train_fwd_lse != 0, st);
        if (r) return r;
    }
    // ---- final_layer_norm: g32a = dL/d(out) -> g32b = dL/d(s2); dbr16 = dL/d(fc2 output) ----
    CKL(e, er_ln_bwd(t->g32a, act.s2, nullptr, C, e->ln2w + (size_t)l * C, t->g32b, t->dbr16, t->mean, t->rstd, M, C, p, seed, 2 * l + 1, st));
    CKL(e, er_ln_param_grad(t->g32a, act.s2, nullptr, C, t->mean, t->rstd, M, C, t->partial, t->gln2w + (size_t)l * C, t->gln2b + (size_t)l * C, st));
    // ---- fc2 ----
    CKL(e, er_colsum_f16(t->dbr16, C, M, C, t->partial, t->gb2 + (size_t)l * C, st));
    { int r = dgrad(e, t->dbr16, C, C, e->w2 + (size_t)l * C * F, F, M, er::GEMM_F16, t->dwide16, nullptr, nullptr, st); if (r) return r; }
    { int r = wgrad(e, t->dbr16, C, C, act.h16, F, F, M, t->gw2 + (size_t)l * C * F, F, st); if (r) return r; }
    CKL(e, er_relu_bwd(t->dwide16, act.h16, (size_t)M * F, st));
    // ---- fc1: g32a = g32b + dh W1 ----
    CKL(e, er_colsum_f16(t->dwide16, F, M, F, t->partial, t->gb1 + (size_t)l * F, st));
    { int r = dgrad(e, t->dwide16, F, F, e->w1 + (size_t)l * F * C, C, M, er::GEMM_F32_RES32, nullptr, t->g32a, t->g32b, st); if (r) return r; }
    { int r = wgrad(e, t->dwide16, F, F, act.x1_16, C, C, M, t->gw1 + (size_t)l * F * C, C, st); if (r) return r; }
    // ---- self_attn_layer_norm ----
    CKL(e, er_ln_bwd(t->g32a, act.s1, nullptr, C, e->ln1w + (size_t)l * C, t->g32b, t->dbr16, t->mean, t->rstd, M, C, p, seed, 2 * l, st));
    CKL(e, er_ln_param_grad(t->g32a, act.s1, nullptr, C, t->mean, t->rstd, M, C, t->partial, t->gln1w + (size_t)l * C, t->gln1b + (size_t)l * C, st));
    // ---- out_proj ----
    CKL(e, er_colsum_f16(t->dbr16, C, M, C, t->partial, t->gbo + (size_t)l * C, st));
    { int r = dgrad(e, t->dbr16, C, C, e->wo + (size_t)l * C * C, C, M, er::GEMM_F16, t->da16, nullptr, nullptr, st); if (r) return r; }
    { int r = wgrad(e, t->dbr16, C, C, act.a16, C, C, M, t->gwo + (size_t)l * C * C, C, st); if (r) return r; }
    if (row_mask) CKL(e, er_zero_masked_rows(t->da16, row_mask, M, C, st));         // backward of pad_input's zero rows
    // ---- attention ----
    er::AttnArgs a{};
    a.q = act.qkv16; a.k = act.qkv16 + C; a.v = act.qkv16 + 2 * C; a.out = act.a16;
    a.ldq = a.ldk = a.ldv = 3 * C; a.ldo = C; a.q_bs = a.k_bs = a.v_bs = (long long)N * 3 * C; a.o_bs = (long long)N * C;
    a.B = B; a.H = H; a.Nq = N; a.Nk = N; a.D = 96; a.causal = 1;
    if (g_er_train_fwd_lse) a.lse2 = act.lse2;          // written by the forward kernel of this layer: no statistics pass
    e->launches += 2;
    CKL(e, er_attention_bwd(a, t->da16, t->dwide16, t->dwide16 + C, t->dwide16 + 2 * C, 3 * C, 3 * C, 3 * C, a.q_bs, a.q_bs, a.q_bs, act.lse2, t->dsum, st));
    // ---- q/k/v projections: g32a = g32b + dqkv Wqkv ----
    CKL(e, er_colsum_f16(t->dwide16, 3 * C, M, 3 * C, t->partial, t->gbqkv + (size_t)l * 3 * C, st));
    { int r = dgrad(e, t->dwide16, 3 * C, 3 * C, e->wqkv + (size_t)l * 3 * C * C, C, M, er::GEMM_F32_RES32, nullptr, t->g32a, t->g32b, st); if (r) return r; }
    { int r = wgrad(e, t->dwide16, 3 * C, 3 * C, e->x16, C, C, M, t->gwqkv + (size_t)l * 3 * C * C, C, st); if (r) return r; }
    return ER_OK;
}

// dW += dY^T X (the point encoder's weights collect one cloud at a time): wgrad into the fp32 scratch, then a plain add
static int wgrad_acc(er_engine* e, const __half* dy, int ldy, int Nout, const __half* x, int ldx, int Kin, int M, float* dW, int ldw, cudaStream_t st) {
    er_train* t = e->train;
    int r = wgrad(e, dy, ldy, Nout, x, ldx, Kin, M, t->gtmp, ldw, st);
    if (r) return r;
    CKL(e, er_add_f32(dW, t->gtmp, (size_t)Nout * ldw, st));
    return ER_OK;
}

// Backward of PointEncoderEmbed.forward (core/transformer/point.py:186-206, 117-126, 74-84) for ONE cloud: the forward is re-run into the encoder
// workspace (encode_points), then walked back.  dlat: gradient of this cloud's latents [LQ][round64(LDP)] (fp16, loss-scaled).  Weight gradients
// accumulate over the clouds of the batch (the gradient buffer was zeroed at the start of the step).
static int encoder_bwd(er_engine* e, const float* pts, int n, const __half* dlat, cudaStream_t st) {
    er_train* t = e->train;
    const int E = e->E, LQ = e->LQ, EH = e->EH, LDP = e->LDP, LDPp = round64(e->LDP);
    { int r = encode_points(e, pts, n, e->pc16 /* scratch: the latents are not needed again */, st); if (r) return r; }
    // ---- linear: lat = ex2 W^T + b ----
    CKL(e, er_colsum_f16(dlat, LDPp, LQ, LDP, t->partial, t->glinb, st, 1));
    { int r = wgrad_acc(e, dlat, LDPp, LDP, e->ex2, E, E, LQ, t->glinw, E, st); if (r) return r; }
    { int r = dgrad(e, dlat, LDPp, LDP, e->lin_w, E, LQ, er::GEMM_F16, t->ed16a, nullptr, nullptr, st); if (r) return r; }       // ed16a = d ex2
    CKL(e, er_f16_to_f32(t->ed16a, t->eacc32, LQ * E, st));                                                                  // eacc32 = d ex1 (residual path)
    // ---- mlp.net.2: ex2 = ex1 + egg W^T + b ----
    CKL(e, er_colsum_f16(t->ed16a, E, LQ, E, t->partial, t->gff2b, st, 1));
    { int r = wgrad_acc(e, t->ed16a, E, E, e->egg, 4 * E, 4 * E, LQ, t->gff2w, 4 * E, st); if (r) return r; }
    { int r = dgrad(e, t->ed16a, E, E, e->ff2w, 4 * E, LQ, er::GEMM_F16, t->ed16b, nullptr, nullptr, st); if (r) return r; }     // ed16b = d egg
    CKL(e, er_geglu_bwd(e->eff, t->ed16b, t->ed16c, LQ, 4 * E, st));                                                            // ed16c = d eff
    // ---- mlp.net.0: eff = ex1ln W^T + b ----
    CKL(e, er_colsum_f16(t->ed16c, 8 * E, LQ, 8 * E, t->partial, t->gff0b, st, 1));
    { int r = wgrad_acc(e, t->ed16c, 8 * E, 8 * E, e->ex1ln, E, E, LQ, t->gff0w, E, st); if (r) return r; }
    { int r = dgrad(e, t->ed16c, 8 * E, 8 * E, e->ff0w, E, LQ, er::GEMM_F16, t->ed16a, nullptr, nullptr, st); if (r) return r; } // ed16a = d ex1ln
    // ---- cross_att.ln2 (input ex1, fp16) ----
    CKL(e, er_f16_to_f32(t->ed16a, t->ef32a, LQ * E, st));
    CKL(e, er_ln_bwd(t->ef32a, nullptr, e->ex1, E, e->cl2w, t->ef32b, nullptr, t->mean, t->rstd, LQ, E, 0.f, 0, 0, st));
    CKL(e, er_ln_param_grad(t->ef32a, nullptr, e->ex1, E, t->mean, t->rstd, LQ, E, t->partial, t->gcl2w, t->gcl2b, st, 1));
    CKL(e, er_add_f32(t->eacc32, t->ef32b, (size_t)LQ * E, st));
    // ---- ex1 = query_embed + out_proj(attention) ----
    CKL(e, er_add_f32(t->gqe, t->eacc32, (size_t)LQ * E, st));
    CKL(e, er_f32_to_f16(t->eacc32, t->ed16a, (size_t)LQ * E, st));                                                            // ed16a = d (out_proj output)
    CKL(e, er_colsum_f16(t->ed16a, E, LQ, E, t->partial, t->gcob, st, 1));
    { int r = wgrad_acc(e, t->ed16a, E, E, e->ea16, E, E, LQ, t->gcow, E, st); if (r) return r; }
    { int r = dgrad(e, t->ed16a, E, E, e->co_w, E, LQ, er::GEMM_F16, t->ed16b, nullptr, nullptr, st); if (r) return r; }         // ed16b = d ea
    // ---- cross attention: q = qq16 [LQ][E], k | v = kvo16 [n][2E] ----
    er::AttnArgs a{};
    a.q = e->qq16; a.k = e->kvo16; a.v = e->kvo16 + E; a.out = e->ea16;
    a.ldq = E; a.ldk = 2 * E; a.ldv = 2 * E; a.ldo = E; a.B = 1; a.H = EH; a.Nq = LQ; a.Nk = n; a.D = 64; a.causal = 0;
    e->launches += 2;
    CKL(e, er_attention_bwd(a, t->ed16b, t->ed16a, t->ed16c, t->ed16c + E, E, 2 * E, 2 * E, 0, 0, 0, t->lse2, t->dsum, st));      // ed16a = d qq, ed16c = d k|v
    // ---- q_proj and cross_att.ln1 (input query_embed) ----
    CKL(e, er_colsum_f16(t->ed16a, E, LQ, E, t->partial, t->gcqb, st, 1));
    { int r = wgrad_acc(e, t->ed16a, E, E, e->qln16, E, E, LQ, t->gcqw, E, st); if (r) return r; }
    { int r = dgrad(e, t->ed16a, E, E, e->cq_w, E, LQ, er::GEMM_F16, t->ed16b, nullptr, nullptr, st); if (r) return r; }         // ed16b = d qln
    CKL(e, er_f16_to_f32(t->ed16b, t->ef32a, LQ * E, st));
    CKL(e, er_ln_bwd(t->ef32a, nullptr, e->qe, E, e->cl1w, t->ef32b, nullptr, t->mean, t->rstd, LQ, E, 0.f, 0, 0, st));
    CKL(e, er_ln_param_grad(t->ef32a, nullptr, e->qe, E, t->mean, t->rstd, LQ, E, t->partial, t->gcl1w, t->gcl1b, st, 1));
    CKL(e, er_add_f32(t->gqe, t->ef32b, (size_t)LQ * E, st));
    // ---- k_proj | v_proj: kvo = kvx W^T + b ----
    CKL(e, er_colsum_f16(t->ed16c, 2 * E, n, 2 * E, t->partial, t->gckvb, st, 1));
    { int r = wgrad_acc(e, t->ed16c, 2 * E, 2 * E, e->kvx16, E, E, n, t->gckvw, E, st); if (r) return r; }
    { int r = dgrad(e, t->ed16c, 2 * E, 2 * E, e->ckv_w, E, n, er::GEMM_F16, t->ed16a, nullptr, nullptr, st); if (r) return r; } // ed16a = d kvx [n][E]
    // ---- ln (input pf16) and point_embed.mlp ----
    CKL(e, er_f16_to_f32(t->ed16a, t->ef32a, n * E, st));
    CKL(e, er_ln_bwd(t->ef32a, nullptr, e->pf16, E, e->ln_w, nullptr, t->ed16b, t->mean, t->rstd, n, E, 0.f, 0, 0, st));          // ed16b = d pf [n][E]
    CKL(e, er_ln_param_grad(t->ef32a, nullptr, e->pf16, E, t->mean, t->rstd, n, E, t->partial, t->glnw, t->glnb, st, 1));
    CKL(e, er_colsum_f16(t->ed16b, E, n, E, t->partial, t->gmlpb, st, 1));
    { int r = wgrad_acc(e, t->ed16b, E, E, e->emb16, 64, 64, n, t->gmlpw, 64, st); if (r) return r; }
    return ER_OK;
}

extern "C" int er_train_step(er_engine* e, const float* conds_dev, int32_t n_points, int32_t is_latent, const int32_t* tokens_dev, const int64_t* labels_dev,
                             const uint8_t* mask_dev, const int32_t* num_faces_host, int32_t B, int32_t T, float kl_weight, float dropout_p, uint64_t seed,
                             float loss_scale, int32_t train_encoder, float* losses_dev, double* sums_dev, void* stream) {
    if (!e || !conds_dev || !tokens_dev || !labels_dev || !num_faces_host || !losses_dev) return set_err(ER_ERR_INVALID, "null argument");
    if (!e->finalized) return set_err(ER_ERR_STATE, "weights not finalized");
    if (B < 1 || T < 1) return set_err(ER_ERR_INVALID, "bad batch shape");
    if (train_encoder && (is_latent || !e->cfg.has_point_encoder)) return set_err(ER_ERR_INVALID, "train_encoder needs cond_mode 'point' (an engine with a point encoder)");
    if (!(dropout_p >= 0.f && dropout_p < 1.f) || !(loss_scale > 0.f)) return set_err(ER_ERR_INVALID, "dropout_p must be in [0, 1), loss_scale > 0");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = e->C, P = e->P, N = P + T, M = B * N, V = e->V, NL = e->NL;
    if (!e->logits_all || B > e->lat_batch_cap) return set_err(ER_ERR_CAPACITY, "engine was created with max_tf_rows too small for a batch of %d samples", B);
    if (N > e->cfg.max_positions) return set_err(ER_ERR_CAPACITY, "sequence longer than the position table");
    { int r0 = ensure_dense_rows(e, M); if (r0) return r0; }
    { int r0 = ensure_train(e, M, B); if (r0) return r0; }
    er_train* t = e->train;
    t->loss_scale = loss_scale;
    t->have_grads = false;
    // ---- forward (training mode) ------------------------------------------------------------------------------------------------------------
    const size_t cstride = is_latent ? (size_t)e->LQ * e->LD : (size_t)n_points * 3;
    std::vector<int32_t> bucket(B);
    for (int b = 0; b < B; b++) {
        float* cond_rows = t->ckpt + (size_t)b * N * C;        // checkpoint 0 = the embeddings
        int r = encode_one(e, conds_dev + b * cstride, n_points, is_latent, num_faces_host[b], e->lat16 + (size_t)b * e->LQ * e->LDP, cond_rows, st);
        if (r) return r;
        CKL(e, er_embed_prefix(cond_rows, P, tokens_dev + (size_t)b * T, T, e->embd, e->pos, C, cond_rows, e->x16 + (size_t)b * N * C, st));
        bucket[b] = er_quantize_num_faces(num_faces_host[b]);
    }
    CK(cudaMemcpyAsync(t->bucket_dev, bucket.data(), (size_t)B * 4, cudaMemcpyHostToDevice, st));
    if (!t->store) t->acts[0] = LayerAct{e->qkv16, e->a16, t->x1_16, e->h16, t->s1, t->s2, t->lse2};     // the dense workspace (it may have been re-allocated)
    for (int l = 0; l < NL; l++) {
        int r = layer_fwd(e, l, t->store ? t->acts[l] : t->acts[0], t->ckpt + (size_t)l * M * C, e->x16, t->ckpt + (size_t)(l + 1) * M * C, e->x16, B, N,
                          mask_dev, dropout_p, seed, t->store && g_er_train_fwd_lse, st);
        if (r) return r;
    }
    er::GemmArgs g = mk_gemm(e->x16, C, e->lm_head, C, nullptr, M, V, C, er::GEMM_F32);
    g.out32 = e->logits_all; g.ldo = V; CKL(e, er_gemm(g, st));
    CK(cudaMemsetAsync(e->tf_acc, 0, 32, st));
    CK(cudaMemsetAsync(e->tf_cnt, 0, 16, st));
    for (int b = 0; b < B; b++) {
        e->launches++;
        CKL(e, er_cross_entropy(e->logits_all + (size_t)b * N * V, V, labels_dev + (size_t)b * N + 1, N - 1, V, e->tf_rows, e->tf_valid, e->tf_acc, e->tf_cnt, st));
    }
    const int has_kl = !is_latent;
    if (has_kl) { e->launches++; CKL(e, er_sum_squares(e->lat16, (size_t)B * e->LQ * e->LDP, e->tf_part, e->tf_acc + 1, st)); }
    e->launches++;
    train_losses_kernel<<<1, 1, 0, st>>>(e->tf_acc, e->tf_cnt, e->tf_acc + 1, kl_weight, has_kl, losses_dev, sums_dev);
    CK(cudaGetLastError());
    // ---- backward ----------------------------------------------------------------------------------------------------------------------------------
    CK(cudaMemsetAsync(t->gflat, 0, t->gtotal * 4, st));
    // lm_head + cross-entropy: dlogits (x loss_scale / valid rows) -> g32a = dL/d(final hidden)
    CKL(e, er_ce_bwd(e->logits_all, V, labels_dev, B, N, V, e->tf_cnt, loss_scale, t->dl16, t->Vp, st));
    { int r = dgrad(e, t->dl16, t->Vp, V, e->lm_head, C, M, er::GEMM_F32, nullptr, t->g32a, nullptr, st); if (r) return r; }
    { int r = wgrad(e, t->dl16, t->Vp, V, e->x16, C, C, M, t->glm, C, st); if (r) return r; }
    for (int l = NL - 1; l >= 0; l--) {
        int r = layer_bwd(e, l, B, N, mask_dev, dropout_p, seed, st);
        if (r) return r;
    }
    // embeddings: g32a = dL/d(embeddings + positions)
    e->launches += 2;
    CKL(e, er_embed_bwd(t->g32a, tokens_dev, t->bucket_dev, B, T, N, P, C, V, e->LQ, t->gpos, t->gembd, e->cfg.use_num_face_cond ? t->genf : nullptr, st));
    // conditioner: cond rows [0, LQ) of every sample = norm_cond(proj_cond(latents)) (models.py:124, 128-129); the latents carry no gradient (frozen encoder)
    const int LQ = e->LQ, RL = B * LQ;
    for (int b = 0; b < B; b++)
        CK(cudaMemcpyAsync(t->dcond32 + (size_t)b * LQ * C, t->g32a + (size_t)b * N * C, (size_t)LQ * C * 4, cudaMemcpyDeviceToDevice, st));
    g = mk_gemm(e->lat16, e->LDP, e->pc_w, e->LDP, e->pc_b, RL, C, e->LDP, er::GEMM_F16);
    g.out16 = t->pc16_all; g.ldo = C; CKL(e, er_gemm(g, st));
    CKL(e, er_ln_bwd(t->dcond32, nullptr, t->pc16_all, C, e->ncw, nullptr, t->dpc16, t->mean, t->rstd, RL, C, 0.f, 0, 0, st));
    CKL(e, er_ln_param_grad(t->dcond32, nullptr, t->pc16_all, C, t->mean, t->rstd, RL, C, t->partial, t->gncw, t->gncb, st));
    CKL(e, er_colsum_f16(t->dpc16, C, RL, C, t->partial, t->gpcb, st));
    { int r = wgrad(e, t->dpc16, C, C, e->lat16, e->LDP, e->LDP, RL, t->gpcw, e->LDP, st); if (r) return r; }
    t->encoder_grads = train_encoder != 0;
    if (train_encoder) {
        // d latents = dpc W_proj_cond + kl_weight * latents (DummyLatent.kl = 0.5 sum lat^2, point.py:33-35; models.py:191-197), loss-scaled like everything else
        const int LDPp = round64(e->LDP);
        // written with row pitch round64(LDP): the pad columns (zero since allocation) are the K tail of the encoder's first dgrad GEMM
        { int r = dgrad(e, t->dpc16, C, C, e->pc_w, e->LDP, RL, er::GEMM_F16, t->dlat16, nullptr, nullptr, st, LDPp); if (r) return r; }
        CKL(e, er_axpy_f16(t->dlat16, LDPp, e->lat16, e->LDP, RL, e->LDP, kl_weight * loss_scale, st));
        for (int b = 0; b < B; b++) {
            int r = encoder_bwd(e, conds_dev + b * cstride, n_points, t->dlat16 + (size_t)b * LQ * LDPp, st);
            if (r) return r;
        }
    }
    t->have_grads = true;
    return ER_OK;
}

// Copy the gradient of one state-dict entry (dense [rows][cols] fp32, loss scale removed) to out_dev.  Keys of the frozen point encoder have none.
extern "C" int er_grad_get(er_engine* e, const char* name, float* out_dev, int64_t numel, void* stream) {
    if (!e || !name || !out_dev) return set_err(ER_ERR_INVALID, "null argument");
    if (!e->train || !e->train->have_grads) return set_err(ER_ERR_STATE, "er_train_step has not run");
    er_train* t = e->train;
    auto it = t->gslots.find(name);
    if (it == t->gslots.end()) return set_err(ER_ERR_INVALID, "no gradient for '%s' (unknown key, or a frozen point-encoder tensor)", name);
    const GradSlot& s = it->second;
    if (!t->encoder_grads && strncmp(name, "point_encoder.", 14) == 0) return set_err(ER_ERR_STATE, "'%s': the last er_train_step ran with the point encoder frozen", name);
    if (numel != (int64_t)s.rows * s.cols) return set_err(ER_ERR_INVALID, "'%s' has %d x %d elements, buffer holds %lld", name, s.rows, s.cols, (long long)numel);
    CKL(e, er_export_f32(s.ptr, s.ld, s.rows, s.cols, 1.f / t->loss_scale, out_dev, (cudaStream_t)stream));
    return ER_OK;
}

// 1 if the entry receives a gradient from er_train_step, else 0
extern "C" int32_t er_grad_has(er_engine* e, const char* name) {
    if (!e || !name) return 0;
    if (!e->train) { if (create_train(e)) return 0; }
    return e->train->gslots.count(name) ? 1 : 0;
}

// Backward of the `attention()` op seam (core/transformer/attention.py:27-95; forward: er_attention_bnhd): q / dq [B][Nq][H][D], k, v / dk, dv
// [B][Nk][H][D], out / dout [B][Nq][H][D] fp16, dense or causal (Nq == Nk); `out` is the forward result.  No loss scaling here.
extern "C" int er_attention_bwd_bnhd(const void* q_dev, const void* k_dev, const void* v_dev, const void* out_dev, const void* dout_dev, void* dq_dev,
                                     void* dk_dev, void* dv_dev, int32_t B, int32_t Nq, int32_t Nk, int32_t H, int32_t D, int32_t causal, void* stream) {
    if (!q_dev || !k_dev || !v_dev || !out_dev || !dout_dev || !dq_dev || !dk_dev || !dv_dev) return set_err(ER_ERR_INVALID, "null argument");
    if (D != 64 && D != 96) return set_err(ER_ERR_INVALID, "head_dim %d not supported (64 or 96)", D);
    if (causal && Nq != Nk) return set_err(ER_ERR_INVALID, "causal attention needs Nq == Nk");
    if (B < 1 || Nq < 1 || Nk < 1 || H < 1) return set_err(ER_ERR_INVALID, "bad shape");
    er::AttnArgs a{};
    a.q = (const __half*)q_dev; a.k = (const __half*)k_dev; a.v = (const __half*)v_dev; a.out = (__half*)out_dev;
    a.ldq = a.ldk = a.ldv = a.ldo = H * D;
    a.q_bs = a.o_bs = (long long)Nq * H * D; a.k_bs = a.v_bs = (long long)Nk * H * D;
    a.B = B; a.H = H; a.Nq = Nq; a.Nk = Nk; a.D = D; a.causal = causal;
    float* stats = nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMallocAsync(&stats, (size_t)2 * B * H * Nq * sizeof(float), st));
    const cudaError_t err = er_attention_bwd(a, (const __half*)dout_dev, (__half*)dq_dev, (__half*)dk_dev, (__half*)dv_dev, H * D, H * D, H * D, a.q_bs, a.k_bs,
                                             a.v_bs, stats, stats + (size_t)B * H * Nq, st);
    cudaFreeAsync(stats, st);
    if (err != cudaSuccess) return set_err(ER_ERR_CUDA, "er_attention_bwd -> %s", cudaGetErrorString(err));
    return ER_OK;
}
