// Engine: owns packed fp16 weights, the KV cache and workspaces; implements the C ABI of include/edgerunner_b200.h
// by orchestrating the kernels of this directory.  No CPU fallback anywhere: every entry point launches CUDA work
// or fails with an error string.
#include "engine_internal.h"

thread_local char g_err[512] = "";
int g_poison_alloc = 0;
int er_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace {

__global__ void convert_copy_kernel(const void* src, int dtype, int rows, int cols, int src_ld, __half* dst, int dst_ld) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)rows * dst_ld) return;
    const int r = i / dst_ld, c = i % dst_ld;
    float v = 0.f;
    if (c < cols) v = dtype == ER_DTYPE_F16 ? __half2float(((const __half*)src)[(size_t)r * src_ld + c]) : ((const float*)src)[(size_t)r * src_ld + c];
    dst[i] = __float2half_rn(v);
}
__global__ void init_state_kernel(er::DecodeState* st, int L) {
    st->t = 0; st->L = L; st->counter = 0; st->last_tok = 0; st->done = 0;
}
__global__ void finish_decode_kernel(const er::DecodeState* st, int32_t* out_len) { *out_len = st->t; }
// losses[0..2] = {loss, mean CE, KL}; sums[0..2] (optional) = {sum of token CEs, supervised tokens, KL}: what a data-parallel run all-reduces
__global__ void tf_losses_kernel(const double* loss_sum, const int* count, const double* sq_sum, float kl_weight, int has_kl, float* losses, double* sums) {
    const double ce = *loss_sum / (double)max(*count, 1);                    // all in fp64, rounded once: the data-parallel path (dist.py) does the same
    const double kl = has_kl ? 0.5 * *sq_sum : 0.0;
    losses[1] = (float)ce; losses[2] = (float)kl; losses[0] = (float)(ce + (has_kl ? (double)kl_weight * kl : 0.0));
    if (sums) { sums[0] = *loss_sum; sums[1] = (double)*count; sums[2] = has_kl ? 0.5 * *sq_sum : 0.0; }
}

// row-major W[rows][K] -> decode units: unit (row * K/C + q) holds W[row][q*C .. (q+1)*C) followed by (ustride - C) zeros
__global__ void pack_units_kernel(const __half* src, int rows, int K, int C, __half* dst, int ustride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // one 16-byte vector of the destination each
    const int vpu = ustride >> 3;
    const size_t total = (size_t)rows * (K / C) * vpu;
    if (i >= total) return;
    const size_t unit = i / vpu; const int v = i % vpu;
    const size_t row = unit / (K / C); const int q = unit % (K / C);
    uint4 val = make_uint4(0, 0, 0, 0);
    if (v * 8 < C) val = *reinterpret_cast<const uint4*>(src + row * K + (size_t)q * C + v * 8);
    *reinterpret_cast<uint4*>(dst + unit * ustride + v * 8) = val;
}

// fused decode phases (experimental): out_proj re-cut per head, Wo[C][C] -> dst[h][row][0..HD) (+ pad to hs), one unit per (head, row)
__global__ void pack_head_cols_kernel(const __half* wo, int C, int H, int HDim, __half* dst, int hs) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;       // one destination element each
    const size_t total = (size_t)H * C * hs;
    if (i >= total) return;
    const int d = (int)(i % hs); const size_t hr = i / hs;
    const int row = (int)(hr % C), h = (int)(hr / C);
    dst[i] = d < HDim ? wo[(size_t)row * C + (size_t)h * HDim + d] : __float2half_rn(0.f);
}
// fc2 transposed: W2[C][F] -> dst[j][0..C) (+ pad to ustride), one unit per fc2 input column j
__global__ void pack_transposed_kernel(const __half* w2, int C, int F, __half* dst, int ustride) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)F * ustride;
    if (i >= total) return;
    const int c = (int)(i % ustride); const size_t j = i / ustride;
    dst[i] = c < C ? w2[(size_t)c * F + j] : __float2half_rn(0.f);
}

}  // namespace

// Dense (N > 1 rows) workspace: sized for the 2050-row generate prefix at creation and grown on demand, so that a long resume prompt
// (LMM.generate(resume_ids=...), infer.py --test_resume_tokens: up to max_seq_length rows in the reference) is not refused.
int ensure_dense_rows(er_engine* e, int rows) {
    if (rows <= e->maxrows) return ER_OK;
    CK(cudaDeviceSynchronize());
    dev_free(e, &e->x32); dev_free(e, &e->x16); dev_free(e, &e->qkv16); dev_free(e, &e->a16); dev_free(e, &e->h16);
    const bool had_logits = e->logits_all != nullptr;
    dev_free(e, &e->logits_all);
    const int maxrows = rows + 8;
    const int C = e->C, F = e->F;
    e->maxrows = 0;
    ALLOC(e->x32, (size_t)maxrows * C); ALLOC(e->x16, (size_t)maxrows * C); ALLOC(e->qkv16, (size_t)maxrows * 3 * C);
    ALLOC(e->a16, (size_t)maxrows * C); ALLOC(e->h16, (size_t)maxrows * F);
    if (had_logits) {
        ALLOC(e->logits_all, (size_t)maxrows * e->V);
        dev_free(e, &e->tf_rows); dev_free(e, &e->tf_valid);
        ALLOC(e->tf_rows, maxrows); ALLOC(e->tf_valid, maxrows);
    }
    e->maxrows = maxrows;
    return ER_OK;
}

// rows in the KV cache: exact after a decode too (an early EOS leaves fewer rows than max_new_tokens would; read back from the device)
static int sync_cache_rows(er_engine* e) {
    if (e->cache_rows_stale) {
        CK(cudaDeviceSynchronize());
        int L = 0;
        CK(cudaMemcpy(&L, &e->st->L, 4, cudaMemcpyDeviceToHost));
        e->cache_rows = L;
        e->cache_rows_stale = false;
    }
    return ER_OK;
}

extern "C" const char* er_last_error(void) { return g_err; }
extern "C" int er_version(void) { return 100; }

static void add_slot(er_engine* e, const std::string& name, __half* dst, int rows, int cols, int dst_ld = 0) {
    e->slots[name] = Slot{dst, rows, cols, dst_ld ? dst_ld : cols};
}

static int create_impl(er_engine* e, const er_config* cfg) {
    const int C = e->C = cfg->hidden_dim, H = e->H = cfg->num_heads, F = e->F = cfg->ffn_dim, V = e->V = cfg->vocab_size;
    const int NL = e->NL = cfg->num_layers, P = e->P = cfg->num_cond_tokens;
    e->D = C / H;
    if (e->D != 96 || C % 8 || F % 8) { return set_err(ER_ERR_INVALID, "decoder head_dim must be 96 (got %d), C,F multiples of 8", e->D); }
    const int E = e->E = cfg->point_hidden_dim, LQ = e->LQ = cfg->point_latent_size, LD = e->LD = cfg->point_latent_dim;
    e->EH = cfg->point_num_heads;
    e->LDP = (LD + 7) / 8 * 8;
    if (cfg->has_point_encoder && (E / e->EH != 64 || E % 8)) { return set_err(ER_ERR_INVALID, "encoder head_dim must be 64"); }
    if (P != LQ + (cfg->use_num_face_cond ? 1 : 0)) { return set_err(ER_ERR_INVALID, "num_cond_tokens %d != latent_size %d + num_face token", P, LQ); }
    const int Lmax = e->Lmax = (cfg->max_seq_rows + 31) / 32 * 32;
    e->nkb = Lmax / 32;
    if (Lmax > cfg->max_positions) {}  // positions are checked per call
    // ---- weights ----------------------------------------------------------------------------------------------------------
    ALLOC(e->wqkv, (size_t)NL * 3 * C * C); ALLOC(e->bqkv, (size_t)NL * 3 * C);
    ALLOC(e->wo, (size_t)NL * C * C); ALLOC(e->bo, (size_t)NL * C);
    ALLOC(e->ln1w, (size_t)NL * C); ALLOC(e->ln1b, (size_t)NL * C);
    ALLOC(e->w1, (size_t)NL * F * C); ALLOC(e->b1, (size_t)NL * F);
    ALLOC(e->w2, (size_t)NL * C * F); ALLOC(e->b2, (size_t)NL * C);
    ALLOC(e->ln2w, (size_t)NL * C); ALLOC(e->ln2b, (size_t)NL * C);
    ALLOC(e->lm_head, (size_t)V * C); ALLOC(e->embd, (size_t)V * C); ALLOC(e->pos, (size_t)cfg->max_positions * C);
    e->ustride = C + 8;
    e->use_mma = (C % 256 == 0) ? 1 : 0;
    e->upstage = e->use_mma ? 8 : er_decode_stage_bytes() / (e->ustride * 2) / 8 * 8;   // multiple of 8 (and so of F/C)
    ALLOC(e->wdec, ((size_t)NL * (4 * (size_t)C + 2 * (size_t)F) + V) * e->ustride + 64);
    char nm[256];
    for (int i = 0; i < NL; i++) {
        const char* pj[3] = {"q_proj", "k_proj", "v_proj"};
        for (int j = 0; j < 3; j++) {
            snprintf(nm, sizeof nm, "mesh_decoder.model.layers.%d.self_attn.%s.weight", i, pj[j]);
            add_slot(e, nm, e->wqkv + ((size_t)i * 3 + j) * C * C, C, C);
            snprintf(nm, sizeof nm, "mesh_decoder.model.layers.%d.self_attn.%s.bias", i, pj[j]);
            add_slot(e, nm, e->bqkv + ((size_t)i * 3 + j) * C, 1, C);
        }
        auto L = [&](const char* s) { snprintf(nm, sizeof nm, "mesh_decoder.model.layers.%d.%s", i, s); return std::string(nm); };
        add_slot(e, L("self_attn.out_proj.weight"), e->wo + (size_t)i * C * C, C, C);
        add_slot(e, L("self_attn.out_proj.bias"), e->bo + (size_t)i * C, 1, C);
        add_slot(e, L("self_attn_layer_norm.weight"), e->ln1w + (size_t)i * C, 1, C);
        add_slot(e, L("self_attn_layer_norm.bias"), e->ln1b + (size_t)i * C, 1, C);
        add_slot(e, L("fc1.weight"), e->w1 + (size_t)i * F * C, F, C);
        add_slot(e, L("fc1.bias"), e->b1 + (size_t)i * F, 1, F);
        add_slot(e, L("fc2.weight"), e->w2 + (size_t)i * C * F, C, F);
        add_slot(e, L("fc2.bias"), e->b2 + (size_t)i * C, 1, C);
        add_slot(e, L("final_layer_norm.weight"), e->ln2w + (size_t)i * C, 1, C);
        add_slot(e, L("final_layer_norm.bias"), e->ln2b + (size_t)i * C, 1, C);
    }
    add_slot(e, "mesh_decoder.lm_head.weight", e->lm_head, V, C);
    add_slot(e, "mesh_decoder.model.embd.weight", e->embd, V, C);
    add_slot(e, "mesh_decoder.model.embed_positions.weight", e->pos, cfg->max_positions, C);
    ALLOC(e->pc_w, (size_t)C * e->LDP); ALLOC(e->pc_b, C); ALLOC(e->ncw, C); ALLOC(e->ncb, C);
    add_slot(e, "proj_cond.weight", e->pc_w, C, LD, e->LDP);
    add_slot(e, "proj_cond.bias", e->pc_b, 1, C);
    add_slot(e, "norm_cond.weight", e->ncw, 1, C);
    add_slot(e, "norm_cond.bias", e->ncb, 1, C);
    if (cfg->use_num_face_cond) { ALLOC(e->enf, (size_t)10 * C); add_slot(e, "embed_num_face.weight", e->enf, 10, C); }
    if (cfg->has_point_encoder) {
        ALLOC(e->qe, (size_t)LQ * E); ALLOC(e->basis, 72); ALLOC(e->mlp_w, (size_t)E * 64); ALLOC(e->mlp_b, E);
        ALLOC(e->ln_w, E); ALLOC(e->ln_b, E); ALLOC(e->cl1w, E); ALLOC(e->cl1b, E); ALLOC(e->cl2w, E); ALLOC(e->cl2b, E);
        ALLOC(e->cq_w, (size_t)E * E); ALLOC(e->cq_b, E); ALLOC(e->ckv_w, (size_t)2 * E * E); ALLOC(e->ckv_b, 2 * E);
        ALLOC(e->co_w, (size_t)E * E); ALLOC(e->co_b, E);
        ALLOC(e->ff0w, (size_t)8 * E * E); ALLOC(e->ff0b, 8 * E); ALLOC(e->ff2w, (size_t)4 * E * E); ALLOC(e->ff2b, E);
        ALLOC(e->lin_w, (size_t)e->LDP * E); ALLOC(e->lin_b, e->LDP);
        CK(cudaMemset(e->lin_w, 0, (size_t)e->LDP * E * 2)); CK(cudaMemset(e->lin_b, 0, e->LDP * 2));
        const std::string pe = "point_encoder.";
        add_slot(e, pe + "query_embed", e->qe, LQ, E);
        add_slot(e, pe + "point_embed.basis", e->basis, 3, 24);
        add_slot(e, pe + "point_embed.mlp.weight", e->mlp_w, E, 51, 64);
        add_slot(e, pe + "point_embed.mlp.bias", e->mlp_b, 1, E);
        add_slot(e, pe + "ln.weight", e->ln_w, 1, E); add_slot(e, pe + "ln.bias", e->ln_b, 1, E);
        add_slot(e, pe + "cross_att.ln1.weight", e->cl1w, 1, E); add_slot(e, pe + "cross_att.ln1.bias", e->cl1b, 1, E);
        add_slot(e, pe + "cross_att.ln2.weight", e->cl2w, 1, E); add_slot(e, pe + "cross_att.ln2.bias", e->cl2b, 1, E);
        add_slot(e, pe + "cross_att.att.q_proj.weight", e->cq_w, E, E); add_slot(e, pe + "cross_att.att.q_proj.bias", e->cq_b, 1, E);
        add_slot(e, pe + "cross_att.att.k_proj.weight", e->ckv_w, E, E); add_slot(e, pe + "cross_att.att.k_proj.bias", e->ckv_b, 1, E);
        add_slot(e, pe + "cross_att.att.v_proj.weight", e->ckv_w + (size_t)E * E, E, E); add_slot(e, pe + "cross_att.att.v_proj.bias", e->ckv_b + E, 1, E);
        add_slot(e, pe + "cross_att.att.out_proj.weight", e->co_w, E, E); add_slot(e, pe + "cross_att.att.out_proj.bias", e->co_b, 1, E);
        add_slot(e, pe + "cross_att.mlp.net.0.weight", e->ff0w, 8 * E, E); add_slot(e, pe + "cross_att.mlp.net.0.bias", e->ff0b, 1, 8 * E);
        add_slot(e, pe + "cross_att.mlp.net.2.weight", e->ff2w, E, 4 * E); add_slot(e, pe + "cross_att.mlp.net.2.bias", e->ff2b, 1, E);
        add_slot(e, pe + "linear.weight", e->lin_w, LD, E); add_slot(e, pe + "linear.bias", e->lin_b, 1, LD);
    }
    // ---- KV cache + decode scratch ---------------------------------------------------------------------------------------------
    ALLOC(e->kc, (size_t)NL * H * e->nkb * 32 * 96); ALLOC(e->vc, (size_t)NL * H * Lmax * 96);
    ALLOC(e->q16, 3 * (size_t)C); ALLOC(e->y1, 8 * (size_t)C); ALLOC(e->h1, 8 * (size_t)F); ALLOC(e->y2, 8 * (size_t)C); ALLOC(e->attn16, 8 * (size_t)C);
    ALLOC(e->logits, V); ALLOC(e->st, 1); ALLOC(e->bar, 64 + 64); ALLOC(e->cond32, (size_t)P * C);
    ALLOC(e->ids_dev, 65536); ALLOC(e->gen_ids_dev, cfg->max_seq_rows + 8); ALLOC(e->gen_len_dev, 4);
    ALLOC(e->conds_dev_buf, (size_t)(cfg->max_points > LQ * LD ? cfg->max_points * 3 : LQ * LD) + 16);
    ALLOC(e->prof, 8192);
    CK(cudaMemset(e->prof, 0, 8192 * 8));
    CK(cudaMemset(e->st, 0, sizeof(er::DecodeState)));
    // decode launch geometry
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device));
    e->grid = sms;
    e->S = sms / H; if (e->S > 16) e->S = 16; if (e->S < 1) { return set_err(ER_ERR_INVALID, "need >= num_heads SMs"); }
    ALLOC(e->part, (size_t)H * 16 * 100);
    e->ll_words = (size_t)3 * C / 2 + (size_t)H * 16 * 100 + (size_t)sms * C;       // flagged words: q|k|v [3C/2], split partials [H][16][100], group-of-4 exchange [grid][4][C/4]
    ALLOC(e->ll, e->ll_words);
    ALLOC(e->acc, 4 * (size_t)C);                                   // four copies of C counting accumulators
    // tensor-parallel layer: S in {12, 9, 6} so that a CTA's 288 / S qkv rows stay inside one of q | k | v; needs the tensor-core
    // GEMV shapes (C % 256), <= 64 fc1 rows and <= 32 accumulator words per CTA
    e->S_fuse = 0;
    for (int cand : {12, 9, 6}) if (cand * H <= sms) { e->S_fuse = cand; break; }
    if (C % 256 || (F + sms - 1) / sms + 1 > 64 || (C + sms - 1) / sms + 1 > 32 || (C & 1) || sms > 254) e->S_fuse = 0;
    if (!e->S_fuse) e->use_fuse = 0;
    e->sc_len = er::score_scratch_len(e->nkb, e->S, V);
    if (e->S_fuse) e->sc_len = std::max(e->sc_len, er::score_scratch_len(e->nkb, e->S_fuse, V));
    {
        er::DecodeParams p{}; p.C = C; p.F = F; p.H = H; p.V = V; p.S = e->S; p.sc_len = e->sc_len;
        int smem_max = 0;
        CK(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device));
        e->nstage = er_decode_pick_stages(p, (size_t)smem_max - 1024);
        if (e->nstage < 2) { return set_err(ER_ERR_CAPACITY, "not enough shared memory for the decode ring (limit %d)", smem_max); }
        p.nstage = e->nstage;
        e->dec_smem = er_decode_smem_bytes(p);
        if (er_decode_max_grid(e->dec_smem) < sms) { return set_err(ER_ERR_CAPACITY, "decode kernel cannot be co-resident on %d SMs (smem %zu)", sms, e->dec_smem); }
    }
    // ---- dense workspace ----------------------------------------------------------------------------------------------------------
    const int maxrows = e->maxrows = std::max(P + 64, cfg->max_tf_rows) + 8;
    ALLOC(e->x32, (size_t)maxrows * C); ALLOC(e->x16, (size_t)maxrows * C); ALLOC(e->qkv16, (size_t)maxrows * 3 * C);
    ALLOC(e->a16, (size_t)maxrows * C); ALLOC(e->h16, (size_t)maxrows * F);
    e->logits_all = nullptr;
    if (cfg->max_tf_rows > 0) ALLOC(e->logits_all, (size_t)maxrows * V);
    ALLOC(e->tf_acc, 4); ALLOC(e->tf_cnt, 4); ALLOC(e->tf_part, 296);
    if (cfg->max_tf_rows > 0) { ALLOC(e->tf_rows, maxrows); ALLOC(e->tf_valid, maxrows); }
    e->lat_batch_cap = cfg->max_tf_rows > 0 ? std::max(1, cfg->max_tf_rows / (P + 2)) : 1;
    ALLOC(e->lat16, (size_t)e->lat_batch_cap * LQ * e->LDP); ALLOC(e->pc16, (size_t)LQ * C);
    if (cfg->has_point_encoder) {
        const int np = cfg->max_points;
        ALLOC(e->emb16, (size_t)np * 64); ALLOC(e->pf16, (size_t)np * E); ALLOC(e->kvx16, (size_t)np * E); ALLOC(e->kvo16, (size_t)np * 2 * E);
        ALLOC(e->qln16, (size_t)LQ * E); ALLOC(e->qq16, (size_t)LQ * E); ALLOC(e->ea16, (size_t)LQ * E); ALLOC(e->ex1, (size_t)LQ * E);
        ALLOC(e->ex1ln, (size_t)LQ * E); ALLOC(e->eff, (size_t)LQ * 8 * E); ALLOC(e->egg, (size_t)LQ * 4 * E); ALLOC(e->ex2, (size_t)LQ * E);
    }
    CK(cudaDeviceSynchronize());
    return ER_OK;
}

extern "C" int er_create(const er_config* cfg, er_engine** out) {
    if (!cfg || !out) return set_err(ER_ERR_INVALID, "null argument");
    CK(cudaSetDevice(cfg->device));
    er_engine* e = new er_engine();
    e->cfg = *cfg;
    const int r = create_impl(e, cfg);
    if (r) { er_destroy(e); return r; }      // frees every device allocation recorded so far (g_err keeps the message)
    *out = e;
    return ER_OK;
}

extern "C" void er_destroy(er_engine* e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    cudaDeviceSynchronize();
    er_train_destroy(e);
    for (void* p : e->allocs) cudaFree(p);
    delete e;
}

extern "C" int er_load_weight(er_engine* e, const char* name, const void* data_dev, int32_t dtype, const int64_t* shape, int32_t ndim, void* stream) {
    if (!e || !name || !data_dev) return set_err(ER_ERR_INVALID, "null argument");
    auto it = e->slots.find(name);
    if (it == e->slots.end()) return set_err(ER_ERR_INVALID, "unknown state-dict key '%s'", name);
    const Slot& s = it->second;
    int64_t n = 1;
    for (int i = 0; i < ndim; i++) n *= shape[i];
    if (n != (int64_t)s.rows * s.cols) return set_err(ER_ERR_INVALID, "shape mismatch for '%s': %lld elements, expected %d x %d", name, (long long)n, s.rows, s.cols);
    if (dtype != ER_DTYPE_F16 && dtype != ER_DTYPE_F32) return set_err(ER_ERR_INVALID, "dtype");
    const size_t total = (size_t)s.rows * s.dst_ld;
    e->launches++;
    convert_copy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(data_dev, dtype, s.rows, s.cols, s.cols, s.dst, s.dst_ld);
    CK(cudaGetLastError());
    e->loaded.insert(name);
    e->finalized = false;
    return ER_OK;
}

extern "C" int er_finalize_weights(er_engine* e, void* stream) {
    if (!e) return set_err(ER_ERR_INVALID, "null engine");
    for (auto& kv : e->slots)
        if (!e->loaded.count(kv.first)) return set_err(ER_ERR_STATE, "weight '%s' was never loaded", kv.first.c_str());
    // decode-stream copy: pack every decoder matrix into padded units (see decode_kernel.cu)
    {
        cudaStream_t st = (cudaStream_t)stream;
        const int C = e->C, F = e->F, us = e->ustride;
        const size_t UL = (size_t)4 * C + 2 * (size_t)F;
        auto pack = [&](const __half* src, int rows, int K, size_t unit0) -> cudaError_t {
            const size_t total = (size_t)rows * (K / C) * (us >> 3);
            e->launches++;
            pack_units_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, rows, K, C, e->wdec + unit0 * us, us);
            return cudaGetLastError();
        };
        for (int l = 0; l < e->NL; l++) {
            CK(pack(e->wqkv + (size_t)l * 3 * C * C, 3 * C, C, l * UL));
            CK(pack(e->wo + (size_t)l * C * C, C, C, l * UL + 3 * (size_t)C));
            CK(pack(e->w1 + (size_t)l * F * C, F, C, l * UL + 4 * (size_t)C));
            CK(pack(e->w2 + (size_t)l * C * F, C, F, l * UL + 4 * (size_t)C + F));
        }
        CK(pack(e->lm_head, e->V, C, (size_t)e->NL * UL));
        if (e->use_fuse) {
            if (!e->wfuse) ALLOC(e->wfuse, (size_t)e->NL * ((size_t)e->H * C * 104 + (size_t)F * us));
            const size_t per_layer = (size_t)e->H * C * 104 + (size_t)F * us;
            for (int l = 0; l < e->NL; l++) {
                __half* dst = e->wfuse + (size_t)l * per_layer;
                const size_t n1 = (size_t)e->H * C * 104, n2 = (size_t)F * us;
                e->launches += 2;
                pack_head_cols_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(e->wo + (size_t)l * C * C, C, e->H, 96, dst, 104);
                CK(cudaGetLastError());
                pack_transposed_kernel<<<(unsigned)((n2 + 255) / 256), 256, 0, st>>>(e->w2 + (size_t)l * C * F, C, F, dst + n1, us);
                CK(cudaGetLastError());
            }
        }
    }
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    e->finalized = true;
    return ER_OK;
}

er::GemmArgs mk_gemm(const __half* A, int lda, const __half* W, int ldw, const __half* bias, int M, int N, int K, int mode) {
    er::GemmArgs g{};
    g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.M = M; g.N = N; g.K = K; g.mode = mode;
    return g;
}

// PointEncoderEmbed.forward for one cloud -> lat16 [LQ][LDP] (point.py:186-206)
int encode_points(er_engine* e, const float* pts, int n, __half* lat, cudaStream_t st) {
    const int E = e->E, LQ = e->LQ, EH = e->EH;
    if (n > e->cfg.max_points) return set_err(ER_ERR_CAPACITY, "n_points %d > max_points %d", n, e->cfg.max_points);
    CKL(e, er_point_embed(pts, e->basis, e->emb16, 64, n, st));
    er::GemmArgs g = mk_gemm(e->emb16, 64, e->mlp_w, 64, e->mlp_b, n, E, 64, er::GEMM_F16);
    g.out16 = e->pf16; g.ldo = E; CKL(e, er_gemm(g, st));
    CKL(e, er_layernorm(nullptr, e->pf16, E, e->ln_w, e->ln_b, nullptr, e->kvx16, E, n, E, st));
    CKL(e, er_layernorm(nullptr, e->qe, E, e->cl1w, e->cl1b, nullptr, e->qln16, E, LQ, E, st));
    g = mk_gemm(e->qln16, E, e->cq_w, E, e->cq_b, LQ, E, E, er::GEMM_F16); g.out16 = e->qq16; g.ldo = E; CKL(e, er_gemm(g, st));
    g = mk_gemm(e->kvx16, E, e->ckv_w, E, e->ckv_b, n, 2 * E, E, er::GEMM_F16); g.out16 = e->kvo16; g.ldo = 2 * E; CKL(e, er_gemm(g, st));
    er::AttnArgs a{};
    a.q = e->qq16; a.k = e->kvo16; a.v = e->kvo16 + E; a.out = e->ea16;
    a.ldq = E; a.ldk = 2 * E; a.ldv = 2 * E; a.ldo = E; a.B = 1; a.H = EH; a.Nq = LQ; a.Nk = n; a.D = 64; a.causal = 0;
    CKL(e, er_attention(a, st));
    g = mk_gemm(e->ea16, E, e->co_w, E, e->co_b, LQ, E, E, er::GEMM_F16_RES16); g.out16 = e->ex1; g.ldo = E; g.res16 = e->qe; g.ldr = E; CKL(e, er_gemm(g, st));
    CKL(e, er_layernorm(nullptr, e->ex1, E, e->cl2w, e->cl2b, nullptr, e->ex1ln, E, LQ, E, st));
    g = mk_gemm(e->ex1ln, E, e->ff0w, E, e->ff0b, LQ, 8 * E, E, er::GEMM_F16); g.out16 = e->eff; g.ldo = 8 * E; CKL(e, er_gemm(g, st));
    CKL(e, er_geglu(e->eff, e->egg, LQ, 4 * E, st));
    g = mk_gemm(e->egg, 4 * E, e->ff2w, 4 * E, e->ff2b, LQ, E, 4 * E, er::GEMM_F16_RES16); g.out16 = e->ex2; g.ldo = E; g.res16 = e->ex1; g.ldr = E; CKL(e, er_gemm(g, st));
    g = mk_gemm(e->ex2, E, e->lin_w, E, e->lin_b, LQ, e->LDP, E, er::GEMM_F16); g.out16 = lat; g.ldo = e->LDP; CKL(e, er_gemm(g, st));
    return ER_OK;
}

// latents -> cond32 [P][C]: norm_cond(proj_cond(lat)) ++ embed_num_face[bucket]   (models.py:124,135-141)
int er_quantize_num_faces(int n) { return n <= 0 ? 0 : n <= 1000 ? 1 : n <= 2000 ? 2 : n <= 4000 ? 3 : n <= 8000 ? 4 : 5; }
static int latents_to_cond(er_engine* e, const __half* lat, int num_faces, float* cond32, cudaStream_t st) {
    const int C = e->C, LQ = e->LQ;
    er::GemmArgs g = mk_gemm(lat, e->LDP, e->pc_w, e->LDP, e->pc_b, LQ, C, e->LDP, er::GEMM_F16);
    g.out16 = e->pc16; g.ldo = C; CKL(e, er_gemm(g, st));
    CKL(e, er_layernorm(nullptr, e->pc16, C, e->ncw, e->ncb, cond32, nullptr, C, LQ, C, st));
    if (e->cfg.use_num_face_cond) CKL(e, er_f16_to_f32(e->enf + (size_t)er_quantize_num_faces(num_faces) * C, cond32 + (size_t)LQ * C, C, st));
    return ER_OK;
}

int encode_one(er_engine* e, const float* conds_dev, int n_points, int is_latent, int num_faces, __half* lat, float* cond32, cudaStream_t st) {
    if (is_latent) {
        if (n_points != e->LQ) return set_err(ER_ERR_INVALID, "latent rows %d != latent_size %d", n_points, e->LQ);
        e->launches++;
        convert_copy_kernel<<<(unsigned)(((size_t)e->LQ * e->LDP + 255) / 256), 256, 0, st>>>(conds_dev, ER_DTYPE_F32, e->LQ, e->LD, e->LD, lat, e->LDP);
        CK(cudaGetLastError());
    } else {
        if (!e->cfg.has_point_encoder) return set_err(ER_ERR_STATE, "engine was created without a point encoder");
        int r = encode_points(e, conds_dev, n_points, lat, st);
        if (r) return r;
    }
    return latents_to_cond(e, lat, num_faces, cond32, st);
}

extern "C" int er_encode_cond(er_engine* e, const float* conds_dev, int32_t n_points, int32_t is_latent, int32_t num_faces,
                              float* cond_embeds_out_dev, void* latents_out_dev, void* stream) {
    if (!e || !conds_dev) return set_err(ER_ERR_INVALID, "null argument");
    if (!e->finalized) return set_err(ER_ERR_STATE, "weights not finalized");
    cudaStream_t st = (cudaStream_t)stream;
    int r = encode_one(e, conds_dev, n_points, is_latent, num_faces, e->lat16, e->cond32, st);
    if (r) return r;
    if (cond_embeds_out_dev) CK(cudaMemcpyAsync(cond_embeds_out_dev, e->cond32, (size_t)e->P * e->C * 4, cudaMemcpyDeviceToDevice, st));
    if (latents_out_dev) CK(cudaMemcpy2DAsync(latents_out_dev, e->LD * 2, e->lat16, e->LDP * 2, e->LD * 2, e->LQ, cudaMemcpyDeviceToDevice, st));
    return ER_OK;
}

// 24 x OPTDecoderLayer on M = B*N rows held in x32/x16 (modeling_opt.py:264-288, 185-232); store_kv: fill the decode cache
static int decoder_layers(er_engine* e, int B, int N, bool store_kv, cudaStream_t st, const unsigned char* row_mask = nullptr) {
    const int C = e->C, F = e->F, H = e->H, M = B * N;
    for (int l = 0; l < e->NL; l++) {
        er::GemmArgs g = mk_gemm(e->x16, C, e->wqkv + (size_t)l * 3 * C * C, C, e->bqkv + (size_t)l * 3 * C, M, 3 * C, C, er::GEMM_F16);
        g.out16 = e->qkv16; g.ldo = 3 * C; CKL(e, er_gemm(g, st));
        if (store_kv) CKL(e, er_kv_store(e->qkv16, N, C, H, l, 0, e->Lmax, e->nkb, e->kc, e->vc, st));
        er::AttnArgs a{};
        a.q = e->qkv16; a.k = e->qkv16 + C; a.v = e->qkv16 + 2 * C; a.out = e->a16;
        a.ldq = a.ldk = a.ldv = 3 * C; a.ldo = C; a.q_bs = a.k_bs = a.v_bs = (long long)N * 3 * C; a.o_bs = (long long)N * C;
        a.B = B; a.H = H; a.Nq = N; a.Nk = N; a.D = 96; a.causal = 1;
        CKL(e, er_attention(a, st));
        if (row_mask) CKL(e, er_zero_masked_rows(e->a16, row_mask, M, C, st));      // flash_attn pad_input: masked rows come back as zeros
        g = mk_gemm(e->a16, C, e->wo + (size_t)l * C * C, C, e->bo + (size_t)l * C, M, C, C, er::GEMM_F32_RES32);
        g.out32 = e->x32; g.ldo = C; g.res32 = e->x32; g.ldr = C; CKL(e, er_gemm(g, st));
        CKL(e, er_layernorm(e->x32, nullptr, C, e->ln1w + (size_t)l * C, e->ln1b + (size_t)l * C, e->x32, e->x16, C, M, C, st));
        g = mk_gemm(e->x16, C, e->w1 + (size_t)l * F * C, C, e->b1 + (size_t)l * F, M, F, C, er::GEMM_F16_RELU);
        g.out16 = e->h16; g.ldo = F; CKL(e, er_gemm(g, st));
        g = mk_gemm(e->h16, F, e->w2 + (size_t)l * C * F, F, e->b2 + (size_t)l * C, M, C, F, er::GEMM_F32_RES32);
        g.out32 = e->x32; g.ldo = C; g.res32 = e->x32; g.ldr = C; CKL(e, er_gemm(g, st));
        CKL(e, er_layernorm(e->x32, nullptr, C, e->ln2w + (size_t)l * C, e->ln2b + (size_t)l * C, e->x32, e->x16, C, M, C, st));
    }
    return ER_OK;
}

extern "C" int er_prefill(er_engine* e, const int32_t* prompt_ids_host, int32_t n_prompt, void* stream) {
    if (!e || !prompt_ids_host || n_prompt < 1) return set_err(ER_ERR_INVALID, "bad prompt");
    if (!e->finalized) return set_err(ER_ERR_STATE, "weights not finalized");
    cudaStream_t st = (cudaStream_t)stream;
    const int N = e->P + n_prompt, C = e->C;
    if (N > e->Lmax || N > e->cfg.max_positions || n_prompt > 65536) return set_err(ER_ERR_CAPACITY, "prefix of %d rows exceeds the cache capacity %d", N, e->Lmax);
    { int r = ensure_dense_rows(e, N); if (r) return r; }
    for (int i = 0; i < n_prompt; i++)
        if (prompt_ids_host[i] < 0 || prompt_ids_host[i] >= e->V) return set_err(ER_ERR_INVALID, "prompt id %d out of range", prompt_ids_host[i]);
    CK(cudaMemcpyAsync(e->ids_dev, prompt_ids_host, (size_t)n_prompt * 4, cudaMemcpyHostToDevice, st));
    CKL(e, er_embed_prefix(e->cond32, e->P, e->ids_dev, n_prompt, e->embd, e->pos, C, e->x32, e->x16, st));
    int r = decoder_layers(e, 1, N, true, st);
    if (r) return r;
    er::GemmArgs g = mk_gemm(e->x16 + (size_t)(N - 1) * C, C, e->lm_head, C, nullptr, 1, e->V, C, er::GEMM_F32);
    g.out32 = e->logits; g.ldo = e->V; CKL(e, er_gemm(g, st));
    e->launches++;
    init_state_kernel<<<1, 1, 0, st>>>(e->st, N);
    CK(cudaGetLastError());
    e->cache_rows = N;
    e->cache_rows_stale = false;
    return ER_OK;
}

extern "C" int er_decode(er_engine* e, int32_t max_new_tokens, int32_t mode, int32_t top_k, uint64_t seed, int32_t use_tokenizer_fsm,
                         int32_t tokens_per_launch, int32_t* out_ids_dev, int32_t* out_len_dev, float* out_logits_dev,
                         const int32_t* forced_ids_dev, void* stream) {
    if (!e || !out_ids_dev || max_new_tokens < 1) return set_err(ER_ERR_INVALID, "bad argument");
    if (e->cache_rows <= 0) return set_err(ER_ERR_STATE, "er_prefill has not run");
    { int r = sync_cache_rows(e); if (r) return r; }
    if (e->cache_rows + max_new_tokens > e->Lmax || e->cache_rows + max_new_tokens > e->cfg.max_positions)
        return set_err(ER_ERR_CAPACITY, "cache rows %d + max_new_tokens %d exceed capacity %d", e->cache_rows, max_new_tokens, e->Lmax);
    if (e->dbg_nosync && !forced_ids_dev) return set_err(ER_ERR_STATE, "debug 'nosync' produces garbage logits: a forced token stream is required");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = e->C, F = e->F, V = e->V, G = e->grid;
    er::DecodeParams p{};
    const bool fuse = e->use_fuse && e->wfuse != nullptr && e->S_fuse > 0;
    p.C = C; p.H = e->H; p.F = F; p.V = V; p.layers = e->NL; p.S = fuse ? e->S_fuse : e->S; p.Lmax = e->Lmax; p.nkb = e->nkb;
    p.nstage = e->nstage; p.sc_len = e->sc_len;
    // red_units holds 2 partial sums per unit of C elements: the widest phases are qkv (rows) and fc2 (rows * F/C)
    const int max_units = std::max(std::max((3 * C + G - 1) / G + 1, (F + G - 1) / G + 1), ((C + G - 1) / G + 1) * (F / C));
    if (max_units > 64 || F % C || C % 16 || C > 1536 || (8 % (F / C)) || (e->upstage % (F / C)) || e->H > 64) return set_err(ER_ERR_CAPACITY, "model shape not supported by the decode kernel on %d SMs (units %d)", G, max_units);
    p.wqkv = e->wqkv; p.bqkv = e->bqkv; p.wo = e->wo; p.bo = e->bo; p.ln1_w = e->ln1w; p.ln1_b = e->ln1b;
    p.w1 = e->w1; p.b1 = e->b1; p.w2 = e->w2; p.b2 = e->b2; p.ln2_w = e->ln2w; p.ln2_b = e->ln2b;
    p.lm_head = e->lm_head; p.embd = e->embd; p.pos = e->pos;
    p.wdec = e->wdec; p.ustride = e->ustride; p.upstage = e->upstage; p.use_mma = e->use_mma;
    // the last KV split also owns the new key; in the five-exchange layer it usually merges the head as well (worth ~4 blocks of streaming), in
    // the tensor-parallel layer every split merges, so the splits are even
    p.split_handicap = fuse ? e->split_handicap_fuse : e->split_handicap;
    p.kc = e->kc; p.vc = e->vc; p.attn16 = e->attn16; p.head_cnt = e->bar + 64; p.q16 = e->q16; p.y1 = e->y1; p.h1 = e->h1; p.y2 = e->y2; p.part = e->part; p.logits = e->logits;
    p.ll_q = e->ll; p.ll_part = p.ll_q + 3 * C / 2;
    p.poll_rounds = e->poll_rounds;
    p.use_fuse = fuse; p.wfuse = e->wfuse; p.acc = e->acc;
    p.xq = p.ll_part + (size_t)e->H * 16 * 100; p.red_group = (G % 4 == 0 && C % 4 == 0 && e->red_group4) ? 4 : 1;
    p.xrep = e->xrep;
    p.pf_dist = e->pf_dist; p.dbg_nosync = e->dbg_nosync;
    p.st = e->st; p.bar = e->bar;
    p.out_ids = out_ids_dev; p.out_logits = out_logits_dev; p.forced = forced_ids_dev;
    p.max_new = max_new_tokens; p.mode = mode; p.top_k = top_k > 0 ? top_k : 10; p.use_fsm = use_tokenizer_fsm; p.eos = e->cfg.eos_token_id;
    p.seed = seed;
    p.prof = e->prof_token >= 0 ? e->prof : nullptr; p.prof_token = e->prof_token; p.prof_cta = e->prof_cta;
    const int chunk = tokens_per_launch > 0 ? tokens_per_launch : max_new_tokens;
    for (int done = 0; done < max_new_tokens; done += chunk) {
        p.steps = std::min(chunk, max_new_tokens - done);
        CK(cudaMemsetAsync(e->bar, 0, (64 + 64) * 4, st));
        if (p.use_fuse) { CK(cudaMemsetAsync(e->ll, 0, e->ll_words * 8, st)); CK(cudaMemsetAsync(e->acc, 0, 4 * (size_t)C * 8, st)); }
        CKL(e, er_decode_launch(p, G, e->dec_smem, st));
    }
    if (out_len_dev) {
        e->launches++;
        finish_decode_kernel<<<1, 1, 0, st>>>(e->st, out_len_dev);
        CK(cudaGetLastError());
    }
    e->cache_rows += max_new_tokens - 1;   // upper bound until the next query reads the exact value back from the device state
    e->cache_rows_stale = true;
    return ER_OK;
}

extern "C" int er_generate_host(er_engine* e, const float* conds_host, int32_t n_points, int32_t is_latent, int32_t num_faces,
                                const int32_t* resume_ids_host, int32_t n_resume, int32_t max_new_tokens, int32_t mode, int32_t top_k,
                                uint64_t seed, int32_t use_tokenizer_fsm, int32_t* out_ids_host, int32_t* out_len_host) {
    if (!e || !conds_host || !out_ids_host || !out_len_host) return set_err(ER_ERR_INVALID, "null argument");
    cudaStream_t st = 0;
    const size_t nfl = is_latent ? (size_t)e->LQ * e->LD : (size_t)n_points * 3;
    if (!is_latent && n_points > e->cfg.max_points) return set_err(ER_ERR_CAPACITY, "n_points");
    CK(cudaMemcpyAsync(e->conds_dev_buf, conds_host, nfl * 4, cudaMemcpyHostToDevice, st));
    int r = er_encode_cond(e, e->conds_dev_buf, n_points, is_latent, num_faces, nullptr, nullptr, st);
    if (r) return r;
    std::vector<int32_t> prompt;
    prompt.push_back(e->cfg.bos_token_id);
    for (int i = 0; i < n_resume; i++) prompt.push_back(resume_ids_host[i]);
    r = er_prefill(e, prompt.data(), (int)prompt.size(), st);
    if (r) return r;
    if (max_new_tokens > e->cfg.max_seq_rows) return set_err(ER_ERR_CAPACITY, "max_new_tokens");
    r = er_decode(e, max_new_tokens, mode, top_k, seed, use_tokenizer_fsm, 0, e->gen_ids_dev, e->gen_len_dev, nullptr, nullptr, st);
    if (r) return r;
    CK(cudaMemcpyAsync(out_len_host, e->gen_len_dev, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(out_ids_host, e->gen_ids_dev, (size_t)(*out_len_host) * 4, cudaMemcpyDeviceToHost));
    return ER_OK;
}

extern "C" int er_forward_tf2(er_engine* e, const float* conds_dev, int32_t n_points, int32_t is_latent, const int32_t* tokens_dev,
                              const int64_t* labels_dev, const uint8_t* mask_dev, const int32_t* num_faces_host, int32_t B, int32_t T, float kl_weight,
                              float* losses_dev, double* sums_dev, float* logits_out_dev, void* stream) {
    if (!e || !conds_dev || !tokens_dev || !labels_dev || !num_faces_host || !losses_dev) return set_err(ER_ERR_INVALID, "null argument");
    if (!e->finalized) return set_err(ER_ERR_STATE, "weights not finalized");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = e->C, P = e->P, N = P + T, M = B * N, V = e->V;
    if (!e->logits_all || B > e->lat_batch_cap) return set_err(ER_ERR_CAPACITY, "engine was created with max_tf_rows too small for a batch of %d samples", B);
    { int r0 = ensure_dense_rows(e, M); if (r0) return r0; }
    if (N > e->cfg.max_positions) return set_err(ER_ERR_CAPACITY, "sequence longer than the position table");
    const size_t cstride = is_latent ? (size_t)e->LQ * e->LD : (size_t)n_points * 3;
    for (int b = 0; b < B; b++) {
        // cond rows go straight into x32 rows [b*N, b*N+P); embed_prefix then adds positions in place
        float* cond_rows = e->x32 + (size_t)b * N * C;
        int r = encode_one(e, conds_dev + b * cstride, n_points, is_latent, num_faces_host[b], e->lat16 + (size_t)b * e->LQ * e->LDP, cond_rows, st);
        if (r) return r;
        CKL(e, er_embed_prefix(cond_rows, P, tokens_dev + (size_t)b * T, T, e->embd, e->pos, C, cond_rows, e->x16 + (size_t)b * N * C, st));
    }
    int r = decoder_layers(e, B, N, false, st, mask_dev);
    if (r) return r;
    er::GemmArgs g = mk_gemm(e->x16, C, e->lm_head, C, nullptr, M, V, C, er::GEMM_F32);
    g.out32 = e->logits_all; g.ldo = V; CKL(e, er_gemm(g, st));
    CK(cudaMemsetAsync(e->tf_acc, 0, 32, st));
    CK(cudaMemsetAsync(e->tf_cnt, 0, 16, st));
    for (int b = 0; b < B; b++) {       // shifted: row i of sample b predicts label i + 1 (modeling_opt.py:500-505); samples in order, fixed reduction order
        e->launches++;
        CKL(e, er_cross_entropy(e->logits_all + (size_t)b * N * V, V, labels_dev + (size_t)b * N + 1, N - 1, V, e->tf_rows, e->tf_valid, e->tf_acc, e->tf_cnt, st));
    }
    const int has_kl = !is_latent;
    if (has_kl) { e->launches++; CKL(e, er_sum_squares(e->lat16, (size_t)B * e->LQ * e->LDP, e->tf_part, e->tf_acc + 1, st)); }
    e->launches++;
    tf_losses_kernel<<<1, 1, 0, st>>>(e->tf_acc, e->tf_cnt, e->tf_acc + 1, kl_weight, has_kl, losses_dev, sums_dev);
    CK(cudaGetLastError());
    if (logits_out_dev) CK(cudaMemcpyAsync(logits_out_dev, e->logits_all, (size_t)M * V * 4, cudaMemcpyDeviceToDevice, st));
    return ER_OK;
}

extern "C" int er_forward_tf(er_engine* e, const float* conds_dev, int32_t n_points, int32_t is_latent, const int32_t* tokens_dev,
                             const int64_t* labels_dev, const int32_t* num_faces_host, int32_t B, int32_t T, float kl_weight,
                             float* losses_dev, float* logits_out_dev, void* stream) {
    return er_forward_tf2(e, conds_dev, n_points, is_latent, tokens_dev, labels_dev, nullptr, num_faces_host, B, T, kl_weight, losses_dev, nullptr,
                          logits_out_dev, stream);
}

extern "C" int64_t er_weight_bytes_per_token(const er_engine* e) {
    const int64_t C = e->C, F = e->F;
    const int64_t per_layer = 4 * (C * C + C) + F * C + F + C * F + C + 4 * C;
    return 2 * (per_layer * e->NL + (int64_t)e->V * C);
}
extern "C" int64_t er_kv_bytes_per_row(const er_engine* e) { return (int64_t)e->NL * 2 * e->C * 2; }
extern "C" int32_t er_cache_rows(const er_engine* e) { sync_cache_rows(const_cast<er_engine*>(e)); return e->cache_rows; }
extern "C" int64_t er_kernel_launches(const er_engine* e) { return e->launches; }

extern "C" int er_attention_bnhd(const void* q_dev, const void* k_dev, const void* v_dev, void* out_dev, int32_t B, int32_t Nq, int32_t Nk,
                                 int32_t H, int32_t D, int32_t causal, void* stream) {
    if (!q_dev || !k_dev || !v_dev || !out_dev) return set_err(ER_ERR_INVALID, "null argument");
    if (D != 64 && D != 96) return set_err(ER_ERR_INVALID, "head_dim %d not supported (64 or 96)", D);
    if (causal && Nq != Nk) return set_err(ER_ERR_INVALID, "causal attention needs Nq == Nk");
    er::AttnArgs a{};
    a.q = (const __half*)q_dev; a.k = (const __half*)k_dev; a.v = (const __half*)v_dev; a.out = (__half*)out_dev;
    a.ldq = a.ldk = a.ldv = a.ldo = H * D;
    a.q_bs = a.o_bs = (long long)Nq * H * D; a.k_bs = a.v_bs = (long long)Nk * H * D;
    a.B = B; a.H = H; a.Nq = Nq; a.Nk = Nk; a.D = D; a.causal = causal;
    CK(er_attention(a, (cudaStream_t)stream));
    return ER_OK;
}

// Experiment / diagnostic switches of the decode kernel.  The product path never reads the environment: everything that is not the
// production configuration is set explicitly through this entry point (scripts/, tests).  Keys that change the weight packing
// ("decode_fuse", "gemv_cuda") must be set before er_finalize_weights.
extern "C" int er_debug_set(er_engine* e, const char* key, int64_t value) {
    if (!key) return set_err(ER_ERR_INVALID, "null key");
    const std::string k = key;
    if (k == "poison_alloc") { g_poison_alloc = (int)value; return ER_OK; }      // process-wide; e may be NULL
    if (k == "train_recompute") { extern int g_er_train_recompute; g_er_train_recompute = value != 0; return ER_OK; }   // process-wide: per-layer recomputation (opt.checkpointing)
    if (k == "train_fwd_lse") { extern int g_er_train_fwd_lse; g_er_train_fwd_lse = value != 0; return ER_OK; }   // process-wide: forward-written softmax statistic in the training backward
    if (k == "attn_bwd_wmma") { extern int g_er_attn_bwd_wmma; g_er_attn_bwd_wmma = value != 0; return ER_OK; }   // process-wide: wmma attention backward (A/B)
    if (k == "dense_legacy") { extern int g_er_dense_legacy; g_er_dense_legacy = value != 0; return ER_OK; }   // process-wide: mma.sync GEMM / attention
    if (!e) return set_err(ER_ERR_INVALID, "null engine");
    const int v = (int)value;
    if (k == "decode_fuse") {
        if (v && !e->S_fuse) return set_err(ER_ERR_CAPACITY, "the tensor-parallel decode layer does not support this model shape / SM count");
        if (e->finalized && v && !e->wfuse) return set_err(ER_ERR_STATE, "decode_fuse must be set before er_finalize_weights");
        e->use_fuse = v != 0;
    }
    else if (k == "gemv_cuda") { if (e->finalized) return set_err(ER_ERR_STATE, "gemv_cuda must be set before er_finalize_weights"); e->use_mma = (v == 0 && e->C % 256 == 0) ? 1 : 0; e->upstage = e->use_mma ? 8 : er_decode_stage_bytes() / (e->ustride * 2) / 8 * 8; }
    else if (k == "split_handicap") e->split_handicap = std::max(0, std::min(7, v));
    else if (k == "red_group4") e->red_group4 = v != 0;
    else if (k == "split_handicap_fuse") e->split_handicap_fuse = std::max(0, std::min(7, v));
    else if (k == "xrep") e->xrep = std::max(1, std::min(8, v));
    else if (k == "poll_rounds") e->poll_rounds = std::max(0, v);
    else if (k == "pf_dist") e->pf_dist = std::max(0, v);
    else if (k == "nosync") e->dbg_nosync = v != 0;
    else if (k == "cache_rows") {       // timing experiments at a chosen context length: pretend the cache holds `value` rows (contents: whatever is there)
        if (e->cache_rows <= 0 || v < 1 || v >= e->Lmax) return set_err(ER_ERR_INVALID, "cache_rows: prefill first, 1 <= rows < capacity");
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&e->st->L, &v, 4, cudaMemcpyHostToDevice));
        e->cache_rows = v; e->cache_rows_stale = false;
    }
    else return set_err(ER_ERR_INVALID, "unknown debug key '%s'", key);
    return ER_OK;
}

// Debug/profiling: record the phase timeline (ns, %globaltimer) of CTA `cta` while it generates token `token` of the next
// er_decode call (token < 0 disables).  er_debug_read_timeline copies n slots back (synchronises the device).
extern "C" int er_debug_phase_timeline(er_engine* e, int32_t token, int32_t cta) {
    if (!e) return set_err(ER_ERR_INVALID, "null engine");
    e->prof_token = token; e->prof_cta = cta;
    return ER_OK;
}
extern "C" int er_debug_read_timeline(er_engine* e, uint64_t* out_host, int32_t n) {
    if (!e || !out_host || n < 0 || n > 8192) return set_err(ER_ERR_INVALID, "bad argument");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out_host, e->prof, (size_t)n * 8, cudaMemcpyDeviceToHost));
    return ER_OK;
}
