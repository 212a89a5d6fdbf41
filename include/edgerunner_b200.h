/*
 * edgerunner_b200 — C ABI of the B200-native mesh-token decode path.
 *
 * The reference (NVlabs/EdgeRunner) has no FFI/plugin registry: its seams are Python call sites plus one pybind
 * module.  Each entry point below names the reference interface it stands in for; the Python
 * mirror of those interfaces (core/models.py, meto/__init__.py in this repository) binds them with ctypes, and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes only (no torch types).  `*_dev` pointers are CUDA device pointers in the
 * current context of `cfg.device`; `stream` is a cudaStream_t passed as void* (NULL = default stream); everything is
 * asynchronous on that stream unless the name ends in `_host`.  The caller owns every buffer it passes; the engine
 * owns its packed weights, KV cache and workspace.  Return value: 0 = ok, negative = error, message via
 * er_last_error() (thread-local).  One engine per device per thread; no Python GIL is ever taken.
 */
#ifndef EDGERUNNER_B200_H
#define EDGERUNNER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ER_OK 0
#define ER_ERR_INVALID (-1)
#define ER_ERR_CUDA (-2)
#define ER_ERR_STATE (-3)
#define ER_ERR_CAPACITY (-4)

#define ER_DTYPE_F16 0
#define ER_DTYPE_F32 1

#define ER_MODE_GREEDY 0
#define ER_MODE_SAMPLE 1

typedef struct er_engine er_engine;

/* Model dimensions: the fields of core/options.py:17-148 that the path reads (core/models.py:33-99). */
typedef struct er_config {
    int32_t device;
    /* mesh decoder (ShapeOPTConfig, core/transformer/modeling_opt.py:86-134) */
    int32_t hidden_dim, num_heads, num_layers, ffn_dim, vocab_size, max_positions;
    int32_t num_cond_tokens;      /* P: latent tokens (+1 if use_num_face_cond) */
    int32_t use_num_face_cond;
    int32_t bos_token_id, eos_token_id, pad_token_id;
    /* point encoder (PointEncoderEmbed, core/transformer/point.py:172-206); has_point_encoder = 0 for point_latent */
    int32_t has_point_encoder;
    int32_t point_hidden_dim, point_num_heads, point_latent_size, point_latent_dim;
    /* capacity */
    int32_t max_seq_rows;         /* KV-cache rows per head: prefix + prompt + max_new_tokens */
    int32_t max_points;           /* largest point cloud passed to er_encode_cond */
    int32_t max_tf_rows;          /* rows (batch * seq) of the largest teacher-forced forward; 0 = decode only */
} er_config;

const char* er_last_error(void);
int er_version(void);

/* LMM.__init__ + .half().to(device)  (core/models.py:33-99, infer.py:41-56) */
int er_create(const er_config* cfg, er_engine** out);
void er_destroy(er_engine* e);

/* load_state_dict: one call per tensor of the reference checkpoint schema (infer.py:44-50).
 * `name` is the state-dict key; data may be fp16 or fp32 on the device; it is rounded to fp16 (model.half()) and
 * re-packed into the engine's layouts.  Unknown names return ER_ERR_INVALID (strict); er_finalize_weights checks
 * that every tensor of the schema was provided. */
int er_load_weight(er_engine* e, const char* name, const void* data_dev, int32_t dtype, const int64_t* shape, int32_t ndim, void* stream);
int er_finalize_weights(er_engine* e, void* stream);

/* LMM.encode_cond (core/models.py:101-144) for cond_mode 'point' (is_latent = 0, conds = [n_points][3] fp32) or
 * 'point_latent' (is_latent = 1, conds = [latent_size][latent_dim] fp32).  Result ([P][C] fp32) stays in the engine;
 * cond_embeds_out_dev (optional) receives a copy, latents_out_dev (optional, [latent_size][latent_dim] fp16) the
 * encoder output. */
int er_encode_cond(er_engine* e, const float* conds_dev, int32_t n_points, int32_t is_latent, int32_t num_faces,
                   float* cond_embeds_out_dev, void* latents_out_dev, void* stream);

/* Step 0 of generate (core/models.py:224-233 + ShapeOPT.forward on the 2050-row prefix, modeling_opt.py:464-497):
 * prompt = BOS [+ resume ids] (host array), fills the KV cache and leaves the last row's logits in the engine. */
int er_prefill(er_engine* e, const int32_t* prompt_ids_host, int32_t n_prompt, void* stream);

/* The auto-regressive loop (HF GenerationMixin._sample as driven by core/models.py:286-303 + FSM :245-271).
 * Generates up to max_new_tokens ids into out_ids_dev (new tokens only, EOS included), count into out_len_dev.
 * out_logits_dev (optional, [max_new_tokens][vocab] fp32) receives the lm_head output of every step before its fp16
 * rounding; forced_ids_dev (optional) teacher-forces the fed token (tests).  tokens_per_launch <= 0: one launch. */
int er_decode(er_engine* e, int32_t max_new_tokens, int32_t mode, int32_t top_k, uint64_t seed, int32_t use_tokenizer_fsm,
              int32_t tokens_per_launch, int32_t* out_ids_dev, int32_t* out_len_dev, float* out_logits_dev,
              const int32_t* forced_ids_dev, void* stream);

/* LMM.generate with HOST buffers (infer.py:104-106 call shape): conds_host fp32 -> ids_host; synchronises.
 * This is the end-to-end entry measured by bench.py's `e2e`. */
int er_generate_host(er_engine* e, const float* conds_host, int32_t n_points, int32_t is_latent, int32_t num_faces,
                     const int32_t* resume_ids_host, int32_t n_resume, int32_t max_new_tokens, int32_t mode, int32_t top_k,
                     uint64_t seed, int32_t use_tokenizer_fsm, int32_t* out_ids_host, int32_t* out_len_host);

/* LMM.forward in eval mode (core/models.py:147-202; dense causal, no padding): batch of B samples.
 * conds_dev [B][n_points][3] fp32, tokens_dev [B][T] int32, labels_dev [B][P+T] int64 (-100 ignored), num_faces_host [B].
 * losses_dev[3] = {loss, loss_ce, loss_kl}; logits_out_dev (optional) [B][P+T][vocab] fp32 (pre-rounding). */
int er_forward_tf(er_engine* e, const float* conds_dev, int32_t n_points, int32_t is_latent, const int32_t* tokens_dev,
                  const int64_t* labels_dev, const int32_t* num_faces_host, int32_t B, int32_t T, float kl_weight,
                  float* losses_dev, float* logits_out_dev, void* stream);

/* er_forward_tf with the two things a data-parallel, padded batch needs (BASELINE configs[3]):
 * mask_dev (optional) [B][P+T] bytes, the reference's `masks` (core/models.py:154) with the P condition rows prepended as ones: 1 = real token,
 *   0 = padding.  Only RIGHT-padded masks are accepted by the Python layer (collate_fn pads at the end, provider.py:469-541): for those the
 *   varlen flash path of the reference (attention.py:65-93: unpad -> causal varlen -> pad_input) equals dense causal attention with the
 *   masked rows zeroed afterwards, which is what runs here.
 * sums_dev (optional) double[3] = {sum of the supervised tokens' cross-entropies, number of supervised tokens, KL term} of this call: the ONE
 *   vector a data-parallel run all-reduces (edgerunner_b200/dist.py::dp_reduce_losses); all reductions are fixed-order (bit-reproducible). */
int er_forward_tf2(er_engine* e, const float* conds_dev, int32_t n_points, int32_t is_latent, const int32_t* tokens_dev,
                   const int64_t* labels_dev, const uint8_t* mask_dev, const int32_t* num_faces_host, int32_t B, int32_t T, float kl_weight,
                   float* losses_dev, double* sums_dev, float* logits_out_dev, void* stream);

/* The op seam core/transformer/attention.py:27-62 `attention(q, k, v, causal)` for unmasked fp16 inputs:
 * q [B][Nq][H][D], k/v [B][Nk][H][D] contiguous, out [B][Nq][H][D]; D in {64, 96}; causal requires Nq == Nk
 * (single-query causal attention over a cache is inside er_decode).  Engine-free. */
int er_attention_bnhd(const void* q_dev, const void* k_dev, const void* v_dev, void* out_dev, int32_t B, int32_t Nq, int32_t Nk,
                      int32_t H, int32_t D, int32_t causal, void* stream);

/* Introspection used by bench.py / tests */
int64_t er_weight_bytes_per_token(const er_engine* e);   /* algorithmic weight bytes one decode step reads */
int64_t er_kv_bytes_per_row(const er_engine* e);         /* K+V bytes one cached position adds to a decode step */
int32_t er_cache_rows(const er_engine* e);               /* rows currently in the KV cache */
int64_t er_kernel_launches(const er_engine* e);          /* kernels launched by this engine so far */

/* Experiment / diagnostic switches of the decode kernel (scripts/, tests; the library never reads the environment).  Keys:
 * "decode_fuse" (1: tensor-parallel decode layer, 0: five-exchange layer), "gemv_cuda" (CUDA-core GEMV consumers) — both before
 * er_finalize_weights —, "split_handicap", "xrep", "poll_rounds", "pf_dist" (bytes of L2 run-ahead per CTA), "nosync" (timing diagnostics: grid barriers skipped,
 * results are garbage), "cache_rows" (pretend the cache holds that many rows; timing at a chosen context length),
 * "poison_alloc" (process-wide, e may be NULL: fill later allocations with 0xFF), "dense_legacy" (process-wide: use the mma.sync GEMM /
 * attention kernels instead of the tcgen05 ones; A/B timing).  Unknown key: ER_ERR_INVALID. */
int er_debug_set(er_engine* e, const char* key, int64_t value);

/* Profiling aid (profiles/): phase timeline of one CTA for one generated token of the next er_decode call. Slots (ns):
 * [0] token start, then for each layer 15 stamps (phase / exchange boundaries, see decode_kernel.cu), then (lm_head end,
 * barrier end); slots [4096 + 16 * cta + k]: the same 15 stamps of layer 5 for every CTA.  n <= 8192. */
int er_debug_phase_timeline(er_engine* e, int32_t token, int32_t cta);
int er_debug_read_timeline(er_engine* e, uint64_t* out_host, int32_t n);

/* ---- DiT denoiser: the latent generator of the image-conditioned path (infer_dit.py:104-113) ---------------------------------------------
 * Replaces core/transformer/dit.py (DiT, DiTLayer, Timesteps, TimestepEmbedding) and the device work of core/models_dit.py MDiT.get_cond /
 * MDiT.run.  The CLIP vision tower stays a library model on the Python side (as in the reference); the scheduler's per-step scalars are
 * computed by the Python mirror of diffusers' DDIMScheduler and handed in as a table, the update itself runs in the fused kernel. */
typedef struct er_dit er_dit;
typedef struct er_dit_config {
    int32_t device;
    int32_t hidden_dim, num_heads, num_layers;   /* Options.dit_hidden_dim / dit_num_heads / dit_num_layers (core/options.py:55-59) */
    int32_t latent_size, latent_dim;             /* Options.point_latent_size / point_latent_dim: the denoised tensor is [latent_size][latent_dim] */
    int32_t cond_tokens, cond_dim;               /* CLIP ViT-H/14 @ 224: 257 tokens of 1280 (core/models_dit.py:52-54) */
} er_dit_config;
#define ER_DIT_PRED_EPSILON 0
#define ER_DIT_PRED_V 1                          /* Options.noise_scheduler_predtype (core/options.py:63) */

/* MDiT.__init__ + load_state_dict + .half().to(device) for the keys `dit.*`, `proj_cond.*`, `norm_cond.*` (core/models_dit.py:33-76,
 * infer_dit.py:58-70); data fp16 or fp32 on the device, rounded to fp16.  Unknown names / wrong sizes: ER_ERR_INVALID. */
int er_dit_create(const er_dit_config* cfg, er_dit** out);
void er_dit_destroy(er_dit* e);
int er_dit_load_weight(er_dit* e, const char* name, const void* data_dev, int32_t dtype, int64_t numel, void* stream);
int er_dit_finalize_weights(er_dit* e, void* stream);

/* MDiT.get_cond after the image encoder (core/models_dit.py:116): norm_cond(proj_cond(h)); clip_hidden_dev fp16 [B][cond_tokens][cond_dim]
 * -> cond_out_dev fp32 [B][cond_tokens][hidden_dim]. */
int er_dit_cond(er_dit* e, const void* clip_hidden_dev, int32_t B, float* cond_out_dev, void* stream);

/* DiT.forward(x, c, t) (core/transformer/dit.py:165-196) under autocast(fp16): x_dev fp32 [B][latent_size][latent_dim], cond_dev fp32
 * [B][cond_tokens][hidden_dim], t_dev fp32 [B] -> out_dev fp16 [B][latent_size][latent_dim]. */
int er_dit_forward(er_dit* e, const float* x_dev, const float* cond_dev, const float* t_dev, int32_t B, void* out_dev, void* stream);

/* MDiT.run's loop (core/models_dit.py:209-227): for each of n_steps timesteps: [latents] * 2 -> DiT -> classifier-free guidance ->
 * scheduler.step; latents_dev fp32 [R][latent_size][latent_dim] is updated in place (in: the initial noise, or the re-noised latents).
 * cond_dev fp32 [R][cond_tokens][hidden_dim] is the conditional half (the unconditional half is zeros, as in the reference); guided = 0
 * runs without guidance (one prediction per sample).  timesteps_host [n_steps] fp32; coef_host [n_steps][4] fp32 = {sqrt(alpha_t),
 * sqrt(1 - alpha_t), sqrt(alpha_prev), sqrt(1 - alpha_prev - sigma_t^2)} (DDIM, eta = 0: sigma = 0).  Asynchronous on `stream` after the tables
 * are uploaded; one CUDA graph per step. */
int er_dit_run(er_dit* e, const float* cond_dev, float* latents_dev, int32_t R, int32_t n_steps, const float* timesteps_host,
               const float* coef_host, float guidance_scale, int32_t guided, int32_t prediction_type, void* stream);
/* the same with HOST buffers, synchronous (bench.py's e2e leg of the DiT workload) */
int er_dit_run_host(er_dit* e, const float* cond_host, float* latents_host, int32_t R, int32_t n_steps, const float* timesteps_host,
                    const float* coef_host, float guidance_scale, int32_t guided, int32_t prediction_type);

int64_t er_dit_kernel_launches(const er_dit* e);
double er_dit_flops_per_forward(const er_dit* e, int32_t batch);   /* GEMM + attention FLOPs of one denoiser forward over `batch` samples */
int er_dit_debug_set(er_dit* e, const char* key, int64_t value);   /* "graph": 0 = launch every step's kernels directly; "fuse": 0 = GEGLU / gated
                                                                      residuals as separate kernels instead of GEMM epilogues; "uncond_shortcut": 0 = run the
                                                                      cross-attention of the zero-condition half in full (all A/B timing, bit-identical) */

/* ---- Training step: forward in training mode + backward ---------------------------------------------------------------------
 * Replaces, for one batch, `out = model(data); accelerator.backward(out['loss'])` (main.py:168-172) = torch autograd over LMM.forward
 * (core/models.py:147-202 -> core/transformer/modeling_opt.py:253-298 post-LN layers with F.dropout(p) on both branches, :464-517 lm_head +
 * shifted cross-entropy); the activations of every layer are kept from the forward pass when device memory allows, else every layer is re-run in
 * the backward pass as the reference's opt.checkpointing does.  Arguments as er_forward_tf2, plus dropout_p (ShapeOPTConfig.dropout, 0.1 in the reference; the keep mask is a
 * counter-based function of `seed`, not torch's Philox stream), loss_scale (static scale of the fp16 activation gradients; exported gradients are
 * unscaled), train_encoder (0: opt.freeze_encoder, the point encoder and the KL term carry no gradient; 1: the point encoder of cond_mode 'point' is
 * trained too — its forward is re-run per cloud in the backward pass — and the KL term kl_weight * 0.5 sum(latent^2) contributes).
 * losses_dev[3] = {loss, mean CE, KL} of this rank's batch (the reference's DDP averages per-rank means: no cross-rank loss sums here);
 * sums_dev (optional) as er_forward_tf2.  The engine must have been created with max_tf_rows >= B * (num_cond_tokens + T).
 * er_grad_get: copy the fp32 gradient of one state-dict entry (reference key schema, dense [rows][cols], numel checked) to out_dev after a
 * step (point_encoder.* keys only after a step with train_encoder = 1).  er_grad_has: 1 if the entry can receive a gradient (every parameter of the
 * reference's LMM in cond_mode point / point_latent), 0 for buffers / unknown keys. */
int er_train_step(er_engine* e, const float* conds_dev, int32_t n_points, int32_t is_latent, const int32_t* tokens_dev, const int64_t* labels_dev,
                  const uint8_t* mask_dev, const int32_t* num_faces_host, int32_t B, int32_t T, float kl_weight, float dropout_p, uint64_t seed,
                  float loss_scale, int32_t train_encoder, float* losses_dev, double* sums_dev, void* stream);
int er_grad_get(er_engine* e, const char* name, float* out_dev, int64_t numel, void* stream);
int32_t er_grad_has(er_engine* e, const char* name);
/* Backward of the attention() op seam (core/transformer/attention.py:27-95; forward: er_attention_bnhd): tensors [B][N][H][D] fp16 contiguous,
 * out = the forward result, dout = the gradient of it; writes dq, dk, dv.  D in {64, 96}; causal needs Nq == Nk.  Deterministic (no atomics). */
int er_attention_bwd_bnhd(const void* q_dev, const void* k_dev, const void* v_dev, const void* out_dev, const void* dout_dev, void* dq_dev, void* dk_dev,
                          void* dv_dev, int32_t B, int32_t Nq, int32_t Nk, int32_t H, int32_t D, int32_t causal, void* stream);

/* ---- Optimizer half of the training step ---------------------------------------------------------------------------------------
 * main.py:133 `torch.optim.AdamW(model.parameters(), lr, weight_decay=0.01, betas=(0.9, 0.95))` and main.py:175-177
 * `accelerator.clip_grad_norm_(model.parameters(), opt.gradient_clip)` over flat fp32 buffers (16-byte aligned, n elements).
 * er_grad_norm_clip: *norm_out_dev = ||g||_2, *scale_out_dev = min(1, max_norm / (norm + 1e-6)) (torch.nn.utils.clip_grad_norm_); scratch_dev:
 * 592 doubles.  er_adamw_step: one AdamW update, `step` counted from 1; grad_scale_dev (optional) = the clipping coefficient, applied to the
 * gradient as it is read; param16_out_dev (optional) receives the fp16 copy of the updated parameters (what the forward kernels consume). */
int er_grad_norm_clip(const float* grad_dev, int64_t n, float max_norm, double* scratch_dev, float* norm_out_dev, float* scale_out_dev, void* stream);
int er_adamw_step(float* param_dev, const float* grad_dev, float* exp_avg_dev, float* exp_avg_sq_dev, void* param16_out_dev, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int32_t step, const float* grad_scale_dev, void* stream);

/* meto tokenizer backends of the reference's pybind module `_meto` (meto/src/bindings.cpp:11-28) */
#define ER_METO_LR_ABSCO 0   /* Engine_LR_ABSCO: absolute coordinates, vocabulary bins + 3 (the ArAE / DiT presets) */
#define ER_METO_LR 1         /* Engine_LR: parallelogram residuals, vocabulary 2 * bins + 3 (Options.meto_backend = 'LR') */
#define ER_METO_CLERS 2      /* Engine_CLERS: classic EdgeBreaker C/L/E/R/S + BOM/EOM, parallelogram residuals offset by 2 * bins + 7
                                (meto/include/meto/engine_clers.h; the reference's python wrapper reports 2 * bins + 7 tokens) */

/* Detokenizer (CPU, native): replaces `_meto.Engine_{LR_ABSCO,LR,CLERS}.decode`
 * (meto/include/meto/engine_lr_absco.h:223-295, engine_lr.h:171-254, engine_clers.h:186-283).  tokens are already -3 shifted (provider.py:115).
 * Capacities: verts >= 3*(n/4+3) floats*3, faces >= (n/4+3)*3, face_type >= n/4+3.  Counts are returned through
 * n_verts/n_faces/n_types. */
int er_meto_decode(int32_t backend, int32_t discrete_bins, const int32_t* tokens, int64_t n, float* verts, int32_t* faces,
                   int32_t* face_type, int64_t* n_verts, int64_t* n_faces, int64_t* n_types);

/* Tokenizer, ENCODE side (CPU, native): replaces `_meto.Engine_{LR_ABSCO,LR}.encode` (Mesh::Mesh, meto/include/meto/mesh.h:172-278;
 * Engine_LR_ABSCO::encode, engine_lr_absco.h:66-220; Engine_LR::encode, engine_lr.h:54-169), the tokenizer call of the
 * training-data path (core/provider.py:69-106).  verts [n_verts][3] float32 in [-1, 1], faces [n_faces][3] vertex indices.
 * Outputs (caller-allocated): tokens [tokens_cap] in the _meto alphabet (0 L, 1 R, 2 BOM, coordinates from 3; LR marks an
 * out-of-range residual with -1); face_order / face_type [faces_cap]: input index and type (0 L, 1 R, 2 end-of-strip) of each
 * emitted face.  LR_ABSCO emits every face once (10 * n_faces tokens and n_faces entries always suffice); LR may emit faces more
 * than once.  *n_tokens / *n_faces_out always receive the needed sizes; ER_ERR_CAPACITY if a capacity is too small (nothing is
 * written then), ER_ERR_INVALID for out-of-range vertex indices. */
int er_meto_encode(int32_t backend, int32_t discrete_bins, const float* verts, int64_t n_verts, const int32_t* faces, int64_t n_faces,
                   int32_t* tokens, int64_t tokens_cap, int32_t* face_order, int32_t* face_type, int64_t faces_cap,
                   int64_t* n_tokens, int64_t* n_faces_out);

/* Mesh clean-up at the tail of save_mesh (core/provider.py:55-58; the reference calls third-party trimesh: merge_vertices,
 * update_faces(unique_faces()), fix_normals).  CPU, native: merge vertices equal after rounding to `digits` decimals (first
 * occurrence order), drop faces whose vertex set already occurred, make the winding consistent per edge-connected component and
 * outward (non-negative signed volume).  verts [n_verts][3] float64, faces [n_faces][3]; outputs caller-allocated with the input
 * sizes; the new counts are returned through n_verts_out / n_faces_out. */
int er_mesh_clean(const double* verts, int64_t n_verts, const int32_t* faces, int64_t n_faces, int32_t digits,
                  double* verts_out, int32_t* faces_out, int64_t* n_verts_out, int64_t* n_faces_out);

#ifdef __cplusplus
}
#endif
#endif
