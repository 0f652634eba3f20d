/* jpdvt_b200.h - C ABI of libjpdvt_sm100.so, the B200 (sm_100a) implementation of the JPDVT hot path.
 *
 * The reference (hamzafer/JPDVT-MT-NTNU) is pure Python and has no FFI; its "plugin boundary" for this path is the
 * Python import surface `models.DiT / DiT_models / get_2d_sincos_pos_embed` and
 * `diffusion.create_diffusion(...).p_sample_loop / training_losses`.  The host-side mirror of that
 * surface lives in jpdvt_mt_ntnu_b200/ and reaches the GPU only through the entry points below, so every symbol
 * cites the reference lines whose work it replaces (paths relative to image_model/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the library never allocates or frees
 *     caller-visible memory and never synchronises the host with the device
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued on it and the call returns immediately
 *   - return value: 0 on success, negative jpdvt_status on failure; jpdvt_last_error_string() describes the last
 *     failure on the calling thread; no exceptions cross the boundary
 *   - bf16 tensors are raw uint16_t storage (torch.bfloat16); row-major everywhere
 *   - hidden width 768, 12 heads x 64, latent width 8 (models.py:176-179, 409-410)
 */
#ifndef JPDVT_B200_H_
#define JPDVT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JPDVT_ABI_VERSION 6
#define JPDVT_HIDDEN 768
#define JPDVT_LATENT 8

typedef enum jpdvt_status {
  JPDVT_OK = 0,
  JPDVT_ERR_BAD_ARG = -1,
  JPDVT_ERR_CUDA = -2,
  JPDVT_ERR_UNSUPPORTED = -3,
  JPDVT_ERR_DRIVER = -4
} jpdvt_status;

typedef uint16_t jpdvt_bf16;

int jpdvt_abi_version(void);
const char* jpdvt_last_error_string(void);
/* Kernels this library has launched in the calling process so far (every launch site counts itself). */
int64_t jpdvt_launch_count(void);
/* Checks that the current device is sm_100 (B200); returns JPDVT_ERR_UNSUPPORTED otherwise.  No CPU fallback exists. */
int jpdvt_device_check(void);

/* ---- single kernels ------------------------------------------------------------------------------------------ */

/* If delta != NULL: x_out = x_in + gate[b] * delta first (delta: bf16 branch output of the previous GEMM; gate: 768-wide
 * adaLN gate of sample b at gate + b*mod_stride, NULL = 1; x_out may alias x_in) - the gated residual add of
 * models.py:120-121 rides on this pass.  Then y = LayerNorm(x_out, eps=1e-6, no affine) * (1 + scale[b]) + shift[b],
 * b = row / tokens; sample b reads its vectors at shift + b*mod_stride, scale + b*mod_stride (mod_stride 0 = one
 * conditioning row for the whole batch).  Replaces nn.LayerNorm + modulate: models.py:19-20,107,109,120-121,131,140. */
int jpdvt_ln_modulate_fwd(const float* x_in, float* x_out_or_null, const jpdvt_bf16* delta_or_null, const float* gate_or_null,
                          const float* shift, const float* scale, int64_t mod_stride, jpdvt_bf16* y, int64_t rows, int tokens,
                          void* stream);

/* tcgen05 GEMMs: out[M,N] = a[M,K] . w[N,K]^T + bias, a/w bf16 (w in nn.Linear layout), fp32 accumulate.
 * M arbitrary, K % 64 == 0, N % 128 == 0.  Replace timm Attention.qkv / Mlp.fc1 / FinalLayer.linear (models.py:108,112,132). */
int jpdvt_gemm_bias(const jpdvt_bf16* a, const jpdvt_bf16* w, const float* bias, jpdvt_bf16* out, float* out_f32_or_null,
                    int64_t m, int n, int k, void* stream);
int jpdvt_gemm_bias_f32(const jpdvt_bf16* a, const jpdvt_bf16* w, const float* bias, float* out, int64_t m, int n, int k,
                        void* stream);
/* out = gelu_tanh(a . w^T + bias)  (timm Mlp.fc1 + nn.GELU(approximate="tanh"), models.py:110-112) */
int jpdvt_gemm_bias_gelu(const jpdvt_bf16* a, const jpdvt_bf16* w, const float* bias, jpdvt_bf16* out, int64_t m, int n,
                         int k, void* stream);
/* out[row] = gate[row / tokens] * (a . w^T + bias) in bf16: the gated branch of attn.proj / mlp.fc2 (models.py:120-121);
 * the `x +=` half is applied by the next jpdvt_ln_modulate_fwd(delta = out). */
int jpdvt_gemm_bias_gate(const jpdvt_bf16* a, const jpdvt_bf16* w, const float* bias, const float* gate,
                         int64_t gate_stride, jpdvt_bf16* out, int64_t m, int n, int k, int tokens, void* stream);
/* x[row] += gate[row / tokens] * (a . w^T + bias), in place on the fp32 residual stream: the whole adaLN-Zero gated
 * residual update `x = x + gate.unsqueeze(1) * branch(...)` of attn.proj / mlp.fc2 (models.py:120-121) as the GEMM
 * epilogue (overlapped with the next tile's MMAs): short contractions move the residual tile with TMA through a
 * shared-memory ring, long ones prefetch it into registers.  n must be a multiple of 128; x 16-byte aligned. */
int jpdvt_gemm_bias_gate_residual(const jpdvt_bf16* a, const jpdvt_bf16* w, const float* bias, const float* gate,
                                  int64_t gate_stride, float* x, int64_t m, int n, int k, int tokens, void* stream);
/* jpdvt_gemm_bias_gate_residual followed, in the same kernel, by the LayerNorm-modulate that consumes the updated rows:
 * x[row] += gate[row / tokens] * (a . w^T + bias);  xn[row] = LN(x[row]) * (1 + ln_scale[row / tokens]) + ln_shift[...]
 * i.e. one whole `x = x + gate * branch(...)` line of DiTBlock.forward plus the `modulate(norm(x), shift, scale)` that opens
 * the next line (models.py:19-20,120-121).  n must be 768; each CTA pair owns whole 256-row blocks, so the rows it
 * normalises are the ones it has just written (read back from L2, not HBM).  ln_shift/ln_scale: [n_cond, 768] fp32 with
 * row stride mod_stride (0 = one row for every sample). */
int jpdvt_gemm_bias_gate_residual_ln(const jpdvt_bf16* a, const jpdvt_bf16* w, const float* bias, const float* gate,
                                     int64_t gate_stride, float* x, const float* ln_shift, const float* ln_scale,
                                     int64_t mod_stride, jpdvt_bf16* xn, int64_t m, int n, int k, int tokens, void* stream);
/* The LayerNorm-modulate that FOLLOWS a gated residual update, folded into the GEMM that consumes it (sampling loop: every
 * puzzle of the batch shares the timestep, gaussian_diffusion.py:509, so shift / scale are one vector per block):
 *   modulate(LayerNorm(x), shift, scale) . W^T + b  (models.py:19-20,107-121)
 *     = rstd * (bf16(x) . W'^T) - rstd * mean * u + v,   W' = W (1 + scale), u = rowsum(W'), v = b + W . shift
 * jpdvt_gemm_bias_gate_residual_copy: jpdvt_gemm_bias_gate_residual that also leaves bf16(x) in x_bf16 [m, n] and the
 *   rows' (sum, sum of squares) partials in row_stats [m, 2 * n / 256, 2] fp32 (n a multiple of 256).
 * jpdvt_fold_ln_weights: W', u, v for the qkv (rows 0..2303) and fc1 (rows 2304..5375) matrices of every block from row 0
 *   of the adaLN table `mod` (jpdvt_adaln_table layout): w_fold [depth, 5376, 768] bf16, fold_u / fold_v [depth, 5376].
 * jpdvt_gemm_ln_folded: out = [gelu_tanh](rstd * (x_bf16 . w_fold^T) - rstd * mean * fold_u + fold_v), mean / rstd of each
 *   row from row_stats over the k inputs (eps 1e-6, biased variance). */
int jpdvt_gemm_bias_gate_residual_copy(const jpdvt_bf16* a, const jpdvt_bf16* w, const float* bias, const float* gate,
                                       int64_t gate_stride, float* x, jpdvt_bf16* x_bf16, float* row_stats, int64_t m, int n,
                                       int k, int tokens, void* stream);
int jpdvt_fold_ln_weights(const jpdvt_bf16* w_qkv, const jpdvt_bf16* w_fc1, const float* b_qkv, const float* b_fc1, const float* mod,
                          jpdvt_bf16* w_fold, float* fold_u, float* fold_v, int depth, void* stream);
int jpdvt_gemm_ln_folded(int gelu, const jpdvt_bf16* x_bf16, const float* row_stats, int stats_slots, const jpdvt_bf16* w_fold,
                         const float* fold_u, const float* fold_v, jpdvt_bf16* out, int64_t m, int n, int k, void* stream);
/* x = cols . w_patch^T + bias + pos_embed[row % tokens] + x_t[row] . w_in_t   (PatchEmbed conv as GEMM + time_emb_in +
 * pos_embed, models.py:280-281).  cols = jpdvt_patchify(img); bias = x_embedder.proj.bias + time_emb_in.bias;
 * w_in_t = time_emb_in.weight^T as [8,768] fp32; pos = pos_embed [tokens,768] fp32. */
int jpdvt_gemm_patch_embed(const jpdvt_bf16* cols, const jpdvt_bf16* w_patch, const float* bias, const float* x_t,
                           const float* w_in_t, const float* pos, float* x, int64_t m, int tokens, void* stream);
/* te_out[row,:8] = w2 . silu(y[row] . w1^T + b1) + b2   (time_emb_out1 -> SiLU -> time_emb_out2, models.py:288-290) */
int jpdvt_final_head_fwd(const jpdvt_bf16* y, const jpdvt_bf16* w1, const float* b1, const float* w2, const float* b2,
                         float* te_out, int64_t m, void* stream);

/* softmax(q k^T / 8) v per (sample, head) on the fused QKV tensor [batch*tokens, 2304] -> [batch*tokens, 768]
 * (timm Attention.forward -> F.scaled_dot_product_attention, called from models.py:108,120).  lse2 (nullable, fp32
 * [batch, 12, tokens]): log2-domain log-sum-exp of the scaled scores, kept for jpdvt_attention_bwd. */
int jpdvt_attention_fwd(const jpdvt_bf16* qkv, jpdvt_bf16* out, float* lse2_or_null, int batch, int tokens, void* stream);

/* im2col of 16x16 patches, k = c*256 + py*16 + px (the flattened Conv2d weight order), fp32 -> bf16. */
int jpdvt_patchify(const float* img, jpdvt_bf16* cols, int batch, int image_size, void* stream);
/* models.py:227-240 */
int jpdvt_unpatchify(const float* y, float* img, int batch, int image_size, void* stream);

/* c = Linear(SiLU(Linear(sinusoid_256(t)))) and silu(c)   (TimestepEmbedder, models.py:27-64).
 * t: int64[n] model timesteps, or NULL: then every row uses map[*step_ptr] (or *step_ptr when map is NULL).
 * hid_scratch: [n, 768] fp32 of its own (the hidden activations between the two grid-wide phases; must not alias c / silu_c). */
int jpdvt_timestep_embed(const int64_t* t, int n, const int32_t* step_ptr, const int32_t* map, const float* w0,
                         const float* b0, const float* w2, const float* b2, float* c, float* silu_c, float* hid_scratch,
                         void* stream);
/* mod[r, :] = W_all . silu_c[r] + b_all for all adaLN linears at once (models.py:113-116,133-136);
 * W_all bf16 [n_out, 768] = concat(blocks[i].adaLN_modulation[1].weight ..., final_layer.adaLN_modulation[1].weight). */
int jpdvt_adaln_table(const float* silu_c, int rows, const jpdvt_bf16* w_all, const float* b_all, float* mod, int n_out,
                      void* stream);

/* mean = coef1[t]*x0 + coef2[t]*x_t ; sample = mean + [t != 0] * exp(0.5*logvar[t]) * noise
 * (q_posterior_mean_variance + p_sample, diffusion/gaussian_diffusion.py:234-254,311-312,424-430).  Tables are the
 * fp64 schedule cast to fp32 (the reference casts after the gather, :926).  t: int64[batch] or NULL -> *step_ptr. */
int jpdvt_posterior_step(const float* x0, const float* x_t, const float* noise, const float* coef1, const float* coef2,
                         const float* logvar, const int64_t* t, const int32_t* step_ptr, float* mean_or_null,
                         float* sample_or_null, int64_t n, int64_t per_sample, void* stream);
/* The generator behind noise_key (jpdvt_sampler) on its own: out[i] (fp32, n % 4 == 0) = the normals the posterior kernel
 * draws for loop position `step`; raw (nullable, uint32[n]) = the underlying Philox4x32-10 words (counter = {i/4 lo, i/4 hi,
 * step, key[1]}, key = key[0]).  jpdvt_posterior_step_philox = jpdvt_posterior_step with that noise drawn in the kernel. */
int jpdvt_philox_normal(float* out_or_null, uint32_t* raw_or_null, int64_t n, int step, const int64_t* key, void* stream);
int jpdvt_posterior_step_philox(const float* x0, const float* x_t, const int64_t* noise_key, int noise_step, const float* coef1,
                                const float* coef2, const float* logvar, const int64_t* t, const int32_t* step_ptr,
                                float* sample, int64_t n, int64_t per_sample, void* stream);
/* DDIM update (gaussian_diffusion.py:559-578, with the `condition` argument the reference call at :547 forgot):
 * eps = (recip[t]*x_t - x0)/recipm1[t]; sample = sqrt_abp[t]*x0 + dir[t]*eps + [t != 0]*sigma[t]*noise.  Pinned against
 * the reference's own DDIM code run with that argument supplied (tests/golden/ddim_*.npz). */
int jpdvt_ddim_step(const float* x0, const float* x_t, const float* noise, const float* recip, const float* recipm1,
                    const float* sqrt_abp, const float* dir, const float* sigma, const int64_t* t, const int32_t* step_ptr,
                    float* sample, int64_t n, int64_t per_sample, void* stream);
/* out = sqrt_ac[t]*x0 + sqrt_1mac[t]*noise, optionally blended out*(1-keep) + keep*x0 (gaussian_diffusion.py:217-232,800) */
int jpdvt_q_sample(const float* x0, const float* noise, const float* sqrt_ac, const float* sqrt_1mac, const int64_t* t,
                   const float* keep_or_null, float* out, int64_t n, int64_t per_sample, void* stream);

/* Greedy assignment on fp64 score matrices [batch, n, n] (rows = slots, columns = grid cells): the bit-exact contract
 * (inference.py:113-125,306).  order/pred: int32 [batch, n]. */
int jpdvt_assign_from_scores(const double* scores, int batch, int n, double sentinel, int32_t* order, int32_t* pred,
                             void* stream);
/* Latents [batch, T, 8] (token order p1 h1 p2 w1) -> per-slot mean -> fp64 L1 scores vs canon [G*G, 8] -> greedy
 * (inference.py:294-306).  scores_out (nullable): fp64 [batch, n, n]. */
int jpdvt_assign_greedy_l1(const float* latents, const float* canon, int batch, int grid, int tokens_per_side,
                           double sentinel, int32_t* order, int32_t* pred, double* scores_out_or_null, void* stream);

/* ---- puzzle plumbing either side of the sampling loop, batched on the device --------------------
 * dst slot i (row-major G x G grid of (size/G)-pixel pieces) = src piece perm[b, i]; slots with keep[b, i] == 0 are zeroed
 * (keep may be NULL).  With perm = the scramble indices this is the reference's scramble
 * (inference_ddp.py:382-395 = inference.py:266-278, batched inferencetexmet.py:318-338; masked-puzzle inference zeroes
 * slots); with perm = `order` from the assignment it is the reconstruction reconstructed[pred[i]] = piece i
 * (inference_ddp.py:449-455).  src/dst [batch, channels, size, size] fp32, perm [batch, G*G] int32, keep [batch, G*G] u8. */
int jpdvt_gather_pieces(const float* src, float* dst, const int32_t* perm, const uint8_t* keep_or_null, int batch,
                        int channels, int size, int grid, void* stream);
/* Crop-gap erosion of the training loader (train_JPDVT.py:345-349): every piece of the G x G puzzle is centre-cropped from
 * in_piece to out_piece pixels (offset `off` = torchvision CenterCrop's int(round((in - out) / 2))) and the crops are
 * re-tiled.  src [batch, channels, G*in_piece, G*in_piece] -> dst [batch, channels, G*out_piece, G*out_piece], fp32. */
int jpdvt_crop_pieces(const float* src, float* dst, int batch, int channels, int grid, int in_piece, int out_piece, int off,
                      void* stream);
/* matches[b] = #{i : pred[b,i] == truth[b,i]}, correct[b] = (matches[b] == n); totals (NULL or int64[3]) accumulates
 * (puzzles correct, pieces correct, puzzles) - the counters of inference_ddp.py:431-447 / :485-490. */
int jpdvt_score_placements(const int32_t* pred, const int32_t* truth, int batch, int n, int32_t* correct, int32_t* matches,
                           int64_t* totals_or_null, void* stream);

/* ---- whole denoiser / whole sampling loop --------------------------------------------------------------------- */

typedef struct jpdvt_weights {
  int32_t depth;                /* number of DiT blocks (12 for JPDVT) */
  int32_t tokens;               /* T = (image_size / 16)^2 */
  int32_t image_size;
  int32_t reserved;
  const jpdvt_bf16* w_patch;    /* [768, 768]   x_embedder.proj.weight flattened */
  const float* b_embed;         /* [768]        x_embedder.proj.bias + time_emb_in.bias */
  const float* w_in_t;          /* [8, 768]     time_emb_in.weight^T */
  const float* pos;             /* [T, 768]     pos_embed */
  const float* t_w0;            /* [768, 256]   t_embedder.mlp.0 */
  const float* t_b0;
  const float* t_w2;            /* [768, 768]   t_embedder.mlp.2 */
  const float* t_b2;
  const jpdvt_bf16* w_ada;      /* [depth*4608 + 1536, 768] all adaLN linears, block-major then final */
  const float* b_ada;
  const jpdvt_bf16* w_qkv;      /* [depth, 2304, 768] */
  const float* b_qkv;           /* [depth, 2304] */
  const jpdvt_bf16* w_proj;     /* [depth, 768, 768] */
  const float* b_proj;          /* [depth, 768] */
  const jpdvt_bf16* w_fc1;      /* [depth, 3072, 768] */
  const float* b_fc1;           /* [depth, 3072] */
  const jpdvt_bf16* w_fc2;      /* [depth, 768, 3072] */
  const float* b_fc2;           /* [depth, 768] */
  const jpdvt_bf16* w_final;    /* [768, 768]   final_layer.linear */
  const float* b_final;
  const jpdvt_bf16* w_head1;    /* [64, 768]    time_emb_out1 */
  const float* b_head1;
  const float* w_head2;         /* [8, 64]      time_emb_out2 */
  const float* b_head2;
} jpdvt_weights;

typedef struct jpdvt_workspace {
  int64_t rows;                 /* capacity in token rows (>= batch * tokens) */
  int32_t cond_rows;            /* capacity in conditioning rows (>= batch, or 1 for a batch-uniform timestep) */
  int32_t step_rows;            /* capacity of the per-step conditioning tables below, in diffusion steps (0 = none) */
  float* x;                     /* [rows, 768] fp32 residual stream */
  jpdvt_bf16* xn;               /* [rows, 768]  */
  jpdvt_bf16* qkv;              /* [rows, 2304] */
  jpdvt_bf16* attn;             /* [rows, 768]  */
  jpdvt_bf16* hid;              /* [rows, 3072] (also holds the im2col tile of the patch embed) */
  jpdvt_bf16* y;                /* [rows, 768] gated branch outputs, then the final-layer output */
  float* y32;                   /* [rows, 768] fp32 copy for unpatchify, or NULL */
  float* c;                     /* [cond_rows, 768] */
  float* silu_c;                /* [cond_rows, 768] */
  jpdvt_bf16* silu_c_bf16;      /* [cond_rows, 768] (tensor-core adaLN path, cond_rows > 8) */
  float* mod;                   /* [cond_rows, depth*4608 + 1536] */
  /* LayerNorm folded into the qkv / fc1 GEMMs (batch-uniform timestep, i.e. the sampling loop - all four NULL = off):
   * modulate(LayerNorm(x), shift, scale) . W^T + b (models.py:19-20,120-121) = rstd * (x . W'^T) - rstd * mean * u + v */
  jpdvt_bf16* w_fold;           /* [depth, 2304 + 3072, 768] W' = W * (1 + scale), rebuilt every forward */
  float* fold_u;                /* [depth, 5376] row sums of W' */
  float* fold_v;                /* [depth, 5376] b + W . shift */
  float* row_stats;             /* [rows, 6, 2] per-row (sum, sum of squares) partials of the residual stream */
  /* jpdvt_sample_loop: conditioning of every step of one call, computed ahead of the loop (all three NULL = per step) */
  float* c_steps;               /* [step_rows, 768] */
  float* silu_c_steps;          /* [step_rows, 768] */
  float* mod_steps;             /* [step_rows, depth*4608 + 1536] */
  float* te_hid;                /* [max(cond_rows, step_rows), 768] hidden activations of the timestep MLP (scratch of its own) */
  float* x_embed;               /* [rows, 768] fp32 or NULL: jpdvt_sample_loop keeps the loop-invariant embedding here (NULL = per step) */
} jpdvt_workspace;

/* One DiT.forward (models.py:273-293): (img [B,3,S,S], t, x_t [B,T,8]) -> te_out [B,T,8] and, when img_out != NULL,
 * the unpatchified image head [B,3,S,S].  t: int64[B] model timesteps; NULL = batch-uniform timestep map[*step_ptr]. */
int jpdvt_denoiser_forward(const jpdvt_weights* w_host, const jpdvt_workspace* ws_host, const float* img,
                           const int64_t* t, const int32_t* step_ptr, const int32_t* map, const float* x_t,
                           float* te_out, float* img_out_or_null, int batch, void* stream);

typedef struct jpdvt_sampler {
  int32_t num_steps;            /* respaced step count (250) */
  int32_t chain;                /* 0 = reference behaviour: every step is fed the INITIAL noise
                                   (gaussian_diffusion.py:518-529); 1 = feed the running sample */
  const int32_t* step_ids;      /* [num_steps] device: num_steps-1, ..., 0 */
  const int32_t* timestep_map;  /* [num_steps] device: respaced index -> original timestep (respace.py:117-129) */
  const float* coef1;           /* [num_steps] posterior_mean_coef1 (fp32) */
  const float* coef2;           /* [num_steps] */
  const float* logvar;          /* [num_steps] posterior_log_variance_clipped */
  const float* step_noise;      /* [num_steps or 1, B, T, 8] noise drawn per step (torch-generated in parity mode), or NULL */
  int64_t step_noise_stride;    /* elements between consecutive steps' noise (0 = reuse one tensor) */
  float* x0;                    /* [B, T, 8] scratch: pred_xstart of the current step */
  float* sample;                /* [B, T, 8] scratch / result: sample of the current step */
  float* traj_x0;               /* optional [num_steps, B, T, 8] record of every pred_xstart, or NULL */
  float* traj_sample;           /* optional [num_steps, B, T, 8] record of every sample, or NULL */
  const int64_t* noise_key;     /* device int64[2] {seed, call counter}: with step_noise == NULL the per-step noise of p_sample
                                   (gaussian_diffusion.py:424) is drawn inside the posterior kernel (Philox4x32-10 + Box-Muller) */
} jpdvt_sampler;

/* SpacedDiffusion.p_sample_loop (gaussian_diffusion.py:433-529 through respace.py:89-129): runs steps
 * [first_step, last_step) of the reverse loop (0 = the first executed step, i.e. respaced index num_steps-1)
 * without any host round trip; the result of the last executed step is left in sampler->sample. */
int jpdvt_sample_loop(const jpdvt_weights* w_host, const jpdvt_workspace* ws_host, const jpdvt_sampler* s_host,
                      const float* condition, const float* noise, int batch, int first_step, int last_step, void* stream);


/* ---- training step (train_JPDVT.py:357-372: training_losses -> backward -> AdamW -> EMA) ----------------------------- */

/* dW[out_rows, n_cols] (fp32, nn.Linear layout) = P[m, out_rows]^T . Q[m, n_cols]: weight gradients as a tcgen05 GEMM with
 * MN-major operands and a split contraction.  scratch: jpdvt_wgrad_scratch_floats(...) floats (may be NULL when 0). */
int jpdvt_gemm_wgrad(const jpdvt_bf16* p, const jpdvt_bf16* q, float* dw, float* scratch, int64_t m, int out_rows, int n_cols,
                     void* stream);
int64_t jpdvt_wgrad_scratch_floats(int64_t m, int out_rows, int n_cols);
/* out = (a . w^T) * gprime, gprime = gelu_tanh'(fc1 pre-activation) kept by the training forward - the fc2 data gradient
 * fused with the GELU derivative (timm Mlp, models.py:110-112) */
int jpdvt_gemm_dgelu(const jpdvt_bf16* a, const jpdvt_bf16* w, const jpdvt_bf16* gprime, jpdvt_bf16* out, int64_t m, int n, int k,
                     void* stream);
/* Data gradient of a Linear layer from the layer's own weight: dX[m, n_in] = dY[m, k_out] . W with W [k_out, n_in] as nn.Linear
 * stores it (autograd of models.py:108-112's Linear calls).  W is read as an MN-major tensor-core operand, so no transposed
 * copy of the weights exists.  Output fp32 or bf16 (exactly one), optionally multiplied by gprime (the dGELU form), whose
 * epilogue can also accumulate the column sums of its bf16 output into colsum_or_null [n_in] (the bias gradient of the
 * layer below: db_fc1 = column sums of dh) so that no pass re-reads the output.  n_in must be a multiple of 256, k_out of 64. */
int jpdvt_gemm_dgrad(const jpdvt_bf16* dy, const jpdvt_bf16* w, const jpdvt_bf16* gprime_or_null, jpdvt_bf16* out_bf16_or_null,
                     float* out_f32_or_null, float* colsum_or_null, int64_t m, int n_in, int k_out, void* stream);
/* dQ, dK, dV of softmax(q k^T / 8) v into dqkv [batch*tokens, 2304]; o / d_o: [batch*tokens, 768]; lse2 from the forward.
 * dbias_or_null: the qkv Linear's bias gradient [2304], dbias[c] += sum_rows dqkv[row, c] (column sums of the bf16 values
 * just written, folded into the kernel's epilogue; accumulated with atomics - zero it once per backward pass). */
int jpdvt_attention_bwd(const jpdvt_bf16* qkv, const jpdvt_bf16* o, const jpdvt_bf16* d_o, const float* lse2, jpdvt_bf16* dqkv,
                        float* dbias_or_null, int batch, int tokens, void* stream);
/* x_out = x_in + gate[b]*y  =>  dy = gate[b]*dx (bf16); dgate[b] += sum_t dx*y; dbias += sum_rows dy */
int jpdvt_gate_bwd(const float* dx, const jpdvt_bf16* y, const float* gate, int64_t gate_stride, jpdvt_bf16* dy, float* dgate,
                   int64_t dgate_stride, float* dbias_or_null, float* part, int batch, int tokens, void* stream);
/* floats of `part` scratch the two reductions above/below need (per-sample partial sums instead of atomics) */
int64_t jpdvt_bwd_part_floats(int batch, int tokens);
/* backward of jpdvt_ln_modulate_fwd: dx (+)= LN'(dxn * (1 + scale[b])); dshift[b] += sum_t dxn; dscale[b] += sum_t dxn*xhat */
int jpdvt_ln_modulate_bwd(const float* x, const float* dxn, const float* scale, int64_t mod_stride, float* dx, int accumulate,
                          float* dshift, float* dscale, int64_t dmod_stride, jpdvt_bf16* dx_bf16_or_null, float* part, int batch,
                          int tokens, void* stream);
/* jpdvt_ln_modulate_bwd fused with the jpdvt_gate_bwd that consumes its dx (models.py:120-121 read upwards: every
 * LayerNorm backward is followed by the gate backward of the residual branch below it): the rows of dx are finished, gated
 * (dy = gate[b] * dx, bf16) and folded into dgate[b] / dbias in one pass.  y_or_null == NULL: LayerNorm backward only.
 * dshift, dscale, dgate and dbias are ACCUMULATED into with atomics (zero them once per backward pass); no scratch. */
int jpdvt_ln_gate_bwd(const float* x, const float* dxn, const float* scale, int64_t mod_stride, float* dx, int accumulate,
                      float* dshift, float* dscale, int64_t dmod_stride, jpdvt_bf16* dx_bf16_or_null, const jpdvt_bf16* y_or_null,
                      const float* gate, int64_t gate_stride, jpdvt_bf16* dy, float* dgate, int64_t dgate_stride,
                      float* dbias_or_null, int batch, int tokens, void* stream);
/* out[c] += sum_rows src[row, c]  (bias gradients) */
int jpdvt_colsum_bf16(const jpdvt_bf16* src, int64_t rows, int cols, float* out, void* stream);
int jpdvt_colsum_f32(const float* src, int64_t rows, int cols, float* out, void* stream);

typedef struct jpdvt_weights_t {   /* transposed bf16 copies ([in, out]) read by the data-gradient GEMMs */
  const jpdvt_bf16* w_qkv_t;    /* [depth, 768, 2304] */
  const jpdvt_bf16* w_proj_t;   /* [depth, 768, 768]  */
  const jpdvt_bf16* w_fc1_t;    /* [depth, 768, 3072] */
  const jpdvt_bf16* w_fc2_t;    /* [depth, 3072, 768] */
  const jpdvt_bf16* w_final_t;  /* [768, 768] */
  const jpdvt_bf16* w_head1_t;  /* [768, 64]  */
  const jpdvt_bf16* w_ada_t;    /* [768, depth*4608 + 1536] */
  const jpdvt_bf16* t_w2_t;     /* [768, 768] */
} jpdvt_weights_t;

typedef struct jpdvt_tape {        /* activations kept by the training forward for the backward pass */
  int64_t rows;                 /* batch * tokens */
  int32_t batch;
  int32_t reserved;
  jpdvt_bf16* cols;             /* [rows, 768]              im2col of the condition image */
  float* x;                     /* [2*depth + 1, rows, 768] input of every LayerNorm (residual stream snapshots) */
  jpdvt_bf16* xn1;              /* [depth, rows, 768]   */
  jpdvt_bf16* qkv;              /* [depth, rows, 2304]  */
  float* lse2;                  /* [depth, batch, 12, tokens] */
  jpdvt_bf16* att;              /* [depth, rows, 768]   */
  jpdvt_bf16* y1;               /* [depth, rows, 768]   attn.proj output (before the gate) */
  jpdvt_bf16* xn2;              /* [depth, rows, 768]   */
  jpdvt_bf16* hpre;             /* [depth, rows, 3072]  gelu'(fc1 output) (the pre-activations are overwritten in place) */
  jpdvt_bf16* h;                /* [depth, rows, 3072]  */
  jpdvt_bf16* y2;               /* [depth, rows, 768]   mlp.fc2 output (before the gate) */
  jpdvt_bf16* xnf;              /* [rows, 768]  */
  jpdvt_bf16* yfin;             /* [rows, 768]  final_layer.linear output */
  float* yfin32;                /* [rows, 768]  fp32 copy (image head), or NULL */
  float* headpre;               /* [rows, 64]   time_emb_out1 output */
  float* feat;                  /* [batch, 256] sinusoid features */
  float* tpre;                  /* [batch, 768] t_embedder.mlp.0 output */
  float* c;                     /* [batch, 768] */
  float* silu_c;                /* [batch, 768] */
  jpdvt_bf16* silu_c_bf16;      /* [batch, 768] */
  float* mod;                   /* [batch, depth*4608 + 1536] */
  float* thid;                  /* [batch, 768] SiLU(tpre): scratch between the two phases of the timestep MLP */
} jpdvt_tape;

typedef struct jpdvt_grads {       /* fp32 gradients in the parameters' own layouts; zero-filled by the caller before a backward */
  float* w_patch;  float* b_patch;  float* w_in;  float* b_in;      /* [768,768] [768] [768,8] [768] */
  float* t_w0;  float* t_b0;  float* t_w2;  float* t_b2;            /* [768,256] [768] [768,768] [768] */
  float* w_ada;  float* b_ada;                                      /* [n_mod,768] [n_mod] */
  float* w_qkv;  float* b_qkv;  float* w_proj;  float* b_proj;      /* stacked over depth */
  float* w_fc1;  float* b_fc1;  float* w_fc2;  float* b_fc2;
  float* w_final;  float* b_final;  float* w_head1;  float* b_head1;  float* w_head2;  float* b_head2;
} jpdvt_grads;

typedef struct jpdvt_bwd_scratch {
  float* dx;                    /* [rows, 768] gradient of the residual stream */
  float* dxn;                   /* [rows, 768] fp32 scratch (data gradients entering a LayerNorm) */
  jpdvt_bf16* dy;               /* [rows, 768]  */
  jpdvt_bf16* dh;               /* [rows, 3072] */
  jpdvt_bf16* dqkv;             /* [rows, 2304] */
  jpdvt_bf16* datt;             /* [rows, 768]  */
  jpdvt_bf16* dpre;             /* [rows, 64]   */
  float* dmod;                  /* [batch, n_mod] zero-filled by the caller */
  jpdvt_bf16* dmod_bf16;        /* [batch, n_mod] */
  float* small_f32;             /* [4, batch, 768] */
  jpdvt_bf16* small_bf16;       /* [4, batch, 768] */
  float* wgrad_scratch;         /* jpdvt_train_wgrad_scratch_floats(...) floats */
  float* part;                  /* jpdvt_bwd_part_floats(batch, tokens) floats */
  const float* zeros;           /* >= n_mod zeros (bias-free GEMM epilogues) */
} jpdvt_bwd_scratch;

int64_t jpdvt_train_wgrad_scratch_floats(int depth, int batch, int tokens);
/* DiT.forward with every activation the backward needs written to the tape (per-sample timesteps t: int64[batch]). */
int jpdvt_train_forward(const jpdvt_weights* w, const jpdvt_tape* tape, const float* img, const int64_t* t, const float* x_t,
                        float* te_out, float* img_out_or_null, int batch, void* stream);
/* Backward in three stages so the caller can overlap gradient all-reduces with the remaining stages:
 * head (final layer + position head + image head), one call per block from depth-1 down to 0, then the embeddings and
 * the conditioning path (adaLN linears, t_embedder). */
int jpdvt_train_backward_head(const jpdvt_weights* w, const jpdvt_weights_t* wt, const jpdvt_tape* tape,
                              const jpdvt_bwd_scratch* s, const jpdvt_grads* g, const float* d_te, const float* d_img_or_null,
                              void* stream);
int jpdvt_train_backward_block(const jpdvt_weights* w, const jpdvt_weights_t* wt, const jpdvt_tape* tape,
                               const jpdvt_bwd_scratch* s, const jpdvt_grads* g, int block, void* stream);
int jpdvt_train_backward_embed(const jpdvt_weights* w, const jpdvt_weights_t* wt, const jpdvt_tape* tape,
                               const jpdvt_bwd_scratch* s, const jpdvt_grads* g, const float* x_t, void* stream);

/* The loss terms of training_losses (diffusion/gaussian_diffusion.py:835-838 with mean_flat, :18-22):
 *   loss[b] = mean((te_tgt - te_out)^2)  [+ mean((img_tgt - img_out)^2 * (1 - keep[b, slot]))  when img_out != NULL]
 * te_*: [batch, per_te] fp32 (per_te = T*8); img_*: [batch, 3, S, S] fp32; keep: [batch, grid*grid] fp32 slot mask (the
 * reference's `masks`, 1 = slot shown clean, carries no image loss); part: jpdvt_mse_part_floats(batch) floats of scratch.
 * _bwd: d_te = dloss[b] * 2 (te_out - te_tgt) / per_te, d_img = dloss[b] * 2 (img_out - img_tgt)(1 - keep) / (3 S S). */
int64_t jpdvt_mse_part_floats(int batch);
int jpdvt_mse_loss_fwd(const float* te_out, const float* te_tgt, int64_t per_te, const float* img_out_or_null,
                       const float* img_tgt_or_null, const float* keep_or_null, int image_size, int grid, float* part,
                       float* loss, int batch, void* stream);
int jpdvt_mse_loss_bwd(const float* te_out, const float* te_tgt, int64_t per_te, const float* img_out_or_null,
                       const float* img_tgt_or_null, const float* keep_or_null, int image_size, int grid, const float* dloss,
                       float* d_te, float* d_img_or_null, int batch, void* stream);

/* One fused pass of torch.optim.AdamW.step() + update_ema() over a flat fp32 parameter buffer (train_JPDVT.py:281,371-372,
 * 36-46): g is scaled by grad_scale first (1/world_size after a SUM all-reduce), `step` counts from 1; optionally refreshes
 * the bf16 operand copy of the parameters in the same pass.  36 B of HBM traffic per parameter (+2 B for the bf16 copy). */
int jpdvt_adamw_ema(float* p, const float* g, float* m, float* v, float* ema_or_null, jpdvt_bf16* p_bf16_or_null, int64_t n,
                    int64_t step, float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay,
                    float ema_decay, void* stream);
/* ---- data-parallel optimizer step over NVLink / NVSwitch peer memory (world > 1) ----------------------------------------
 * Replaces DistributedDataParallel's gradient all-reduce followed by AdamW.step() and update_ema() (train_JPDVT.py:231,
 * 370-372, 36-46) with ONE kernel per rank: reduce-scatter of the fp32 gradients (multimem.ld_reduce through the NVSwitch
 * multicast object, or plain peer loads), AdamW + EMA on the rank's own contiguous slice of the flat parameter space, and
 * all-gather of the refreshed bf16 operand copy (multimem.st, or peer stores).  The caller maps every rank's buffers into
 * this process (CUDA VMM / symmetric memory) and passes the addresses AS SEEN FROM THIS PROCESS. */
#define JPDVT_MAX_PEERS 8
#define JPDVT_MAX_F32_RANGES 16
typedef struct jpdvt_peer_step {
  int32_t world;                 /* ranks on this NVLink domain, 2..JPDVT_MAX_PEERS */
  int32_t rank;
  int64_t shard_begin;           /* this rank's slice [shard_begin, shard_end) of the flat parameter index space; */
  int64_t shard_end;             /*   multiples of 8; every parameter belongs to exactly one rank */
  uint32_t epoch;                /* barrier token: strictly increasing from call to call, identical on every rank */
  uint32_t timeout_ms;           /* a barrier gives up after this long and sets *status (1: gradients, 2: weights) */
  const float* grads[JPDVT_MAX_PEERS];         /* rank q's flat fp32 gradient buffer (grads[rank] is the local one) */
  jpdvt_bf16* weights_bf16[JPDVT_MAX_PEERS];   /* rank q's flat bf16 operand buffer */
  float* params[JPDVT_MAX_PEERS];              /* rank q's flat fp32 parameter buffer (written only inside f32_ranges) */
  void* signals[JPDVT_MAX_PEERS];              /* rank q's flag pad: uint32[2 * JPDVT_MAX_PEERS], zero before the first call */
  const float* grads_mc;         /* multicast address of the gradient buffers (NULL: plain peer loads / stores) */
  jpdvt_bf16* weights_mc;        /* multicast address of the bf16 operand buffers */
  float* params_mc;              /* multicast address of the fp32 parameter buffers */
  int32_t n_f32_ranges;          /* parameters the kernels read in fp32 (biases, timestep MLP, head): index ranges */
  int32_t reserved;              /*   [f32_ranges[2k], f32_ranges[2k+1]) whose fp32 values are replicated to every rank too */
  int64_t f32_ranges[2 * JPDVT_MAX_F32_RANGES];
  uint32_t* local_sync;          /* device uint32, zero before the first call (CTA counter of this rank) */
  int32_t* status;               /* device int32, stays zero unless a barrier timed out */
  uint32_t* epoch_dev;           /* optional device uint32: when set, the barrier token is *epoch_dev + 1 and the kernel stores it back
                                  * (a CUDA-graph replay cannot carry a changing kernel parameter); must agree across the ranks */
} jpdvt_peer_step;
/* p, m, v, ema: this rank's full-length flat fp32 buffers - only [shard_begin, shard_end) is read and written (the fp32
 * master state of a parameter lives on its owner, except the f32_ranges, which every rank receives; gather the rest for
 * checkpoints).  `step` counts from 1; grad_scale is applied
 * to the summed gradient (1 / world for DDP's mean).  Every rank must call this once per step on its own stream. */
int jpdvt_adamw_ema_peer(const jpdvt_peer_step* px, float* p, float* m, float* v, float* ema_or_null, int64_t step,
                         float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay, float ema_decay,
                         void* stream);
/* The two optimizer steps with the step count (Adam's bias corrections) read from DEVICE memory instead of a host argument -
 * what a CUDA-graph replay of the whole training step needs (train_JPDVT.py:340-376 without the host in the loop).  The caller
 * increments *step_dev (stream ordered) before the call; the peer form also takes its barrier token from px->epoch_dev. */
int jpdvt_adamw_ema_dev(float* p, const float* g, float* m, float* v, float* ema_or_null, jpdvt_bf16* p_bf16_or_null, int64_t n,
                        const int64_t* step_dev, float grad_scale, float lr, float beta1, float beta2, float eps,
                        float weight_decay, float ema_decay, void* stream);
int jpdvt_adamw_ema_peer_dev(const jpdvt_peer_step* px, float* p, float* m, float* v, float* ema_or_null, const int64_t* step_dev,
                             float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay,
                             float ema_decay, void* stream);
/* out[b][c][r] = in[b][r][c]: refreshes the [in, out] weight copies of jpdvt_weights_t after an optimizer step. */
int jpdvt_transpose_bf16(const jpdvt_bf16* in, jpdvt_bf16* out, int batch, int rows, int cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JPDVT_B200_H_ */
