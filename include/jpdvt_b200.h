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

#define JPDVT_ABI_VERSION 1
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
/* Checks that the current device is sm_100 (B200); returns JPDVT_ERR_UNSUPPORTED otherwise.  No CPU fallback exists. */
int jpdvt_device_check(void);

/* ---- single kernels ------------------------------------------------------------------------------------------ */

/* If delta != NULL: x += delta first (the gated branch output of the previous GEMM, bf16; x is updated in place - the
 * residual add of models.py:120-121 rides on this pass).  Then y = LayerNorm(x, eps=1e-6, no affine) * (1 + scale[b]) +
 * shift[b], b = row / tokens; sample b reads its 768-wide vectors at shift + b*mod_stride, scale + b*mod_stride
 * (mod_stride 0 = one conditioning row for the whole batch).
 * Replaces nn.LayerNorm + modulate: models.py:19-20,107,109,120-121,131,140. */
int jpdvt_ln_modulate_fwd(float* x, const jpdvt_bf16* delta_or_null, const float* shift, const float* scale,
                          int64_t mod_stride, jpdvt_bf16* y, int64_t rows, int tokens, void* stream);

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
/* x = cols . w_patch^T + bias + pos_embed[row % tokens] + x_t[row] . w_in_t   (PatchEmbed conv as GEMM + time_emb_in +
 * pos_embed, models.py:280-281).  cols = jpdvt_patchify(img); bias = x_embedder.proj.bias + time_emb_in.bias;
 * w_in_t = time_emb_in.weight^T as [8,768] fp32; pos = pos_embed [tokens,768] fp32. */
int jpdvt_gemm_patch_embed(const jpdvt_bf16* cols, const jpdvt_bf16* w_patch, const float* bias, const float* x_t,
                           const float* w_in_t, const float* pos, float* x, int64_t m, int tokens, void* stream);
/* te_out[row,:8] = w2 . silu(y[row] . w1^T + b1) + b2   (time_emb_out1 -> SiLU -> time_emb_out2, models.py:288-290) */
int jpdvt_final_head_fwd(const jpdvt_bf16* y, const jpdvt_bf16* w1, const float* b1, const float* w2, const float* b2,
                         float* te_out, int64_t m, void* stream);

/* softmax(q k^T / 8) v per (sample, head) on the fused QKV tensor [batch*tokens, 2304] -> [batch*tokens, 768]
 * (timm Attention.forward -> F.scaled_dot_product_attention, called from models.py:108,120). */
int jpdvt_attention_fwd(const jpdvt_bf16* qkv, jpdvt_bf16* out, int batch, int tokens, void* stream);

/* im2col of 16x16 patches, k = c*256 + py*16 + px (the flattened Conv2d weight order), fp32 -> bf16. */
int jpdvt_patchify(const float* img, jpdvt_bf16* cols, int batch, int image_size, void* stream);
/* models.py:227-240 */
int jpdvt_unpatchify(const float* y, float* img, int batch, int image_size, void* stream);

/* c = Linear(SiLU(Linear(sinusoid_256(t)))) and silu(c)   (TimestepEmbedder, models.py:27-64).
 * t: int64[n] model timesteps, or NULL: then every row uses map[*step_ptr] (or *step_ptr when map is NULL). */
int jpdvt_timestep_embed(const int64_t* t, int n, const int32_t* step_ptr, const int32_t* map, const float* w0,
                         const float* b0, const float* w2, const float* b2, float* c, float* silu_c, void* stream);
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
/* DDIM update (gaussian_diffusion.py:559-578, with the `condition` argument the reference call at :547 forgot):
 * eps = (recip[t]*x_t - x0)/recipm1[t]; sample = sqrt_abp[t]*x0 + dir[t]*eps + [t != 0]*sigma[t]*noise.  PARITY UNPINNED:
 * the reference's ddim_sample raises TypeError, so only the oracle restatement of those lines checks this kernel. */
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
  int32_t reserved;
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
  const float* step_noise;      /* [num_steps or 1, B, T, 8] noise drawn per step (torch-generated in parity mode) */
  int64_t step_noise_stride;    /* elements between consecutive steps' noise (0 = reuse one tensor) */
  float* x0;                    /* [B, T, 8] scratch: pred_xstart of the current step */
  float* sample;                /* [B, T, 8] scratch / result: sample of the current step */
  float* traj_x0;               /* optional [num_steps, B, T, 8] record of every pred_xstart, or NULL */
  float* traj_sample;           /* optional [num_steps, B, T, 8] record of every sample, or NULL */
} jpdvt_sampler;

/* SpacedDiffusion.p_sample_loop (gaussian_diffusion.py:433-529 through respace.py:89-129): runs steps
 * [first_step, last_step) of the reverse loop (0 = the first executed step, i.e. respaced index num_steps-1)
 * without any host round trip; the result of the last executed step is left in sampler->sample. */
int jpdvt_sample_loop(const jpdvt_weights* w_host, const jpdvt_workspace* ws_host, const jpdvt_sampler* s_host,
                      const float* condition, const float* noise, int batch, int first_step, int last_step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JPDVT_B200_H_ */
