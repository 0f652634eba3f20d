// Shared declarations between the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace jp {

constexpr int kHidden = 768;   // JPDVT width (reference: models.py:409-410, hard-wired 8->768 / 768->64 heads at :176-179)
constexpr int kHeads = 12;
constexpr int kHeadDim = 64;
constexpr int kLatent = 8;     // width of the positional latent the diffusion runs on

enum Status : int {
  kOk = 0,
  kErrBadArg = -1,
  kErrCuda = -2,
  kErrUnsupported = -3,
  kErrDriver = -4,
};

int set_error(int code, const char* fmt, ...);

// Launch with the programmatic-stream-serialization attribute (JPDVT_PDL=0: plain stream order).  Only for kernels that call
// griddep_wait() before their first global access.
bool pdl_enabled();
template <typename... KP, typename... A>
inline cudaError_t launch_pdl(void (*kern)(KP...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<A&&>(args)...);
}
int check_launch(const char* what);
long long launch_count();

// GEMM epilogues (see gemm.cu)
enum Epilogue : int {
  EPI_BIAS_BF16 = 0,        // out_bf16 = acc + bias
  EPI_BIAS_GELU_BF16 = 1,   // out_bf16 = gelu_tanh(acc + bias)
  EPI_GATE_BF16 = 2,        // out_bf16 = gate[row / tokens] * (acc + bias)   (the residual add is fused into the next LN)
  EPI_PATCH_EMBED_F32 = 3,  // out_f32 = acc + bias + pos[row % tokens] + x_t[row,:8] . w_in_t[:, n]
  EPI_BIAS_F32 = 4,         // out_f32 = acc + bias
  EPI_BIAS_BF16_F32 = 5,    // out_bf16 = acc + bias, and (if out2 != null) out2_f32 = acc + bias
  EPI_HEAD = 6,             // out_f32[row, :8] = w2 . silu(acc[:, :64] + bias) + b2     (N == 64); optional pre-activation copy
  EPI_DGELU_BF16 = 7,       // out_bf16 = acc * aux_bf16[row, col], aux = gelu_tanh'(fc1 pre-activation) kept by the forward
  EPI_WGRAD_F32 = 8,        // MN-major operands, split contraction: partial[s][i][j] = sum_m P[m,i] Q[m,j]   (weight gradients)
  EPI_RESID_F32 = 9,        // out_f32 += gate[row / tokens] * (acc + bias): the adaLN-Zero gated residual update, in place, fp32
  EPI_BIAS_GELU_GRAD_BF16 = 11,   // training fc1: out_bf16 = gelu_tanh(acc + bias) and out_aux_bf16 = gelu_tanh'(acc + bias)
  EPI_RESID_TMA_F32 = 12,   // EPI_RESID_F32 with the residual tile moved by TMA: 32x32 fp32 boxes into a per-warp shared-memory
                            // ring (deep prefetch, no registers), updated in place, TMA-stored back
  EPI_RESID_LN_TMA_F32 = 13,   // EPI_RESID_TMA_F32 + the next LayerNorm-modulate of the same rows (ln_out), boxes re-read by TMA
  EPI_RESID_TMA_XB_F32 = 14,   // EPI_RESID_TMA_F32 + a bf16 copy of the updated rows (ln_out) + per-row (sum, sum of squares) partials
                               // (stats_out): the producer half of the LayerNorm folded into the next GEMM (see fold.cu)
  EPI_RESID_LN_F32 = 10,    // EPI_RESID_F32 (N == 768) + the NEXT LayerNorm-modulate of the updated rows:
                            // ln_out_bf16 = LN(out_f32) * (1 + ln_scale[sample]) + ln_shift[sample]
};

struct GemmParams {
  int M, N, K;
  int tokens;            // rows per sample (row -> sample index = row / tokens)
  const float* bias;     // [N]
  void* out;             // primary output, row-major, leading dimension ldo (elements)
  long long ldo;
  float* out2;           // optional fp32 copy (EPI_BIAS_BF16_F32); EPI_DGELU_BF16: optional [N] column sums of the output (bias gradient)
  const float* gate;     // sample b reads gate + b * gate_stride, [N] contiguous
  long long gate_stride;
  const float* xt;       // [M, 8] fp32                     (patch embed)
  const float* w_in_t;   // [8, N] fp32                     (patch embed)
  const float* pos;      // [tokens, N] fp32                (patch embed)
  const float* w2;       // [8, 64] fp32                    (head)
  const float* b2;       // [8] fp32                        (head)
  const __nv_bfloat16* aux;   // [M, N] bf16 (EPI_DGELU_BF16: fc1 pre-activations), leading dimension ldo
  // EPI_WGRAD_F32: M = contraction length (token rows), N = output columns (in_features), wg_rows = output rows
  int wg_rows, split, split_len;
  // EPI_RESID_LN_F32: the LayerNorm-modulate that consumes the updated residual rows (models.py:120-121, 19-20)
  int tma_out;                // bf16 epilogues (bias, bias+GELU): leave through TMA store boxes instead of per-thread stores (set by launch_gemm)
  __nv_bfloat16* out_aux;     // EPI_BIAS_GELU_GRAD_BF16: [M, N] bf16, leading dimension ldo
  __nv_bfloat16* ln_out;      // [M, N] bf16
  const float* ln_shift;      // sample b reads ln_shift + b * ln_stride, [N]
  const float* ln_scale;
  long long ln_stride;
  // LayerNorm folded into the consuming GEMM (uniform conditioning only; fold.cu): a row's LayerNorm is two scalars, so
  //   modulate(LN(x)) . W^T + b  =  rstd * (bf16(x) . W'^T) - rstd * mean * u + v,   W' = W * (1 + scale), u = rowsum(W'), v = b + W . shift
  // producer (EPI_RESID_TMA_XB_F32): stats_out[row * stats_slots + slot] = (sum, sum of squares) over one column group
  // consumer (EPI_BIAS_BF16 / EPI_BIAS_GELU_BF16 with stats_in != null): bias = v, fold_u = u, K = LayerNorm width
  float2* stats_out;
  const float2* stats_in;
  const float* fold_u;
  int stats_slots;
  int b_mn;                   // the B operand is given as [K, N] row-major (the contraction index is the ROW): data-gradient GEMMs read
                              // nn.Linear's own [out, in] weight, dX = dY . W, instead of a transposed copy (MN-major UMMA operand)
  int reverse_m;              // walk the row blocks from the last to the first (set by launch_gemm from sweep_reverse())
};

// Sweep direction of the next row-streaming launch (LayerNorm, GEMMs, attention).  Every activation of a forward is larger
// than what L2 keeps of it (M = 36,864: x 113 MB, qkv 170 MB, hidden 226 MB against a 126 MB L2), so a consumer that walks
// the rows in its producer's order finds the head of its input evicted and re-reads all of it from HBM; walking them the
// other way round it starts on the rows written last, which are still resident.  forward_impl therefore flips the
// direction with every launch (JPDVT_SWEEP=0: always upwards).  Single-op C-ABI calls leave it at 0.
int sweep_reverse();
void set_sweep_reverse(int reverse);

// dW[wg_rows, n_cols] (fp32) = P[M, wg_rows]^T . Q[M, n_cols]   (both bf16 row-major); `partial` is scratch of
// split * wg_rows * n_cols floats (see wgrad_scratch_floats)
int launch_wgrad(const __nv_bfloat16* pmat, long long ldp, const __nv_bfloat16* qmat, long long ldq, float* out,
                 float* partial, long long m, int out_rows, int n_cols, cudaStream_t stream, bool out_zeroed = false);
long long wgrad_scratch_floats(long long m, int out_rows, int n_cols);

// a: bf16 [M, K] row-major (lda elements); w: bf16 [N, K] row-major (nn.Linear layout)
int launch_gemm(int epi, const __nv_bfloat16* a, long long lda, const __nv_bfloat16* w, long long ldw, const GemmParams& p,
                cudaStream_t stream);

int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int batch, int tokens, cudaStream_t stream);
// attention_tc.cu: tcgen05 / TMEM forward for the sizes attention_tc_supported() names (lse2 optional, as launch_attention)
bool attention_tc_supported(int tokens);
int launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int batch, int tokens, cudaStream_t stream);
// K-major bf16 [rows, cols] tensor map, {64 cols x box_rows} boxes, 128-byte swizzle (gemm.cu)
int make_tmap_bf16_kmajor(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld, int box_rows);
// dbias (nullable): the qkv Linear's bias gradient, dbias[c] += sum_rows dqkv[row, c] (accumulated with atomics)
int launch_attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* o, const __nv_bfloat16* d_o, const float* lse2,
                         __nv_bfloat16* dqkv, float* dbias, int batch, int tokens, cudaStream_t stream);
bool attention_bwd_tc_supported(int tokens);
int launch_attention_bwd_tc(const __nv_bfloat16* qkv, const __nv_bfloat16* o, const __nv_bfloat16* d_o, const float* lse2,
                            __nv_bfloat16* dqkv, float* dbias, int batch, int tokens, cudaStream_t stream);

// fold.cu: W' = W (1 + scale), u = rowsum(W'), v = b + W . shift for the qkv and fc1 matrices of every block (mod row 0)
int launch_fold_ln(const __nv_bfloat16* w_qkv, const __nv_bfloat16* w_fc1, const float* b_qkv, const float* b_fc1, const float* mod,
                   __nv_bfloat16* w_fold, float* fold_u, float* fold_v, int depth, cudaStream_t stream);

// puzzle.cu
int launch_gather_pieces(const float* src, float* dst, const int* perm, const unsigned char* keep, int batch, int channels,
                         int size, int grid, cudaStream_t stream);
int launch_crop_pieces(const float* src, float* dst, int batch, int channels, int grid, int in_piece, int out_piece, int off,
                       cudaStream_t stream);
int launch_score_placements(const int* pred, const int* truth, int batch, int n, int* correct, int* matches, long long* totals,
                            cudaStream_t stream);

// backward.cu
// `part`: scratch of bwd_part_floats(batch, tokens) floats (per-sample partial sums; no atomics on the hot reductions)
long long bwd_part_floats(int batch, int tokens);
int launch_gate_bwd(const float* dx, const __nv_bfloat16* y, const float* gate, long long gate_stride, __nv_bfloat16* dy,
                    float* dgate, long long dgate_stride, float* dbias, float* part, int batch, int tokens, cudaStream_t stream);
int launch_ln_gate_bwd(const float* x, const float* dxn, const float* scale, long long mod_stride, float* dx, int accumulate,
                       float* dshift, float* dscale, long long dmod_stride, __nv_bfloat16* dx_bf16, const __nv_bfloat16* y,
                       const float* gate, long long gate_stride, __nv_bfloat16* dy, float* dgate, long long dgate_stride,
                       float* dbias, int batch, int tokens, cudaStream_t stream);
int launch_ln_modulate_bwd(const float* x, const float* dxn, const float* scale, long long mod_stride, float* dx,
                           int accumulate, float* dshift, float* dscale, long long dmod_stride, __nv_bfloat16* dx_bf16,
                           float* part, int batch, int tokens, cudaStream_t stream);
int launch_colsum_bf16(const __nv_bfloat16* src, long long ld, long long rows, int cols, float* out, cudaStream_t stream);
int launch_colsum_f32(const float* src, long long ld, long long rows, int cols, float* out, cudaStream_t stream);
int launch_head_bwd(const float* dte, const float* pre, const float* w2, __nv_bfloat16* dpre, float* dw2, float* db2,
                    float* db1, long long rows, cudaStream_t stream);
int launch_cast_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t stream);
int launch_silu_bwd(const float* grad, const float* pre, float* out, __nv_bfloat16* out_bf16, long long n, cudaStream_t stream);
int launch_win_grad(const float* dx0, const float* xt, float* dw, long long rows, cudaStream_t stream);
int launch_unpatchify_bwd(const float* dimg, float* dy, int batch, int size, int accumulate, cudaStream_t stream);
int launch_gelu(__nv_bfloat16* pre_to_grad, __nv_bfloat16* out, long long n, cudaStream_t stream);
int launch_silu_fwd_bf16(const float* pre, __nv_bfloat16* out, long long n, cudaStream_t stream);

// x_out = x_in + gate[b] * delta (optional: delta bf16, gate may be null = 1), then y = LN(x_out) * (1 + scale) + shift
int launch_ln_modulate(const float* x_in, float* x_out, const __nv_bfloat16* delta, const float* gate, long long gate_stride,
                       const float* shift, const float* scale, long long mod_stride, __nv_bfloat16* y, long long rows,
                       int tokens, cudaStream_t stream);
int launch_patchify(const float* img, __nv_bfloat16* cols, int batch, int size, cudaStream_t stream);
int launch_unpatchify(const float* y, float* img, int batch, int size, cudaStream_t stream);
int launch_timestep_embed(const long long* t, int n, const int* step_ptr, const int* map, const float* w0, const float* b0,
                          const float* w2, const float* b2, float* c, float* silu_c, float* hid_scratch, float* feat_out,
                          float* pre_out, cudaStream_t stream, int step_stride = 0);
int launch_adaln_gemv(const float* silu_c, int rows, const __nv_bfloat16* w, const float* bias, float* out, int n_out,
                      cudaStream_t stream);
int launch_posterior(const float* x0, const float* xt, const float* noise, const float* coef1, const float* coef2,
                     const float* logvar, const long long* t, const int* step_ptr, float* mean, float* sample, long long n,
                     long long per_sample, cudaStream_t stream, const long long* noise_key = nullptr, int noise_step = 0);
// Philox4x32-10 normals (the generator posterior_kernel uses when no noise tensor is given); raw (nullable): the 32-bit stream
int launch_philox_normal(float* out, unsigned* raw, long long n, int step, const long long* key, cudaStream_t stream);
// loss.cu: per-sample MSE terms of training_losses and their gradient
long long mse_part_floats(int batch);
int launch_mse_loss_fwd(const float* te_out, const float* te_tgt, long long per_te, const float* img_out, const float* img_tgt,
                        const float* keep, int size, int grid, float* part, float* loss, int batch, cudaStream_t stream);
int launch_mse_loss_bwd(const float* te_out, const float* te_tgt, long long per_te, const float* img_out, const float* img_tgt,
                        const float* keep, int size, int grid, const float* dloss, float* d_te, float* d_img, int batch,
                        cudaStream_t stream);
int launch_ddim(const float* x0, const float* xt, const float* noise, const float* recip, const float* recipm1,
                const float* sqrt_abp, const float* dir, const float* sigma, const long long* t, const int* step_ptr,
                float* sample, long long n, long long per_sample, cudaStream_t stream);
int launch_q_sample(const float* x0, const float* noise, const float* sqrt_ac, const float* sqrt_1mac, const long long* t,
                    const float* keep_mask, float* out, long long n, long long per_sample, cudaStream_t stream);
int launch_assign_scores(const double* scores, int batch, int n, double sentinel, int* order, int* pred,
                         cudaStream_t stream);
int launch_assign_latents(const float* latents, const float* canon, int batch, int grid, int tok, double sentinel,
                          int* order, int* pred, double* scores_out, cudaStream_t stream);

}  // namespace jp
