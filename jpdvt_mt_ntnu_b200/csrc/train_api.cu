// Training step on the device: the denoiser forward that keeps its activations (tape) and the three-stage backward.
// Mirrors what autograd does for the reference's `loss.backward()` through DiT.forward (image_model/models.py:273-293,
// called from diffusion/gaussian_diffusion.py:817 inside train_JPDVT.py:357-370), as explicit kernel launches.
#include <cstdlib>

#include "../../include/jpdvt_b200.h"
#include "common.cuh"

using namespace jp;

namespace {

typedef const __nv_bfloat16* bfp;
typedef __nv_bfloat16* bfm;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BFM(p) reinterpret_cast<__nv_bfloat16*>(p)
#define JP_TRY(expr) do { int rc_ = (expr); if (rc_ != kOk) return rc_; } while (0)

inline long long n_mod_of(int depth) { return static_cast<long long>(depth) * 6 * kHidden + 2 * kHidden; }

// out = a[M,K] . w[N,K]^T (+ bias or zeros) with the given epilogue
int gemm(int epi, bfp a, long long lda, bfp w, long long ldw, const float* bias, void* out, long long ldo, long long m, int n,
         int k, cudaStream_t st, float* out2 = nullptr, bfp aux = nullptr, const float* w2 = nullptr, const float* b2 = nullptr,
         int b_mn = 0) {
  GemmParams p{};
  p.M = static_cast<int>(m); p.N = n; p.K = k; p.tokens = 1;
  p.bias = bias; p.out = out; p.ldo = ldo; p.out2 = out2; p.aux = aux; p.w2 = w2; p.b2 = b2; p.b_mn = b_mn;
  return launch_gemm(epi, a, lda, w, ldw, p, st);
}

// JPDVT_DGRAD_MN=0: data-gradient GEMMs read transposed weight copies ([in, out], refreshed after every optimizer step) as
// K-major operands (A/B knob); default: they read nn.Linear's own [out, in] weight as an MN-major B operand - no copies.
bool dgrad_mn() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("JPDVT_DGRAD_MN"); v = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
// dX[m, n_in] = dY[m, k_out] . W, W = the layer's [k_out, n_in] weight (w) or its transposed copy (w_t)
// colsum (dGELU form only, nullable): [n_in] column sums of the output, accumulated - the bias gradient of the layer below
int dgrad(int epi, bfp dy, bfp w, bfp w_t, void* out, long long m, int n_in, int k_out, const float* zeros, cudaStream_t st,
          bfp aux = nullptr, float* colsum = nullptr) {
  if (dgrad_mn()) return gemm(epi, dy, k_out, w, n_in, zeros, out, n_in, m, n_in, k_out, st, colsum, aux, nullptr, nullptr, 1);
  return gemm(epi, dy, k_out, w_t, k_out, zeros, out, n_in, m, n_in, k_out, st, colsum, aux);
}

// JPDVT_BWD_FUSED=0: separate LayerNorm-backward / gate-backward / partial-sum launches (A/B knob); default: each
// LayerNorm backward also runs the gate backward of the residual branch below it (backward.cu: ln_gate_bwd_kernel)
bool bwd_fused() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("JPDVT_BWD_FUSED"); v = (e != nullptr && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

}  // namespace

extern "C" {

int jpdvt_gemm_wgrad(const jpdvt_bf16* p, const jpdvt_bf16* q, float* dw, float* scratch, int64_t m, int out_rows, int n_cols,
                     void* stream) {
  if (!p || !q || !dw) return set_error(kErrBadArg, "gemm_wgrad: null pointer");
  return launch_wgrad(BF(p), out_rows, BF(q), n_cols, dw, scratch, m, out_rows, n_cols, ST(stream));
}
int64_t jpdvt_wgrad_scratch_floats(int64_t m, int out_rows, int n_cols) { return wgrad_scratch_floats(m, out_rows, n_cols); }

int jpdvt_gemm_dgelu(const jpdvt_bf16* a, const jpdvt_bf16* w, const jpdvt_bf16* pre, jpdvt_bf16* out, int64_t m, int n, int k,
                     void* stream) {
  if (m == 0) return kOk;
  if (!a || !w || !pre || !out) return set_error(kErrBadArg, "gemm_dgelu: null pointer");
  return gemm(EPI_DGELU_BF16, BF(a), k, BF(w), k, nullptr, out, n, m, n, k, ST(stream), nullptr, BF(pre));
}
int jpdvt_gemm_dgrad(const jpdvt_bf16* dy, const jpdvt_bf16* w, const jpdvt_bf16* gprime_or_null, jpdvt_bf16* out_bf16_or_null,
                     float* out_f32_or_null, float* colsum_or_null, int64_t m, int n_in, int k_out, void* stream) {
  if (m == 0) return kOk;
  if (!dy || !w || (!out_bf16_or_null == !out_f32_or_null)) return set_error(kErrBadArg, "gemm_dgrad: null pointer / exactly one output");
  if (gprime_or_null && !out_bf16_or_null) return set_error(kErrBadArg, "gemm_dgrad: the dGELU form writes bf16");
  if (colsum_or_null && !gprime_or_null) return set_error(kErrBadArg, "gemm_dgrad: column sums come with the dGELU form only");
  static float* zeros_of[64] = {};     // the bias slot of the shared epilogues: 4 x 768 zeros, allocated once per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return set_error(kErrUnsupported, "gemm_dgrad: device index %d", dev);
  if (zeros_of[dev] == nullptr) {
    if (cudaMalloc(&zeros_of[dev], 4 * kHidden * sizeof(float)) != cudaSuccess ||
        cudaMemset(zeros_of[dev], 0, 4 * kHidden * sizeof(float)) != cudaSuccess)
      return set_error(kErrCuda, "gemm_dgrad: cannot allocate the zero bias");
  }
  const float* zeros = zeros_of[dev];
  if (n_in > 4 * kHidden) return set_error(kErrBadArg, "gemm_dgrad: n_in=%d > %d", n_in, 4 * kHidden);
  const int epi = gprime_or_null ? EPI_DGELU_BF16 : (out_bf16_or_null ? EPI_BIAS_BF16 : EPI_BIAS_F32);
  void* out = out_bf16_or_null ? static_cast<void*>(out_bf16_or_null) : static_cast<void*>(out_f32_or_null);
  return gemm(epi, BF(dy), k_out, BF(w), n_in, zeros, out, n_in, m, n_in, k_out, ST(stream), colsum_or_null, BF(gprime_or_null), nullptr, nullptr, 1);
}
int jpdvt_attention_bwd(const jpdvt_bf16* qkv, const jpdvt_bf16* o, const jpdvt_bf16* d_o, const float* lse2, jpdvt_bf16* dqkv,
                        float* dbias_or_null, int batch, int tokens, void* stream) {
  if (!qkv || !o || !d_o || !lse2 || !dqkv) return set_error(kErrBadArg, "attention_bwd: null pointer");
  return launch_attention_bwd(BF(qkv), BF(o), BF(d_o), lse2, BFM(dqkv), dbias_or_null, batch, tokens, ST(stream));
}
int64_t jpdvt_bwd_part_floats(int batch, int tokens) { return bwd_part_floats(batch, tokens); }

int jpdvt_gate_bwd(const float* dx, const jpdvt_bf16* y, const float* gate, int64_t gate_stride, jpdvt_bf16* dy, float* dgate,
                   int64_t dgate_stride, float* dbias_or_null, float* part, int batch, int tokens, void* stream) {
  if (!dx || !y || !gate || !dy || !dgate || !part) return set_error(kErrBadArg, "gate_bwd: null pointer");
  return launch_gate_bwd(dx, BF(y), gate, gate_stride, BFM(dy), dgate, dgate_stride, dbias_or_null, part, batch, tokens, ST(stream));
}
int jpdvt_ln_modulate_bwd(const float* x, const float* dxn, const float* scale, int64_t mod_stride, float* dx, int accumulate,
                          float* dshift, float* dscale, int64_t dmod_stride, jpdvt_bf16* dx_bf16_or_null, float* part, int batch,
                          int tokens, void* stream) {
  if (!x || !dxn || !scale || !dx || !dshift || !dscale || !part) return set_error(kErrBadArg, "ln_modulate_bwd: null pointer");
  return launch_ln_modulate_bwd(x, dxn, scale, mod_stride, dx, accumulate, dshift, dscale, dmod_stride, BFM(dx_bf16_or_null),
                                part, batch, tokens, ST(stream));
}
int jpdvt_ln_gate_bwd(const float* x, const float* dxn, const float* scale, int64_t mod_stride, float* dx, int accumulate,
                      float* dshift, float* dscale, int64_t dmod_stride, jpdvt_bf16* dx_bf16_or_null, const jpdvt_bf16* y_or_null,
                      const float* gate, int64_t gate_stride, jpdvt_bf16* dy, float* dgate, int64_t dgate_stride,
                      float* dbias_or_null, int batch, int tokens, void* stream) {
  if (!x || !dxn || !scale || !dx || !dshift || !dscale) return set_error(kErrBadArg, "ln_gate_bwd: null pointer");
  return launch_ln_gate_bwd(x, dxn, scale, mod_stride, dx, accumulate, dshift, dscale, dmod_stride, BFM(dx_bf16_or_null), BF(y_or_null),
                            gate, gate_stride, BFM(dy), dgate, dgate_stride, dbias_or_null, batch, tokens, ST(stream));
}
int jpdvt_colsum_bf16(const jpdvt_bf16* src, int64_t rows, int cols, float* out, void* stream) {
  if (rows == 0) return kOk;
  if (!src || !out) return set_error(kErrBadArg, "colsum: null pointer");
  return launch_colsum_bf16(BF(src), cols, rows, cols, out, ST(stream));
}
int jpdvt_colsum_f32(const float* src, int64_t rows, int cols, float* out, void* stream) {
  if (rows == 0) return kOk;
  if (!src || !out) return set_error(kErrBadArg, "colsum: null pointer");
  return launch_colsum_f32(src, cols, rows, cols, out, ST(stream));
}

int64_t jpdvt_train_wgrad_scratch_floats(int depth, int batch, int tokens) {
  const long long m = static_cast<long long>(batch) * tokens;
  long long need = 0;
  auto upd = [&](long long mm, int r, int c) { const long long v = wgrad_scratch_floats(mm, r, c); if (v > need) need = v; };
  upd(m, 3 * kHidden, kHidden); upd(m, kHidden, kHidden); upd(m, 4 * kHidden, kHidden); upd(m, kHidden, 4 * kHidden);
  upd(m, 64, kHidden);
  upd(batch, static_cast<int>(n_mod_of(depth)), kHidden); upd(batch, kHidden, kHidden); upd(batch, kHidden, 256);
  return need;
}

// ------------------------------------------------------------------------------------------------ forward with tape
int jpdvt_train_forward(const jpdvt_weights* w, const jpdvt_tape* tp, const float* img, const int64_t* t, const float* x_t,
                        float* te_out, float* img_out, int batch, void* stream) {
  if (!w || !tp || !img || !t || !x_t || !te_out) return set_error(kErrBadArg, "train_forward: null pointer");
  cudaStream_t st = ST(stream);
  const int T = w->tokens, depth = w->depth, S = w->image_size;
  const long long M = static_cast<long long>(batch) * T;
  if (M != tp->rows || batch != tp->batch) return set_error(kErrBadArg, "train_forward: tape sized for %lld rows / %d samples, got %lld / %d", (long long)tp->rows, tp->batch, M, batch);
  if (M > 0x7fffffffLL) return set_error(kErrUnsupported, "train_forward: too many token rows");
  if (img_out != nullptr && tp->yfin32 == nullptr) return set_error(kErrBadArg, "train_forward: image output needs tape.yfin32");
  const long long n_mod = n_mod_of(depth);
  const long long X = M * kHidden;

  // embeddings (models.py:280-281)
  JP_TRY(launch_patchify(img, BFM(tp->cols), batch, S, st));
  {
    GemmParams p{};
    p.M = static_cast<int>(M); p.N = kHidden; p.K = kHidden; p.tokens = T;
    p.bias = w->b_embed; p.out = tp->x; p.ldo = kHidden; p.xt = x_t; p.w_in_t = w->w_in_t; p.pos = w->pos;
    JP_TRY(launch_gemm(EPI_PATCH_EMBED_F32, BF(tp->cols), kHidden, BF(w->w_patch), kHidden, p, st));
  }
  // conditioning (models.py:282-284,119,134): per-sample timesteps -> tensor-core adaLN over bf16 silu(c) (GEMV for <= 8 rows)
  JP_TRY(launch_timestep_embed(reinterpret_cast<const long long*>(t), batch, nullptr, nullptr, w->t_w0, w->t_b0, w->t_w2, w->t_b2,
                               tp->c, tp->silu_c, tp->thid, tp->feat, tp->tpre, st));
  JP_TRY(launch_cast_bf16(tp->silu_c, BFM(tp->silu_c_bf16), static_cast<long long>(batch) * kHidden, st));
  if (batch <= 8) {
    JP_TRY(launch_adaln_gemv(tp->silu_c, batch, BF(w->w_ada), w->b_ada, tp->mod, static_cast<int>(n_mod), st));
  } else {
    JP_TRY(gemm(EPI_BIAS_F32, BF(tp->silu_c_bf16), kHidden, BF(w->w_ada), kHidden, w->b_ada, tp->mod, n_mod, batch,
                static_cast<int>(n_mod), kHidden, st));
  }
  const __nv_bfloat16* pending = nullptr;
  const float* pending_gate = nullptr;
  for (int i = 0; i < depth; ++i) {
    const float* mod = tp->mod + static_cast<long long>(i) * 6 * kHidden;
    bfm xn1 = BFM(tp->xn1) + i * X, qkv = BFM(tp->qkv) + i * M * 3 * kHidden, att = BFM(tp->att) + i * X;
    bfm y1 = BFM(tp->y1) + i * X, xn2 = BFM(tp->xn2) + i * X, hpre = BFM(tp->hpre) + i * M * 4 * kHidden;
    bfm h = BFM(tp->h) + i * M * 4 * kHidden, y2 = BFM(tp->y2) + i * X;
    float* x_a = tp->x + static_cast<long long>(2 * i) * X;       // LN1 input (block input after the pending residual add)
    float* x_b = tp->x + static_cast<long long>(2 * i + 1) * X;   // LN2 input
    const float* x_prev = (i == 0) ? tp->x : tp->x + static_cast<long long>(2 * i - 1) * X;
    if (pending == nullptr) {
      JP_TRY(launch_ln_modulate(x_a, nullptr, nullptr, nullptr, 0, mod, mod + kHidden, n_mod, xn1, M, T, st));
    } else {
      JP_TRY(launch_ln_modulate(x_prev, x_a, pending, pending_gate, n_mod, mod, mod + kHidden, n_mod, xn1, M, T, st));
    }
    JP_TRY(gemm(EPI_BIAS_BF16, xn1, kHidden, BF(w->w_qkv) + static_cast<long long>(i) * 3 * kHidden * kHidden, kHidden,
                w->b_qkv + static_cast<long long>(i) * 3 * kHidden, qkv, 3 * kHidden, M, 3 * kHidden, kHidden, st));
    JP_TRY(launch_attention(qkv, att, tp->lse2 + static_cast<long long>(i) * batch * kHeads * T, batch, T, st));
    JP_TRY(gemm(EPI_BIAS_BF16, att, kHidden, BF(w->w_proj) + static_cast<long long>(i) * kHidden * kHidden, kHidden,
                w->b_proj + static_cast<long long>(i) * kHidden, y1, kHidden, M, kHidden, kHidden, st));
    JP_TRY(launch_ln_modulate(x_a, x_b, y1, mod + 2 * kHidden, n_mod, mod + 3 * kHidden, mod + 4 * kHidden, n_mod, xn2, M, T, st));
    // fc1: pre-activations (kept for gelu') and activations
    {   // fc1 with the activation AND its derivative (kept for the backward's dGELU epilogue) as one epilogue
      GemmParams p{};
      p.M = static_cast<int>(M); p.N = 4 * kHidden; p.K = kHidden; p.tokens = 1;
      p.bias = w->b_fc1 + static_cast<long long>(i) * 4 * kHidden; p.out = h; p.ldo = 4 * kHidden; p.out_aux = hpre;
      JP_TRY(launch_gemm(EPI_BIAS_GELU_GRAD_BF16, xn2, kHidden, BF(w->w_fc1) + static_cast<long long>(i) * 4 * kHidden * kHidden,
                         kHidden, p, st));
    }
    JP_TRY(gemm(EPI_BIAS_BF16, h, 4 * kHidden, BF(w->w_fc2) + static_cast<long long>(i) * 4 * kHidden * kHidden, 4 * kHidden,
                w->b_fc2 + static_cast<long long>(i) * kHidden, y2, kHidden, M, kHidden, 4 * kHidden, st));
    pending = y2;
    pending_gate = mod + 5 * kHidden;
  }
  {
    const float* mod = tp->mod + static_cast<long long>(depth) * 6 * kHidden;
    float* x_f = tp->x + static_cast<long long>(2 * depth) * X;
    const float* x_prev = (depth == 0) ? tp->x : tp->x + static_cast<long long>(2 * depth - 1) * X;
    if (pending == nullptr) {
      JP_TRY(launch_ln_modulate(x_f, nullptr, nullptr, nullptr, 0, mod, mod + kHidden, n_mod, BFM(tp->xnf), M, T, st));
    } else {
      JP_TRY(launch_ln_modulate(x_prev, x_f, pending, pending_gate, n_mod, mod, mod + kHidden, n_mod, BFM(tp->xnf), M, T, st));
    }
    JP_TRY(gemm(EPI_BIAS_BF16_F32, BF(tp->xnf), kHidden, BF(w->w_final), kHidden, w->b_final, tp->yfin, kHidden, M, kHidden,
                kHidden, st, img_out != nullptr ? tp->yfin32 : nullptr));
    GemmParams hp{};
    hp.M = static_cast<int>(M); hp.N = 64; hp.K = kHidden; hp.tokens = T;
    hp.bias = w->b_head1; hp.out = te_out; hp.ldo = kLatent; hp.w2 = w->w_head2; hp.b2 = w->b_head2; hp.out2 = tp->headpre;
    JP_TRY(launch_gemm(EPI_HEAD, BF(tp->yfin), kHidden, BF(w->w_head1), kHidden, hp, st));
    if (img_out != nullptr) JP_TRY(launch_unpatchify(tp->yfin32, img_out, batch, S, st));
  }
  return kOk;
}

// ------------------------------------------------------------------------------------------------ backward: head
int jpdvt_train_backward_head(const jpdvt_weights* w, const jpdvt_weights_t* wt, const jpdvt_tape* tp,
                              const jpdvt_bwd_scratch* s, const jpdvt_grads* g, const float* d_te, const float* d_img,
                              void* stream) {
  if (!w || !wt || !tp || !s || !g || !d_te) return set_error(kErrBadArg, "train_backward_head: null pointer");
  cudaStream_t st = ST(stream);
  const int T = w->tokens, depth = w->depth, batch = tp->batch;
  const long long M = tp->rows, X = M * kHidden, n_mod = n_mod_of(depth);
  // te = W2 silu(pre) + b2, pre = W1 y + b1 (models.py:288-290)
  JP_TRY(launch_head_bwd(d_te, tp->headpre, w->w_head2, BFM(s->dpre), g->w_head2, g->b_head2, g->b_head1, M, st));
  JP_TRY(launch_wgrad(BF(s->dpre), 64, BF(tp->yfin), kHidden, g->w_head1, s->wgrad_scratch, M, 64, kHidden, st, true));
  // dy = dpre . W1 (+ the image head's gradient through unpatchify)
  JP_TRY(dgrad(EPI_BIAS_F32, BF(s->dpre), BF(w->w_head1), BF(wt->w_head1_t), s->dxn, M, kHidden, 64, s->zeros, st));
  if (d_img != nullptr) JP_TRY(launch_unpatchify_bwd(d_img, s->dxn, batch, w->image_size, 1, st));
  JP_TRY(launch_cast_bf16(s->dxn, BFM(s->dy), X, st));
  JP_TRY(launch_colsum_f32(s->dxn, kHidden, M, kHidden, g->b_final, st));
  JP_TRY(launch_wgrad(BF(s->dy), kHidden, BF(tp->xnf), kHidden, g->w_final, s->wgrad_scratch, M, kHidden, kHidden, st, true));
  JP_TRY(dgrad(EPI_BIAS_F32, BF(s->dy), BF(w->w_final), BF(wt->w_final_t), s->dxn, M, kHidden, kHidden, s->zeros, st));
  const float* mod = tp->mod + static_cast<long long>(depth) * 6 * kHidden;
  float* dmod = s->dmod + static_cast<long long>(depth) * 6 * kHidden;
  if (bwd_fused()) {
    if (depth == 0)
      return launch_ln_gate_bwd(tp->x, s->dxn, mod + kHidden, n_mod, s->dx, 0, dmod, dmod + kHidden, n_mod, BFM(s->dy), nullptr,
                                nullptr, 0, nullptr, nullptr, 0, nullptr, batch, T, st);
    // ... and the gate backward of the last block's MLP branch (models.py:121), whose dy the block stage starts from
    const int l = depth - 1;
    const float* mod_l = tp->mod + static_cast<long long>(l) * 6 * kHidden;
    float* dmod_l = s->dmod + static_cast<long long>(l) * 6 * kHidden;
    return launch_ln_gate_bwd(tp->x + static_cast<long long>(2 * depth) * X, s->dxn, mod + kHidden, n_mod, s->dx, 0, dmod,
                              dmod + kHidden, n_mod, nullptr, BF(tp->y2) + l * X, mod_l + 5 * kHidden, n_mod, BFM(s->dy),
                              dmod_l + 5 * kHidden, n_mod, g->b_fc2 + static_cast<long long>(l) * kHidden, batch, T, st);
  }
  JP_TRY(launch_ln_modulate_bwd(tp->x + static_cast<long long>(2 * depth) * X, s->dxn, mod + kHidden, n_mod, s->dx, 0, dmod,
                                dmod + kHidden, n_mod, depth == 0 ? BFM(s->dy) : nullptr, s->part, batch, T, st));
  return kOk;
}

// ------------------------------------------------------------------------------------------------ backward: one block
int jpdvt_train_backward_block(const jpdvt_weights* w, const jpdvt_weights_t* wt, const jpdvt_tape* tp,
                               const jpdvt_bwd_scratch* s, const jpdvt_grads* g, int i, void* stream) {
  if (!w || !wt || !tp || !s || !g) return set_error(kErrBadArg, "train_backward_block: null pointer");
  if (i < 0 || i >= w->depth) return set_error(kErrBadArg, "train_backward_block: block %d out of range", i);
  cudaStream_t st = ST(stream);
  const int T = w->tokens, depth = w->depth, batch = tp->batch;
  const long long M = tp->rows, X = M * kHidden, n_mod = n_mod_of(depth);
  const long long H4 = 4LL * kHidden, H3 = 3LL * kHidden;
  const float* mod = tp->mod + static_cast<long long>(i) * 6 * kHidden;
  float* dmod = s->dmod + static_cast<long long>(i) * 6 * kHidden;
  bfp xn1 = BF(tp->xn1) + i * X, qkv = BF(tp->qkv) + i * M * H3, att = BF(tp->att) + i * X, y1 = BF(tp->y1) + i * X;
  bfp xn2 = BF(tp->xn2) + i * X, hpre = BF(tp->hpre) + i * M * H4, h = BF(tp->h) + i * M * H4, y2 = BF(tp->y2) + i * X;
  bfm dy = BFM(s->dy), dh = BFM(s->dh), dqkv = BFM(s->dqkv), datt = BFM(s->datt);

  // ---- MLP branch: x_out = x_mid + gate_mlp * fc2(gelu(fc1(xn2)))            (models.py:121)
  const bool fused = bwd_fused();      // then dy = gate_mlp * dx came with the LayerNorm backward of the stage before
  if (!fused) JP_TRY(launch_gate_bwd(s->dx, y2, mod + 5 * kHidden, n_mod, dy, dmod + 5 * kHidden, n_mod, g->b_fc2 + static_cast<long long>(i) * kHidden, s->part, batch, T, st));
  JP_TRY(launch_wgrad(dy, kHidden, h, H4, g->w_fc2 + static_cast<long long>(i) * kHidden * H4, s->wgrad_scratch, M, kHidden, static_cast<int>(H4), st, true));
  JP_TRY(dgrad(EPI_DGELU_BF16, dy, BF(w->w_fc2) + static_cast<long long>(i) * kHidden * H4, BF(wt->w_fc2_t) + static_cast<long long>(i) * H4 * kHidden,
               dh, M, static_cast<int>(H4), kHidden, nullptr, st, hpre, g->b_fc1 + static_cast<long long>(i) * H4));   // + db_fc1 = column sums of dh
  JP_TRY(launch_wgrad(dh, H4, xn2, kHidden, g->w_fc1 + static_cast<long long>(i) * H4 * kHidden, s->wgrad_scratch, M, static_cast<int>(H4), kHidden, st, true));
  JP_TRY(dgrad(EPI_BIAS_F32, dh, BF(w->w_fc1) + static_cast<long long>(i) * H4 * kHidden, BF(wt->w_fc1_t) + static_cast<long long>(i) * kHidden * H4,
               s->dxn, M, kHidden, static_cast<int>(H4), s->zeros, st));
  // ---- attention branch: x_mid = x_in + gate_msa * proj(attn(qkv(xn1)))     (models.py:120)
  if (fused) {
    JP_TRY(launch_ln_gate_bwd(tp->x + static_cast<long long>(2 * i + 1) * X, s->dxn, mod + 4 * kHidden, n_mod, s->dx, 1,
                              dmod + 3 * kHidden, dmod + 4 * kHidden, n_mod, nullptr, y1, mod + 2 * kHidden, n_mod, dy,
                              dmod + 2 * kHidden, n_mod, g->b_proj + static_cast<long long>(i) * kHidden, batch, T, st));
  } else {
    JP_TRY(launch_ln_modulate_bwd(tp->x + static_cast<long long>(2 * i + 1) * X, s->dxn, mod + 4 * kHidden, n_mod, s->dx, 1,
                                  dmod + 3 * kHidden, dmod + 4 * kHidden, n_mod, nullptr, s->part, batch, T, st));
    JP_TRY(launch_gate_bwd(s->dx, y1, mod + 2 * kHidden, n_mod, dy, dmod + 2 * kHidden, n_mod, g->b_proj + static_cast<long long>(i) * kHidden, s->part, batch, T, st));
  }
  JP_TRY(launch_wgrad(dy, kHidden, att, kHidden, g->w_proj + static_cast<long long>(i) * kHidden * kHidden, s->wgrad_scratch, M, kHidden, kHidden, st, true));
  JP_TRY(dgrad(EPI_BIAS_BF16, dy, BF(w->w_proj) + static_cast<long long>(i) * kHidden * kHidden, BF(wt->w_proj_t) + static_cast<long long>(i) * kHidden * kHidden,
               datt, M, kHidden, kHidden, s->zeros, st));
  // dQ, dK, dV and (folded into the epilogue) the qkv bias gradient = column sums of dqkv
  JP_TRY(launch_attention_bwd(qkv, att, datt, tp->lse2 + static_cast<long long>(i) * batch * kHeads * T, dqkv,
                              g->b_qkv + static_cast<long long>(i) * H3, batch, T, st));
  JP_TRY(launch_wgrad(dqkv, H3, xn1, kHidden, g->w_qkv + static_cast<long long>(i) * H3 * kHidden, s->wgrad_scratch, M, static_cast<int>(H3), kHidden, st, true));
  JP_TRY(dgrad(EPI_BIAS_F32, dqkv, BF(w->w_qkv) + static_cast<long long>(i) * H3 * kHidden, BF(wt->w_qkv_t) + static_cast<long long>(i) * kHidden * H3,
               s->dxn, M, kHidden, static_cast<int>(H3), s->zeros, st));
  if (fused) {
    if (i == 0)     // bf16(dx0) for the patch-embedding weight gradient
      return launch_ln_gate_bwd(tp->x, s->dxn, mod + kHidden, n_mod, s->dx, 1, dmod, dmod + kHidden, n_mod, dy, nullptr, nullptr, 0,
                                nullptr, nullptr, 0, nullptr, batch, T, st);
    // ... and the gate backward of block i - 1's MLP branch
    const float* mod_l = mod - 6 * kHidden;
    float* dmod_l = dmod - 6 * kHidden;
    return launch_ln_gate_bwd(tp->x + static_cast<long long>(2 * i) * X, s->dxn, mod + kHidden, n_mod, s->dx, 1, dmod, dmod + kHidden,
                              n_mod, nullptr, BF(tp->y2) + (i - 1) * X, mod_l + 5 * kHidden, n_mod, dy, dmod_l + 5 * kHidden, n_mod,
                              g->b_fc2 + static_cast<long long>(i - 1) * kHidden, batch, T, st);
  }
  JP_TRY(launch_ln_modulate_bwd(tp->x + static_cast<long long>(2 * i) * X, s->dxn, mod + kHidden, n_mod, s->dx, 1, dmod,
                                dmod + kHidden, n_mod, i == 0 ? dy : nullptr, s->part, batch, T, st));
  return kOk;
}

// ------------------------------------------------------------------------------------------------ backward: embeddings + conditioning
int jpdvt_train_backward_embed(const jpdvt_weights* w, const jpdvt_weights_t* wt, const jpdvt_tape* tp,
                               const jpdvt_bwd_scratch* s, const jpdvt_grads* g, const float* x_t, void* stream) {
  if (!w || !wt || !tp || !s || !g || !x_t) return set_error(kErrBadArg, "train_backward_embed: null pointer");
  cudaStream_t st = ST(stream);
  const int depth = w->depth, batch = tp->batch;
  const long long M = tp->rows, n_mod = n_mod_of(depth);
  const long long BH = static_cast<long long>(batch) * kHidden;
  // x0 = cols . Wp^T + b_patch + b_in + pos + x_t . Win^T                      (models.py:280-281); s->dy holds bf16(dx0)
  JP_TRY(launch_wgrad(BF(s->dy), kHidden, BF(tp->cols), kHidden, g->w_patch, s->wgrad_scratch, M, kHidden, kHidden, st, true));
  JP_TRY(launch_colsum_f32(s->dx, kHidden, M, kHidden, g->b_patch, st));
  JP_TRY(cudaMemcpyAsync(g->b_in, g->b_patch, kHidden * sizeof(float), cudaMemcpyDeviceToDevice, st) == cudaSuccess
             ? kOk : set_error(kErrCuda, "train_backward_embed: bias copy failed"));
  JP_TRY(launch_win_grad(s->dx, x_t, g->w_in, M, st));

  // adaLN linears: mod = W_ada . silu(c) + b_ada                                (models.py:113-116,133-136)
  float* dsilu = s->small_f32;            // [batch, 768]
  float* dc = s->small_f32 + BH;
  float* dhid = s->small_f32 + 2 * BH;
  float* dtpre = s->small_f32 + 3 * BH;
  bfm dc_bf = BFM(s->small_bf16), hid_bf = BFM(s->small_bf16) + BH, dtpre_bf = BFM(s->small_bf16) + 2 * BH, feat_bf = BFM(s->small_bf16) + 3 * BH;
  JP_TRY(launch_colsum_f32(s->dmod, n_mod, batch, static_cast<int>(n_mod), g->b_ada, st));
  JP_TRY(launch_cast_bf16(s->dmod, BFM(s->dmod_bf16), static_cast<long long>(batch) * n_mod, st));
  JP_TRY(launch_wgrad(BF(s->dmod_bf16), n_mod, BF(tp->silu_c_bf16), kHidden, g->w_ada, s->wgrad_scratch, batch, static_cast<int>(n_mod), kHidden, st, true));
  JP_TRY(dgrad(EPI_BIAS_F32, BF(s->dmod_bf16), BF(w->w_ada), BF(wt->w_ada_t), dsilu, batch, kHidden, static_cast<int>(n_mod), s->zeros, st));
  // c = W2 . silu(tpre) + b2 ; tpre = W0 . feat + b0                            (models.py:61-64)
  JP_TRY(launch_silu_bwd(dsilu, tp->c, dc, dc_bf, BH, st));
  JP_TRY(launch_colsum_f32(dc, kHidden, batch, kHidden, g->t_b2, st));
  JP_TRY(launch_silu_fwd_bf16(tp->tpre, hid_bf, BH, st));
  JP_TRY(launch_wgrad(dc_bf, kHidden, hid_bf, kHidden, g->t_w2, s->wgrad_scratch, batch, kHidden, kHidden, st, true));
  JP_TRY(gemm(EPI_BIAS_F32, dc_bf, kHidden, BF(wt->t_w2_t), kHidden, s->zeros, dhid, kHidden, batch, kHidden, kHidden, st));
  JP_TRY(launch_silu_bwd(dhid, tp->tpre, dtpre, dtpre_bf, BH, st));
  JP_TRY(launch_colsum_f32(dtpre, kHidden, batch, kHidden, g->t_b0, st));
  JP_TRY(launch_cast_bf16(tp->feat, feat_bf, static_cast<long long>(batch) * 256, st));
  JP_TRY(launch_wgrad(dtpre_bf, kHidden, feat_bf, 256, g->t_w0, s->wgrad_scratch, batch, kHidden, 256, st, true));
  return kOk;
}

}  // extern "C"
