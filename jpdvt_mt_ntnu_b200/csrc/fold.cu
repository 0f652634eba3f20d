// LayerNorm-modulate folded into the GEMM that consumes it (sampling path, batch-uniform timestep).
//
// Reference: x -> modulate(LayerNorm(x), shift, scale) -> Linear(W, b)   (image_model/models.py:19-20, 107-121).  In the
// sampling loop every puzzle of the batch sees the same timestep (gaussian_diffusion.py:509), so shift / scale are one
// vector per block and step, and for a token row x with mean m and rstd r = (var + 1e-6)^-1/2
//
//     y[n] = sum_j ((x[j] - m) r (1 + scale[j]) + shift[j]) W[n, j] + b[n]
//          = r * sum_j x[j] W'[n, j]  -  r m u[n]  +  v[n]
//     W'[n, j] = W[n, j] (1 + scale[j]),   u[n] = sum_j W'[n, j],   v[n] = b[n] + sum_j shift[j] W[n, j]
//
// The GEMM therefore contracts bf16(x) - written by the epilogue of the GEMM that produced x (EPI_RESID_TMA_XB_F32,
// together with the row's (sum, sum of squares)) - against W', and its epilogue applies r, r m, u and v: the stand-alone
// LayerNorm kernel (and its fp32 re-read of the residual stream) disappears for 23 of the 25 LayerNorms of a forward.
// This file folds the weights: one launch per diffusion step over all 12 x (qkv + fc1) matrices (99 MB in, 99 MB out).
// u is summed over the ROUNDED W', which makes  acc - m u  exactly the contraction of (x - m) with the operand the
// tensor cores see.
#include "common.cuh"
#include "ptx.cuh"

namespace jp {

constexpr int kFoldWarps = 8;

__global__ void __launch_bounds__(kFoldWarps * 32)
fold_ln_kernel(const __nv_bfloat16* __restrict__ w_qkv, const __nv_bfloat16* __restrict__ w_fc1, const float* __restrict__ b_qkv,
               const float* __restrict__ b_fc1, const float* __restrict__ mod, __nv_bfloat16* __restrict__ w_fold,
               float* __restrict__ fold_u, float* __restrict__ fold_v, int depth) {
  constexpr int kRows = 7 * kHidden;                       // 3 * 768 qkv rows + 4 * 768 fc1 rows per block
  const int lane = threadIdx.x & 31;
  const long long gw = static_cast<long long>(blockIdx.x) * kFoldWarps + (threadIdx.x >> 5);
  if (gw >= static_cast<long long>(depth) * kRows) return;
  const int blk = static_cast<int>(gw / kRows), n = static_cast<int>(gw % kRows);
  const float* m = mod + static_cast<long long>(blk) * 6 * kHidden;   // shift_msa scale_msa gate_msa shift_mlp scale_mlp gate_mlp
  const __nv_bfloat16* src;
  const float *shift, *scale;
  float b;
  if (n < 3 * kHidden) {
    src = w_qkv + (static_cast<long long>(blk) * 3 * kHidden + n) * kHidden;
    shift = m; scale = m + kHidden;
    b = b_qkv[blk * 3 * kHidden + n];
  } else {
    const int r = n - 3 * kHidden;
    src = w_fc1 + (static_cast<long long>(blk) * 4 * kHidden + r) * kHidden;
    shift = m + 3 * kHidden; scale = m + 4 * kHidden;
    b = b_fc1[blk * 4 * kHidden + r];
  }
  __nv_bfloat16* dst = w_fold + gw * kHidden;
  float usum = 0.f, vsum = 0.f;
#pragma unroll
  for (int k = 0; k < kHidden / 256; ++k) {
    const int j0 = (k * 32 + lane) * 8;
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src + j0));
    const float4 sc0 = __ldg(reinterpret_cast<const float4*>(scale + j0)), sc1 = __ldg(reinterpret_cast<const float4*>(scale + j0 + 4));
    const float4 sh0 = __ldg(reinterpret_cast<const float4*>(shift + j0)), sh1 = __ldg(reinterpret_cast<const float4*>(shift + j0 + 4));
    const float sc[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
    const float sh[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
    const uint32_t in[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t out[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float w0 = __uint_as_float(in[e] << 16), w1 = __uint_as_float(in[e] & 0xffff0000u);
      vsum = fmaf(w0, sh[2 * e], vsum);
      vsum = fmaf(w1, sh[2 * e + 1], vsum);
      out[e] = pack_bf16(w0 * (1.0f + sc[2 * e]), w1 * (1.0f + sc[2 * e + 1]));
      usum += __uint_as_float(out[e] << 16) + __uint_as_float(out[e] & 0xffff0000u);
    }
    *reinterpret_cast<uint4*>(dst + j0) = make_uint4(out[0], out[1], out[2], out[3]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    usum += __shfl_xor_sync(0xffffffffu, usum, o);
    vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
  }
  if (lane == 0) {
    fold_u[gw] = usum;
    fold_v[gw] = b + vsum;
  }
}

int launch_fold_ln(const __nv_bfloat16* w_qkv, const __nv_bfloat16* w_fc1, const float* b_qkv, const float* b_fc1, const float* mod,
                   __nv_bfloat16* w_fold, float* fold_u, float* fold_v, int depth, cudaStream_t stream) {
  const long long rows = static_cast<long long>(depth) * 7 * kHidden;
  const unsigned blocks = static_cast<unsigned>((rows + kFoldWarps - 1) / kFoldWarps);
  fold_ln_kernel<<<blocks, kFoldWarps * 32, 0, stream>>>(w_qkv, w_fc1, b_qkv, b_fc1, mod, w_fold, fold_u, fold_v, depth);
  return check_launch("fold_ln_kernel");
}

}  // namespace jp
