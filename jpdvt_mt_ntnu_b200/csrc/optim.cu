// Fused AdamW + EMA (+ bf16 operand refresh) over flat parameter buffers, and the bf16 transpose used to refresh the
// [in, out] weight copies the data-gradient GEMMs read.
//
// Replaces torch.optim.AdamW(lr 1e-4, betas 0.9/0.999, eps 1e-8, weight_decay 0).step() followed by update_ema(ema,
// model, 0.9999) (image_model/train_JPDVT.py:281,371-372,36-46: ~150 tensors, ~450 small launches) with ONE pass:
// 20 B read + 18 B written per parameter (p, g, m, v, ema -> p, m, v, ema, bf16(p)).
#include "../../include/jpdvt_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace jp {

// AdamW exactly as torch.optim.AdamW (decoupled weight decay, bias correction via step-dependent scalars computed on the
// host): p *= 1 - lr*wd; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
// step_ptr != null: the step count lives on the device (a CUDA-graph replay of the training step cannot carry step-dependent
// scalars as kernel parameters) and the two bias-correction factors are derived from it, once per block
__global__ void __launch_bounds__(256)
adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 float* __restrict__ ema, __nv_bfloat16* __restrict__ p_bf16, long long n4, float grad_scale, float lr,
                 float beta1, float beta2, float eps, float weight_decay, float step_size, float inv_sqrt_bc2, float ema_decay,
                 const long long* __restrict__ step_ptr) {
  if (step_ptr != nullptr) {
    __shared__ float s_corr[2];
    if (threadIdx.x == 0) {
      const double st = static_cast<double>(*step_ptr);
      s_corr[0] = static_cast<float>(static_cast<double>(lr) / (1.0 - pow(static_cast<double>(beta1), st)));
      s_corr[1] = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(beta2), st)));
    }
    __syncthreads();
    step_size = s_corr[0];
    inv_sqrt_bc2 = s_corr[1];
  }
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 pp = reinterpret_cast<float4*>(p)[i];
  const float4 gg4 = __ldcs(reinterpret_cast<const float4*>(g) + i);
  float4 mm = reinterpret_cast<float4*>(m)[i];
  float4 vv = reinterpret_cast<float4*>(v)[i];
  float pa[4] = {pp.x, pp.y, pp.z, pp.w};
  const float ga[4] = {gg4.x * grad_scale, gg4.y * grad_scale, gg4.z * grad_scale, gg4.w * grad_scale};
  float ma[4] = {mm.x, mm.y, mm.z, mm.w};
  float va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    pa[k] *= 1.0f - lr * weight_decay;
    ma[k] = beta1 * ma[k] + (1.0f - beta1) * ga[k];
    va[k] = beta2 * va[k] + (1.0f - beta2) * ga[k] * ga[k];
    const float denom = sqrtf(va[k]) * inv_sqrt_bc2 + eps;
    pa[k] -= step_size * (ma[k] / denom);
  }
  reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
  reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
  reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
  if (ema != nullptr) {   // update_ema (train_JPDVT.py:36-46): ema = decay * ema + (1 - decay) * p
    float4 ee = reinterpret_cast<float4*>(ema)[i];
    ee.x = ema_decay * ee.x + (1.0f - ema_decay) * pa[0]; ee.y = ema_decay * ee.y + (1.0f - ema_decay) * pa[1];
    ee.z = ema_decay * ee.z + (1.0f - ema_decay) * pa[2]; ee.w = ema_decay * ee.w + (1.0f - ema_decay) * pa[3];
    reinterpret_cast<float4*>(ema)[i] = ee;
  }
  if (p_bf16 != nullptr) {
    uint2 u;
    u.x = pack_bf16(pa[0], pa[1]); u.y = pack_bf16(pa[2], pa[3]);
    reinterpret_cast<uint2*>(p_bf16)[i] = u;
  }
}

// out[b][c][r] = in[b][r][c], bf16, 32x32 tiles through padded shared memory
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][34];
  const long long base = static_cast<long long>(blockIdx.z) * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8)
    if (r0 + j < rows && c0 + tx < cols) tile[j][tx] = in[base + static_cast<long long>(r0 + j) * cols + c0 + tx];
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
    if (c0 + j < cols && r0 + tx < rows) out[base + static_cast<long long>(c0 + j) * rows + r0 + tx] = tile[tx][j];
}

}  // namespace jp

using namespace jp;

extern "C" {

int jpdvt_adamw_ema(float* p, const float* g, float* m, float* v, float* ema_or_null, jpdvt_bf16* p_bf16_or_null, int64_t n,
                    int64_t step, float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay,
                    float ema_decay, void* stream) {
  if (n == 0) return kOk;
  if (!p || !g || !m || !v) return set_error(kErrBadArg, "adamw_ema: null pointer");
  if (n & 3) return set_error(kErrBadArg, "adamw_ema: element count must be a multiple of 4");
  if (step < 1) return set_error(kErrBadArg, "adamw_ema: step counts from 1");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  const float step_size = static_cast<float>(lr / bc1);
  const float inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  const long long n4 = n / 4;
  adamw_ema_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, ema_or_null, reinterpret_cast<__nv_bfloat16*>(p_bf16_or_null), n4, grad_scale, lr, beta1, beta2, eps,
      weight_decay, step_size, inv_sqrt_bc2, ema_decay, nullptr);
  return check_launch("adamw_ema_kernel");
}

int jpdvt_adamw_ema_dev(float* p, const float* g, float* m, float* v, float* ema_or_null, jpdvt_bf16* p_bf16_or_null, int64_t n,
                        const int64_t* step_dev, float grad_scale, float lr, float beta1, float beta2, float eps,
                        float weight_decay, float ema_decay, void* stream) {
  if (n == 0) return kOk;
  if (!p || !g || !m || !v || !step_dev) return set_error(kErrBadArg, "adamw_ema_dev: null pointer");
  if (n & 3) return set_error(kErrBadArg, "adamw_ema_dev: element count must be a multiple of 4");
  const long long n4 = n / 4;
  adamw_ema_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, ema_or_null, reinterpret_cast<__nv_bfloat16*>(p_bf16_or_null), n4, grad_scale, lr, beta1, beta2, eps,
      weight_decay, 0.f, 0.f, ema_decay, reinterpret_cast<const long long*>(step_dev));
  return check_launch("adamw_ema_kernel");
}

int jpdvt_transpose_bf16(const jpdvt_bf16* in, jpdvt_bf16* out, int batch, int rows, int cols, void* stream) {
  if (batch <= 0 || rows <= 0 || cols <= 0) return kOk;
  if (!in || !out) return set_error(kErrBadArg, "transpose: null pointer");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
  transpose_bf16_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), rows, cols);
  return check_launch("transpose_bf16_kernel");
}

}  // extern "C"
