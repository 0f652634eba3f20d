// Training loss of the JPDVT diffusion step and its gradient (diffusion/gaussian_diffusion.py:18-22 `mean_flat`, :835-838):
//   mse[b] = mean_{T*8}((te_tgt - te_out)^2)  [+ mean_{3*S*S}((img_tgt - img_out)^2 * (1 - keep[b, slot]))  when add_mask]
// keep[b, slot] is the per-puzzle slot mask the reference expands to a full-size image (`masks`, :760-779): 1 = the slot is
// shown clean to the network and carries no image loss, 0 = the slot was noised and its reconstruction is scored.
// HBM-bound streaming reductions: one CTA per (sample, chunk) walks its slice with 16-byte loads; partial sums meet in a
// fixed order (per-CTA tree, then chunk 0..n in a second tiny kernel) so the result does not depend on scheduling.
#include "common.cuh"

namespace jp {
namespace {

constexpr int kLossThreads = 256;
constexpr int kLossChunks = 8;      // CTAs per sample over the image term (batch 128 -> 1024 CTAs)

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  if (warp == 0) {
    s = (lane < kLossThreads / 32) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  __syncthreads();
  return s;   // valid in warp 0
}

// slot of pixel (row, col) in a G x G grid of (size / G)-pixel pieces
__device__ __forceinline__ int slot_of(int row, int col, int piece, int grid) { return (row / piece) * grid + (col / piece); }

// part[b, 0] = sum of squared latent errors, part[b, 1 + chunk] = masked image partial sums
__global__ void __launch_bounds__(kLossThreads)
mse_partial_kernel(const float* __restrict__ te_out, const float* __restrict__ te_tgt, long long per_te,
                   const float* __restrict__ img_out, const float* __restrict__ img_tgt, const float* __restrict__ keep,
                   int size, int grid, float* __restrict__ part) {
  __shared__ float red[kLossThreads / 32];
  const int b = blockIdx.x, chunk = blockIdx.y;
  float acc = 0.f;
  if (chunk == 0) {
    const float4* o = reinterpret_cast<const float4*>(te_out + b * per_te);
    const float4* g = reinterpret_cast<const float4*>(te_tgt + b * per_te);
    for (long long i = threadIdx.x; i < per_te / 4; i += kLossThreads) {
      const float4 a = o[i], t = g[i];
      const float dx = t.x - a.x, dy = t.y - a.y, dz = t.z - a.z, dw = t.w - a.w;
      acc += dx * dx + dy * dy + dz * dz + dw * dw;
    }
  } else {
    const long long per_img = 3LL * size * size;
    const long long n4 = per_img / 4, lo = n4 * (chunk - 1) / kLossChunks, hi = n4 * chunk / kLossChunks;
    const float4* o = reinterpret_cast<const float4*>(img_out + b * per_img);
    const float4* g = reinterpret_cast<const float4*>(img_tgt + b * per_img);
    const int piece = size / grid;
    const float* kb = keep + static_cast<long long>(b) * grid * grid;
    for (long long i = lo + threadIdx.x; i < hi; i += kLossThreads) {
      const long long e = i * 4;                       // 4 consecutive pixels of one row (size % 4 == 0) -> one slot when piece % 4 == 0
      const int pix = static_cast<int>(e % (static_cast<long long>(size) * size));
      const int row = pix / size, col = pix - row * size;
      const float4 a = o[i], t = g[i];
      const float d[4] = {t.x - a.x, t.y - a.y, t.z - a.z, t.w - a.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) acc += d[j] * d[j] * (1.0f - __ldg(kb + slot_of(row, col + j, piece, grid)));
    }
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) part[b * (1 + kLossChunks) + chunk] = s;
}

__global__ void mse_finish_kernel(const float* __restrict__ part, int batch, int chunks, float inv_te, float inv_img,
                                  float* __restrict__ loss) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const float* p = part + b * (1 + kLossChunks);
  float v = p[0] * inv_te;
  if (chunks > 0) {
    float s = 0.f;
    for (int c = 1; c <= chunks; ++c) s += p[c];
    v += s * inv_img;
  }
  loss[b] = v;
}

// d_te = dloss[b] * 2 (te_out - te_tgt) / (T*8);   d_img = dloss[b] * 2 (img_out - img_tgt) (1 - keep) / (3*S*S)
__global__ void __launch_bounds__(kLossThreads)
mse_bwd_kernel(const float* __restrict__ te_out, const float* __restrict__ te_tgt, long long per_te,
               const float* __restrict__ img_out, const float* __restrict__ img_tgt, const float* __restrict__ keep, int size,
               int grid, const float* __restrict__ dloss, float* __restrict__ d_te, float* __restrict__ d_img, long long te4,
               long long total4) {
  const long long i = static_cast<long long>(blockIdx.x) * kLossThreads + threadIdx.x;
  if (i >= total4) return;
  if (i < te4) {
    const long long e = i * 4;
    const int b = static_cast<int>(e / per_te);
    const float s = __ldg(dloss + b) * (2.0f / static_cast<float>(per_te));
    const float4 a = reinterpret_cast<const float4*>(te_out)[i], t = reinterpret_cast<const float4*>(te_tgt)[i];
    reinterpret_cast<float4*>(d_te)[i] = make_float4(s * (a.x - t.x), s * (a.y - t.y), s * (a.z - t.z), s * (a.w - t.w));
    return;
  }
  const long long k = i - te4, e = k * 4;
  const long long per_img = 3LL * size * size;
  const int b = static_cast<int>(e / per_img);
  const int pix = static_cast<int>((e - b * per_img) % (static_cast<long long>(size) * size));
  const int row = pix / size, col = pix - row * size, piece = size / grid;
  const float s = __ldg(dloss + b) * (2.0f / static_cast<float>(per_img));
  const float* kb = keep + static_cast<long long>(b) * grid * grid;
  const float4 a = reinterpret_cast<const float4*>(img_out)[k], t = reinterpret_cast<const float4*>(img_tgt)[k];
  float4 o;
  o.x = s * (a.x - t.x) * (1.0f - __ldg(kb + slot_of(row, col, piece, grid)));
  o.y = s * (a.y - t.y) * (1.0f - __ldg(kb + slot_of(row, col + 1, piece, grid)));
  o.z = s * (a.z - t.z) * (1.0f - __ldg(kb + slot_of(row, col + 2, piece, grid)));
  o.w = s * (a.w - t.w) * (1.0f - __ldg(kb + slot_of(row, col + 3, piece, grid)));
  reinterpret_cast<float4*>(d_img)[k] = o;
}

}  // namespace

long long mse_part_floats(int batch) { return static_cast<long long>(batch) * (1 + kLossChunks); }

int launch_mse_loss_fwd(const float* te_out, const float* te_tgt, long long per_te, const float* img_out, const float* img_tgt,
                        const float* keep, int size, int grid, float* part, float* loss, int batch, cudaStream_t stream) {
  if (batch <= 0) return kOk;
  if (per_te <= 0 || (per_te & 3)) return set_error(kErrBadArg, "mse_loss: latent elements per sample must be a positive multiple of 4");
  const bool img = img_out != nullptr;
  if (img && (img_tgt == nullptr || keep == nullptr || grid <= 0 || size <= 0 || size % grid != 0 || (size & 3)))
    return set_error(kErrBadArg, "mse_loss: the image term needs img_tgt, keep and size %% grid == 0, size %% 4 == 0");
  dim3 g(batch, img ? 1 + kLossChunks : 1);
  mse_partial_kernel<<<g, kLossThreads, 0, stream>>>(te_out, te_tgt, per_te, img_out, img_tgt, keep, size, grid, part);
  int rc = check_launch("mse_partial_kernel");
  if (rc != kOk) return rc;
  mse_finish_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(part, batch, img ? kLossChunks : 0, 1.0f / static_cast<float>(per_te),
                                                           img ? 1.0f / (3.0f * size * size) : 0.f, loss);
  return check_launch("mse_finish_kernel");
}

int launch_mse_loss_bwd(const float* te_out, const float* te_tgt, long long per_te, const float* img_out, const float* img_tgt,
                        const float* keep, int size, int grid, const float* dloss, float* d_te, float* d_img, int batch,
                        cudaStream_t stream) {
  if (batch <= 0) return kOk;
  if (per_te <= 0 || (per_te & 3)) return set_error(kErrBadArg, "mse_loss_bwd: latent elements per sample must be a positive multiple of 4");
  const bool img = d_img != nullptr;
  if (img && (img_out == nullptr || img_tgt == nullptr || keep == nullptr || grid <= 0 || size <= 0 || size % grid != 0 || (size & 3)))
    return set_error(kErrBadArg, "mse_loss_bwd: the image term needs img_out, img_tgt, keep and size %% grid == 0, size %% 4 == 0");
  const long long te4 = per_te * batch / 4;
  const long long total4 = te4 + (img ? 3LL * size * size * batch / 4 : 0);
  mse_bwd_kernel<<<static_cast<unsigned>((total4 + kLossThreads - 1) / kLossThreads), kLossThreads, 0, stream>>>(
      te_out, te_tgt, per_te, img_out, img_tgt, keep, size, grid, dloss, d_te, d_img, te4, total4);
  return check_launch("mse_bwd_kernel");
}

}  // namespace jp
