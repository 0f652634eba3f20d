// Memory-bound kernels of the JPDVT denoiser / diffusion step (sm_100a).
//   ln_modulate      : LayerNorm(eps 1e-6, no affine) fused with adaLN modulate, fp32 in -> bf16 out
//                      (image_model/models.py:19-20,107,109,120-121,131,140)
//   patchify         : [B,3,S,S] fp32 -> im2col bf16 [B*T, 768] for the patch-embed GEMM (timm PatchEmbed, models.py:169)
//   unpatchify       : models.py:227-240
//   timestep_embed   : models.py:27-64 (sinusoid -> Linear -> SiLU -> Linear) + SiLU(c) for the adaLN linears
//   adaln_gemv       : all 13 adaLN modulation linears for a handful of conditioning rows (models.py:113-116,133-136)
//   posterior        : q_posterior mean + p_sample noise update (diffusion/gaussian_diffusion.py:234-254,424-430)
//   q_sample         : gaussian_diffusion.py:217-232 (+ the masked blend of :800)
#include "common.cuh"
#include "ptx.cuh"

namespace jp {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------- LN + modulate
// One warp per 768-wide row: 6 float4 per lane (coalesced 512 B per warp load), two-pass statistics in registers,
// 8-byte bf16x4 stores.  When `delta` is given the gated branch output of the previous GEMM is added to the fp32
// residual stream first (x += delta, models.py:120-121) and x is written back, so the residual add costs no extra pass.
// Algorithmic traffic per row: 3072 B read + 1536 B written (+ 1536 B read + 3072 B written with delta).
constexpr int kLnWarps = 8;

template <bool HAS_DELTA, bool COPY_X = false>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_modulate_kernel(const float* __restrict__ x_in, float* __restrict__ x_out, const __nv_bfloat16* __restrict__ delta,
                   const float* __restrict__ gate, long long gate_stride, const float* __restrict__ shift,
                   const float* __restrict__ scale, long long mod_stride, __nv_bfloat16* __restrict__ y, long long rows,
                   int tokens, int reverse) {
  griddep_wait();
  griddep_launch_dependents();
  // blocks are dispatched in index order: `reverse` makes the kernel walk the rows downwards (sweep_reverse(), common.cuh)
  const long long blk = reverse ? static_cast<long long>(gridDim.x) - 1 - blockIdx.x : blockIdx.x;
  const long long row = blk * kLnWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long sample = row / tokens;
  const float4* xr = reinterpret_cast<const float4*>(x_in + row * kHidden);
  float4 v[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) v[j] = __ldcs(xr + lane + 32 * j);
  if constexpr (HAS_DELTA) {
    // residual update fused into this pass: x_out = x_in + gate[b] * delta   (models.py:120-121)
    const uint2* dr = reinterpret_cast<const uint2*>(delta + row * kHidden);
    float4* xo = reinterpret_cast<float4*>(x_out + row * kHidden);
    uint2 d[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) d[j] = __ldcs(dr + lane + 32 * j);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      float4 g = make_float4(1.f, 1.f, 1.f, 1.f);
      if (gate != nullptr) g = __ldg(reinterpret_cast<const float4*>(gate + sample * gate_stride) + lane + 32 * j);
      v[j].x = fmaf(g.x, __uint_as_float(d[j].x << 16), v[j].x); v[j].y = fmaf(g.y, __uint_as_float(d[j].x & 0xffff0000u), v[j].y);
      v[j].z = fmaf(g.z, __uint_as_float(d[j].y << 16), v[j].z); v[j].w = fmaf(g.w, __uint_as_float(d[j].y & 0xffff0000u), v[j].w);
      xo[lane + 32 * j] = v[j];
    }
  }
  if constexpr (COPY_X) {
    // x_out = x_in: the sampling loop's hoisted embedding enters the residual stream on this pass (api.cu: jpdvt_sample_loop)
    float4* xo = reinterpret_cast<float4*>(x_out + row * kHidden);
#pragma unroll
    for (int j = 0; j < 6; ++j) xo[lane + 32 * j] = v[j];
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mean = warp_sum(s) * (1.0f / kHidden);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / kHidden) + 1e-6f);
  const float4* sh = reinterpret_cast<const float4*>(shift + sample * mod_stride);
  const float4* sc = reinterpret_cast<const float4*>(scale + sample * mod_stride);
  uint2* yr = reinterpret_cast<uint2*>(y + row * kHidden);
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const float4 a = __ldg(sh + lane + 32 * j);
    const float4 b = __ldg(sc + lane + 32 * j);
    float o0 = fmaf(v[j].x * rstd, 1.0f + b.x, a.x);
    float o1 = fmaf(v[j].y * rstd, 1.0f + b.y, a.y);
    float o2 = fmaf(v[j].z * rstd, 1.0f + b.z, a.z);
    float o3 = fmaf(v[j].w * rstd, 1.0f + b.w, a.w);
    uint2 u;
    u.x = pack_bf16(o0, o1);
    u.y = pack_bf16(o2, o3);
    yr[lane + 32 * j] = u;
  }
}

// x_out may alias x_in (in-place residual update) or be a fresh buffer (training keeps every LN input for the backward).
int launch_ln_modulate(const float* x_in, float* x_out, const __nv_bfloat16* delta, const float* gate, long long gate_stride,
                       const float* shift, const float* scale, long long mod_stride, __nv_bfloat16* y, long long rows,
                       int tokens, cudaStream_t stream) {
  if (rows <= 0) return kOk;
  if (tokens <= 0) return set_error(kErrBadArg, "ln_modulate: tokens must be positive");
  if (delta != nullptr && x_out == nullptr) return set_error(kErrBadArg, "ln_modulate: residual update needs an output buffer");
  const unsigned blocks = static_cast<unsigned>((rows + kLnWarps - 1) / kLnWarps);
  cudaError_t e;
  if (delta != nullptr)
    e = launch_pdl(ln_modulate_kernel<true>, dim3(blocks), dim3(kLnWarps * 32), 0, stream, x_in, x_out, delta, gate, gate_stride, shift, scale,
                   mod_stride, y, rows, tokens, sweep_reverse());
  else if (x_out != nullptr && x_out != x_in)
    e = launch_pdl(ln_modulate_kernel<false, true>, dim3(blocks), dim3(kLnWarps * 32), 0, stream, x_in, x_out, delta, gate, gate_stride, shift,
                   scale, mod_stride, y, rows, tokens, sweep_reverse());
  else
    e = launch_pdl(ln_modulate_kernel<false>, dim3(blocks), dim3(kLnWarps * 32), 0, stream, x_in, x_out, delta, gate, gate_stride, shift, scale,
                   mod_stride, y, rows, tokens, sweep_reverse());
  if (e != cudaSuccess) return set_error(kErrCuda, "ln_modulate_kernel: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return check_launch("ln_modulate_kernel");
}

// ---------------------------------------------------------------------------------------------- patchify / unpatchify
// cols[(b, ty, tx), c*256 + py*16 + px] = img[b, c, ty*16 + py, tx*16 + px]; 8 elements (32 B in, 16 B out) per thread.
__global__ void patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ cols, long long total_chunks,
                                int size) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total_chunks) return;
  const int g = size / 16;
  const int chunk = static_cast<int>(idx % 96);   // 96 chunks of 8 per 768-wide row
  const long long m = idx / 96;
  const int tx = static_cast<int>(m % g);
  const int ty = static_cast<int>((m / g) % g);
  const long long b = m / (g * g);
  const int k0 = chunk * 8;
  const int c = k0 >> 8, py = (k0 >> 4) & 15, px = k0 & 15;
  const float* src = img + ((b * 3 + c) * size + (ty * 16 + py)) * static_cast<long long>(size) + tx * 16 + px;
  const float4 u = __ldg(reinterpret_cast<const float4*>(src));
  const float4 v = __ldg(reinterpret_cast<const float4*>(src) + 1);
  uint4 o;
  o.x = pack_bf16(u.x, u.y); o.y = pack_bf16(u.z, u.w); o.z = pack_bf16(v.x, v.y); o.w = pack_bf16(v.z, v.w);
  *reinterpret_cast<uint4*>(cols + m * kHidden + k0) = o;
}

int launch_patchify(const float* img, __nv_bfloat16* cols, int batch, int size, cudaStream_t stream) {
  if (size % 16 != 0 || size <= 0) return set_error(kErrBadArg, "patchify: image size %d is not a multiple of 16", size);
  const long long g = size / 16;
  const long long total = static_cast<long long>(batch) * g * g * 96;
  if (total == 0) return kOk;
  const int threads = 256;
  patchify_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, stream>>>(img, cols, total, size);
  return check_launch("patchify_kernel");
}

// img[b, c, h*16 + p, w*16 + q] = y[(b, h, w), (p*16 + q)*3 + c]
__global__ void unpatchify_kernel(const float* __restrict__ y, float* __restrict__ img, long long total, int size) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = size / 16;
  const int col = static_cast<int>(idx % size);
  const int row = static_cast<int>((idx / size) % size);
  const int c = static_cast<int>((idx / (static_cast<long long>(size) * size)) % 3);
  const long long b = idx / (3LL * size * size);
  const int h = row >> 4, p = row & 15, w = col >> 4, q = col & 15;
  img[idx] = __ldg(y + ((b * g + h) * g + w) * static_cast<long long>(kHidden) + (p * 16 + q) * 3 + c);
}

int launch_unpatchify(const float* y, float* img, int batch, int size, cudaStream_t stream) {
  const long long total = 3LL * batch * size * size;
  if (total == 0) return kOk;
  unpatchify_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(y, img, total, size);
  return check_launch("unpatchify_kernel");
}

// ---------------------------------------------------------------------------------------------- timestep embedding
// Two grid-wide phases (hidden = SiLU(W0 . sinusoid + b0), then c = W2 . hidden + b2) so the 2.4 MB + 0.8 MB of fp32
// weights are streamed by ~96 CTAs instead of one.  Up to kTeRows conditioning rows per blockIdx.y share every weight
// read; each warp owns one output feature, lanes stride the reduction dimension (coalesced), warp-shuffle reduction.
constexpr int kTeRows = 8;
constexpr int kTeWarps = 8;

// step_stride 0: one device-resident step index for every row; 1: row r reads step_ptr[r] (a whole schedule at once)
__device__ __forceinline__ long long te_timestep(const long long* t, int r, const int* step_ptr, const int* map, int step_stride) {
  if (t != nullptr) return t[r];
  const int idx = step_ptr[r * step_stride];
  return (map != nullptr) ? static_cast<long long>(map[idx]) : static_cast<long long>(idx);
}

__global__ void __launch_bounds__(kTeWarps * 32)
timestep_hidden_kernel(const long long* __restrict__ t, int n, const int* __restrict__ step_ptr, int step_stride,
                       const int* __restrict__ map, const float* __restrict__ w0, const float* __restrict__ b0, float* __restrict__ hid,
                       float* __restrict__ feat_out, float* __restrict__ pre_out) {
  __shared__ float feat[kTeRows][256];
  const int r0 = blockIdx.y * kTeRows;
  const int nr = min(kTeRows, n - r0);
  for (int i = threadIdx.x; i < kTeRows * 128; i += kTeWarps * 32) {
    const int r = i >> 7, k = i & 127;
    float c = 0.f, s = 0.f;
    if (r < nr) {
      const long long tv = te_timestep(t, r0 + r, step_ptr, map, step_stride);
      // models.py:52-56: freqs = exp(-ln(10000) * arange(128, fp32) / 128) in fp32, args = t.float() * freqs
      const float f = expf((-9.210340371976184f * static_cast<float>(k)) / 128.0f);
      const float a = static_cast<float>(tv) * f;
      c = cosf(a); s = sinf(a);
    }
    feat[r][k] = c;          // cos first, then sin (models.py:56)
    feat[r][128 + k] = s;
    if (feat_out != nullptr && blockIdx.x == 0 && r < nr) {   // training keeps the sinusoid features for the backward
      feat_out[static_cast<long long>(r0 + r) * 256 + k] = c;
      feat_out[static_cast<long long>(r0 + r) * 256 + 128 + k] = s;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * kTeWarps + warp;
  if (o >= kHidden) return;
  float acc[kTeRows];
#pragma unroll
  for (int r = 0; r < kTeRows; ++r) acc[r] = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = lane + 32 * j;
    const float w = __ldg(w0 + o * 256 + k);
#pragma unroll
    for (int r = 0; r < kTeRows; ++r) acc[r] = fmaf(w, feat[r][k], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kTeRows; ++r) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
    const float b = __ldg(b0 + o);
    for (int r = 0; r < nr; ++r) {
      const float v = acc[r] + b;
      hid[static_cast<long long>(r0 + r) * kHidden + o] = v / (1.0f + expf(-v));
      if (pre_out != nullptr) pre_out[static_cast<long long>(r0 + r) * kHidden + o] = v;
    }
  }
}

__global__ void __launch_bounds__(kTeWarps * 32)
timestep_out_kernel(const float* __restrict__ hid, int n, const float* __restrict__ w2, const float* __restrict__ b2,
                    float* __restrict__ c_out, float* __restrict__ silu_out) {
  __shared__ float h[kTeRows][kHidden];
  const int r0 = blockIdx.y * kTeRows;
  const int nr = min(kTeRows, n - r0);
  for (int i = threadIdx.x; i < kTeRows * kHidden; i += kTeWarps * 32) {
    const int r = i / kHidden;
    h[r][i - r * kHidden] = (r < nr) ? hid[static_cast<long long>(r0) * kHidden + i] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * kTeWarps + warp;
  if (o >= kHidden) return;
  float acc[kTeRows];
#pragma unroll
  for (int r = 0; r < kTeRows; ++r) acc[r] = 0.f;
#pragma unroll 4
  for (int j = 0; j < kHidden / 32; ++j) {
    const int k = lane + 32 * j;
    const float w = __ldg(w2 + o * kHidden + k);
#pragma unroll
    for (int r = 0; r < kTeRows; ++r) acc[r] = fmaf(w, h[r][k], acc[r]);
  }
#pragma unroll
  for (int r = 0; r < kTeRows; ++r) acc[r] = warp_sum(acc[r]);
  if (lane == 0) {
    const float b = __ldg(b2 + o);
    for (int r = 0; r < nr; ++r) {
      const float v = acc[r] + b;
      c_out[static_cast<long long>(r0 + r) * kHidden + o] = v;
      silu_out[static_cast<long long>(r0 + r) * kHidden + o] = v / (1.0f + expf(-v));
    }
  }
}

// `hid` ([n, 768] fp32, caller-provided scratch distinct from c and silu_c) carries the hidden activations between the two
// phases: several CTAs (blockIdx.x) of one row group read all 768 hidden features while others already write their outputs,
// so the scratch must not alias either output.
int launch_timestep_embed(const long long* t, int n, const int* step_ptr, const int* map, const float* w0, const float* b0,
                          const float* w2, const float* b2, float* c, float* silu_c, float* hid, float* feat_out, float* pre_out,
                          cudaStream_t stream, int step_stride) {
  if (n <= 0) return kOk;
  if (t == nullptr && step_ptr == nullptr) return set_error(kErrBadArg, "timestep_embed: need t or step_ptr");
  if (hid == nullptr || hid == c || hid == silu_c) return set_error(kErrBadArg, "timestep_embed: hidden scratch must be a buffer of its own");
  dim3 grid(kHidden / kTeWarps, (n + kTeRows - 1) / kTeRows);
  timestep_hidden_kernel<<<grid, kTeWarps * 32, 0, stream>>>(t, n, step_ptr, step_stride, map, w0, b0, hid, feat_out, pre_out);
  int rc = check_launch("timestep_hidden_kernel");
  if (rc != kOk) return rc;
  timestep_out_kernel<<<grid, kTeWarps * 32, 0, stream>>>(hid, n, w2, b2, c, silu_c);
  return check_launch("timestep_out_kernel");
}

// ---------------------------------------------------------------------------------------------- adaLN modulation GEMV
// out[r, n] = bias[n] + sum_k silu_c[r, k] * W[n, k]   for r < rows <= kGvRows; W bf16 [n_out, 768] streamed once.
// Persistent grid: each warp walks output features with a grid stride; per feature 3 x 16-byte loads per lane cover the
// 768-wide weight row (1536 B, coalesced).  Algorithmic traffic: n_out * 1536 B of weights (87 MB for JPDVT).
constexpr int kGvRows = 8;
constexpr int kGvWarps = 8;

template <int ROWS>
__global__ void __launch_bounds__(kGvWarps * 32)
adaln_gemv_kernel(const float* __restrict__ silu_c, const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
                  float* __restrict__ out, int n_out) {
  __shared__ float sc[ROWS][kHidden];
  for (int i = threadIdx.x; i < ROWS * kHidden; i += kGvWarps * 32) sc[i / kHidden][i % kHidden] = silu_c[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * kGvWarps;
  for (int n = blockIdx.x * kGvWarps + (threadIdx.x >> 5); n < n_out; n += stride) {
    float acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.f;
    const uint4* wr = reinterpret_cast<const uint4*>(w + static_cast<long long>(n) * kHidden);
    uint4 u[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) u[j] = __ldcs(wr + lane + 32 * j);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int k0 = (lane + 32 * j) * 8;
      const uint32_t words[4] = {u[j].x, u[j].y, u[j].z, u[j].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float lo = __uint_as_float(words[e] << 16);
        const float hi = __uint_as_float(words[e] & 0xffff0000u);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          acc[r] = fmaf(lo, sc[r][k0 + 2 * e], acc[r]);
          acc[r] = fmaf(hi, sc[r][k0 + 2 * e + 1], acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
      const float b = __ldg(bias + n);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) out[static_cast<long long>(r) * n_out + n] = acc[r] + b;
    }
  }
}

int launch_adaln_gemv(const float* silu_c, int rows, const __nv_bfloat16* w, const float* bias, float* out, int n_out,
                      cudaStream_t stream) {
  if (rows <= 0 || n_out <= 0) return kOk;
  int blocks = (n_out + kGvWarps - 1) / kGvWarps;
  if (blocks > 148 * 8) blocks = 148 * 8;
  for (int r0 = 0; r0 < rows;) {
    const int left = rows - r0;
    const float* s = silu_c + static_cast<long long>(r0) * kHidden;
    float* o = out + static_cast<long long>(r0) * n_out;
    int used;
    if (left >= 8) { adaln_gemv_kernel<8><<<blocks, kGvWarps * 32, 0, stream>>>(s, w, bias, o, n_out); used = 8; }
    else if (left >= 4) { adaln_gemv_kernel<4><<<blocks, kGvWarps * 32, 0, stream>>>(s, w, bias, o, n_out); used = 4; }
    else if (left >= 2) { adaln_gemv_kernel<2><<<blocks, kGvWarps * 32, 0, stream>>>(s, w, bias, o, n_out); used = 2; }
    else { adaln_gemv_kernel<1><<<blocks, kGvWarps * 32, 0, stream>>>(s, w, bias, o, n_out); used = 1; }
    int rc = check_launch("adaln_gemv_kernel");
    if (rc != kOk) return rc;
    r0 += used;
  }
  return kOk;
}

// ---------------------------------------------------------------------------------------------- diffusion elementwise
// Counter-based normals for the per-step noise of p_sample (gaussian_diffusion.py:424 `th.randn_like(x)`): Philox4x32-10
// keyed by (seed), counter = (element group, diffusion step, call counter) -> 4 uniforms -> 2 Box-Muller pairs, so a
// sampling step needs no noise buffer and no generator launch.  philox_normal_kernel fills a buffer with exactly the same
// numbers (tests compare the two bit for bit; the raw 32-bit stream is checked against a numpy Philox in tests/).
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}
__device__ __forceinline__ float4 philox_normal4(long long group, int step, const long long* __restrict__ key) {
  const unsigned long long seed = static_cast<unsigned long long>(key[0]), call = static_cast<unsigned long long>(key[1]);
  const uint4 r = philox4x32_10(make_uint4(static_cast<unsigned>(group), static_cast<unsigned>(static_cast<unsigned long long>(group) >> 32),
                                           static_cast<unsigned>(step), static_cast<unsigned>(call)),
                                make_uint2(static_cast<unsigned>(seed), static_cast<unsigned>(seed >> 32)));
  // uniforms strictly inside (0, 1): 24 random bits + half an ulp
  const float u0 = static_cast<float>(r.x >> 8) * 5.9604645e-8f + 2.9802322e-8f;
  const float u1 = static_cast<float>(r.y >> 8) * 5.9604645e-8f + 2.9802322e-8f;
  const float u2 = static_cast<float>(r.z >> 8) * 5.9604645e-8f + 2.9802322e-8f;
  const float u3 = static_cast<float>(r.w >> 8) * 5.9604645e-8f + 2.9802322e-8f;
  const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
  float sa, ca, sb, cb;
  sincosf(6.2831853071795865f * u1, &sa, &ca);
  sincosf(6.2831853071795865f * u3, &sb, &cb);
  return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

__global__ void philox_normal_kernel(float* __restrict__ out, long long n4, int step, const long long* __restrict__ key,
                                     unsigned* __restrict__ raw) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  if (out != nullptr) reinterpret_cast<float4*>(out)[i] = philox_normal4(i, step, key);
  if (raw != nullptr) {
    const unsigned long long seed = static_cast<unsigned long long>(key[0]), call = static_cast<unsigned long long>(key[1]);
    reinterpret_cast<uint4*>(raw)[i] = philox4x32_10(
        make_uint4(static_cast<unsigned>(i), static_cast<unsigned>(static_cast<unsigned long long>(i) >> 32), static_cast<unsigned>(step),
                   static_cast<unsigned>(call)), make_uint2(static_cast<unsigned>(seed), static_cast<unsigned>(seed >> 32)));
  }
}

int launch_philox_normal(float* out, unsigned* raw, long long n, int step, const long long* key, cudaStream_t stream) {
  if (n <= 0) return kOk;
  if (n & 3) return set_error(kErrBadArg, "philox_normal: element count must be a multiple of 4");
  if (key == nullptr || (out == nullptr && raw == nullptr)) return set_error(kErrBadArg, "philox_normal: null pointer");
  const long long n4 = n / 4;
  philox_normal_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(out, n4, step, key, raw);
  return check_launch("philox_normal_kernel");
}

// mean = coef1[t] * x0 + coef2[t] * x_t ; sample = mean + [t != 0] * exp(0.5 * logvar[t]) * noise
// t is per-sample (`t` != null) or one device-resident step index for the whole batch (`step_ptr`).
// noise == null: drawn in place from Philox (noise_key = {seed, call counter} on the device, noise_step = loop position).
__global__ void posterior_kernel(const float* __restrict__ x0, const float* __restrict__ xt, const float* __restrict__ noise,
                                 const float* __restrict__ coef1, const float* __restrict__ coef2,
                                 const float* __restrict__ logvar, const long long* __restrict__ t,
                                 const int* __restrict__ step_ptr, float* __restrict__ mean, float* __restrict__ sample,
                                 long long n4, long long per_sample4, const long long* __restrict__ noise_key, int noise_step) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const long long ti = (t != nullptr) ? t[i / per_sample4] : static_cast<long long>(*step_ptr);
  const float c1 = __ldg(coef1 + ti), c2 = __ldg(coef2 + ti);
  const float sigma = (ti != 0) ? expf(0.5f * __ldg(logvar + ti)) : 0.f;
  const float4 a = reinterpret_cast<const float4*>(x0)[i];
  const float4 b = reinterpret_cast<const float4*>(xt)[i];
  float4 m;
  // same association as the reference: (c1 * x0) + (c2 * x_t), no fused multiply-add across the sum
  m.x = __fadd_rn(__fmul_rn(c1, a.x), __fmul_rn(c2, b.x));
  m.y = __fadd_rn(__fmul_rn(c1, a.y), __fmul_rn(c2, b.y));
  m.z = __fadd_rn(__fmul_rn(c1, a.z), __fmul_rn(c2, b.z));
  m.w = __fadd_rn(__fmul_rn(c1, a.w), __fmul_rn(c2, b.w));
  if (mean != nullptr) reinterpret_cast<float4*>(mean)[i] = m;
  if (sample != nullptr) {
    float4 s = m;
    if (ti != 0) {
      const float4 e = (noise != nullptr) ? reinterpret_cast<const float4*>(noise)[i] : philox_normal4(i, noise_step, noise_key);
      s.x = __fadd_rn(m.x, __fmul_rn(sigma, e.x));
      s.y = __fadd_rn(m.y, __fmul_rn(sigma, e.y));
      s.z = __fadd_rn(m.z, __fmul_rn(sigma, e.z));
      s.w = __fadd_rn(m.w, __fmul_rn(sigma, e.w));
    }
    reinterpret_cast<float4*>(sample)[i] = s;
  }
}

int launch_posterior(const float* x0, const float* xt, const float* noise, const float* coef1, const float* coef2,
                     const float* logvar, const long long* t, const int* step_ptr, float* mean, float* sample, long long n,
                     long long per_sample, cudaStream_t stream, const long long* noise_key, int noise_step) {
  if (n <= 0) return kOk;
  if ((n & 3) || (per_sample & 3)) return set_error(kErrBadArg, "posterior: element counts must be multiples of 4");
  if (t == nullptr && step_ptr == nullptr) return set_error(kErrBadArg, "posterior: need t or step_ptr");
  if (sample != nullptr && noise == nullptr && noise_key == nullptr) return set_error(kErrBadArg, "posterior: sample requested without noise or a Philox key");
  const long long n4 = n / 4;
  posterior_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(x0, xt, noise, coef1, coef2, logvar, t,
                                                                                step_ptr, mean, sample, n4, per_sample / 4,
                                                                                noise_key, noise_step);
  return check_launch("posterior_kernel");
}

// out = sqrt_ac[t] * x0 + sqrt_1mac[t] * noise ; optionally out = out * (1 - keep) + keep * x0 (keep == 1 -> clean pixel)
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                const float* __restrict__ sqrt_ac, const float* __restrict__ sqrt_1mac,
                                const long long* __restrict__ t, const float* __restrict__ keep, float* __restrict__ out,
                                long long n4, long long per_sample4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const long long ti = t[i / per_sample4];
  const float a = __ldg(sqrt_ac + ti), b = __ldg(sqrt_1mac + ti);
  const float4 x = reinterpret_cast<const float4*>(x0)[i];
  const float4 e = reinterpret_cast<const float4*>(noise)[i];
  float4 o;
  o.x = __fadd_rn(__fmul_rn(a, x.x), __fmul_rn(b, e.x));
  o.y = __fadd_rn(__fmul_rn(a, x.y), __fmul_rn(b, e.y));
  o.z = __fadd_rn(__fmul_rn(a, x.z), __fmul_rn(b, e.z));
  o.w = __fadd_rn(__fmul_rn(a, x.w), __fmul_rn(b, e.w));
  if (keep != nullptr) {
    const float4 k = reinterpret_cast<const float4*>(keep)[i];
    o.x = __fadd_rn(__fmul_rn(o.x, 1.0f - k.x), __fmul_rn(k.x, x.x));
    o.y = __fadd_rn(__fmul_rn(o.y, 1.0f - k.y), __fmul_rn(k.y, x.y));
    o.z = __fadd_rn(__fmul_rn(o.z, 1.0f - k.z), __fmul_rn(k.z, x.z));
    o.w = __fadd_rn(__fmul_rn(o.w, 1.0f - k.w), __fmul_rn(k.w, x.w));
  }
  reinterpret_cast<float4*>(out)[i] = o;
}

int launch_q_sample(const float* x0, const float* noise, const float* sqrt_ac, const float* sqrt_1mac, const long long* t,
                    const float* keep_mask, float* out, long long n, long long per_sample, cudaStream_t stream) {
  if (n <= 0) return kOk;
  if ((n & 3) || (per_sample & 3)) return set_error(kErrBadArg, "q_sample: element counts must be multiples of 4");
  const long long n4 = n / 4;
  q_sample_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(x0, noise, sqrt_ac, sqrt_1mac, t, keep_mask,
                                                                               out, n4, per_sample / 4);
  return check_launch("q_sample_kernel");
}

// DDIM update (diffusion/gaussian_diffusion.py:559-578): eps = (recip[t]*x_t - x0) / recipm1[t];
// sample = sqrt_abp[t]*x0 + dir[t]*eps + [t != 0]*sigma[t]*noise, with dir = sqrt(1 - abar_prev - sigma^2).
__global__ void ddim_kernel(const float* __restrict__ x0, const float* __restrict__ xt, const float* __restrict__ noise,
                            const float* __restrict__ recip, const float* __restrict__ recipm1,
                            const float* __restrict__ sqrt_abp, const float* __restrict__ dir,
                            const float* __restrict__ sigma, const long long* __restrict__ t,
                            const int* __restrict__ step_ptr, float* __restrict__ sample, long long n4,
                            long long per_sample4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const long long ti = (t != nullptr) ? t[i / per_sample4] : static_cast<long long>(*step_ptr);
  const float r = __ldg(recip + ti), rm1 = __ldg(recipm1 + ti), sa = __ldg(sqrt_abp + ti), dr = __ldg(dir + ti);
  const float sg = (ti != 0) ? __ldg(sigma + ti) : 0.f;
  const float4 a = reinterpret_cast<const float4*>(x0)[i];
  const float4 b = reinterpret_cast<const float4*>(xt)[i];
  float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
  if (sg != 0.f) e = reinterpret_cast<const float4*>(noise)[i];
  float4 o;
  o.x = a.x * sa + dr * ((r * b.x - a.x) / rm1) + sg * e.x;
  o.y = a.y * sa + dr * ((r * b.y - a.y) / rm1) + sg * e.y;
  o.z = a.z * sa + dr * ((r * b.z - a.z) / rm1) + sg * e.z;
  o.w = a.w * sa + dr * ((r * b.w - a.w) / rm1) + sg * e.w;
  reinterpret_cast<float4*>(sample)[i] = o;
}

int launch_ddim(const float* x0, const float* xt, const float* noise, const float* recip, const float* recipm1,
                const float* sqrt_abp, const float* dir, const float* sigma, const long long* t, const int* step_ptr,
                float* sample, long long n, long long per_sample, cudaStream_t stream) {
  if (n <= 0) return kOk;
  if ((n & 3) || (per_sample & 3)) return set_error(kErrBadArg, "ddim: element counts must be multiples of 4");
  if (t == nullptr && step_ptr == nullptr) return set_error(kErrBadArg, "ddim: need t or step_ptr");
  const long long n4 = n / 4;
  ddim_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(x0, xt, noise, recip, recipm1, sqrt_abp, dir, sigma,
                                                                           t, step_ptr, sample, n4, per_sample / 4);
  return check_launch("ddim_kernel");
}

}  // namespace jp
