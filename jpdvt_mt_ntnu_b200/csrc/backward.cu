// Memory-bound backward kernels of the JPDVT denoiser (training step, image_model/train_JPDVT.py:357-370 ->
// loss.backward() through image_model/models.py:273-293).  The tensor-core parts of the backward pass (dgrad / wgrad
// GEMMs) live in gemm.cu, the attention backward in attention.cu.
//
//   gate_bwd        : x_out = x_in + gate[b] * y           ->  dy = gate[b] * dx (bf16), dgate[b] += sum_t dx * y,
//                                                               dbias += sum_rows dy                (models.py:120-121)
//   ln_modulate_bwd : xn = LN(x) * (1 + scale[b]) + shift[b] ->  dx += LN'(.), dshift[b] += sum_t dxn,
//                                                               dscale[b] += sum_t dxn * xhat      (models.py:19-20)
//   colsum          : bias gradients  db[c] = sum_rows dY[row, c]
//   head_bwd        : te = W2 silu(pre) + b2                 ->  dpre (bf16), dW2, db2, db1          (models.py:288-290)
//   small helpers   : fp32 -> bf16 casts, silu', time_emb_in weight gradient, unpatchify transpose
#include "common.cuh"
#include "ptx.cuh"

namespace jp {

int launch_colsum_f32(const float* src, long long ld, long long rows, int cols, float* out, cudaStream_t stream);

__device__ __forceinline__ float warp_sum_b(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------- per-sample partial sums
// part: [batch][slices][768] -> out[b][c] (+)= sum_slices part[b][s][c]; one block per sample, float4 per thread
__global__ void __launch_bounds__(kHidden / 4)
sum_parts_kernel(const float* __restrict__ part0, float* __restrict__ out0, const float* __restrict__ part1,
                 float* __restrict__ out1, int slices, long long out_stride) {
  const int b = blockIdx.x, c4 = threadIdx.x;
  const float* part = blockIdx.y == 0 ? part0 : part1;
  float* out = blockIdx.y == 0 ? out0 : out1;
  const float4* src = reinterpret_cast<const float4*>(part + static_cast<long long>(b) * slices * kHidden) + c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < slices; ++s) {
    const float4 v = src[static_cast<long long>(s) * (kHidden / 4)];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  float4* dst = reinterpret_cast<float4*>(out + b * out_stride) + c4;
  float4 o = *dst;
  o.x += acc.x; o.y += acc.y; o.z += acc.z; o.w += acc.w;
  *dst = o;
}

// ---------------------------------------------------------------------------------------------- gate backward
// grid (B, slices); 192 threads, each owns 4 consecutive columns of the 768-wide row and walks a slice of the sample's
// tokens.  Per-(sample, slice) partial sums are written to scratch (no atomics) and folded by sum_parts / colsum.
constexpr int kGateThreads = kHidden / 4;

__global__ void __launch_bounds__(kGateThreads)
gate_bwd_kernel(const float* __restrict__ dx, const __nv_bfloat16* __restrict__ y, const float* __restrict__ gate,
                long long gate_stride, __nv_bfloat16* __restrict__ dy, float* __restrict__ part_gate,
                float* __restrict__ part_bias, int tokens, int rows_per_block) {
  const int b = blockIdx.x;
  const int t0 = blockIdx.y * rows_per_block;
  const int t1 = min(tokens, t0 + rows_per_block);
  const int c4 = threadIdx.x;
  const float4 g = __ldg(reinterpret_cast<const float4*>(gate + b * gate_stride) + c4);
  float4 sg = make_float4(0.f, 0.f, 0.f, 0.f), sb = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = t0; t < t1; ++t) {
    const long long row = static_cast<long long>(b) * tokens + t;
    const float4 d = __ldcs(reinterpret_cast<const float4*>(dx + row * kHidden) + c4);
    const uint2 yv = __ldcs(reinterpret_cast<const uint2*>(y + row * kHidden) + c4);
    const float y0 = __uint_as_float(yv.x << 16), y1 = __uint_as_float(yv.x & 0xffff0000u);
    const float y2 = __uint_as_float(yv.y << 16), y3 = __uint_as_float(yv.y & 0xffff0000u);
    sg.x = fmaf(d.x, y0, sg.x); sg.y = fmaf(d.y, y1, sg.y); sg.z = fmaf(d.z, y2, sg.z); sg.w = fmaf(d.w, y3, sg.w);
    const float o0 = g.x * d.x, o1 = g.y * d.y, o2 = g.z * d.z, o3 = g.w * d.w;
    sb.x += o0; sb.y += o1; sb.z += o2; sb.w += o3;
    uint2 o;
    o.x = pack_bf16(o0, o1); o.y = pack_bf16(o2, o3);
    reinterpret_cast<uint2*>(dy + row * kHidden)[c4] = o;
  }
  const long long slot = (static_cast<long long>(b) * gridDim.y + blockIdx.y) * (kHidden / 4) + c4;
  reinterpret_cast<float4*>(part_gate)[slot] = sg;
  reinterpret_cast<float4*>(part_bias)[slot] = sb;
}

// part: scratch of 2 * batch * gate_bwd_slices(batch, tokens) * 768 floats
int gate_bwd_slices(int batch, int tokens) {
  int splits = (8 * 148 + batch - 1) / batch;
  if (splits > tokens) splits = tokens;
  if (splits < 1) splits = 1;
  const int rows = (tokens + splits - 1) / splits;
  return (tokens + rows - 1) / rows;
}

int launch_gate_bwd(const float* dx, const __nv_bfloat16* y, const float* gate, long long gate_stride, __nv_bfloat16* dy,
                    float* dgate, long long dgate_stride, float* dbias, float* part, int batch, int tokens,
                    cudaStream_t stream) {
  if (batch <= 0 || tokens <= 0) return kOk;
  if (part == nullptr) return set_error(kErrBadArg, "gate_bwd: scratch buffer required");
  const int slices = gate_bwd_slices(batch, tokens);
  const int rows = (tokens + slices - 1) / slices;
  float* part_gate = part;
  float* part_bias = part + static_cast<long long>(batch) * slices * kHidden;
  dim3 grid(batch, slices);
  gate_bwd_kernel<<<grid, kGateThreads, 0, stream>>>(dx, y, gate, gate_stride, dy, part_gate, part_bias, tokens, rows);
  int rc = check_launch("gate_bwd_kernel");
  if (rc != kOk) return rc;
  sum_parts_kernel<<<dim3(batch, 1), kHidden / 4, 0, stream>>>(part_gate, dgate, nullptr, nullptr, slices, dgate_stride);
  rc = check_launch("sum_parts_kernel");
  if (rc != kOk) return rc;
  if (dbias != nullptr) return launch_colsum_f32(part_bias, kHidden, static_cast<long long>(batch) * slices, kHidden, dbias, stream);
  return kOk;
}

// ---------------------------------------------------------------------------------------------- LN + modulate backward
// One warp per (sample, chunk of kLnbRows tokens): rows are processed one after the other so the per-sample sums for
// dshift / dscale stay in registers (24 columns per lane) and hit global memory once per warp.
constexpr int kLnbWarps = 8;
constexpr int kLnbRows = 4;

__global__ void __launch_bounds__(kLnbWarps * 32, 2)
ln_modulate_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dxn, const float* __restrict__ scale,
                       long long mod_stride, float* __restrict__ dx, int accumulate, float* __restrict__ part_shift,
                       float* __restrict__ part_scale, __nv_bfloat16* __restrict__ dx_bf16, int batch, int tokens) {
  const int chunks = (tokens + kLnbRows - 1) / kLnbRows;
  const long long wid = static_cast<long long>(blockIdx.x) * kLnbWarps + (threadIdx.x >> 5);
  if (wid >= static_cast<long long>(batch) * chunks) return;
  const int b = static_cast<int>(wid / chunks), ch = static_cast<int>(wid % chunks);
  const int lane = threadIdx.x & 31;
  const int t0 = ch * kLnbRows, t1 = min(tokens, t0 + kLnbRows);
  float4 sc[6], ssh[6], ssc[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    sc[j] = __ldg(reinterpret_cast<const float4*>(scale + b * mod_stride) + lane + 32 * j);
    sc[j].x += 1.0f; sc[j].y += 1.0f; sc[j].z += 1.0f; sc[j].w += 1.0f;
    ssh[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    ssc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int t = t0; t < t1; ++t) {
    const long long row = static_cast<long long>(b) * tokens + t;
    const float4* xr = reinterpret_cast<const float4*>(x + row * kHidden);
    const float4* gr = reinterpret_cast<const float4*>(dxn + row * kHidden);
    float4 v[6], g[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { v[j] = __ldcs(xr + lane + 32 * j); g[j] = __ldcs(gr + lane + 32 * j); }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum_b(s) * (1.0f / kHidden);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
      q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    const float rstd = rsqrtf(warp_sum_b(q) * (1.0f / kHidden) + 1e-6f);
    // xhat = v * rstd;  a = dxn * (1 + scale);  dx_ln = rstd * (a - mean(a) - xhat * mean(a * xhat))
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      v[j].x *= rstd; v[j].y *= rstd; v[j].z *= rstd; v[j].w *= rstd;
      ssh[j].x += g[j].x; ssh[j].y += g[j].y; ssh[j].z += g[j].z; ssh[j].w += g[j].w;
      ssc[j].x = fmaf(g[j].x, v[j].x, ssc[j].x); ssc[j].y = fmaf(g[j].y, v[j].y, ssc[j].y);
      ssc[j].z = fmaf(g[j].z, v[j].z, ssc[j].z); ssc[j].w = fmaf(g[j].w, v[j].w, ssc[j].w);
      g[j].x *= sc[j].x; g[j].y *= sc[j].y; g[j].z *= sc[j].z; g[j].w *= sc[j].w;
      s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
      s2 += (g[j].x * v[j].x + g[j].y * v[j].y) + (g[j].z * v[j].z + g[j].w * v[j].w);
    }
    const float c1 = warp_sum_b(s1) * (1.0f / kHidden), c2 = warp_sum_b(s2) * (1.0f / kHidden);
    float4* dr = reinterpret_cast<float4*>(dx + row * kHidden);
    uint2* db = dx_bf16 != nullptr ? reinterpret_cast<uint2*>(dx_bf16 + row * kHidden) : nullptr;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      float4 o;
      o.x = rstd * (g[j].x - c1 - v[j].x * c2); o.y = rstd * (g[j].y - c1 - v[j].y * c2);
      o.z = rstd * (g[j].z - c1 - v[j].z * c2); o.w = rstd * (g[j].w - c1 - v[j].w * c2);
      if (accumulate) {
        const float4 prev = dr[lane + 32 * j];
        o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
      }
      dr[lane + 32 * j] = o;
      if (db != nullptr) { uint2 u; u.x = pack_bf16(o.x, o.y); u.y = pack_bf16(o.z, o.w); db[lane + 32 * j] = u; }
    }
  }
  // per-(sample, chunk) partial sums to scratch (no atomics); folded by sum_parts_kernel
  float4* psh = reinterpret_cast<float4*>(part_shift + wid * kHidden);
  float4* psc = reinterpret_cast<float4*>(part_scale + wid * kHidden);
#pragma unroll
  for (int j = 0; j < 6; ++j) { psh[lane + 32 * j] = ssh[j]; psc[lane + 32 * j] = ssc[j]; }
}

// part: scratch of 2 * batch * ceil(tokens / kLnbRows) * 768 floats
int launch_ln_modulate_bwd(const float* x, const float* dxn, const float* scale, long long mod_stride, float* dx,
                           int accumulate, float* dshift, float* dscale, long long dmod_stride, __nv_bfloat16* dx_bf16,
                           float* part, int batch, int tokens, cudaStream_t stream) {
  if (batch <= 0 || tokens <= 0) return kOk;
  if (part == nullptr) return set_error(kErrBadArg, "ln_modulate_bwd: scratch buffer required");
  const int chunks = (tokens + kLnbRows - 1) / kLnbRows;
  const long long warps = static_cast<long long>(batch) * chunks;
  float* part_shift = part;
  float* part_scale = part + warps * kHidden;
  ln_modulate_bwd_kernel<<<static_cast<unsigned>((warps + kLnbWarps - 1) / kLnbWarps), kLnbWarps * 32, 0, stream>>>(
      x, dxn, scale, mod_stride, dx, accumulate, part_shift, part_scale, dx_bf16, batch, tokens);
  int rc = check_launch("ln_modulate_bwd_kernel");
  if (rc != kOk) return rc;
  sum_parts_kernel<<<dim3(batch, 2), kHidden / 4, 0, stream>>>(part_shift, dshift, part_scale, dscale, chunks, dmod_stride);
  return check_launch("sum_parts_kernel");
}

long long bwd_part_floats(int batch, int tokens) {
  const long long a = 2LL * batch * ((tokens + kLnbRows - 1) / kLnbRows) * kHidden;
  const long long b = 2LL * batch * gate_bwd_slices(batch, tokens) * kHidden;
  return a > b ? a : b;
}

// ---------------------------------------------------------------------------------------------- LN backward + the next gate backward
// In the backward pass every LayerNorm-modulate backward is followed by the gate backward of the residual branch below it
// (models.py:120-121 read upwards): the second kernel re-reads the dx rows the first has just written.  Fused here: the
// warp that finishes a row of dx also emits dy = gate[b] * dx (bf16) and folds the row into dgate[b] / dbias.  A CTA of
// eight warps owns a slice of ONE sample's tokens, so the per-sample sums (dshift, dscale, dgate) leave through a
// cross-warp reduction in shared memory and one atomicAdd per column and CTA - no partial buffers, no sum_parts /
// colsum launches (ten launches per block become two).
constexpr int kLgWarps = 4;
constexpr int kLgRowBytes = 3 * kHidden * 4 + kHidden * 2;      // x, dxn, previous dx (fp32) + y (bf16) of one token row: 10,752 B

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}

// Four warps per CTA, two CTAs per SM (eight warps with up to 255 registers each): a warp keeps its running column sums AND
// the whole row in registers, and the NEXT row of the warp is already on its way into a shared-memory buffer (cp.async, two
// buffers per warp) while the current one is processed, so no row waits on a memory latency.  History: v1 fetched the previous
// dx / y chunk by chunk inside the output loop (ncu long_scoreboard 73 %, 94 us per launch at M = 18,432); v2 requested every
// operand of a row up front but had to keep the sums in shared memory for lack of registers at 16 warps per SM (short_scoreboard
// + mio 49 %, 69 us).
template <bool GATE>
__global__ void __launch_bounds__(kLgWarps * 32, 2)
ln_gate_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dxn, const float* __restrict__ scale,
                   long long mod_stride, float* __restrict__ dx, int accumulate, float* __restrict__ dshift,
                   float* __restrict__ dscale, long long dmod_stride, __nv_bfloat16* __restrict__ dx_bf16,
                   const __nv_bfloat16* __restrict__ y, const float* __restrict__ gate, long long gate_stride,
                   __nv_bfloat16* __restrict__ dy, float* __restrict__ dgate, long long dgate_stride,
                   float* __restrict__ dbias, int tokens, int rows_per_cta) {
  extern __shared__ __align__(16) uint8_t lg_smem[];   // [warp][2] row buffers, then the sample's (1 + scale) and gate rows
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.y * rows_per_cta, t1 = min(tokens, t0 + rows_per_cta);
  float* row_sc = reinterpret_cast<float*>(lg_smem + 2 * kLgWarps * kLgRowBytes);
  float* row_gt = row_sc + kHidden;
  for (int c = threadIdx.x; c < kHidden; c += kLgWarps * 32) {
    row_sc[c] = __ldg(scale + b * mod_stride + c) + 1.0f;
    if constexpr (GATE) row_gt[c] = __ldg(gate + b * gate_stride + c);
  }
  const uint32_t buf0 = smem_u32(lg_smem + warp * 2 * kLgRowBytes);
  // lane's 16-byte chunks: fp32 arrays chunk lane + 32 j (j < 6), bf16 y chunk lane + 32 j (j < 3)
  auto prefetch = [&](int t, int which) {
    const long long row = static_cast<long long>(b) * tokens + t;
    const uint32_t dst = buf0 + which * kLgRowBytes + 16u * lane;
    const float4* xr = reinterpret_cast<const float4*>(x + row * kHidden) + lane;
    const float4* gr = reinterpret_cast<const float4*>(dxn + row * kHidden) + lane;
    const float4* pr = reinterpret_cast<const float4*>(dx + row * kHidden) + lane;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      cp_async16(dst + 512u * j, xr + 32 * j);
      cp_async16(dst + 3072u + 512u * j, gr + 32 * j);
      if (accumulate) cp_async16(dst + 6144u + 512u * j, pr + 32 * j);
    }
    if constexpr (GATE) {
      const uint4* yr = reinterpret_cast<const uint4*>(y + row * kHidden) + lane;
#pragma unroll
      for (int j = 0; j < 3; ++j) cp_async16(dst + 9216u + 512u * j, yr + 32 * j);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float4 ssh[6], ssc[6], sg[6], sb[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    ssh[j] = make_float4(0.f, 0.f, 0.f, 0.f); ssc[j] = ssh[j]; sg[j] = ssh[j]; sb[j] = ssh[j];
  }
  int t = t0 + warp, it = 0;
  if (t < t1) prefetch(t, 0);
  __syncthreads();                                            // row_sc / row_gt are in place
  const uint32_t sc_s = smem_u32(row_sc) + 16u * lane, gt_s = smem_u32(row_gt) + 16u * lane;
  for (; t < t1; t += kLgWarps, ++it) {
    const int which = it & 1;
    __syncwarp();                                             // every lane is done with the buffer the next row goes into (the y
    if (t + kLgWarps < t1) {                                  // halves are read by a different lane than the one that copied them)
      prefetch(t + kLgWarps, which ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();                                             // ... and every lane's copies of the current row have landed
    const long long row = static_cast<long long>(b) * tokens + t;
    const uint32_t src = buf0 + which * kLgRowBytes + 16u * lane;
    float4 v[6], g[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) { v[j] = lds_f4(src + 512u * j); g[j] = lds_f4(src + 3072u + 512u * j); }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    const float mean = warp_sum_b(s) * (1.0f / kHidden);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
      q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
    }
    const float rstd = rsqrtf(warp_sum_b(q) * (1.0f / kHidden) + 1e-6f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {        // same arithmetic, in the same order, as ln_modulate_bwd_kernel
      const float4 sc = lds_f4(sc_s + 512u * j);                                   // 1 + scale
      v[j].x *= rstd; v[j].y *= rstd; v[j].z *= rstd; v[j].w *= rstd;
      ssh[j].x += g[j].x; ssh[j].y += g[j].y; ssh[j].z += g[j].z; ssh[j].w += g[j].w;
      ssc[j].x = fmaf(g[j].x, v[j].x, ssc[j].x); ssc[j].y = fmaf(g[j].y, v[j].y, ssc[j].y);
      ssc[j].z = fmaf(g[j].z, v[j].z, ssc[j].z); ssc[j].w = fmaf(g[j].w, v[j].w, ssc[j].w);
      g[j].x *= sc.x; g[j].y *= sc.y; g[j].z *= sc.z; g[j].w *= sc.w;
      s1 += (g[j].x + g[j].y) + (g[j].z + g[j].w);
      s2 += (g[j].x * v[j].x + g[j].y * v[j].y) + (g[j].z * v[j].z + g[j].w * v[j].w);
    }
    const float c1 = warp_sum_b(s1) * (1.0f / kHidden), c2 = warp_sum_b(s2) * (1.0f / kHidden);
    float4* dr = reinterpret_cast<float4*>(dx + row * kHidden);
    uint2* db = dx_bf16 != nullptr ? reinterpret_cast<uint2*>(dx_bf16 + row * kHidden) : nullptr;
    uint2* dyr = GATE ? reinterpret_cast<uint2*>(dy + row * kHidden) : nullptr;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      float4 o;
      o.x = rstd * (g[j].x - c1 - v[j].x * c2); o.y = rstd * (g[j].y - c1 - v[j].y * c2);
      o.z = rstd * (g[j].z - c1 - v[j].z * c2); o.w = rstd * (g[j].w - c1 - v[j].w * c2);
      if (accumulate) {
        const float4 prev = lds_f4(src + 6144u + 512u * j);
        o.x += prev.x; o.y += prev.y; o.z += prev.z; o.w += prev.w;
      }
      dr[lane + 32 * j] = o;
      if (db != nullptr) { uint2 u; u.x = pack_bf16(o.x, o.y); u.y = pack_bf16(o.z, o.w); db[lane + 32 * j] = u; }
      if constexpr (GATE) {              // gate_bwd_kernel's arithmetic on the finished row
        const float4 gt = lds_f4(gt_s + 512u * j);
        // y: bf16 chunk (lane + 32 j') holds columns 8 (lane + 32 j') ..; this lane's fp32 columns 4 (lane + 32 j) .. +3 are the
        // low or high half of y chunk (lane + 32 j) / 2 -> read the 8-byte half directly
        uint2 yv;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(yv.x), "=r"(yv.y) : "r"(buf0 + which * kLgRowBytes + 9216u + 8u * (lane + 32 * j)));
        const float y0 = __uint_as_float(yv.x << 16), y1 = __uint_as_float(yv.x & 0xffff0000u);
        const float y2 = __uint_as_float(yv.y << 16), y3 = __uint_as_float(yv.y & 0xffff0000u);
        sg[j].x = fmaf(o.x, y0, sg[j].x); sg[j].y = fmaf(o.y, y1, sg[j].y); sg[j].z = fmaf(o.z, y2, sg[j].z); sg[j].w = fmaf(o.w, y3, sg[j].w);
        const float o0 = gt.x * o.x, o1 = gt.y * o.y, o2 = gt.z * o.z, o3 = gt.w * o.w;
        sb[j].x += o0; sb[j].y += o1; sb[j].z += o2; sb[j].w += o3;
        uint2 u; u.x = pack_bf16(o0, o1); u.y = pack_bf16(o2, o3);
        dyr[lane + 32 * j] = u;
      }
    }
  }
  // cross-warp sums through the (now idle) row buffers: [quantity][warp][768] floats
  __syncthreads();
  float* red = reinterpret_cast<float*>(lg_smem);
  constexpr int kQ = GATE ? 4 : 2;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    reinterpret_cast<float4*>(red + (0 * kLgWarps + warp) * kHidden)[lane + 32 * j] = ssh[j];
    reinterpret_cast<float4*>(red + (1 * kLgWarps + warp) * kHidden)[lane + 32 * j] = ssc[j];
    if constexpr (GATE) {
      reinterpret_cast<float4*>(red + (2 * kLgWarps + warp) * kHidden)[lane + 32 * j] = sg[j];
      reinterpret_cast<float4*>(red + (3 * kLgWarps + warp) * kHidden)[lane + 32 * j] = sb[j];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kHidden; c += kLgWarps * 32) {
    float a[kQ];
#pragma unroll
    for (int qq = 0; qq < kQ; ++qq) {
      a[qq] = 0.f;
#pragma unroll
      for (int w = 0; w < kLgWarps; ++w) a[qq] += red[(qq * kLgWarps + w) * kHidden + c];
    }
    atomicAdd(dshift + b * dmod_stride + c, a[0]);
    atomicAdd(dscale + b * dmod_stride + c, a[1]);
    if constexpr (GATE) {
      atomicAdd(dgate + b * dgate_stride + c, a[2]);
      if (dbias != nullptr) atomicAdd(dbias + c, a[3]);
    }
  }
}

// dshift / dscale / dgate / dbias are ACCUMULATED into (atomicAdd): the caller zeroes them once per backward pass.
int launch_ln_gate_bwd(const float* x, const float* dxn, const float* scale, long long mod_stride, float* dx, int accumulate,
                       float* dshift, float* dscale, long long dmod_stride, __nv_bfloat16* dx_bf16, const __nv_bfloat16* y,
                       const float* gate, long long gate_stride, __nv_bfloat16* dy, float* dgate, long long dgate_stride,
                       float* dbias, int batch, int tokens, cudaStream_t stream) {
  if (batch <= 0 || tokens <= 0) return kOk;
  const bool with_gate = y != nullptr;
  if (with_gate && (!gate || !dy || !dgate)) return set_error(kErrBadArg, "ln_gate_bwd: the gate half needs gate, dy and dgate");
  int slices = (2 * 148) / batch;                              // one wave of two CTAs per SM (rounding up would leave a short second wave)
  const int max_slices = (tokens + kLgWarps - 1) / kLgWarps;   // at least one row per warp
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  const int rows = (tokens + slices - 1) / slices;
  slices = (tokens + rows - 1) / rows;
  // row buffers (2 per warp; they also hold the 4 x kLgWarps x 768 floats of the closing reduction) + the scale / gate rows
  static_assert(2 * kLgWarps * kLgRowBytes >= 4 * kLgWarps * kHidden * 4, "the reduction reuses the row buffers");
  const size_t smem = static_cast<size_t>(2 * kLgWarps) * kLgRowBytes + 2 * kHidden * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(ln_gate_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess ||
        cudaFuncSetAttribute(ln_gate_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return set_error(kErrCuda, "ln_gate_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
    attr_set = true;
  }
  dim3 grid(batch, slices);
  if (with_gate)
    ln_gate_bwd_kernel<true><<<grid, kLgWarps * 32, smem, stream>>>(x, dxn, scale, mod_stride, dx, accumulate, dshift, dscale, dmod_stride,
                                                                    dx_bf16, y, gate, gate_stride, dy, dgate, dgate_stride, dbias, tokens, rows);
  else
    ln_gate_bwd_kernel<false><<<grid, kLgWarps * 32, smem, stream>>>(x, dxn, scale, mod_stride, dx, accumulate, dshift, dscale, dmod_stride,
                                                                     dx_bf16, nullptr, nullptr, 0, nullptr, nullptr, 0, nullptr, tokens, rows);
  return check_launch("ln_gate_bwd_kernel");
}

// ---------------------------------------------------------------------------------------------- column sums (bias grads)
template <typename T>
__device__ __forceinline__ float2 load2(const T* p);
template <>
__device__ __forceinline__ float2 load2<float>(const float* p) { return __ldcs(reinterpret_cast<const float2*>(p)); }
template <>
__device__ __forceinline__ float2 load2<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint32_t u = __ldcs(reinterpret_cast<const unsigned int*>(p));
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// grid (cols / 256, row slices); 128 threads x 2 columns
template <typename T>
__global__ void __launch_bounds__(128)
colsum_kernel(const T* __restrict__ src, long long ld, long long rows, int cols, float* __restrict__ out, int rows_per_block) {
  const int c = blockIdx.x * 256 + threadIdx.x * 2;
  if (c >= cols) return;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float2 acc = make_float2(0.f, 0.f);
  for (long long r = r0; r < r1; ++r) {
    const float2 v = load2<T>(src + r * ld + c);
    acc.x += v.x; acc.y += v.y;
  }
  atomicAdd(out + c, acc.x);
  atomicAdd(out + c + 1, acc.y);
}

template <typename T>
static int launch_colsum_t(const T* src, long long ld, long long rows, int cols, float* out, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return kOk;
  if (cols % 2) return set_error(kErrBadArg, "colsum: cols must be even");
  const int col_blocks = (cols + 255) / 256;
  long long slices = (8 * 148 + col_blocks - 1) / col_blocks;
  if (slices > rows) slices = rows;
  const int rpb = static_cast<int>((rows + slices - 1) / slices);
  dim3 grid(col_blocks, static_cast<unsigned>((rows + rpb - 1) / rpb));
  colsum_kernel<T><<<grid, 128, 0, stream>>>(src, ld, rows, cols, out, rpb);
  return check_launch("colsum_kernel");
}
int launch_colsum_bf16(const __nv_bfloat16* src, long long ld, long long rows, int cols, float* out, cudaStream_t stream) {
  return launch_colsum_t<__nv_bfloat16>(src, ld, rows, cols, out, stream);
}
int launch_colsum_f32(const float* src, long long ld, long long rows, int cols, float* out, cudaStream_t stream) {
  return launch_colsum_t<float>(src, ld, rows, cols, out, stream);
}

// ---------------------------------------------------------------------------------------------- position head backward
// te[m, :8] = W2 . silu(pre[m, :64]) + b2.  One thread per row; block-level partial sums in shared memory.
constexpr int kHeadThreads = 128;

__global__ void __launch_bounds__(kHeadThreads)
head_bwd_kernel(const float* __restrict__ dte, const float* __restrict__ pre, const float* __restrict__ w2,
                __nv_bfloat16* __restrict__ dpre, float* __restrict__ dw2, float* __restrict__ db2, float* __restrict__ db1,
                long long rows) {
  __shared__ float s_w2[kLatent * 64];
  __shared__ float s_dw2[kLatent * 64];
  __shared__ float s_db1[64];
  __shared__ float s_db2[kLatent];
  for (int i = threadIdx.x; i < kLatent * 64; i += kHeadThreads) { s_w2[i] = w2[i]; s_dw2[i] = 0.f; }
  if (threadIdx.x < 64) s_db1[threadIdx.x] = 0.f;
  if (threadIdx.x < kLatent) s_db2[threadIdx.x] = 0.f;
  __syncthreads();
  const long long m = static_cast<long long>(blockIdx.x) * kHeadThreads + threadIdx.x;
  const int lane = threadIdx.x & 31;
  float d[kLatent];
#pragma unroll
  for (int k = 0; k < kLatent; ++k) d[k] = 0.f;
  if (m < rows) {
    const float4 a = reinterpret_cast<const float4*>(dte + m * kLatent)[0], b = reinterpret_cast<const float4*>(dte + m * kLatent)[1];
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  }
#pragma unroll
  for (int k = 0; k < kLatent; ++k) {
    const float s = warp_sum_b(d[k]);
    if (lane == 0) atomicAdd(&s_db2[k], s);
  }
#pragma unroll 4
  for (int j = 0; j < 64; ++j) {
    float p = 0.f, sg = 0.f, ds = 0.f;
    if (m < rows) {
      p = pre[m * 64 + j];
      const float sig = 1.0f / (1.0f + __expf(-p));
      sg = p * sig;                                   // silu(pre)
      ds = sig * (1.0f + p * (1.0f - sig));           // silu'(pre)
    }
    float up = 0.f;
#pragma unroll
    for (int k = 0; k < kLatent; ++k) up = fmaf(d[k], s_w2[k * 64 + j], up);
    const float dp = up * ds;
    if (m < rows) dpre[m * 64 + j] = __float2bfloat16_rn(dp);
    const float sdp = warp_sum_b(dp);
    if (lane == 0) atomicAdd(&s_db1[j], sdp);
#pragma unroll
    for (int k = 0; k < kLatent; ++k) {
      const float v = warp_sum_b(d[k] * sg);
      if (lane == 0) atomicAdd(&s_dw2[k * 64 + j], v);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kLatent * 64; i += kHeadThreads) atomicAdd(dw2 + i, s_dw2[i]);
  if (threadIdx.x < 64) atomicAdd(db1 + threadIdx.x, s_db1[threadIdx.x]);
  if (threadIdx.x < kLatent) atomicAdd(db2 + threadIdx.x, s_db2[threadIdx.x]);
}

int launch_head_bwd(const float* dte, const float* pre, const float* w2, __nv_bfloat16* dpre, float* dw2, float* db2,
                    float* db1, long long rows, cudaStream_t stream) {
  if (rows <= 0) return kOk;
  head_bwd_kernel<<<static_cast<unsigned>((rows + kHeadThreads - 1) / kHeadThreads), kHeadThreads, 0, stream>>>(
      dte, pre, w2, dpre, dw2, db2, db1, rows);
  return check_launch("head_bwd_kernel");
}

// ---------------------------------------------------------------------------------------------- small helpers
__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(in)[i];
  uint2 u;
  u.x = pack_bf16(v.x, v.y); u.y = pack_bf16(v.z, v.w);
  reinterpret_cast<uint2*>(out)[i] = u;
}
int launch_cast_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t stream) {
  if (n <= 0) return kOk;
  if (n & 3) return set_error(kErrBadArg, "cast: element count must be a multiple of 4");
  cast_f32_bf16_kernel<<<static_cast<unsigned>((n / 4 + 255) / 256), 256, 0, stream>>>(in, out, n / 4);
  return check_launch("cast_f32_bf16_kernel");
}

// out = grad * silu'(pre)   (fp32 and an optional bf16 copy)
__global__ void silu_bwd_kernel(const float* __restrict__ grad, const float* __restrict__ pre, float* __restrict__ out,
                                __nv_bfloat16* __restrict__ out_bf16, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p = pre[i];
  const float sig = 1.0f / (1.0f + expf(-p));
  const float v = grad[i] * sig * (1.0f + p * (1.0f - sig));
  if (out != nullptr) out[i] = v;
  if (out_bf16 != nullptr) out_bf16[i] = __float2bfloat16_rn(v);
}
int launch_silu_bwd(const float* grad, const float* pre, float* out, __nv_bfloat16* out_bf16, long long n, cudaStream_t stream) {
  if (n <= 0) return kOk;
  silu_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(grad, pre, out, out_bf16, n);
  return check_launch("silu_bwd_kernel");
}

// dW_in[c, j] = sum_m dx0[m, c] * x_t[m, j]   (time_emb_in.weight [768, 8]);  grid (3, row slices), 256 threads = columns
__global__ void __launch_bounds__(256)
win_grad_kernel(const float* __restrict__ dx0, const float* __restrict__ xt, float* __restrict__ dw, long long rows,
                int rows_per_block) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const long long r0 = static_cast<long long>(blockIdx.y) * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float acc[kLatent];
#pragma unroll
  for (int j = 0; j < kLatent; ++j) acc[j] = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float d = __ldcs(dx0 + r * kHidden + c);
    const float4 a = __ldg(reinterpret_cast<const float4*>(xt + r * kLatent)), b = __ldg(reinterpret_cast<const float4*>(xt + r * kLatent) + 1);
    acc[0] = fmaf(d, a.x, acc[0]); acc[1] = fmaf(d, a.y, acc[1]); acc[2] = fmaf(d, a.z, acc[2]); acc[3] = fmaf(d, a.w, acc[3]);
    acc[4] = fmaf(d, b.x, acc[4]); acc[5] = fmaf(d, b.y, acc[5]); acc[6] = fmaf(d, b.z, acc[6]); acc[7] = fmaf(d, b.w, acc[7]);
  }
#pragma unroll
  for (int j = 0; j < kLatent; ++j) atomicAdd(dw + c * kLatent + j, acc[j]);
}
int launch_win_grad(const float* dx0, const float* xt, float* dw, long long rows, cudaStream_t stream) {
  if (rows <= 0) return kOk;
  long long slices = 2 * 148;
  if (slices > rows) slices = rows;
  const int rpb = static_cast<int>((rows + slices - 1) / slices);
  dim3 grid(kHidden / 256, static_cast<unsigned>((rows + rpb - 1) / rpb));
  win_grad_kernel<<<grid, 256, 0, stream>>>(dx0, xt, dw, rows, rpb);
  return check_launch("win_grad_kernel");
}

// transpose of unpatchify (models.py:227-240): dy[(b,h,w), (p*16+q)*3 + c] (+)= dimg[b, c, h*16+p, w*16+q]
__global__ void unpatchify_bwd_kernel(const float* __restrict__ dimg, float* __restrict__ dy, long long total, int size,
                                      int accumulate) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = size / 16;
  const int n = static_cast<int>(idx % kHidden);
  const long long m = idx / kHidden;
  const int w = static_cast<int>(m % g), h = static_cast<int>((m / g) % g);
  const long long b = m / (g * g);
  const int c = n % 3, q = (n / 3) % 16, p = n / 48;
  const float v = __ldg(dimg + ((b * 3 + c) * size + (h * 16 + p)) * static_cast<long long>(size) + w * 16 + q);
  dy[idx] = accumulate ? dy[idx] + v : v;
}
int launch_unpatchify_bwd(const float* dimg, float* dy, int batch, int size, int accumulate, cudaStream_t stream) {
  const long long g = size / 16;
  const long long total = static_cast<long long>(batch) * g * g * kHidden;
  if (total == 0) return kOk;
  unpatchify_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(dimg, dy, total, size, accumulate);
  return check_launch("unpatchify_bwd_kernel");
}

// h = gelu_tanh(pre) and, in place of pre, g' = gelu_tanh'(pre)  (bf16).  Training keeps g' so that the fc2 data-gradient
// GEMM epilogue is a plain multiply instead of two transcendentals per element.
__global__ void gelu_bf16_kernel(__nv_bfloat16* __restrict__ pre_to_grad, __nv_bfloat16* __restrict__ out, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 a = __ldcs(reinterpret_cast<const uint4*>(pre_to_grad) + i);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
  uint32_t ow[4], gw[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float x0 = __uint_as_float(aw[e] << 16), x1 = __uint_as_float(aw[e] & 0xffff0000u);
    ow[e] = pack_bf16(gelu_tanh(x0), gelu_tanh(x1));
    gw[e] = pack_bf16(gelu_tanh_grad(x0), gelu_tanh_grad(x1));
  }
  reinterpret_cast<uint4*>(out)[i] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  reinterpret_cast<uint4*>(pre_to_grad)[i] = make_uint4(gw[0], gw[1], gw[2], gw[3]);
}
int launch_gelu(__nv_bfloat16* pre_to_grad, __nv_bfloat16* out, long long n, cudaStream_t stream) {
  if (n <= 0) return kOk;
  if (n & 7) return set_error(kErrBadArg, "gelu: element count must be a multiple of 8");
  gelu_bf16_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, stream>>>(pre_to_grad, out, n / 8);
  return check_launch("gelu_bf16_kernel");
}

__global__ void silu_fwd_bf16_kernel(const float* __restrict__ pre, __nv_bfloat16* __restrict__ out, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p = pre[i];
  out[i] = __float2bfloat16_rn(p / (1.0f + expf(-p)));
}
int launch_silu_fwd_bf16(const float* pre, __nv_bfloat16* out, long long n, cudaStream_t stream) {
  if (n <= 0) return kOk;
  silu_fwd_bf16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(pre, out, n);
  return check_launch("silu_fwd_bf16_kernel");
}

}  // namespace jp
