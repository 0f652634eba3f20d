// Short-sequence fused multi-head attention for the JPDVT piece tokens (T <= ~1024, head_dim 64, no mask).
//
// Replaces timm Attention.forward's reshape/permute + F.scaled_dot_product_attention (called from
// image_model/models.py:108,120): softmax(Q K^T / 8) V per (sample, head), reading Q/K/V straight out of the fused
// QKV GEMM output [B*T, 2304] (row layout [q(12x64) | k(12x64) | v(12x64)]) and writing [B*T, 768] head-major columns,
// i.e. exactly the tensor the out-projection GEMM consumes.
//
// One CTA per (q-block, head, sample).  The whole K and V of the head live in shared memory (XOR-swizzled 128-byte
// rows, conflict-free ldmatrix); each warp owns 16 query rows, keeps its Q fragments in registers and runs an
// online (flash-style) softmax over 48-column KV chunks with warp-level quad reductions.  Tensor-core math is
// mma.sync m16n8k16 bf16 -> fp32.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace jp {

constexpr int kQkvLd = 3 * kHidden;   // 2304
constexpr int kChunkGroups = 3;       // KV chunk = 3 x 16 = 48 columns

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk `ch` (0..7) of row `row` inside a [rows][64] bf16 tile with XOR swizzle
__device__ __forceinline__ uint32_t swz(int row, int ch) { return static_cast<uint32_t>(row * 128 + ((ch ^ (row & 7)) << 4)); }

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One CTA per (sample, head): K and V of the head in shared memory, 3 or 4 warps that each walk their share of the
// 16-row query tiles (tile index = warp, warp + nw, ...).  Q tiles and the normalised output tile pass through a
// private 2 KB staging buffer per warp (cp.async in, 16-byte coalesced stores out).  128-thread CTAs keep the
// register file allocation exact (warp granularity 4) and let 5 CTAs share an SM at T = 144.
constexpr int kAttThreadsMax = 128;

__global__ void __launch_bounds__(kAttThreadsMax, 4)
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse2, int T,
                 int Tp) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int nthreads = blockDim.x;
  const int nw = nthreads >> 5;
  uint8_t* sK = att_smem;
  uint8_t* sV = sK + Tp * 128;
  uint8_t* sQ = sV + Tp * 128;                       // nw private [16][64] bf16 tiles
  const int b = blockIdx.y, h = blockIdx.x;
  const __nv_bfloat16* base = qkv + static_cast<long long>(b) * T * kQkvLd + h * kHeadDim;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* myQ = sQ + warp * 2048;

  // ---- stage K, V (rows >= T are zero filled) and this warp's first Q tile
  for (int i = threadIdx.x; i < Tp * 8; i += nthreads) {
    const int r = i >> 3, ch = i & 7;
    void* dk = sK + swz(r, ch);
    void* dv = sV + swz(r, ch);
    if (r < T) {
      const __nv_bfloat16* src = base + static_cast<long long>(r) * kQkvLd + ch * 8;
      cp_async16(dk, src + kHidden);
      cp_async16(dv, src + 2 * kHidden);
    } else {
      *reinterpret_cast<uint4*>(dk) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(dv) = make_uint4(0, 0, 0, 0);
    }
  }
  const int mt = Tp >> 4;
  auto load_q = [&](int tile) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 32 + lane;
      const int r = idx >> 3, ch = idx & 7;
      void* dst = myQ + swz(r, ch);
      const int row = tile * 16 + r;
      if (row < T) cp_async16(dst, base + static_cast<long long>(row) * kQkvLd + ch * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  };
  if (warp < mt) load_q(warp);
  cp_async_wait_all();
  __syncthreads();

  const int g = lane >> 2, tq = lane & 3;
  const int li = lane >> 3, lr = lane & 7;   // ldmatrix: matrix index / row inside the 8x8 matrix
  const uint32_t k_base = smem_u32(sK), v_base = smem_u32(sV), q_base = smem_u32(myQ);
  const float sl2 = 0.125f * 1.4426950408889634f;   // head_dim^-0.5 * log2(e)
  __nv_bfloat16* obase = out + static_cast<long long>(b) * T * kHidden + h * kHeadDim;

  for (int tile = warp; tile < mt; tile += nw) {
    // ---- Q fragments: 4 k-steps of 16 along d
    uint32_t qf[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) ldsm_x4(qf[ks], q_base + swz(lr + ((li & 1) ? 8 : 0), ks * 2 + (li >> 1)));

    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    for (int kv0 = 0; kv0 < Tp; kv0 += 16 * kChunkGroups) {
      float s[2 * kChunkGroups][4];
#pragma unroll
      for (int j = 0; j < 2 * kChunkGroups; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
      // ---- S = Q K^T for up to 48 kv columns; k-step outermost so consecutive MMAs hit independent accumulators
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
        for (int gi = 0; gi < kChunkGroups; ++gi) {
          const int kvg = kv0 + 16 * gi;
          if (kvg < Tp) {
            uint32_t kf[4];
            ldsm_x4(kf, k_base + swz(kvg + lr + ((li >> 1) ? 8 : 0), ks * 2 + (li & 1)));
            mma_bf16(s[2 * gi], qf[ks], kf[0], kf[1]);
            mma_bf16(s[2 * gi + 1], qf[ks], kf[2], kf[3]);
          }
        }
      }
      // ---- mask the padded tail (last chunk only), chunk row-max
      if (kv0 + 16 * kChunkGroups > T) {
#pragma unroll
        for (int j = 0; j < 2 * kChunkGroups; ++j) {
          const int col = kv0 + 8 * j + 2 * tq;
          if (col >= T) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
          if (col + 1 >= T) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 2 * kChunkGroups; ++j) {
        mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
        mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every processed chunk has >= 1 valid column
      const float a0 = ex2((m0 - mn0) * sl2), a1 = ex2((m1 - mn1) * sl2);
      m0 = mn0; m1 = mn1;
      const float ms0 = mn0 * sl2, ms1 = mn1 * sl2;
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t pf[kChunkGroups][4];
#pragma unroll
      for (int j = 0; j < 2 * kChunkGroups; ++j) {
        const float p0 = ex2(fmaf(s[j][0], sl2, -ms0));
        const float p1 = ex2(fmaf(s[j][1], sl2, -ms0));
        const float p2 = ex2(fmaf(s[j][2], sl2, -ms1));
        const float p3 = ex2(fmaf(s[j][3], sl2, -ms1));
        rs0 += p0 + p1; rs1 += p2 + p3;
        pf[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
        pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
      l0 = l0 * a0 + rs0; l1 = l1 * a1 + rs1;
#pragma unroll
      for (int j = 0; j < 8; ++j) { o[j][0] *= a0; o[j][1] *= a0; o[j][2] *= a1; o[j][3] *= a1; }
      // ---- O += P V
#pragma unroll
      for (int gi = 0; gi < kChunkGroups; ++gi) {
        const int kvg = kv0 + 16 * gi;
        if (kvg < Tp) {
#pragma unroll
          for (int dp = 0; dp < 4; ++dp) {   // pairs of 8-wide d tiles
            uint32_t vf[4];
            ldsm_x4_trans(vf, v_base + swz(kvg + lr + ((li & 1) ? 8 : 0), dp * 2 + (li >> 1)));
            mma_bf16(o[2 * dp], pf[gi], vf[0], vf[1]);
            mma_bf16(o[2 * dp + 1], pf[gi], vf[2], vf[3]);
          }
        }
      }
    }
    // ---- finalize: quad-reduce the row sums, normalise, stage through the warp's Q buffer, coalesced 16-byte stores
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    if (lse2 != nullptr && tq == 0) {   // training: log2-domain log-sum-exp of the scaled scores, [B, 12, T]
      float* lrow = lse2 + (static_cast<long long>(b) * kHeads + h) * T;
      if (tile * 16 + g < T) lrow[tile * 16 + g] = fmaf(m0, sl2, log2f(l0));
      if (tile * 16 + g + 8 < T) lrow[tile * 16 + g + 8] = fmaf(m1, sl2, log2f(l1));
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      *reinterpret_cast<uint32_t*>(myQ + swz(g, j) + tq * 4) = pack_bf16(o[j][0] * i0, o[j][1] * i0);
      *reinterpret_cast<uint32_t*>(myQ + swz(g + 8, j) + tq * 4) = pack_bf16(o[j][2] * i1, o[j][3] * i1);
    }
    __syncwarp();
    const int r0 = tile * 16;
    uint4 ov[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 32 + lane;
      ov[it] = *reinterpret_cast<const uint4*>(myQ + swz(idx >> 3, idx & 7));
    }
    __syncwarp();
    // next Q tile of this warp can stream in while the output tile drains
    if (tile + nw < mt) load_q(tile + nw);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = it * 32 + lane;
      const int r = idx >> 3, ch = idx & 7;
      if (r0 + r < T) *reinterpret_cast<uint4*>(obase + static_cast<long long>(r0 + r) * kHidden + ch * 8) = ov[it];
    }
    if (tile + nw < mt) {
      cp_async_wait_all();
      __syncwarp();
    }
  }
}

int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int batch, int tokens, cudaStream_t stream) {
  if (batch <= 0 || tokens <= 0) return kOk;
  if (attention_tc_supported(tokens)) {
    static int legacy = -1;            // JPDVT_ATTN_LEGACY=1 keeps the mma.sync kernel for every size (A/B timing knob)
    if (legacy < 0) { const char* e = getenv("JPDVT_ATTN_LEGACY"); legacy = (e != nullptr && e[0] == '1') ? 1 : 0; }
    if (!legacy) return launch_attention_tc(qkv, out, lse2, batch, tokens, stream);
  }
  if (batch > 65535) return set_error(kErrBadArg, "attention: batch %d exceeds gridDim.y limit", batch);
  const int mt = (tokens + 15) / 16;        // 16-row query tiles
  int nw = (mt % 3 == 0) ? 3 : 4;           // balanced tile counts per warp (144 -> 3x3, 256 -> 4x4, 324 -> 3x7)
  if (nw > mt) nw = mt;
  const int Tp = mt * 16;
  const size_t smem = static_cast<size_t>(2 * Tp) * 128 + static_cast<size_t>(nw) * 2048;
  if (smem > 227 * 1024) return set_error(kErrUnsupported, "attention: %d tokens need %zu B of shared memory", tokens, smem);
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return set_error(kErrCuda, "attention: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
    configured = smem;
    // ask for the full shared-memory carve-out so several CTAs (5 at T=144) are resident per SM
    cudaFuncSetAttribute(attention_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  dim3 grid(kHeads, batch);
  attention_kernel<<<grid, nw * 32, smem, stream>>>(qkv, out, lse2, tokens, Tp);
  return check_launch("attention_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// Attention backward (training): dQ, dK, dV for softmax(Q K^T / 8) V, recomputing the probabilities from the forward's
// log-sum-exp (flash-attention style).  One CTA per (sample, head); Q, K, V and dO of the head live in shared memory.
//   phase A: each warp owns 16-key tiles and walks all query tiles  -> dK, dV   (S^T = K Q^T, dP^T = V dO^T)
//   phase B: each warp owns 16-query tiles and walks all key groups  -> dQ       (S = Q K^T,  dP  = dO V^T)
// Zero padding of rows >= T makes every padded contribution vanish, so no masks are needed.
// qkv / dqkv: [B*T, 2304] bf16; o / d_o: [B*T, 768] bf16; lse2: [B, 12, T] fp32 (log2 domain, from the forward).
// accumulator tile (16 rows x 64 columns in mma.sync C layout) -> bf16 global rows; 4-byte stores, each 32-byte sector is
// completed by the two d-tiles that share it (no shared-memory staging: the backward kernel needs the space for residency)
__device__ __forceinline__ void store_tile_bf16(const float (&acc)[8][4], __nv_bfloat16* gdst, long long ld, int row0, int T,
                                                int lane) {
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (row0 + g < T)
      *reinterpret_cast<uint32_t*>(gdst + static_cast<long long>(row0 + g) * ld + 8 * j + 2 * tq) = pack_bf16(acc[j][0], acc[j][1]);
    if (row0 + g + 8 < T)
      *reinterpret_cast<uint32_t*>(gdst + static_cast<long long>(row0 + g + 8) * ld + 8 * j + 2 * tq) = pack_bf16(acc[j][2], acc[j][3]);
  }
}

__global__ void __launch_bounds__(kAttThreadsMax, 3)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ o,
                     const __nv_bfloat16* __restrict__ d_o, const float* __restrict__ lse2, __nv_bfloat16* __restrict__ dqkv,
                     int T, int Tp) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  const int nthreads = blockDim.x;
  const int nw = nthreads >> 5;
  uint8_t* sQ = att_smem;
  uint8_t* sK = sQ + Tp * 128;
  uint8_t* sV = sK + Tp * 128;
  uint8_t* sdO = sV + Tp * 128;
  float* sD = reinterpret_cast<float*>(sdO + Tp * 128);
  float* sL = sD + Tp;
  const int b = blockIdx.y, h = blockIdx.x;
  const long long tok0 = static_cast<long long>(b) * T;
  const __nv_bfloat16* base = qkv + tok0 * kQkvLd + h * kHeadDim;
  const __nv_bfloat16* obase = o + tok0 * kHidden + h * kHeadDim;
  const __nv_bfloat16* dobase = d_o + tok0 * kHidden + h * kHeadDim;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < Tp * 8; i += nthreads) {
    const int r = i >> 3, ch = i & 7;
    const uint32_t off = swz(r, ch);
    if (r < T) {
      const __nv_bfloat16* src = base + static_cast<long long>(r) * kQkvLd + ch * 8;
      cp_async16(sQ + off, src);
      cp_async16(sK + off, src + kHidden);
      cp_async16(sV + off, src + 2 * kHidden);
      cp_async16(sdO + off, dobase + static_cast<long long>(r) * kHidden + ch * 8);
    } else {
      const uint4 z = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(sQ + off) = z; *reinterpret_cast<uint4*>(sK + off) = z;
      *reinterpret_cast<uint4*>(sV + off) = z; *reinterpret_cast<uint4*>(sdO + off) = z;
    }
  }
  // D[r] = sum_d dO[r, d] * O[r, d]; one warp per row, 2 elements per lane
  for (int r = warp; r < Tp; r += nw) {
    float acc = 0.f;
    if (r < T) {
      const uint32_t a = *reinterpret_cast<const uint32_t*>(dobase + static_cast<long long>(r) * kHidden + 2 * lane);
      const uint32_t c = *reinterpret_cast<const uint32_t*>(obase + static_cast<long long>(r) * kHidden + 2 * lane);
      acc = __uint_as_float(a << 16) * __uint_as_float(c << 16) + __uint_as_float(a & 0xffff0000u) * __uint_as_float(c & 0xffff0000u);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
      sD[r] = acc;
      sL[r] = (r < T) ? lse2[(static_cast<long long>(b) * kHeads + h) * T + r] : 0.f;
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int g = lane >> 2, tq = lane & 3;
  const int li = lane >> 3, lr = lane & 7;
  const uint32_t q_base = smem_u32(sQ), k_base = smem_u32(sK), v_base = smem_u32(sV), do_base = smem_u32(sdO);
  const float sl2 = 0.125f * 1.4426950408889634f;
  const int mt = Tp >> 4;
  __nv_bfloat16* dq_out = dqkv + tok0 * kQkvLd + h * kHeadDim;

  // ------------------------------------------------------------------ phase A: dK, dV
  for (int jt = warp; jt < mt; jt += nw) {
    const int kb0 = jt * 16;
    uint32_t kfA[4][4], vfA[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t off = swz(kb0 + lr + ((li & 1) ? 8 : 0), ks * 2 + (li >> 1));
      ldsm_x4(kfA[ks], k_base + off);
      ldsm_x4(vfA[ks], v_base + off);
    }
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f; dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f; }
    for (int it = 0; it < mt; ++it) {
      const int q0 = it * 16;
      float st[2][4], dpt[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) { st[n][0] = st[n][1] = st[n][2] = st[n][3] = 0.f; dpt[n][0] = dpt[n][1] = dpt[n][2] = dpt[n][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t qB[4], oB[4];
        const uint32_t off = swz(q0 + lr + ((li >> 1) ? 8 : 0), ks * 2 + (li & 1));
        ldsm_x4(qB, q_base + off);
        ldsm_x4(oB, do_base + off);
        mma_bf16(st[0], kfA[ks], qB[0], qB[1]);
        mma_bf16(st[1], kfA[ks], qB[2], qB[3]);
        mma_bf16(dpt[0], vfA[ks], oB[0], oB[1]);
        mma_bf16(dpt[1], vfA[ks], oB[2], oB[3]);
      }
      uint32_t pA[4], dsA[4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int qc = q0 + 8 * n + 2 * tq;
        const float l0 = sL[qc], l1 = sL[qc + 1], d0 = sD[qc], d1 = sD[qc + 1];
        const float p0 = ex2(fmaf(st[n][0], sl2, -l0)), p1 = ex2(fmaf(st[n][1], sl2, -l1));
        const float p2 = ex2(fmaf(st[n][2], sl2, -l0)), p3 = ex2(fmaf(st[n][3], sl2, -l1));
        pA[2 * n] = pack_bf16(p0, p1);
        pA[2 * n + 1] = pack_bf16(p2, p3);
        dsA[2 * n] = pack_bf16(p0 * (dpt[n][0] - d0) * 0.125f, p1 * (dpt[n][1] - d1) * 0.125f);
        dsA[2 * n + 1] = pack_bf16(p2 * (dpt[n][2] - d0) * 0.125f, p3 * (dpt[n][3] - d1) * 0.125f);
      }
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t of[4], qf[4];
        const uint32_t off = swz(q0 + lr + ((li & 1) ? 8 : 0), dp * 2 + (li >> 1));
        ldsm_x4_trans(of, do_base + off);
        ldsm_x4_trans(qf, q_base + off);
        mma_bf16(dv[2 * dp], pA, of[0], of[1]);
        mma_bf16(dv[2 * dp + 1], pA, of[2], of[3]);
        mma_bf16(dk[2 * dp], dsA, qf[0], qf[1]);
        mma_bf16(dk[2 * dp + 1], dsA, qf[2], qf[3]);
      }
    }
    store_tile_bf16(dk, dq_out + kHidden, kQkvLd, kb0, T, lane);
    store_tile_bf16(dv, dq_out + 2 * kHidden, kQkvLd, kb0, T, lane);
  }

  // ------------------------------------------------------------------ phase B: dQ
  for (int it = warp; it < mt; it += nw) {
    const int q0 = it * 16;
    uint32_t qfA[4][4], ofA[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t off = swz(q0 + lr + ((li & 1) ? 8 : 0), ks * 2 + (li >> 1));
      ldsm_x4(qfA[ks], q_base + off);
      ldsm_x4(ofA[ks], do_base + off);
    }
    const float l0 = sL[q0 + g], l1 = sL[q0 + g + 8], d0 = sD[q0 + g], d1 = sD[q0 + g + 8];
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f; }
    for (int jt = 0; jt < mt; ++jt) {
      const int kb0 = jt * 16;
      float sc[2][4], dp_[2][4];
#pragma unroll
      for (int n = 0; n < 2; ++n) { sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f; dp_[n][0] = dp_[n][1] = dp_[n][2] = dp_[n][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t kB[4], vB[4];
        const uint32_t off = swz(kb0 + lr + ((li >> 1) ? 8 : 0), ks * 2 + (li & 1));
        ldsm_x4(kB, k_base + off);
        ldsm_x4(vB, v_base + off);
        mma_bf16(sc[0], qfA[ks], kB[0], kB[1]);
        mma_bf16(sc[1], qfA[ks], kB[2], kB[3]);
        mma_bf16(dp_[0], ofA[ks], vB[0], vB[1]);
        mma_bf16(dp_[1], ofA[ks], vB[2], vB[3]);
      }
      uint32_t dsA[4];
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const float p0 = ex2(fmaf(sc[n][0], sl2, -l0)), p1 = ex2(fmaf(sc[n][1], sl2, -l0));
        const float p2 = ex2(fmaf(sc[n][2], sl2, -l1)), p3 = ex2(fmaf(sc[n][3], sl2, -l1));
        dsA[2 * n] = pack_bf16(p0 * (dp_[n][0] - d0) * 0.125f, p1 * (dp_[n][1] - d0) * 0.125f);
        dsA[2 * n + 1] = pack_bf16(p2 * (dp_[n][2] - d1) * 0.125f, p3 * (dp_[n][3] - d1) * 0.125f);
      }
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t kf[4];
        ldsm_x4_trans(kf, k_base + swz(kb0 + lr + ((li & 1) ? 8 : 0), dp * 2 + (li >> 1)));
        mma_bf16(dq[2 * dp], dsA, kf[0], kf[1]);
        mma_bf16(dq[2 * dp + 1], dsA, kf[2], kf[3]);
      }
    }
    store_tile_bf16(dq, dq_out, kQkvLd, q0, T, lane);
  }
}

int launch_attention_bwd(const __nv_bfloat16* qkv, const __nv_bfloat16* o, const __nv_bfloat16* d_o, const float* lse2,
                         __nv_bfloat16* dqkv, float* dbias, int batch, int tokens, cudaStream_t stream) {
  if (batch <= 0 || tokens <= 0) return kOk;
  if (attention_bwd_tc_supported(tokens)) return launch_attention_bwd_tc(qkv, o, d_o, lse2, dqkv, dbias, batch, tokens, stream);
  if (batch > 65535) return set_error(kErrBadArg, "attention_bwd: batch %d exceeds gridDim.y limit", batch);
  const int mt = (tokens + 15) / 16;
  int nw = (mt % 3 == 0) ? 3 : 4;
  if (nw > mt) nw = mt;
  const int Tp = mt * 16;
  const size_t smem = static_cast<size_t>(4 * Tp) * 128 + static_cast<size_t>(Tp) * 8;
  if (smem > 227 * 1024) return set_error(kErrUnsupported, "attention_bwd: %d tokens need %zu B of shared memory", tokens, smem);
  static size_t configured = 0;
  if (smem > configured) {
    if (cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return set_error(kErrCuda, "attention_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
    configured = smem;
    cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  dim3 grid(kHeads, batch);
  attention_bwd_kernel<<<grid, nw * 32, smem, stream>>>(qkv, o, d_o, lse2, dqkv, tokens, Tp);
  const int rc = check_launch("attention_bwd_kernel");
  if (rc != kOk || dbias == nullptr) return rc;
  // the mma.sync kernel has no fused column sums: one pass over dqkv for the qkv bias gradient
  return launch_colsum_bf16(dqkv, 3 * kHidden, static_cast<long long>(batch) * tokens, 3 * kHidden, dbias, stream);
}

}  // namespace jp
