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
attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, int T, int Tp) {
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

int launch_attention(const __nv_bfloat16* qkv, __nv_bfloat16* out, int batch, int tokens, cudaStream_t stream) {
  if (batch <= 0 || tokens <= 0) return kOk;
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
  attention_kernel<<<grid, nw * 32, smem, stream>>>(qkv, out, tokens, Tp);
  return check_launch("attention_kernel");
}

}  // namespace jp
