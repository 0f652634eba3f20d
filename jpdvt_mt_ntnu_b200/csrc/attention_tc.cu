// tcgen05 attention for the JPDVT piece tokens: softmax(Q K^T / 8) V per (sample, head) with both contractions on the
// 5th-generation tensor cores and the scores / outputs living in TMEM.
//
// Replaces timm Attention.forward's F.scaled_dot_product_attention (image_model/models.py:108,120) for the inference
// path, T <= 256 tokens and T % 16 == 0 (144 @192 px, 256 @256 px); the mma.sync kernel in attention.cu keeps the other
// sizes and the training forward (which also wants the log-sum-exp).
//
// One CTA works on one (sample, head) unit at a time, several units per CTA (persistent grid):
//   warps 0-3 : softmax + output epilogue; warp w owns TMEM lanes [32w, 32w+32) = 32 query rows of a 128-row tile
//   warp 4    : TMA producer - Q, K, V of the head straight out of the fused QKV activation [B*T, 2304]
//               ({64 cols x T rows} boxes, 128-byte swizzle -> canonical K-major tiles)
//   warp 5    : MMA issuer   - S = Q K^T  (M = 128 query rows, N = T keys, K = 64)  -> TMEM columns [0, T)
//                              O = P V    (M = 128, N = 64, K = T; P is the bf16 probability tile the softmax warps
//                              wrote to shared memory in K-major swizzled form, V is read MN-major as loaded)
// T = 144 is 128 + 16 query rows.  The 16-row remainder is a second 128-row MMA tile whose A operand starts
// 32 q rows before the end (q = unit index mod 4), so the 16 live rows land in TMEM lanes [32q, 32q+16) and the warp
// that pays for the remainder rotates from unit to unit; its P tile is stored compactly (only those rows exist).
// TMEM per CTA: S [0,T) shared by both tiles in turn, O0 [T,T+64), O1 aliases S[0,64) -> 256 columns at T = 144, two
// CTAs per SM, so one CTA's loads / MMA latencies hide behind the other's softmax.
#include "common.cuh"
#include "ptx.cuh"

namespace jp {

namespace {

constexpr int kTcThreads = 192;
constexpr int kQkvCols = 3 * kHidden;

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int T>
struct TcCfg {
  static_assert(T % 16 == 0 && T >= 16 && T <= 256, "tokens must be a multiple of 16, at most 256");
  static constexpr int kTiles = T > 128 ? 2 : 1;
  static constexpr int kRem = T - 128;                       // live rows of tile 1
  static constexpr bool kCompact = kTiles == 2 && kRem <= 32;   // remainder fits one warp: rotate it, store P1 compactly
  static constexpr int kKBlocks = (T + 63) / 64;             // 64-key blocks of the P tile (128 B per row and block)
  static constexpr int kTileBytes = T * 128;                 // one of Q / K / V
  static constexpr int kPBytes = kKBlocks * 16384;           // P0: 128 rows x kKBlocks x 128 B
  static constexpr int kOffQ = 0, kOffK = kTileBytes, kOffV = 2 * kTileBytes, kOffP = 3 * kTileBytes;
  // the tile-1 A operand reads Q rows up to 255: keep that inside the allocation (it may run into K / V / P, read only)
  static constexpr int kDataBytes = (kOffP + kPBytes) > 256 * 128 ? (kOffP + kPBytes) : 256 * 128;
  static constexpr int kBarOff = kDataBytes;
  static constexpr int kSmemBytes = kDataBytes + 128 + 1024; // + barriers / TMEM slot + alignment slack
  static constexpr int kTmemCols = (T + 64 <= 256) ? 256 : 512;
  static constexpr int kCtasPerSm = (T + 64 <= 256 && 2 * (kSmemBytes + 1024) <= 227 * 1024) ? 2 : 1;
  static_assert(kOffP % 1024 == 0 && kTileBytes % 1024 == 0, "operand tiles must stay 1024-byte aligned (swizzle atoms)");
};

// One softmax pass of the thread's S row (TMEM lane = row, columns [0,T)): exact row maximum, then
// p = 2^((s - max) * log2(e) / 8) written as bf16 into the K-major swizzled P tile; returns sum(p) (fp32, unrounded p).
template <int T>
__device__ __forceinline__ float softmax_row_to_p(uint32_t t_row, uint32_t p_row_addr, uint32_t blk_stride, int sw, bool store) {
  constexpr float sl2 = 0.125f * 1.4426950408889634f;       // head_dim^-0.5 * log2(e)
  constexpr int kFull = T / 32, kTail = T % 32;              // kTail is 0 or 16
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < kFull; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(t_row + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; j += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])));
  }
  if constexpr (kTail != 0) {
    uint32_t r[16];
    tmem_ld_32x16(t_row + kFull * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; j += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])));
  }
  const float ms = mx * sl2;
  float sum = 0.f;
  auto emit8 = [&](const uint32_t* r, int chunk) {            // 8 consecutive keys -> one 16-byte chunk of the P row
    float p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = ex2f(fmaf(__uint_as_float(r[j]), sl2, -ms));
    sum += ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
    const uint4 u = make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
    if (store) sts_u4(p_row_addr + static_cast<uint32_t>(chunk >> 3) * blk_stride + static_cast<uint32_t>(((chunk & 7) ^ sw) << 4), u);
  };
#pragma unroll
  for (int c = 0; c < kFull; ++c) {
    uint32_t r[32];
    tmem_ld_32x32(t_row + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) emit8(r + 8 * g, c * 4 + g);
  }
  if constexpr (kTail != 0) {
    uint32_t r[16];
    tmem_ld_32x16(t_row + kFull * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 2; ++g) emit8(r + 8 * g, kFull * 4 + g);
  }
  return sum;
}

// O tile row (64 fp32 columns in TMEM) * inv -> bf16 -> 128 contiguous bytes of the output row
__device__ __forceinline__ void store_o_row(uint32_t t_row, float inv, __nv_bfloat16* dst, bool live) {
  uint32_t a[32], b[32];
  tmem_ld_32x32(t_row, a);
  tmem_ld_32x32(t_row + 32, b);
  tmem_ld_wait();
  if (live) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      d4[j] = make_uint4(pack_bf16(__uint_as_float(a[8 * j]) * inv, __uint_as_float(a[8 * j + 1]) * inv),
                         pack_bf16(__uint_as_float(a[8 * j + 2]) * inv, __uint_as_float(a[8 * j + 3]) * inv),
                         pack_bf16(__uint_as_float(a[8 * j + 4]) * inv, __uint_as_float(a[8 * j + 5]) * inv),
                         pack_bf16(__uint_as_float(a[8 * j + 6]) * inv, __uint_as_float(a[8 * j + 7]) * inv));
#pragma unroll
    for (int j = 0; j < 4; ++j)
      d4[4 + j] = make_uint4(pack_bf16(__uint_as_float(b[8 * j]) * inv, __uint_as_float(b[8 * j + 1]) * inv),
                             pack_bf16(__uint_as_float(b[8 * j + 2]) * inv, __uint_as_float(b[8 * j + 3]) * inv),
                             pack_bf16(__uint_as_float(b[8 * j + 4]) * inv, __uint_as_float(b[8 * j + 5]) * inv),
                             pack_bf16(__uint_as_float(b[8 * j + 6]) * inv, __uint_as_float(b[8 * j + 7]) * inv));
  }
}

template <int T>
__global__ void __launch_bounds__(kTcThreads, TcCfg<T>::kCtasPerSm)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, int num_units) {
  using Cfg = TcCfg<T>;
  extern __shared__ uint8_t att_tc_smem[];
  uint8_t* smem = att_tc_smem + ((1024u - (smem_u32(att_tc_smem) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* qk_full = bars + 0;        // TMA: Q and K landed
  uint64_t* v_full = bars + 1;         // TMA: V landed
  uint64_t* s_full = bars + 2;         // [2] MMA: scores of tile t are in TMEM
  uint64_t* p_full = bars + 4;         // [2] softmax warps: P tile t is in shared memory (4 arrivals)
  uint64_t* o_full = bars + 6;         // [2] MMA: O tile t is in TMEM
  uint64_t* epi_done = bars + 8;       // softmax warps: every TMEM read of the unit is done (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(qk_full, 1); mbar_init(v_full, 1);
    for (int t = 0; t < 2; ++t) { mbar_init(&s_full[t], 1); mbar_init(&p_full[t], 4); mbar_init(&o_full[t], 1); }
    mbar_init(epi_done, 4);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  if (warp == 4 && lane == 0) tma_prefetch_desc(&tm_qkv);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV),
                 sP = smem_u32(smem + Cfg::kOffP);
  constexpr int kLast = Cfg::kTiles - 1;

  if (warp == 4) {
    // ---------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int b = unit / kHeads, h = unit - b * kHeads;
        const uint32_t prev = static_cast<uint32_t>((it - 1) & 1);
        if (it > 0) mbar_wait(&s_full[kLast], prev);          // every score MMA of the previous unit has read Q, K
        mbar_expect_tx(qk_full, 2 * Cfg::kTileBytes);
        tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffQ, h * kHeadDim, b * T);
        tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffK, kHidden + h * kHeadDim, b * T);
        if (it > 0) mbar_wait(&o_full[kLast], prev);          // ... and every P V MMA has read V
        mbar_expect_tx(v_full, Cfg::kTileBytes);
        tma_load_2d(&tm_qkv, v_full, smem + Cfg::kOffV, 2 * kHidden + h * kHeadDim, b * T);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, T);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);   // B = V, MN-major (keys are the strided index)
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const uint32_t ph = static_cast<uint32_t>(it & 1);
        const int q = Cfg::kCompact ? (unit & 3) : 0;
        mbar_wait(qk_full, ph);
        if (it > 0) mbar_wait(epi_done, static_cast<uint32_t>((it - 1) & 1));   // O of the previous unit has left TMEM
        tc_fence_after();
        auto issue_s = [&](int row0) {
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k)
            umma_bf16(tmem_base, umma_desc_k_sw128(sQ + row0 * 128 + k * 32), umma_desc_k_sw128(sK + k * 32), idesc_s, k != 0);
        };
        auto issue_o = [&](uint32_t d_col, uint32_t p_addr, uint32_t blk_stride) {
#pragma unroll
          for (int j = 0; j < T / 16; ++j)
            umma_bf16(tmem_base + d_col, umma_desc_k_sw128(p_addr + (j >> 2) * blk_stride + (j & 3) * 32),
                      umma_desc_mn_sw128(sV + j * 2048), idesc_o, j != 0);
        };
        issue_s(0);
        umma_commit(&s_full[0]);
        mbar_wait(&p_full[0], ph);                            // softmax has consumed S and written P0
        mbar_wait(v_full, ph);
        tc_fence_after();
        issue_o(T, sP, 16384);
        umma_commit(&o_full[0]);
        if constexpr (Cfg::kTiles == 2) {
          issue_s(Cfg::kCompact ? 128 - 32 * q : 128);
          umma_commit(&s_full[1]);
          mbar_wait(&p_full[1], ph);
          tc_fence_after();
          issue_o(0, sP, Cfg::kCompact ? 2048 : 16384);
          umma_commit(&o_full[1]);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- softmax + epilogue
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int r_tile = warp * 32 + lane;                      // row inside a 128-row tile = TMEM lane
    int it = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const uint32_t ph = static_cast<uint32_t>(it & 1);
      const int b = unit / kHeads, h = unit - b * kHeads;
      const int q = Cfg::kCompact ? (unit & 3) : 0;
      __nv_bfloat16* obase = out + static_cast<long long>(b) * T * kHidden + h * kHeadDim;
      // ---- tile 0
      mbar_wait(&s_full[0], ph);
      tc_fence_after();
      const float sum0 = softmax_row_to_p<T>(t_lane, sP + r_tile * 128, 16384, r_tile & 7, true);
      fence_proxy_async_smem();                               // generic-proxy stores -> visible to the tensor core's reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[0]);
      // ---- tile 1 (the 16- or 128-row remainder)
      float sum1 = 1.f;
      bool mine1 = false;                                     // does this WARP take part in tile 1
      int row1 = 0;                                           // query row of this thread in tile 1
      if constexpr (Cfg::kTiles == 2) {
        mine1 = Cfg::kCompact ? (warp == q) : (warp * 32 < Cfg::kRem);
        row1 = Cfg::kCompact ? 128 + lane : 128 + r_tile;
        if (mine1) {
          mbar_wait(&s_full[1], ph);
          mbar_wait(&o_full[0], ph);                          // P1 reuses P0's shared memory: the O0 MMAs must have read it
          tc_fence_after();
          // compact P1: the 2 KB blocks of successive key groups abut, so only live rows may be written
          sum1 = softmax_row_to_p<T>(t_lane, sP + r_tile * 128, Cfg::kCompact ? 2048 : 16384, r_tile & 7, row1 < T);
          fence_proxy_async_smem();
          tc_fence_before();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[1]);
      }
      // ---- outputs
      mbar_wait(&o_full[0], ph);
      tc_fence_after();
      store_o_row(t_lane + T, 1.0f / sum0, obase + static_cast<long long>(r_tile) * kHidden, r_tile < T);
      if constexpr (Cfg::kTiles == 2) {
        if (mine1) {
          mbar_wait(&o_full[1], ph);
          tc_fence_after();
          store_o_row(t_lane, 1.0f / sum1, obase + static_cast<long long>(row1) * kHidden, row1 < T);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_done);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int T>
int launch_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int batch, cudaStream_t stream) {
  using Cfg = TcCfg<T>;
  static bool configured = false;
  auto kern = attention_tc_kernel<T>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "attention_tc: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
                       cudaGetErrorString(cudaGetLastError()));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  CUtensorMap tm;
  const long long rows = static_cast<long long>(batch) * T;
  int rc = make_tmap_bf16_kmajor(&tm, qkv, rows, kQkvCols, kQkvCols, T);
  if (rc != kOk) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int units = batch * kHeads;
  const int slots = sms * Cfg::kCtasPerSm;
  kern<<<units < slots ? units : slots, kTcThreads, Cfg::kSmemBytes, stream>>>(tm, out, units);
  return check_launch("attention_tc_kernel");
}

}  // namespace

bool attention_tc_supported(int tokens) { return tokens == 144 || tokens == 256; }

int launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, int batch, int tokens, cudaStream_t stream) {
  if (batch <= 0) return kOk;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(kErrBadArg, "attention_tc: pointers must be 16-byte aligned");
  switch (tokens) {
    case 144: return launch_tc<144>(qkv, out, batch, stream);
    case 256: return launch_tc<256>(qkv, out, batch, stream);
    default: return set_error(kErrUnsupported, "attention_tc: %d tokens not instantiated", tokens);
  }
}

}  // namespace jp
