// Persistent warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue).
//
// CTA pairs (thread-block cluster of 2, tcgen05 cta_group::2): one pair owns a 256 x BN output tile; each CTA loads its
// own 128 rows of A and HALF of the W tile (BN/2 rows) with TMA and both halves feed one 256 x BN x 16 MMA, so the
// L2->SM operand traffic per FLOP is 2/3 of a single-CTA 128 x 256 tile (the round-1 profile showed the single-CTA
// kernel pinned at ~38 B/clk/SM of L2 reads, 48 % tensor-pipe active).
//
//   warp 0    : TMA producer   (cp.async.bulk.tensor 2-D, 128B swizzle; both CTAs signal the LEADER's full barrier)
//   warp 1    : MMA issuer     (leader CTA only: one lane issues tcgen05.mma.cta_group::2, accumulators in the TMEM
//                               of both CTAs, double buffered; tcgen05.commit multicasts to both CTAs' barriers)
//   warps 2-5 : epilogue       (tcgen05.ld 32 lanes x 32 columns, fused math, transpose through padded smem so every
//                               global access is a full 128-byte row segment)
//
// Replaces the cuBLAS calls behind nn.Linear / Conv2d in the reference denoiser (image_model/models.py:108-121,132,
// 169,176-179; timm Attention.qkv/proj, Mlp.fc1/fc2, PatchEmbed.proj) - see SURVEY.md 2.1.
// A and W are both K-major bf16 ({64 x rows} TMA boxes -> canonical K-major SWIZZLE_128B smem tiles), fp32 accumulate.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace jp {

constexpr int BM = 128;         // rows per CTA (256 per pair)
constexpr int BK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kMaxSmem = 227 * 1024;
constexpr int kStageRowBytes = 144;                       // 128 B of payload + 16 B pad: conflict-free 16-byte accesses
constexpr int kStageWarpBytes = 32 * kStageRowBytes;      // one 32-row transpose buffer per epilogue warp

constexpr int kXBoxBytes = 32 * 32 * 4;                   // one residual box: 32 rows x 32 fp32 columns (128-byte swizzled rows)

// NB > 0: per-warp ring of NB residual boxes instead of the staging tiles; LNS: room for the LayerNorm row statistics
// SB: staging tiles per epilogue warp (2 for the epilogue that hands TWO 32 x 64 bf16 boxes per chunk to the TMA)
template <int BN, int CS, int EW, int NB = 0, bool LNS = false, int SB = 1>
struct GemmCfg {
  static constexpr int kThreads = 64 + EW * 32;
  static constexpr int kBRows = BN / CS;                  // W rows this CTA loads
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kColsPerWarp = BN / (EW / 4);      // EW/4 warps share a TMEM lane quadrant and split the columns
  // per warp and tile: bias slice + gate slices of <= 2 samples (the residual-ring variant reads them through L1 instead:
  // its ring needs the room to keep the five pipeline stages)
  static constexpr int kVecWarpBytes = NB > 0 ? 0 : 3 * kColsPerWarp * 4;
  static constexpr int kLnStatsBytes = LNS ? 2 * 2 * BM * 8 : 0;      // EPI_RESID_LN_F32: (mean, M2) per row, column half, M-block parity
  static constexpr int kStagingBytes = NB > 0 ? EW * NB * kXBoxBytes : EW * SB * kStageWarpBytes;
  // staging + barriers + align slack (+ the head epilogue's static arrays, absent from the ring variant)
  static constexpr int kFixedBytes = kStagingBytes + EW * kVecWarpBytes + kLnStatsBytes + 256 + 1024 + (NB > 0 ? 0 : 2304);
  static constexpr int kStagesRaw = (kMaxSmem - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  // two accumulator buffers in a power-of-two allocation; BN = 192 takes 512 columns with the buffers 256 apart
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kAccStride = kTmemCols / 2;
  static constexpr int kBarBytes = (2 * kStages + 4 + EW * NB) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + EW * kVecWarpBytes + kLnStatsBytes + kBarBytes + 1024;
};

// ---- cluster helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // relaxed: the only thing this arrive publishes is "my tcgen05.ld reads of the accumulator are done", which the
  // preceding tcgen05.wait::ld + tcgen05.fence::before_thread_sync already order; a release here would drain every
  // outstanding global store of the epilogue first (ncu r1c: 18 % of the kernel's warp samples sat in that MEMBAR)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load issued by either CTA of a pair; the completion bytes are credited to the barrier at `bar_cluster_addr`
// (a shared::cluster address - the leader CTA's full barrier).
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int32_t c0,
                                                 int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all previously issued MMAs retire) on the barrier at the same smem offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}

__device__ __forceinline__ void sts_u4_shared(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---- epilogue --------------------------------------------------------------------------------------------------------
// The thread that owns accumulator row `row` holds fp32 values v[0..NC) for columns [n0, n0 + NC).
// bf16 outputs are staged 64 columns (128 B per row) at a time, fp32 outputs 32 columns (128 B) at a time.

template <int EPI, typename GateFn>
__device__ __forceinline__ void bf16_math(float (&v)[64], const float* bias_smem, GateFn gate4, int col4) {
  if constexpr (EPI == EPI_BIAS_GELU_BF16) {
    // bias + GELU on packed fp32 pairs: 6 FP32 issue slots + 2 MUFU per two elements instead of 9 + 1 per element
    const float4* b4 = reinterpret_cast<const float4*>(bias_smem);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 b = b4[j];
      bias_gelu_tanh_x2(v[4 * j + 0], v[4 * j + 1], b.x, b.y);
      bias_gelu_tanh_x2(v[4 * j + 2], v[4 * j + 3], b.z, b.w);
    }
  } else if constexpr (EPI != EPI_DGELU_BF16) {
    const float4* b4 = reinterpret_cast<const float4*>(bias_smem);   // broadcast reads of the staged bias slice
#pragma unroll
    for (int j = 0; j < 16; ++j) {             // packed pairs: half the issue slots of 64 scalar adds
      const float4 b = b4[j];
      f2_unpack(f2_add(f2_pack(v[4 * j + 0], v[4 * j + 1]), f2_pack(b.x, b.y)), v[4 * j + 0], v[4 * j + 1]);
      f2_unpack(f2_add(f2_pack(v[4 * j + 2], v[4 * j + 3]), f2_pack(b.z, b.w)), v[4 * j + 2], v[4 * j + 3]);
    }
  }
  if constexpr (EPI == EPI_GATE_BF16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 g = gate4(col4 + j);
      v[4 * j + 0] *= g.x; v[4 * j + 1] *= g.y; v[4 * j + 2] *= g.z; v[4 * j + 3] *= g.w;
    }
  }
}

// 32 rows x 64 bf16 columns of this warp -> global, every store instruction writes 4 full 128-byte row segments.
// colsum (DGELU only, nullable): colsum[n0 + c] += sum over the live rows of the bf16 values written - the bias gradient of the
// layer whose pre-activation gradient this tile is (fc1: db1 = column sums of dh), folded in so that no pass re-reads dh.
template <bool DGELU>
__device__ __forceinline__ void store_bf16_tile(uint8_t* stage, const float (&v)[64], __nv_bfloat16* out,
                                                const __nv_bfloat16* aux, long long ldo, int row0, int n0, int M, int lane,
                                                float* colsum = nullptr) {
  const int sub = lane >> 3, ch = lane & 7;
  uint4 auxv[8];
  if constexpr (DGELU) {   // issue the coalesced gelu' loads first so their latency hides behind the staging round trip
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int r = it * 4 + sub;
      auxv[it] = make_uint4(0, 0, 0, 0);
      if (row0 + r < M) auxv[it] = __ldg(reinterpret_cast<const uint4*>(aux + static_cast<long long>(row0 + r) * ldo + n0 + 8 * ch));
    }
  }
  [[maybe_unused]] float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint8_t* mine = stage + lane * kStageRowBytes;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 u;
    u.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]); u.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
    u.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]); u.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(mine + 16 * j) = u;
  }
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + sub;
    uint4 u = *reinterpret_cast<const uint4*>(stage + r * kStageRowBytes + 16 * ch);
    if (row0 + r < M) {
      if constexpr (DGELU) {   // d(pre-activation) = d(activation) * gelu'(pre-activation), 8 bf16 per lane
        uint32_t uw[4] = {u.x, u.y, u.z, u.w};
        const uint32_t aw[4] = {auxv[it].x, auxv[it].y, auxv[it].z, auxv[it].w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          uw[e] = pack_bf16(__uint_as_float(uw[e] << 16) * __uint_as_float(aw[e] << 16),
                            __uint_as_float(uw[e] & 0xffff0000u) * __uint_as_float(aw[e] & 0xffff0000u));
        u = make_uint4(uw[0], uw[1], uw[2], uw[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) { cs[2 * e] += __uint_as_float(uw[e] << 16); cs[2 * e + 1] += __uint_as_float(uw[e] & 0xffff0000u); }
      }
      *reinterpret_cast<uint4*>(out + static_cast<long long>(row0 + r) * ldo + n0 + 8 * ch) = u;
    }
  }
  if constexpr (DGELU) {
    if (colsum != nullptr) {             // lane (sub, ch): 8 columns x 8 of the warp's 32 rows -> sum over sub -> 2 vector red.adds
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);
        cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
      }
      if (sub == 0) {
        float* dst = colsum + n0 + 8 * ch;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(cs[0]), "f"(cs[1]), "f"(cs[2]), "f"(cs[3]) : "memory");
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 4), "f"(cs[4]), "f"(cs[5]), "f"(cs[6]), "f"(cs[7]) : "memory");
      }
    }
  }
  __syncwarp();
}

// same transposition for values that are already packed bf16 pairs (32 words = 64 columns of this thread's row)
__device__ __forceinline__ void store_packed_tile(uint8_t* stage, const uint32_t (&u)[32], __nv_bfloat16* out, long long ldo,
                                                  int row0, int n0, int M, int lane) {
  const int sub = lane >> 3, ch = lane & 7;
  uint8_t* mine = stage + lane * kStageRowBytes;
#pragma unroll
  for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(mine + 16 * j) = make_uint4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + sub;
    const uint4 q = *reinterpret_cast<const uint4*>(stage + r * kStageRowBytes + 16 * ch);
    if (row0 + r < M) *reinterpret_cast<uint4*>(out + static_cast<long long>(row0 + r) * ldo + n0 + 8 * ch) = q;
  }
  __syncwarp();
}

// 32 rows x 32 fp32 columns of this warp -> staging buffer (row-per-thread in, coalesced row segments out).
__device__ __forceinline__ void stage_f32_tile(uint8_t* stage, const float (&v)[32], int lane) {
  uint8_t* mine = stage + lane * kStageRowBytes;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(mine + 16 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
}

template <int BN, int EPI, int CS, int EW, int NB = 0, bool BMN = false>
__global__ void __cluster_dims__(CS, 1, 1) __launch_bounds__(64 + EW * 32, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
            const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_y, const GemmParams p) {
  constexpr int SB = (EPI == EPI_BIAS_GELU_GRAD_BF16) ? 2 : 1;
  using Cfg = GemmCfg<BN, CS, EW, NB, EPI == EPI_RESID_LN_F32, SB>;
  static_assert((NB > 0) == (EPI == EPI_RESID_TMA_F32 || EPI == EPI_RESID_LN_TMA_F32 || EPI == EPI_RESID_TMA_XB_F32),
                "the residual ring belongs to the TMA residual epilogues");
  constexpr int kColsPerWarp = Cfg::kColsPerWarp;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSETTING the shared-space pointer (integer round-trips make the compiler lose the address
  // space and emit generic LD/ST for every staging access)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage_buf = smem + Cfg::kStages * Cfg::kStageBytes;
  uint8_t* bias_buf = stage_buf + Cfg::kStagingBytes;     // stage_buf: staging tiles, or (NB > 0) the residual rings, 1024-aligned
  float2* ln_stats = reinterpret_cast<float2*>(bias_buf + EW * Cfg::kVecWarpBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bias_buf + EW * Cfg::kVecWarpBytes + Cfg::kLnStatsBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;   // [2] accumulator ready   (own CTA)
  uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained (leader's is the one waited on)
  uint64_t* xfull_bar = tempty_bar + 2;             // [EW * NB] residual box landed (TMA residual epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xfull_bar + EW * NB);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CS == 2) ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  constexpr bool WG = (EPI == EPI_WGRAD_F32);              // MN-major operands + split contraction (weight gradients)
  const int out_rows = WG ? p.wg_rows : p.M;
  const int num_m = (out_rows + CS * BM - 1) / (CS * BM);  // tiles of CS*128 rows
  const int num_n = p.N / BN;
  const int tiles_mn = num_m * num_n;
  const int num_tiles = WG ? tiles_mn * p.split : tiles_mn;
  const int num_kb = WG ? p.split_len / BK : p.K / BK;
  const int group = blockIdx.x / CS, num_groups = gridDim.x / CS;
  constexpr bool kLN = (EPI == EPI_RESID_LN_F32);           // residual update + the next LayerNorm of the same rows
  constexpr bool kLNX = (EPI == EPI_RESID_LN_TMA_F32);      // ... with both the residual tile and the LayerNorm pass through the TMA ring
  constexpr bool kXB = (EPI == EPI_RESID_TMA_XB_F32);       // ... plus a bf16 copy of the updated rows and their (sum, sum of squares)
  constexpr bool kXR = (EPI == EPI_RESID_TMA_F32) || kXB;   // residual tile through the TMA ring
  constexpr bool kFoldIn = (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16);   // may consume a folded LayerNorm (p.stats_in)
  constexpr bool kResid = (EPI == EPI_RESID_F32) || kLN;
  // s-th work item of this CTA group.  Normally tiles are dealt round robin; the LayerNorm-fused epilogue needs whole rows,
  // so there a group owns M-blocks and walks the N tiles of each one in turn.
  auto tile_at = [&](int s, int& split_idx, int& m_blk, int& n_blk) -> bool {
    if constexpr (kLN || kLNX) {
      split_idx = 0;
      m_blk = group + (s / num_n) * num_groups;
      n_blk = s % num_n;
      return m_blk < num_m;
    } else {
      const int tile = group + s * num_groups;
      if (tile >= num_tiles) return false;
      split_idx = tile / tiles_mn;
      const int rem = tile - split_idx * tiles_mn;
      m_blk = rem / num_n;
      n_blk = rem % num_n;
      if (p.reverse_m) m_blk = num_m - 1 - m_blk;            // walk the row blocks downwards (sweep_reverse(), common.cuh)
      return true;
    }
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], CS * EW); }
    for (int i = 0; i < EW * NB; ++i) mbar_init(&xfull_bar[i], 1);
    if constexpr (kLNX) tma_prefetch_desc(&tma_y);
    if constexpr (NB > 0) tma_prefetch_desc(&tma_x);
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CS == 2) { tmem_alloc2(tmem_slot, Cfg::kTmemCols); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (CS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                  // everything above overlapped the previous kernel's tail; from here on global memory is touched
  griddep_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane per CTA)
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int s = 0;; ++s) {
        int split_idx, m_blk, n_blk;
        if (!tile_at(s, split_idx, m_blk, n_blk)) break;
        const int row_a = (m_blk * CS + static_cast<int>(rank)) * BM;
        const int row_b = n_blk * BN + static_cast<int>(rank) * Cfg::kBRows;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if constexpr (WG) {
            // operands are [contraction rows, features] row-major: {64 features x 64 rows} boxes stacked along the
            // feature (MN) dimension, 8 KB each
            static_assert(!WG || CS == 2, "wgrad mode is pair-only");
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            const uint32_t bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
            const int m0 = split_idx * p.split_len + kb * BK;
#pragma unroll
            for (int hh = 0; hh < BM / 64; ++hh) tma_load_2d_pair(&tma_a, bar, sa + hh * 8192, row_a + 64 * hh, m0);
#pragma unroll
            for (int hh = 0; hh < Cfg::kBRows / 64; ++hh) tma_load_2d_pair(&tma_b, bar, sb + hh * 8192, row_b + 64 * hh, m0);
          } else if constexpr (BMN) {
            // B = W [K, N] row-major (contraction rows, feature columns): {64 features x 64 contraction rows} boxes stacked
            // along the feature dimension, exactly the weight-gradient mode's operand layout; A stays K-major
            static_assert(!BMN || CS == 2, "the MN-major B operand is pair-only");
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            const uint32_t bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_pair(&tma_a, bar, sa, kb * BK, row_a);
#pragma unroll
            for (int hh = 0; hh < Cfg::kBRows / 64; ++hh) tma_load_2d_pair(&tma_b, bar, sb + hh * 8192, row_b + 64 * hh, kb * BK);
          } else if constexpr (CS == 2) {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);   // both CTAs' bytes land on the leader's barrier
            const uint32_t bar = map_to_cta(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_pair(&tma_a, bar, sa, kb * BK, row_a);
            tma_load_2d_pair(&tma_b, bar, sb, kb * BK, row_b);
          } else {
            mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d(&tma_a, &full_bar[stage], sa, kb * BK, row_a);
            tma_load_2d(&tma_b, &full_bar[stage], sb, kb * BK, row_b);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();   // reconverge before the aligned cluster barrier at teardown
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, one lane)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(CS * BM, BN, WG ? 1 : 0, (WG || BMN) ? 1 : 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int s = 0;; ++s) {
        int split_idx, m_blk, n_blk;
        if (!tile_at(s, split_idx, m_blk, n_blk)) break;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * Cfg::kAccStride);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 16 elements = 32 B inside the swizzle row; MN-major: 16 contraction rows = two 1024 B atoms
            const uint64_t da = WG ? umma_desc_mn_sw128(a_addr + k * 2048) : umma_desc_k_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t db = (WG || BMN) ? umma_desc_mn_sw128(b_addr + k * 2048) : umma_desc_k_sw128(b_addr + k * UMMA_K * 2);
            if constexpr (CS == 2) umma_bf16_pair(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (CS == 2) umma_commit_pair(&empty_bar[stage], 3); else umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if constexpr (CS == 2) umma_commit_pair(&tfull_bar[acc], 3); else umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2 .. 2+EW)
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    uint8_t* stage = stage_buf + (warp - 2) * SB * kStageWarpBytes;
    const int col_base = ((warp - 2) >> 2) * kColsPerWarp;   // first accumulator column of this warp
    int acc = 0; uint32_t acc_phase = 0;
    __shared__ float s_w2[kLatent * 64];
    __shared__ float s_b2[kLatent];
    if constexpr (EPI == EPI_HEAD) {
      for (int i = threadIdx.x - 64; i < kLatent * 64; i += EW * 32) s_w2[i] = p.w2[i];
      if (threadIdx.x - 64 < kLatent) s_b2[threadIdx.x - 64] = p.b2[threadIdx.x - 64];
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
    }
    const uint32_t tempty_remote = (CS == 2) ? map_to_cta(smem_u32(&tempty_bar[0]), 0) : 0u;
    if constexpr (kLNX) {
      // ============================================================================================================
      // Gated residual update + the LayerNorm-modulate of the same rows, everything through this warp's TMA ring.
      // The CTA pair owns whole 256-row blocks (tile_at walks the N tiles of an M-block in turn).  Per M-block the warp
      // (lane quadrant `quad`, column half `half`) handles 32 rows x 128 columns of each of the kNT N tiles:
      //   phase R: box b = (tile j, chunk c): TMA-load x, x += gate * (acc + bias) in place, TMA-store; the row's running
      //            (count, mean, M2) is updated from the registers that hold the new values (Chan's merge per 32 values)
      //   phase L: statistics of the two column halves meet through shared memory; the same boxes are re-read (L2 hits;
      //            this warp wrote them itself), normalised + modulated, and leave as bf16 boxes of ln_out.
      // Boxes are consumed in the fixed order [R 0..kB) [L 0..kB) per M-block; box `pos` sits in ring slot pos % NB and is
      // requested up to NB - 1 positions ahead, but an L box never before the stores of its M-block have completed.
      constexpr int kNC = kColsPerWarp / 32;               // 32-column chunks per tile and warp
      const int kNT = num_n;                               // N tiles per M-block (3 at N = 768)
      const int kB = kNT * kNC;                            // boxes per phase
      const int wi = warp - 2, half = wi >> 2;
      uint8_t* ring = stage_buf + wi * NB * kXBoxBytes;
      uint64_t* xbar = xfull_bar + wi * NB;
      const int bar_id = 2 + quad;                         // named barrier of the two warps that share a lane quadrant
      long long issued = 0, pos = 0;                       // boxes requested / consumed so far (this warp)
      const int my_blocks = (num_m > group) ? (num_m - group + num_groups - 1) / num_groups : 0;
      const long long total = static_cast<long long>(my_blocks) * 2 * kB;
      // coordinates of the box consumed at position q
      auto box_at = [&](long long q, bool& is_l, int& col, int& row) {
        const int mb_it = static_cast<int>(q / (2 * kB));
        int k = static_cast<int>(q - static_cast<long long>(mb_it) * 2 * kB);
        is_l = k >= kB;
        if (is_l) k -= kB;
        const int mb = group + mb_it * num_groups;
        row = (mb * CS + static_cast<int>(rank)) * BM + quad * 32;
        col = (k / kNC) * BN + col_base + (k % kNC) * 32;
      };
      // lane 0: request boxes up to position `upto` (exclusive); L boxes only when `l_ok` (their M-block's stores are complete)
      auto issue_to = [&](long long upto, bool l_ok) {
        if (upto > total) upto = total;
        while (issued < upto) {
          bool is_l; int col, row;
          box_at(issued, is_l, col, row);
          if (is_l && !l_ok) break;
          const int slot = static_cast<int>(issued % NB);
          mbar_expect_tx(&xbar[slot], kXBoxBytes);
          tma_load_2d_hint(&tma_x, &xbar[slot], ring + slot * kXBoxBytes, col, row, is_l ? kEvictFirst : kEvictNormal);
          ++issued;
        }
      };
      if (lane == 0) issue_to(NB - 1, false);
      __syncwarp();
      for (int mb_it = 0; mb_it < my_blocks; ++mb_it) {
        const int m_blk = group + mb_it * num_groups;
        const int row0 = (m_blk * CS + static_cast<int>(rank)) * BM + quad * 32;
        const int row = row0 + lane;
        const bool row_ok = row < out_rows;
        const long long sample = (row_ok ? row : (out_rows - 1)) / p.tokens;
        const float* gate_row = p.gate + sample * p.gate_stride + col_base;
        float mean = 0.f, m2 = 0.f, cnt = 0.f;             // this row, this warp's columns
        // ---------------------------------------------------------------- phase R
        for (int j = 0; j < kNT; ++j) {
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * Cfg::kAccStride);
#pragma unroll
          for (int c = 0; c < kNC; ++c, ++pos) {
            const int slot = static_cast<int>(pos % NB);
            if (lane == 0) {
              tma_store_wait_read<0>();                      // the store that used the slot about to be refilled has read it
              issue_to(pos + NB, false);
            }
            const int n0 = j * BN + c * 32;                  // column offset relative to col_base
            float4 bb[8], gg[8];                             // bias / gate of this chunk: requested before the waits below
            {
              const float4* b4 = reinterpret_cast<const float4*>(p.bias + col_base + n0);
              const float4* g4 = reinterpret_cast<const float4*>(gate_row + n0);
#pragma unroll
              for (int i = 0; i < 8; ++i) { bb[i] = __ldg(b4 + i); gg[i] = __ldg(g4 + i); }
            }
            uint32_t r[32];
            tmem_ld_32x32(t_row + col_base + c * 32, r);
            tmem_ld_wait();
            if (c == kNC - 1) {                              // accumulator fully read: hand the TMEM buffer back
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if constexpr (CS == 2) mbar_arrive_cluster(tempty_remote + static_cast<uint32_t>(acc) * 8u);
                else mbar_arrive(&tempty_bar[acc]);
              }
            }
            float v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = bb[i], gt = gg[i];
              v[4 * i + 0] = gt.x * (__uint_as_float(r[4 * i + 0]) + b.x); v[4 * i + 1] = gt.y * (__uint_as_float(r[4 * i + 1]) + b.y);
              v[4 * i + 2] = gt.z * (__uint_as_float(r[4 * i + 2]) + b.z); v[4 * i + 3] = gt.w * (__uint_as_float(r[4 * i + 3]) + b.w);
            }
            mbar_wait(&xbar[slot], static_cast<uint32_t>((pos / NB) & 1));
            const uint32_t mine = smem_u32(ring + slot * kXBoxBytes) + lane * 128, sw = lane & 7;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t a = mine + ((i ^ sw) << 4);
              const float4 q = lds_f4(a);
              v[4 * i + 0] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[4 * i]), "f"(v[4 * i + 1]), "f"(v[4 * i + 2]),
                           "f"(v[4 * i + 3]) : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {   // keep the updated rows in L2: phase L of this M-block reads them back
              tma_store_2d_hint(&tma_x, ring + slot * kXBoxBytes, col_base + n0, row0, kEvictLast);
              tma_store_commit();
            }
            // exact two-pass moments of these 32 new values, merged into the row's running (count, mean, M2)
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 2) { s0 += v[i]; s1 += v[i + 1]; }
            const float mc = (s0 + s1) * (1.0f / 32.0f);
            float q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float d0 = v[i] - mc, d1 = v[i + 1] - mc;
              q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
            }
            const float delta = mc - mean, n_new = cnt + 32.0f;
            mean = fmaf(delta, 32.0f / n_new, mean);
            m2 += (q0 + q1) + delta * delta * (cnt * 32.0f / n_new);
            cnt = n_new;
          }
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        // ---------------------------------------------------------------- phase L
        if (lane == 0) tma_store_wait<0>();                  // every box of this M-block is in global memory (and no slot is being read)
        __syncwarp();
        float2* xs = reinterpret_cast<float2*>(ring);        // slot 0 doubles as the exchange buffer: (mean, M2) per row
        xs[lane] = make_float2(mean, m2);
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
        const float2 other = reinterpret_cast<const float2*>(stage_buf + (wi ^ 4) * NB * kXBoxBytes)[lane];
        asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");   // both have read before the ring is refilled
        const float dm = other.x - mean;
        const float row_mean = 0.5f * (mean + other.x);      // equal column counts in both halves
        const float rstd = rsqrtf((m2 + other.y + dm * dm * (0.5f * cnt)) / (2.0f * cnt) + 1e-6f);
        const float* sh_row = p.ln_shift + sample * p.ln_stride + col_base;
        const float* sc_row = p.ln_scale + sample * p.ln_stride + col_base;
        fence_proxy_async_smem();                            // generic reads / writes of slot 0 before the TMA refills it
        if (lane == 0) issue_to(pos + NB - 1, true);
        __syncwarp();
        for (int k = 0; k < kB; ++k, ++pos) {
          const int slot = static_cast<int>(pos % NB);
          if (lane == 0) {
            tma_store_wait_read<0>();
            issue_to(pos + NB, true);                        // further L boxes, then the next M-block's R boxes
          }
          const int n0 = (k / kNC) * BN + (k % kNC) * 32;
          float4 sa[8], sb[8];                               // shift / scale of this box: in flight while the box arrives
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            sa[i] = __ldg(reinterpret_cast<const float4*>(sh_row + n0) + i);
            sb[i] = __ldg(reinterpret_cast<const float4*>(sc_row + n0) + i);
          }
          mbar_wait(&xbar[slot], static_cast<uint32_t>((pos / NB) & 1));
          const uint32_t base = smem_u32(ring + slot * kXBoxBytes);
          const uint32_t mine = base + lane * 128, sw = lane & 7;
          float y[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = lds_f4(mine + ((i ^ sw) << 4));
            const float4 a = sa[i], b = sb[i];
            y[4 * i + 0] = fmaf((q.x - row_mean) * rstd, 1.0f + b.x, a.x); y[4 * i + 1] = fmaf((q.y - row_mean) * rstd, 1.0f + b.y, a.y);
            y[4 * i + 2] = fmaf((q.z - row_mean) * rstd, 1.0f + b.z, a.z); y[4 * i + 3] = fmaf((q.w - row_mean) * rstd, 1.0f + b.w, a.w);
          }
          __syncwarp();                                      // every lane has read its row before the slot is rewritten
          // bf16 result: 32 rows x 64 bytes, SWIZZLE_64B (16-byte chunk i of row r at (i ^ ((r >> 1) & 3)))
          const uint32_t orow = base + lane * 64, osw = (lane >> 1) & 3;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            sts_u4_shared(orow + ((i ^ osw) << 4), pack_bf16(y[8 * i], y[8 * i + 1]), pack_bf16(y[8 * i + 2], y[8 * i + 3]),
                          pack_bf16(y[8 * i + 4], y[8 * i + 5]), pack_bf16(y[8 * i + 6], y[8 * i + 7]));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tma_y, ring + slot * kXBoxBytes, col_base + n0, row0);
            tma_store_commit();
          }
        }
      }
    } else {
    float4 fst[3] = {};                   // kFoldIn: this thread's row statistics of the NEXT tile, requested one tile ahead
    float ln_mean = 0.f, ln_m2 = 0.f;     // kLN: running (mean, sum of squared deviations) of this thread's row over its columns
    int ln_cnt = 0;
    // kXR: this warp's ring of residual boxes.  Box g (g = s * NC + c, the c-th 32-column chunk of the warp's s-th tile)
    // lives in slot g % NB; its load is issued NB - 1 chunks ahead, as soon as the TMA store of box g - 1 has read its slot.
    constexpr int kXNC = kColsPerWarp / 32;
    uint8_t* xring = stage_buf + (warp - 2) * (NB > 0 ? NB : 1) * kXBoxBytes;
    uint64_t* xbar = xfull_bar + (warp - 2) * NB;
    auto x_coords = [&](int g, int& col, int& row) -> bool {
      int sp, mb, nb;
      if (!tile_at(g / kXNC, sp, mb, nb)) return false;
      row = (mb * CS + static_cast<int>(rank)) * BM + quad * 32;
      col = nb * BN + col_base + (g % kXNC) * 32;
      return true;
    };
    auto x_issue = [&](int g) {            // lane 0 only
      int col, row;
      if (x_coords(g, col, row)) {
        const int slot = g % (NB > 0 ? NB : 1);
        mbar_expect_tx(&xbar[slot], kXBoxBytes);
        tma_load_2d(&tma_x, &xbar[slot], xring + slot * kXBoxBytes, col, row);
      }
    };
    if constexpr (kXR) {
      if (lane == 0) {
        for (int g = 0; g < NB - 1; ++g) x_issue(g);
      }
      __syncwarp();
    }
    for (int s = 0;; ++s) {
      int split_idx, m_blk, n_blk;
      if (!tile_at(s, split_idx, m_blk, n_blk)) break;
      const int row0 = (m_blk * CS + static_cast<int>(rank)) * BM + quad * 32;   // first row of this warp
      const int row = row0 + lane;
      const bool row_ok = row < out_rows;
      float* bias_smem = reinterpret_cast<float*>(bias_buf + (warp - 2) * Cfg::kVecWarpBytes);
      float* gate_smem = bias_smem + kColsPerWarp;        // [2][kColsPerWarp]: gate rows of the first / last sample of the warp
      if constexpr (EPI != EPI_DGELU_BF16 && EPI != EPI_WGRAD_F32 && EPI != EPI_HEAD && !kXR) {
        // this warp's bias (and gate) slices -> smem while the MMAs of the tile are still running, so the per-chunk
        // epilogue math never queues a global load behind the streaming residual prefetch
        for (int i = lane; i < kColsPerWarp / 4; i += 32)
          reinterpret_cast<float4*>(bias_smem)[i] = __ldg(reinterpret_cast<const float4*>(p.bias + n_blk * BN + col_base) + i);
        if constexpr (kFoldIn) {
          if (p.stats_in != nullptr) {        // folded LayerNorm: u slice next to the bias (= v) slice
            for (int i = lane; i < kColsPerWarp / 4; i += 32)
              reinterpret_cast<float4*>(gate_smem)[i] = __ldg(reinterpret_cast<const float4*>(p.fold_u + n_blk * BN + col_base) + i);
          }
        }
        if constexpr (kResid || kXR || EPI == EPI_GATE_BF16) {
          const int last = out_rows - 1;
          const int s_first = (row0 < last ? row0 : last) / p.tokens, s_last = (row0 + 31 < last ? row0 + 31 : last) / p.tokens;
          const float* ga = p.gate + static_cast<long long>(s_first) * p.gate_stride + n_blk * BN + col_base;
          const float* gb = p.gate + static_cast<long long>(s_last) * p.gate_stride + n_blk * BN + col_base;
          for (int i = lane; i < kColsPerWarp / 4; i += 32) {
            reinterpret_cast<float4*>(gate_smem)[i] = __ldg(reinterpret_cast<const float4*>(ga) + i);
            reinterpret_cast<float4*>(gate_smem + kColsPerWarp)[i] = __ldg(reinterpret_cast<const float4*>(gb) + i);
          }
        }
        __syncwarp();
      }
      // gate row of this thread's sample: the staged copy when the warp's 32 rows span at most two samples (tokens >= 32),
      // else straight from global memory (kept as two typed pointers so the staged reads stay LDS)
      uint32_t gate_sm = smem_u32(gate_smem);
      const float* gate_gl = nullptr;
      const bool gate_staged = !kXR && p.tokens >= 32;
      if constexpr (kResid || kXR || EPI == EPI_GATE_BF16) {
        const int last = out_rows - 1;
        const int s_first = (row0 < last ? row0 : last) / p.tokens;
        const int s_mine = row_ok ? row / p.tokens : s_first;
        if (s_mine != s_first) gate_sm += kColsPerWarp * 4;
        gate_gl = p.gate + static_cast<long long>(s_mine) * p.gate_stride + n_blk * BN + col_base;
      }
      auto gate4 = [&](int col4) -> float4 {   // 4 gate values of this thread's sample at warp-local column 4 * col4
        if (gate_staged) return lds_f4(gate_sm + 16u * static_cast<uint32_t>(col4));
        return __ldg(reinterpret_cast<const float4*>(gate_gl) + col4);
      };
      auto release_tmem = [&]() {   // accumulator fully read into registers: hand the TMEM buffer back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CS == 2) mbar_arrive_cluster(tempty_remote + static_cast<uint32_t>(acc) * 8u);
          else mbar_arrive(&tempty_bar[acc]);
        }
      };
      // EPI_RESID_F32: the residual-stream tile this warp updates, as coalesced 128-byte row segments (lane -> row
      // it*4 + lane/8, 16-byte chunk lane%8), double buffered so chunk c+1 streams in while chunk c is processed; the
      // first chunk is requested before the accumulator is even complete
      auto load_x = [&](float4 (&buf)[8], int c) {
        const float* xb = reinterpret_cast<const float*>(p.out) + n_blk * BN + col_base + c * 32 + 4 * (lane & 7);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int grow = row0 + it * 4 + (lane >> 3);
          buf[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (grow < out_rows) buf[it] = __ldcs(reinterpret_cast<const float4*>(xb + static_cast<long long>(grow) * p.ldo));
        }
      };
      float4 xa[8], xb[8];
      if constexpr (kResid) {
        load_x(xa, 0);
      }
      float ln_r = 1.0f, ln_rm = 0.0f;       // kFoldIn: rstd and rstd * mean of this thread's row (LayerNorm over the K inputs)
      if constexpr (kFoldIn) {
        if (p.stats_in != nullptr) {
          // six (sum, sum of squares) slots = three 16-byte loads, all in flight at once (a counted loop serialises them:
          // 6 x L2 latency per tile, qkv 110 vs 95 us); from the second tile on they were requested one tile ahead
          float4 t0 = fst[0], t1 = fst[1], t2 = fst[2];
          if (s == 0 && row_ok) {
            const float4* st = reinterpret_cast<const float4*>(p.stats_in + static_cast<long long>(row) * 6);
            t0 = __ldg(st); t1 = __ldg(st + 1); t2 = __ldg(st + 2);
          }
          const float sum = (t0.x + t0.z) + (t1.x + t1.z) + (t2.x + t2.z), sq = (t0.y + t0.w) + (t1.y + t1.w) + (t2.y + t2.w);
          const float inv = 1.0f / static_cast<float>(p.K);
          const float mean = sum * inv;
          ln_r = rsqrtf(fmaxf(fmaf(-mean, mean, sq * inv), 0.f) + 1e-6f);   // nn.LayerNorm(eps=1e-6), biased variance (models.py:107)
          ln_rm = ln_r * mean;
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * Cfg::kAccStride);
      if constexpr (kFoldIn) {
        if (p.stats_in != nullptr) {          // the next tile's row statistics: in flight behind this tile's epilogue math
          int sp, mb, nb;
          if (tile_at(s + 1, sp, mb, nb)) {
            const int nrow = (mb * CS + static_cast<int>(rank)) * BM + quad * 32 + lane;
            if (nrow < out_rows) {
              const float4* st = reinterpret_cast<const float4*>(p.stats_in + static_cast<long long>(nrow) * 6);
              fst[0] = __ldg(st); fst[1] = __ldg(st + 1); fst[2] = __ldg(st + 2);
            }
          }
        }
      }

      if constexpr (EPI == EPI_HEAD) {
        static_assert(EPI != EPI_HEAD || (BN == 64 && EW == 4), "head epilogue needs the whole 64-wide row in one warp");
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_row, r0);
        tmem_ld_32x32(t_row + 32, r1);
        tmem_ld_wait();
        release_tmem();
        float h[64];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          h[j] = silu(__uint_as_float(r0[j]) + __ldg(p.bias + j));
          h[32 + j] = silu(__uint_as_float(r1[j]) + __ldg(p.bias + 32 + j));
        }
        if (row_ok && p.out2 != nullptr) {   // training: keep the pre-activations (time_emb_out1 output incl. bias)
          float4* pre = reinterpret_cast<float4*>(p.out2 + static_cast<long long>(row) * 64);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            pre[j] = make_float4(__uint_as_float(r0[4 * j]) + __ldg(p.bias + 4 * j), __uint_as_float(r0[4 * j + 1]) + __ldg(p.bias + 4 * j + 1),
                                 __uint_as_float(r0[4 * j + 2]) + __ldg(p.bias + 4 * j + 2), __uint_as_float(r0[4 * j + 3]) + __ldg(p.bias + 4 * j + 3));
            pre[8 + j] = make_float4(__uint_as_float(r1[4 * j]) + __ldg(p.bias + 32 + 4 * j), __uint_as_float(r1[4 * j + 1]) + __ldg(p.bias + 33 + 4 * j),
                                     __uint_as_float(r1[4 * j + 2]) + __ldg(p.bias + 34 + 4 * j), __uint_as_float(r1[4 * j + 3]) + __ldg(p.bias + 35 + 4 * j));
          }
        }
        if (row_ok) {
          float o[kLatent];
#pragma unroll
          for (int d = 0; d < kLatent; ++d) {
            float s = s_b2[d];
#pragma unroll
            for (int j = 0; j < 64; ++j) s = fmaf(h[j], s_w2[d * 64 + j], s);
            o[d] = s;
          }
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + static_cast<long long>(row) * p.ldo);
          dst[0] = make_float4(o[0], o[1], o[2], o[3]);
          dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
      } else if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_GATE_BF16 || EPI == EPI_BIAS_BF16_F32 ||
                           EPI == EPI_DGELU_BF16 || EPI == EPI_BIAS_GELU_GRAD_BF16) {
#pragma unroll 1
        for (int c = 0; c < kColsPerWarp / 64; ++c) {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32(t_row + col_base + c * 64, r0);
          tmem_ld_32x32(t_row + col_base + c * 64 + 32, r1);
          tmem_ld_wait();
          if (c == kColsPerWarp / 64 - 1) release_tmem();
          float v[64];
#pragma unroll
          for (int j = 0; j < 32; ++j) { v[j] = __uint_as_float(r0[j]); v[32 + j] = __uint_as_float(r1[j]); }
          const int n0 = n_blk * BN + col_base + c * 64;
          if constexpr (kFoldIn) {
            if (p.stats_in != nullptr) {      // acc is bf16(x) . W'^T: apply the row's LayerNorm, y = r acc - r m u (+ v via the bias slot)
              const float4* u4 = reinterpret_cast<const float4*>(gate_smem + c * 64);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float4 u = u4[j];
                const uint64_t r2 = f2_pack(ln_r, ln_r), nm2 = f2_pack(-ln_rm, -ln_rm);
                f2_unpack(f2_fma(r2, f2_pack(v[4 * j + 0], v[4 * j + 1]), f2_mul(nm2, f2_pack(u.x, u.y))), v[4 * j + 0], v[4 * j + 1]);
                f2_unpack(f2_fma(r2, f2_pack(v[4 * j + 2], v[4 * j + 3]), f2_mul(nm2, f2_pack(u.z, u.w))), v[4 * j + 2], v[4 * j + 3]);
              }
            }
          }
          bf16_math<EPI>(v, bias_smem + c * 64, gate4, c * 16);
          if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16) {
            if (p.tma_out) {
              // 32 rows x 64 bf16 = one 4 KB box of 128-byte swizzled rows, written row-per-thread (conflict free) and
              // handed to the TMA: no read-back, no per-thread global store, the previous box drains while this one fills
              const uint32_t sbase = smem_u32(stage);
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
              const uint32_t mine = sbase + lane * 128, sw = ((sbase >> 7) + lane) & 7;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sts_u4_shared(mine + ((j ^ sw) << 4), pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                              pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tma_x, stage, n0, row0);
                tma_store_commit();
              }
              continue;
            }
          }
          if constexpr (EPI == EPI_BIAS_GELU_GRAD_BF16) {   // activation and its derivative from one tanh; derivative tile first
            uint32_t gp[32];
#pragma unroll
            for (int j = 0; j < 64; j += 2) {
              float y0, d0, y1, d1;
              gelu_tanh_both_x2(v[j], v[j + 1], y0, y1, d0, d1);
              v[j] = y0; v[j + 1] = y1;
              gp[j >> 1] = pack_bf16(d0, d1);
            }
            if (p.tma_out) {
              // both 32 x 64 bf16 tiles leave as 4 KB boxes of 128-byte swizzled rows (row per thread, conflict free) through
              // the TMA - the form that took the sampling fc1 epilogue off the critical path; the warp's two staging tiles
              // hold the derivative box and the activation box, refilled once the previous chunk's stores have read them
              const uint32_t sb0 = smem_u32(stage), sb1 = sb0 + kStageWarpBytes;
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
              const uint32_t mine0 = sb0 + lane * 128, sw0 = ((sb0 >> 7) + lane) & 7;
              const uint32_t mine1 = sb1 + lane * 128, sw1 = ((sb1 >> 7) + lane) & 7;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                sts_u4_shared(mine0 + ((j ^ sw0) << 4), gp[4 * j], gp[4 * j + 1], gp[4 * j + 2], gp[4 * j + 3]);
                sts_u4_shared(mine1 + ((j ^ sw1) << 4), pack_bf16(v[8 * j], v[8 * j + 1]), pack_bf16(v[8 * j + 2], v[8 * j + 3]),
                              pack_bf16(v[8 * j + 4], v[8 * j + 5]), pack_bf16(v[8 * j + 6], v[8 * j + 7]));
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tma_y, stage, n0, row0);
                tma_store_2d(&tma_x, stage + kStageWarpBytes, n0, row0);
                tma_store_commit();
              }
              continue;
            }
            store_packed_tile(stage, gp, p.out_aux, p.ldo, row0, n0, p.M, lane);
          }
          store_bf16_tile<EPI == EPI_DGELU_BF16>(stage, v, reinterpret_cast<__nv_bfloat16*>(p.out), p.aux, p.ldo, row0, n0, p.M, lane,
                                                 EPI == EPI_DGELU_BF16 ? p.out2 : nullptr);
          if constexpr (EPI == EPI_BIAS_BF16_F32) {
            if (p.out2 != nullptr) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                float w32[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) w32[j] = v[32 * hh + j];
                stage_f32_tile(stage, w32, lane);
                const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  const int r = it * 4 + sub;
                  const float4 u = *reinterpret_cast<const float4*>(stage + r * kStageRowBytes + 16 * ch);
                  if (row0 + r < p.M)
                    *reinterpret_cast<float4*>(p.out2 + static_cast<long long>(row0 + r) * p.ldo + n0 + 32 * hh + 4 * ch) = u;
                }
                __syncwarp();
              }
            }
          }
        }
      } else if constexpr (kXR) {
        // x[row, n] += gate[row / tokens, n] * (acc + bias[n]): the residual box arrives by TMA (swizzled 128-byte rows, row =
        // lane), is updated in place by its row's thread and leaves by TMA store - no per-thread global traffic at all
        float rs = 0.f, rq = 0.f;            // kXB: sum and sum of squares of this thread's updated row over the warp's columns
#pragma unroll
        for (int c = 0; c < kXNC; ++c) {
          const int g = s * kXNC + c;
          const int slot = g % NB;
          if (lane == 0) {
            tma_store_wait_read<0>();                        // the store of box g - 1 has released slot (g - 1) % NB ...
            x_issue(g + NB - 1);                             // ... which is where box g + NB - 1 goes
          }
          uint32_t r[32];
          tmem_ld_32x32(t_row + col_base + c * 32, r);
          tmem_ld_wait();
          if (c == kXNC - 1) release_tmem();
          [[maybe_unused]] float v[32];
          uint64_t v2[16];                                    // gate * (acc + bias) as packed fp32 pairs (FADD2 / FMUL2)
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + n_blk * BN + col_base + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            const float4 gt = gate4(c * 8 + j);
            v2[2 * j + 0] = f2_mul(f2_pack(gt.x, gt.y), f2_add(f2_pack(__uint_as_float(r[4 * j + 0]), __uint_as_float(r[4 * j + 1])), f2_pack(b.x, b.y)));
            v2[2 * j + 1] = f2_mul(f2_pack(gt.z, gt.w), f2_add(f2_pack(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])), f2_pack(b.z, b.w)));
          }
          mbar_wait(&xbar[slot], static_cast<uint32_t>((g / NB) & 1));
          const uint32_t mine = smem_u32(xring + slot * kXBoxBytes) + lane * 128, sw = lane & 7;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = mine + ((j ^ sw) << 4);
            float4 q = lds_f4(a);
            f2_unpack(f2_add(f2_pack(q.x, q.y), v2[2 * j + 0]), q.x, q.y);
            f2_unpack(f2_add(f2_pack(q.z, q.w), v2[2 * j + 1]), q.z, q.w);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(q.x), "f"(q.y), "f"(q.z), "f"(q.w) : "memory");
            if constexpr (kXB) {
              v[4 * j + 0] = q.x; v[4 * j + 1] = q.y; v[4 * j + 2] = q.z; v[4 * j + 3] = q.w;
              rs += (q.x + q.y) + (q.z + q.w);
              rq = fmaf(q.x, q.x, rq); rq = fmaf(q.y, q.y, rq); rq = fmaf(q.z, q.z, rq); rq = fmaf(q.w, q.w, rq);
            }
          }
          fence_proxy_async_smem();                            // generic-proxy writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) {
            int col, row;
            x_coords(g, col, row);
            tma_store_2d(&tma_x, xring + slot * kXBoxBytes, col, row);
            tma_store_commit();
          }
          if constexpr (kXB) {
            // bf16 copy of the updated rows (the next GEMM's A operand): 64 contiguous bytes of this thread's row, straight
            // from registers, behind the box's TMA store.  Measured at M = 36,864: fc2 147 vs 132 us, proj 74-84 vs 54 us; a
            // staging box + TMA store instead costs a pipeline stage and lands at +14 us for both.  Either way the copy eats
            // most of what the folded LayerNorm saves, which is why the fold is opt-in.
            if (row_ok) {
              uint4* xb = reinterpret_cast<uint4*>(p.ln_out + static_cast<long long>(row) * p.ldo + n_blk * BN + col_base + c * 32);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                xb[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                   pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
            }
          }
        }
        if constexpr (kXB) {
          // one (sum, sum of squares) slot per tile column group; with four epilogue warps a warp covers both groups
          if (row_ok) {
            float2* st = p.stats_out + static_cast<long long>(row) * p.stats_slots + n_blk * 2;
            if constexpr (EW == 4) *reinterpret_cast<float4*>(st) = make_float4(rs, rq, 0.f, 0.f);
            else st[(warp - 2) >> 2] = make_float2(rs, rq);
          }
        }
      } else if constexpr (kResid) {
        // x[row, n] += gate[row / tokens, n] * (acc + bias[n])  -  fp32 read-modify-write of the residual stream
        constexpr int NC = kColsPerWarp / 32;
        const int sub = lane >> 3, ch = lane & 7;
        if constexpr (kLN) {
          if (n_blk == 0) { ln_mean = 0.f; ln_m2 = 0.f; ln_cnt = 0; }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          float4 (&cur)[8] = (c & 1) ? xb : xa;
          float4 (&nxt)[8] = (c & 1) ? xa : xb;
          if (c + 1 < NC) load_x(nxt, c + 1);
          uint32_t r[32];
          tmem_ld_32x32(t_row + col_base + c * 32, r);
          tmem_ld_wait();
          if (c == NC - 1) release_tmem();
          float v[32];
          const float4* b4 = reinterpret_cast<const float4*>(bias_smem + c * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = b4[j];
            const float4 g = gate4(c * 8 + j);
            v[4 * j + 0] = g.x * (__uint_as_float(r[4 * j + 0]) + b.x); v[4 * j + 1] = g.y * (__uint_as_float(r[4 * j + 1]) + b.y);
            v[4 * j + 2] = g.z * (__uint_as_float(r[4 * j + 2]) + b.z); v[4 * j + 3] = g.w * (__uint_as_float(r[4 * j + 3]) + b.w);
          }
          stage_f32_tile(stage, v, lane);
          float* ob = reinterpret_cast<float*>(p.out) + n_blk * BN + col_base + c * 32 + 4 * ch;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + sub;
            float4 u = *reinterpret_cast<const float4*>(stage + rr * kStageRowBytes + 16 * ch);
            u.x += cur[it].x; u.y += cur[it].y; u.z += cur[it].z; u.w += cur[it].w;
            if (row0 + rr < out_rows) *reinterpret_cast<float4*>(ob + static_cast<long long>(row0 + rr) * p.ldo) = u;
            if constexpr (kLN) *reinterpret_cast<float4*>(stage + rr * kStageRowBytes + 16 * ch) = u;   // updated values back
          }
          __syncwarp();
          if constexpr (kLN) {
            // row statistics in the row-per-thread view of the staging tile: exact two-pass moments of these 32 values,
            // merged into the running (count, mean, M2) of the row (Chan et al.) - no E[x^2] - mean^2 cancellation
            float w[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 q = *reinterpret_cast<const float4*>(stage + lane * kStageRowBytes + 16 * j);
              w[4 * j] = q.x; w[4 * j + 1] = q.y; w[4 * j + 2] = q.z; w[4 * j + 3] = q.w;
            }
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) { s0 += w[j]; s1 += w[j + 1]; }
            const float mc = (s0 + s1) * (1.0f / 32.0f);
            float q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float d0 = w[j] - mc, d1 = w[j + 1] - mc;
              q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1);
            }
            const float delta = mc - ln_mean;
            const float n_old = static_cast<float>(ln_cnt), n_new = n_old + 32.0f;
            ln_mean = fmaf(delta, 32.0f / n_new, ln_mean);
            ln_m2 += (q0 + q1) + delta * delta * (n_old * 32.0f / n_new);
            ln_cnt += 32;
            __syncwarp();
          }
        }
        if constexpr (kLN) {
          if (n_blk == num_n - 1) {
            // ---- the M-block is complete: LayerNorm-modulate its rows (models.py:19-20,120-121) while they are still in L2
            const int par = (s / num_n) & 1;
            const int half = (warp - 2) >> 2;
            ln_stats[(par * 2 + half) * BM + quad * 32 + lane] = make_float2(ln_mean, ln_m2);
            __threadfence_block();                               // this warp's residual stores before the other warps' reads
            asm volatile("bar.sync 2, %0;" ::"n"(EW * 32) : "memory");
            const int wi = warp - 2;                             // 16 rows per warp
            const float inv_n = 1.0f / static_cast<float>(p.N), half_n = 0.25f * static_cast<float>(p.N);
            long long cur_sample = -1;
            float4 sh[6], sc[6];
#pragma unroll 1
            for (int i = 0; i < BM / EW; ++i) {
              const int rr = wi * (BM / EW) + i;
              const int grow = (m_blk * CS + static_cast<int>(rank)) * BM + rr;
              if (grow >= out_rows) break;
              const float2 a = ln_stats[(par * 2 + 0) * BM + rr], b = ln_stats[(par * 2 + 1) * BM + rr];
              const float dm = b.x - a.x;
              const float mean = 0.5f * (a.x + b.x);
              const float rstd = rsqrtf((a.y + b.y + dm * dm * half_n) * inv_n + 1e-6f);     // two halves of N/2 columns each
              const long long sample = grow / p.tokens;
              if (sample != cur_sample) {
                cur_sample = sample;
                const float4* s4 = reinterpret_cast<const float4*>(p.ln_shift + sample * p.ln_stride);
                const float4* c4 = reinterpret_cast<const float4*>(p.ln_scale + sample * p.ln_stride);
#pragma unroll
                for (int j = 0; j < 6; ++j) { sh[j] = __ldg(s4 + lane + 32 * j); sc[j] = __ldg(c4 + lane + 32 * j); }
              }
              const float4* xr = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.out) + static_cast<long long>(grow) * p.ldo);
              float4 xv[6];
#pragma unroll
              for (int j = 0; j < 6; ++j) xv[j] = __ldcg(xr + lane + 32 * j);
              uint2* yr = reinterpret_cast<uint2*>(p.ln_out + static_cast<long long>(grow) * p.ldo);
#pragma unroll
              for (int j = 0; j < 6; ++j) {
                uint2 o;
                o.x = pack_bf16(fmaf((xv[j].x - mean) * rstd, 1.0f + sc[j].x, sh[j].x), fmaf((xv[j].y - mean) * rstd, 1.0f + sc[j].y, sh[j].y));
                o.y = pack_bf16(fmaf((xv[j].z - mean) * rstd, 1.0f + sc[j].z, sh[j].z), fmaf((xv[j].w - mean) * rstd, 1.0f + sc[j].w, sh[j].w));
                yr[lane + 32 * j] = o;
              }
            }
          }
        }
      } else {
        // fp32 outputs: EPI_BIAS_F32, EPI_PATCH_EMBED_F32, EPI_WGRAD_F32
        float xt_row[kLatent];
        if constexpr (EPI == EPI_PATCH_EMBED_F32) {
#pragma unroll
          for (int d = 0; d < kLatent; ++d) xt_row[d] = 0.f;
          if (row_ok) {
            const float4* x4 = reinterpret_cast<const float4*>(p.xt + static_cast<long long>(row) * kLatent);
            const float4 u = x4[0], w = x4[1];
            xt_row[0] = u.x; xt_row[1] = u.y; xt_row[2] = u.z; xt_row[3] = u.w;
            xt_row[4] = w.x; xt_row[5] = w.y; xt_row[6] = w.z; xt_row[7] = w.w;
          }
        }
#pragma unroll 1
        for (int c = 0; c < kColsPerWarp / 32; ++c) {
          const int n0 = n_blk * BN + col_base + c * 32;
          uint32_t r[32];
          tmem_ld_32x32(t_row + col_base + c * 32, r);
          tmem_ld_wait();
          if (c == kColsPerWarp / 32 - 1) release_tmem();
          float v[32];
          if constexpr (WG) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          } else {
            const float4* b4 = reinterpret_cast<const float4*>(bias_smem + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = b4[j];
              v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b.x; v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
              v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z; v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
            }
          }
          if constexpr (WG) {
            if (p.tma_out) {
              // split contraction: every split ADDS its 32 x 32 fp32 chunk into the gradient through a TMA reduction box -
              // no partial buffers, no reduction pass (the caller zeroes the gradient first)
              const uint32_t sbase = smem_u32(stage);
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
              const uint32_t mine = sbase + lane * 128, sw = ((sbase >> 7) + lane) & 7;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(mine + ((j ^ sw) << 4)), "f"(v[4 * j]), "f"(v[4 * j + 1]),
                             "f"(v[4 * j + 2]), "f"(v[4 * j + 3]) : "memory");
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_reduce_add_2d(&tma_x, stage, n0, row0);
                tma_store_commit();
              }
              continue;
            }
          }
          if constexpr (EPI == EPI_PATCH_EMBED_F32) {
            // + x_t[row, :8] . w_in_t[:, n]  (time_emb_in, models.py:280); 8 FMAs per output, weights broadcast from L1
#pragma unroll
            for (int d = 0; d < kLatent; ++d) {
              const float4* w4 = reinterpret_cast<const float4*>(p.w_in_t + static_cast<long long>(d) * p.N + n0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 w = __ldg(w4 + j);
                v[4 * j + 0] = fmaf(xt_row[d], w.x, v[4 * j + 0]); v[4 * j + 1] = fmaf(xt_row[d], w.y, v[4 * j + 1]);
                v[4 * j + 2] = fmaf(xt_row[d], w.z, v[4 * j + 2]); v[4 * j + 3] = fmaf(xt_row[d], w.w, v[4 * j + 3]);
              }
            }
          }
          stage_f32_tile(stage, v, lane);
          const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + sub;
            float4 u = *reinterpret_cast<const float4*>(stage + rr * kStageRowBytes + 16 * ch);
            const int grow = row0 + rr;
            if (grow < out_rows) {
              if constexpr (EPI == EPI_PATCH_EMBED_F32) {   // + pos_embed[row % tokens]  (coalesced 128-byte segments)
                const float4 q = __ldg(reinterpret_cast<const float4*>(p.pos + static_cast<long long>(grow % p.tokens) * p.N + n0) + ch);
                u.x += q.x; u.y += q.y; u.z += q.z; u.w += q.w;
              }
              float* obase = reinterpret_cast<float*>(p.out);
              if constexpr (WG) obase += static_cast<long long>(split_idx) * p.wg_rows * p.ldo;
              *reinterpret_cast<float4*>(obase + static_cast<long long>(grow) * p.ldo + n0 + 4 * ch) = u;
            }
          }
          __syncwarp();
        }
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    }   // !kLNX
  }

  if constexpr (kXR || kLNX || EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_GELU_GRAD_BF16 || EPI == EPI_WGRAD_F32) {
    if (warp >= 2 && lane == 0) tma_store_wait<0>();          // bulk stores read this CTA's shared memory: drain before exit
  }
  tc_fence_before();
  if constexpr (CS == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    __syncwarp();
    if constexpr (CS == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols); else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------- host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess) {
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// K-major bf16 [rows, cols] with leading dimension ld (elements); box = {64 cols, box_rows}, 128B swizzle.
int make_tmap_bf16_kmajor(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(kErrDriver, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(kErrDriver, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
  return kOk;
}

// fp32 [rows, cols] row-major with leading dimension ld (elements); box = {32 cols, 32 rows} = 128-byte swizzled rows.
static int make_tmap_f32_box32(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(kErrDriver, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(kErrDriver, "cuTensorMapEncodeTiled(fp32 box) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
  return kOk;
}

// bf16 [rows, cols] row-major; box = {32 cols, 32 rows} = 64-byte rows, 64B swizzle (the LayerNorm output boxes).
static int make_tmap_bf16_box32(CUtensorMap* out, const void* base, long long rows, long long cols, long long ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error(kErrDriver, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(kErrDriver, "cuTensorMapEncodeTiled(bf16 box) failed (%d) rows=%lld cols=%lld ld=%lld", (int)r, rows, cols, ld);
  return kOk;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// Epilogue warps per CTA.  The bias / bias+GELU epilogues hand their tiles to TMA store boxes and are best with four
// (qkv 98 vs 101 us, fc1 140 vs 143 us at M = 36,864); the epilogues that still move data per thread (dGELU, fp32 outputs,
// weight gradients) want eight.  JPDVT_GEMM_EPI_WARPS=4|8 forces one value for all of them (A/B knob).
static int epi_warps(int epi) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("JPDVT_GEMM_EPI_WARPS");
    forced = (e == nullptr) ? 0 : (e[0] == '4' ? 4 : 8);
  }
  if (forced) return forced;
  return (epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_BIAS_GELU_GRAD_BF16) ? 4 : 8;
}

template <int BN, int EPI, int EW, int NB = 0, bool BMN = false>
static int launch_cfg(const __nv_bfloat16* a, long long lda, const __nv_bfloat16* w, long long ldw, const GemmParams& p,
                      cudaStream_t stream) {
  constexpr int CS = 2;
  using Cfg = GemmCfg<BN, CS, EW, NB, EPI == EPI_RESID_LN_F32, (EPI == EPI_BIAS_GELU_GRAD_BF16) ? 2 : 1>;
  static_assert(Cfg::kStages >= 3, "pipeline too shallow");
  static_assert(!BMN || (BN / CS) % 64 == 0, "MN-major B boxes are 64 features wide");
  static bool attr_set = false;
  auto kern = gemm_kernel<BN, EPI, CS, EW, NB, BMN>;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes, cudaGetErrorString(cudaGetLastError()));
    attr_set = true;
  }
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_kmajor(&ta, a, p.M, p.K, lda, BM);
  if (rc != kOk) return rc;
  if constexpr (BMN) rc = make_tmap_bf16_kmajor(&tb, w, p.K, p.N, ldw, 64);   // W [K, N]: {64 features x 64 contraction rows} boxes
  else rc = make_tmap_bf16_kmajor(&tb, w, p.N, p.K, ldw, Cfg::kBRows);
  if (rc != kOk) return rc;
  CUtensorMap tx = ta;                                       // residual-stream boxes (TMA residual epilogue only)
  if constexpr (NB > 0) {
    rc = make_tmap_f32_box32(&tx, p.out, p.M, p.N, p.ldo);
    if (rc != kOk) return rc;
  }
  GemmParams pp = p;
  pp.reverse_m = sweep_reverse();
  if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16) {
    static int tma_out = -1;                                 // JPDVT_GEMM_TMA_OUT=0: per-thread coalesced stores instead (A/B knob)
    if (tma_out < 0) { const char* e = getenv("JPDVT_GEMM_TMA_OUT"); tma_out = (e != nullptr && e[0] == '0') ? 0 : 1; }
    pp.tma_out = (tma_out && (p.ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0) ? 1 : 0;
    if (pp.tma_out) {
      rc = make_tmap_bf16_kmajor(&tx, p.out, p.M, p.N, p.ldo, 32);   // {64 cols x 32 rows} boxes, 128-byte swizzle
      if (rc != kOk) return rc;
    }
  }
  CUtensorMap ty = ta;                                       // LayerNorm output boxes (EPI_RESID_LN_TMA_F32 only)
  if constexpr (EPI == EPI_RESID_LN_TMA_F32) {
    rc = make_tmap_bf16_box32(&ty, p.ln_out, p.M, p.N, p.ldo);
    if (rc != kOk) return rc;
  }
  if constexpr (EPI == EPI_BIAS_GELU_GRAD_BF16) {            // training fc1: activation boxes through tx, derivative boxes through ty
    static int tma_out = -1;
    if (tma_out < 0) { const char* e = getenv("JPDVT_GEMM_TMA_OUT"); tma_out = (e != nullptr && e[0] == '0') ? 0 : 1; }
    pp.tma_out = (tma_out && (p.ldo % 8) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && p.out_aux != nullptr &&
                  (reinterpret_cast<uintptr_t>(p.out_aux) & 15) == 0) ? 1 : 0;
    if (pp.tma_out) {
      rc = make_tmap_bf16_kmajor(&tx, p.out, p.M, p.N, p.ldo, 32);
      if (rc != kOk) return rc;
      rc = make_tmap_bf16_kmajor(&ty, p.out_aux, p.M, p.N, p.ldo, 32);
      if (rc != kOk) return rc;
    }
  }
  const int tiles = ((p.M + CS * BM - 1) / (CS * BM)) * (p.N / BN);
  const int max_groups = num_sms() / CS;
  const int groups = tiles < max_groups ? tiles : max_groups;
  if (launch_pdl(kern, dim3(groups * CS), dim3(Cfg::kThreads), Cfg::kSmemBytes, stream, ta, tb, tx, ty, pp) != cudaSuccess)
    return set_error(kErrCuda, "gemm_kernel: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return check_launch("gemm_kernel");
}

template <int BN, int EPI>
static int launch_cs(const __nv_bfloat16* a, long long lda, const __nv_bfloat16* w, long long ldw, const GemmParams& p,
                     cudaStream_t stream) {
  if constexpr (EPI == EPI_HEAD) return launch_cfg<BN, EPI, 4>(a, lda, w, ldw, p, stream);
  else return epi_warps(EPI) == 8 ? launch_cfg<BN, EPI, 8>(a, lda, w, ldw, p, stream) : launch_cfg<BN, EPI, 4>(a, lda, w, ldw, p, stream);
}

// Small problems (serving batches of a few puzzles): a 256x256x64 UMMA step costs the same with 16 valid rows as with 256,
// and N / 256 tiles leave most CTA pairs idle (batch 1: 3 of 74 for proj / fc2, whose 48 k-steps then run back to back on
// those three).  Narrower tiles put four (or two) times as many pairs to work on k-steps a quarter (half) as long; the
// arithmetic per output element - and therefore every result bit - is unchanged.  Returns 0 (keep 256), 128 or 64.
static int small_tile_width(int epi, const GemmParams& p) {
  static int enabled = -1;                                     // JPDVT_GEMM_SMALL_TILES=0: always 256-wide tiles (A/B knob)
  if (enabled < 0) { const char* e = getenv("JPDVT_GEMM_SMALL_TILES"); enabled = (e != nullptr && e[0] == '0') ? 0 : 1; }
  if (!enabled || p.N % 256 != 0) return 0;
  if (epi != EPI_BIAS_BF16 && epi != EPI_BIAS_GELU_BF16 && epi != EPI_RESID_TMA_F32 && epi != EPI_RESID_F32 &&
      epi != EPI_PATCH_EMBED_F32 && epi != EPI_BIAS_BF16_F32) return 0;
  const long long tiles256 = static_cast<long long>((p.M + 2 * BM - 1) / (2 * BM)) * (p.N / 256);
  const int pairs = num_sms() / 2;
  if (tiles256 * 4 <= pairs) return 64;
  if (tiles256 * 2 <= pairs) return 128;
  // Mid-size batches (a few waves of tiles): the time of a launch is waves x tile width.  192-wide tiles (a 256 x 192 x 16
  // UMMA, 512 TMEM columns) fill the last wave where 256-wide ones leave a third of the pairs idle - e.g. 32 puzzles,
  // M = 4,608: proj / fc2 54 -> 72 tiles on 74 pairs, qkv 162 tiles in 3 waves -> 216 tiles in 3 shorter waves.  A tile
  // carries a fixed cost (pipeline fill, epilogue tail), hence the +16; ties go to the wide tile (M = 36,864 keeps 256).
  if (p.N % 192 == 0 && (epi == EPI_BIAS_BF16 || epi == EPI_BIAS_GELU_BF16 || epi == EPI_RESID_TMA_F32)) {
    static int wide192 = -1;                                   // JPDVT_GEMM_TILE192=0: never (A/B knob)
    if (wide192 < 0) { const char* e = getenv("JPDVT_GEMM_TILE192"); wide192 = (e != nullptr && e[0] == '0') ? 0 : 1; }
    const long long tiles192 = static_cast<long long>((p.M + 2 * BM - 1) / (2 * BM)) * (p.N / 192);
    const long long cost256 = ((tiles256 + pairs - 1) / pairs) * (256 + 16), cost192 = ((tiles192 + pairs - 1) / pairs) * (192 + 16);
    if (wide192 && cost192 < cost256) return 192;
  }
  return 0;
}

int launch_gemm(int epi, const __nv_bfloat16* a, long long lda, const __nv_bfloat16* w, long long ldw, const GemmParams& p,
                cudaStream_t stream) {
  if (p.M <= 0) return kOk;
  if (p.K % BK != 0 || p.K <= 0) return set_error(kErrBadArg, "gemm: K=%d must be a positive multiple of %d", p.K, BK);
  if ((lda % 8) != 0 || (ldw % 8) != 0) return set_error(kErrBadArg, "gemm: leading dimensions must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(w) & 15))
    return set_error(kErrBadArg, "gemm: operand pointers must be 16-byte aligned");
  if (p.b_mn) {          // data gradients: dX = dY . W with W as stored ([K, N] here); 256-wide tiles, three epilogues
    if (p.N % 256 != 0 || (ldw % 8) != 0) return set_error(kErrBadArg, "gemm: the MN-major B operand needs N %% 256 == 0 (N=%d)", p.N);
    if (epi == EPI_BIAS_F32) return launch_cfg<256, EPI_BIAS_F32, 8, 0, true>(a, lda, w, ldw, p, stream);
    if (epi == EPI_BIAS_BF16) return launch_cfg<256, EPI_BIAS_BF16, 4, 0, true>(a, lda, w, ldw, p, stream);
    if (epi == EPI_DGELU_BF16) {
      if (p.aux == nullptr) return set_error(kErrBadArg, "gemm: dgelu epilogue needs the pre-activations");
      return launch_cfg<256, EPI_DGELU_BF16, 8, 0, true>(a, lda, w, ldw, p, stream);
    }
    return set_error(kErrBadArg, "gemm: epilogue %d has no MN-major B form", epi);
  }
  const int bn_small = (p.stats_in == nullptr) ? small_tile_width(epi, p) : 0;
  if (bn_small == 192) {                                       // 192-wide tiles: the TMA-store and TMA-ring epilogues as they are
    if (epi == EPI_BIAS_BF16) return launch_cfg<192, EPI_BIAS_BF16, 4>(a, lda, w, ldw, p, stream);
    if (epi == EPI_BIAS_GELU_BF16) return launch_cfg<192, EPI_BIAS_GELU_BF16, 4>(a, lda, w, ldw, p, stream);
    if (p.gate == nullptr || p.tokens <= 0 || (p.ldo % 4) != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15))
      return set_error(kErrBadArg, "gemm: the TMA residual epilogue needs the gate, tokens and a 16-byte aligned residual stream");
    return p.K >= 2048 ? launch_cfg<192, EPI_RESID_TMA_F32, 4, 2>(a, lda, w, ldw, p, stream)
                       : launch_cfg<192, EPI_RESID_TMA_F32, 8, 2>(a, lda, w, ldw, p, stream);
  }
  if (bn_small != 0 && epi == EPI_RESID_TMA_F32) epi = EPI_RESID_F32;     // narrow tiles: register read-modify-write epilogue
  const int bn = (epi == EPI_HEAD) ? 64 : (bn_small != 0 ? bn_small : ((p.N % 256 == 0) ? 256 : 128));
  if (p.N % bn != 0) return set_error(kErrBadArg, "gemm: N=%d is not a multiple of the %d-wide tile", p.N, bn);
  if (p.stats_in != nullptr && ((epi != EPI_BIAS_BF16 && epi != EPI_BIAS_GELU_BF16) || p.fold_u == nullptr || p.stats_slots != 6 ||
                                (reinterpret_cast<uintptr_t>(p.stats_in) & 15)))
    return set_error(kErrBadArg, "gemm: a folded LayerNorm needs the bias / bias+GELU epilogue, fold_u and six 16-byte aligned statistics slots per row");
  if (epi == EPI_HEAD && p.N != 64) return set_error(kErrBadArg, "gemm: head epilogue requires N == 64");
  if ((epi == EPI_GATE_BF16 || epi == EPI_PATCH_EMBED_F32 || epi == EPI_RESID_F32 || epi == EPI_RESID_LN_F32) && p.tokens <= 0) return set_error(kErrBadArg, "gemm: tokens must be positive");
  if ((epi == EPI_RESID_F32 || epi == EPI_RESID_LN_F32) && p.gate == nullptr) return set_error(kErrBadArg, "gemm: residual epilogue needs the gate");
  if (epi == EPI_RESID_TMA_F32) {
    if (p.gate == nullptr || p.tokens <= 0) return set_error(kErrBadArg, "gemm: residual epilogue needs the gate and tokens");
    if (p.N % 256 != 0 || (p.ldo % 4) != 0 || (reinterpret_cast<uintptr_t>(p.out) & 15))
      return set_error(kErrBadArg, "gemm: the TMA residual epilogue needs N %% 256 == 0 and a 16-byte aligned residual stream");
    static int ring = -1, e_forced = 0;                       // JPDVT_RESID_RING: 2 = eight warps x two boxes for every K, 3/4/5 variants
    if (ring < 0) { const char* e = getenv("JPDVT_RESID_RING"); ring = (e != nullptr && e[0] >= '3' && e[0] <= '5') ? e[0] - '0' : 2; e_forced = (e != nullptr) ? 1 : 0; }
    if (ring == 4) return launch_cfg<256, EPI_RESID_TMA_F32, 4, 4>(a, lda, w, ldw, p, stream);   // four warps, four boxes each
    if (ring == 5) return launch_cfg<256, EPI_RESID_TMA_F32, 4, 2>(a, lda, w, ldw, p, stream);   // four warps, two boxes, six stages
    if (ring == 3) return launch_cfg<256, EPI_RESID_TMA_F32, 8, 3>(a, lda, w, ldw, p, stream);
    // default: long contractions (fc2) are MMA-bound - four epilogue warps leave room for a sixth pipeline stage (136.5 vs
    // 139.3 us); short ones (proj) are bound by the residual traffic and want all eight warps moving boxes (56 vs 67 us)
    return (e_forced == 0 && p.K >= 2048) ? launch_cfg<256, EPI_RESID_TMA_F32, 4, 2>(a, lda, w, ldw, p, stream)
                                          : launch_cfg<256, EPI_RESID_TMA_F32, 8, 2>(a, lda, w, ldw, p, stream);
  }
  if (epi == EPI_RESID_TMA_XB_F32) {
    if (p.gate == nullptr || p.tokens <= 0) return set_error(kErrBadArg, "gemm: residual epilogue needs the gate and tokens");
    if (p.N % 256 != 0 || p.ldo != p.N || (reinterpret_cast<uintptr_t>(p.out) & 15) || (reinterpret_cast<uintptr_t>(p.ln_out) & 15))
      return set_error(kErrBadArg, "gemm: the residual + bf16-copy epilogue needs N %% 256 == 0, ldo == N and 16-byte aligned x / copy");
    if (!p.ln_out || !p.stats_out || p.stats_slots != 2 * (p.N / 256))
      return set_error(kErrBadArg, "gemm: residual + bf16-copy epilogue: null copy / statistics pointer or stats_slots != 2 * N / 256");
    // same split as the plain TMA residual epilogue
    return p.K >= 2048 ? launch_cfg<256, EPI_RESID_TMA_XB_F32, 4, 2>(a, lda, w, ldw, p, stream)
                       : launch_cfg<256, EPI_RESID_TMA_XB_F32, 8, 2>(a, lda, w, ldw, p, stream);
  }
  if (epi == EPI_RESID_LN_TMA_F32) {
    if (p.gate == nullptr || p.tokens <= 0) return set_error(kErrBadArg, "gemm: residual epilogue needs the gate and tokens");
    if (p.N != kHidden || p.ldo != kHidden) return set_error(kErrBadArg, "gemm: the LayerNorm-fused epilogue needs N == ldo == %d", kHidden);
    if (!p.ln_out || !p.ln_shift || !p.ln_scale) return set_error(kErrBadArg, "gemm: LayerNorm-fused epilogue: null pointer");
    if ((reinterpret_cast<uintptr_t>(p.out) & 15) || (reinterpret_cast<uintptr_t>(p.ln_out) & 15))
      return set_error(kErrBadArg, "gemm: the TMA LayerNorm epilogue needs 16-byte aligned x and xn");
    return launch_cfg<256, EPI_RESID_LN_TMA_F32, 8, 2>(a, lda, w, ldw, p, stream);
  }
  if (epi == EPI_RESID_LN_F32) {
    if (p.N != kHidden || p.ldo != kHidden) return set_error(kErrBadArg, "gemm: the LayerNorm-fused epilogue needs N == ldo == %d", kHidden);
    if (!p.ln_out || !p.ln_shift || !p.ln_scale) return set_error(kErrBadArg, "gemm: LayerNorm-fused epilogue: null pointer");
    return launch_cfg<256, EPI_RESID_LN_F32, 8>(a, lda, w, ldw, p, stream);     // two warps per lane quadrant split the columns
  }
  if (epi == EPI_BIAS_GELU_GRAD_BF16 && p.out_aux == nullptr) return set_error(kErrBadArg, "gemm: gelu+grad epilogue needs the derivative buffer");
  if (epi == EPI_DGELU_BF16 && p.aux == nullptr) return set_error(kErrBadArg, "gemm: dgelu epilogue needs the pre-activations");
#define JP_CASE(E)                                                             \
  case E:                                                                      \
    return bn == 256 ? launch_cs<256, E>(a, lda, w, ldw, p, stream) : launch_cs<128, E>(a, lda, w, ldw, p, stream);
  if (bn == 64 && epi != EPI_HEAD) {                           // 64-wide tiles: one warp per lane quadrant covers all 64 columns
    if (epi == EPI_BIAS_BF16) return launch_cfg<64, EPI_BIAS_BF16, 4>(a, lda, w, ldw, p, stream);
    if (epi == EPI_BIAS_GELU_BF16) return launch_cfg<64, EPI_BIAS_GELU_BF16, 4>(a, lda, w, ldw, p, stream);
    if (epi == EPI_PATCH_EMBED_F32) return launch_cfg<64, EPI_PATCH_EMBED_F32, 4>(a, lda, w, ldw, p, stream);
    if (epi == EPI_BIAS_BF16_F32) return launch_cfg<64, EPI_BIAS_BF16_F32, 4>(a, lda, w, ldw, p, stream);
    return launch_cfg<64, EPI_RESID_F32, 4>(a, lda, w, ldw, p, stream);
  }
  switch (epi) {
    JP_CASE(EPI_BIAS_BF16)
    JP_CASE(EPI_BIAS_GELU_BF16)
    JP_CASE(EPI_GATE_BF16)
    JP_CASE(EPI_PATCH_EMBED_F32)
    JP_CASE(EPI_BIAS_F32)
    JP_CASE(EPI_BIAS_BF16_F32)
    JP_CASE(EPI_DGELU_BF16)
    JP_CASE(EPI_RESID_F32)
    JP_CASE(EPI_BIAS_GELU_GRAD_BF16)
    case EPI_HEAD:
      return launch_cs<64, EPI_HEAD>(a, lda, w, ldw, p, stream);
    default:
      return set_error(kErrBadArg, "gemm: unknown epilogue %d", epi);
  }
#undef JP_CASE
}

// ------------------------------------------------------------------------------------------- weight gradients
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, long long n4, int split) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 acc = reinterpret_cast<const float4*>(partial)[i];
  for (int s = 1; s < split; ++s) {
    const float4 v = reinterpret_cast<const float4*>(partial)[i + static_cast<long long>(s) * n4];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  reinterpret_cast<float4*>(out)[i] = acc;
}

// Split of the contraction (the M = batch x tokens rows) over work items.  One work item = one 256 x BN tile of dW over
// `kb_per` 64-row blocks; the CTA pairs take items round robin, so the launch lasts rounds x (kb_per + per-item overhead) with
// rounds = ceil(tiles x split / pairs).  The first plan aimed at "about two items per pair" (s = ceil(148 / tiles)) and landed
// between waves: fc1 / fc2 (36 tiles) got 5 splits = 180 items = 2.43 waves on 74 pairs, qkv (27 tiles) 6 splits = 162 items =
// 2.19 waves - a quarter of the SM time idle in the last round (ncu: SMs active 74 % of the launch).  Now every split count
// up to 16 is costed and the cheapest wins (ties: fewer splits = fewer reduction boxes): 36 tiles -> 2 splits (72 items, one
// full wave), 27 tiles -> 8 splits (216 items, 2.92 waves), 9 tiles -> 8 or 16.  JPDVT_WGRAD_SPLIT=<n> forces a count (A/B).
static void wgrad_plan(long long m, int out_rows, int n_cols, int* split, int* split_len) {
  const int bn = (n_cols % 256 == 0) ? 256 : 128;
  const int tiles = ((out_rows + 255) / 256) * (n_cols / bn);
  const int kb_total = static_cast<int>((m + BK - 1) / BK);
  const int pairs = num_sms() / 2;
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("JPDVT_WGRAD_SPLIT"); forced = (e != nullptr) ? atoi(e) : 0; }
  constexpr int kItemOverhead = 2;          // per work item, in k-block times: accumulator hand-over, first operand boxes
  int best_s = 1;
  long long best_cost = -1;
  for (int s = 1; s <= 16 && s <= kb_total; ++s) {
    const int kb_per = (kb_total + s - 1) / s;
    const int n_split = (kb_total + kb_per - 1) / kb_per;
    if (n_split != s) continue;               // the same plan as a smaller s
    const int rounds = (tiles * n_split + pairs - 1) / pairs;
    const long long cost = static_cast<long long>(rounds) * (kb_per + kItemOverhead);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = s; }
  }
  if (forced > 0) best_s = forced < kb_total ? forced : kb_total;
  const int kb_per = (kb_total + best_s - 1) / best_s;
  *split = (kb_total + kb_per - 1) / kb_per;
  *split_len = kb_per * BK;
}

long long wgrad_scratch_floats(long long m, int out_rows, int n_cols) {
  int split, split_len;
  wgrad_plan(m, out_rows, n_cols, &split, &split_len);
  return split > 1 ? static_cast<long long>(split) * out_rows * n_cols : 0;
}

template <int BN>
static int launch_wgrad_cfg(const __nv_bfloat16* pmat, long long ldp, const __nv_bfloat16* qmat, long long ldq,
                            const GemmParams& p, cudaStream_t stream) {
  constexpr int CS = 2, EW = 8;
  using Cfg = GemmCfg<BN, CS, EW>;
  static bool attr_set = false;
  auto kern = gemm_kernel<BN, EPI_WGRAD_F32, CS, EW>;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes, cudaGetErrorString(cudaGetLastError()));
    attr_set = true;
  }
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_kmajor(&ta, pmat, p.M, p.wg_rows, ldp, 64);   // {64 features x 64 contraction rows} boxes
  if (rc != kOk) return rc;
  rc = make_tmap_bf16_kmajor(&tb, qmat, p.M, p.N, ldq, 64);
  if (rc != kOk) return rc;
  CUtensorMap tx = ta;
  if (p.tma_out) {
    rc = make_tmap_f32_box32(&tx, p.out, p.wg_rows, p.N, p.ldo);
    if (rc != kOk) return rc;
  }
  const int tiles = ((p.wg_rows + 255) / 256) * (p.N / BN) * p.split;
  const int max_groups = num_sms() / CS;
  const int groups = tiles < max_groups ? tiles : max_groups;
  kern<<<groups * CS, Cfg::kThreads, Cfg::kSmemBytes, stream>>>(ta, tb, tx, ta, p);
  return check_launch("gemm_kernel<wgrad>");
}

// out_zeroed: the caller has already zero-filled `out` on this stream (the training backward's flat gradient buffer is cleared
// once per step), so the split path need not clear it again before its reduction boxes add into it
int launch_wgrad(const __nv_bfloat16* pmat, long long ldp, const __nv_bfloat16* qmat, long long ldq, float* out,
                 float* partial, long long m, int out_rows, int n_cols, cudaStream_t stream, bool out_zeroed) {
  if (out_rows <= 0 || n_cols <= 0) return kOk;
  if (m <= 0) return set_error(kErrBadArg, "wgrad: empty contraction");
  if (n_cols % 128 != 0) return set_error(kErrBadArg, "wgrad: n_cols=%d must be a multiple of 128", n_cols);
  if ((ldp % 8) || (ldq % 8) || (out_rows % 8)) return set_error(kErrBadArg, "wgrad: leading dimensions / out_rows must be multiples of 8");
  if (m > 0x7fffffffLL) return set_error(kErrUnsupported, "wgrad: contraction too long");
  GemmParams p{};
  p.M = static_cast<int>(m); p.N = n_cols; p.K = BK; p.tokens = 1;
  p.wg_rows = out_rows; p.ldo = n_cols;
  wgrad_plan(m, out_rows, n_cols, &p.split, &p.split_len);
  static int tma_red = -1;          // JPDVT_WGRAD_TMA_REDUCE=0: split partials to scratch + a reduction pass instead (A/B knob)
  if (tma_red < 0) { const char* e = getenv("JPDVT_WGRAD_TMA_REDUCE"); tma_red = (e != nullptr && e[0] == '0') ? 0 : 1; }
  if (p.split > 1 && tma_red && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (n_cols % 4) == 0) {
    // splits accumulate straight into the gradient with TMA reduction boxes
    if (!out_zeroed && cudaMemsetAsync(out, 0, static_cast<size_t>(out_rows) * n_cols * sizeof(float), stream) != cudaSuccess)
      return set_error(kErrCuda, "wgrad: cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
    p.out = out;
    p.tma_out = 1;
    return (n_cols % 256 == 0) ? launch_wgrad_cfg<256>(pmat, ldp, qmat, ldq, p, stream)
                               : launch_wgrad_cfg<128>(pmat, ldp, qmat, ldq, p, stream);
  }
  if (p.split > 1 && partial == nullptr) return set_error(kErrBadArg, "wgrad: scratch buffer required (split=%d)", p.split);
  p.out = (p.split > 1) ? partial : out;
  int rc = (n_cols % 256 == 0) ? launch_wgrad_cfg<256>(pmat, ldp, qmat, ldq, p, stream)
                               : launch_wgrad_cfg<128>(pmat, ldp, qmat, ldq, p, stream);
  if (rc != kOk) return rc;
  if (p.split > 1) {
    const long long n4 = static_cast<long long>(out_rows) * n_cols / 4;
    wgrad_reduce_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(partial, out, n4, p.split);
    return check_launch("wgrad_reduce_kernel");
  }
  return kOk;
}

}  // namespace jp
