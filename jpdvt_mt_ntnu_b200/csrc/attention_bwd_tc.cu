// tcgen05 attention backward for the JPDVT piece tokens (training): dQ, dK, dV of softmax(Q K^T / 8) V per (sample, head),
// all five contractions on the 5th-generation tensor cores with the score-sized tiles living in TMEM.
//
// Replaces the autograd of timm Attention.forward's F.scaled_dot_product_attention (image_model/models.py:108,120, reached
// from loss.backward() in train_JPDVT.py:369) for T = 144 (192 px, BASELINE configs[2]); the mma.sync kernel in attention.cu
// keeps the other sizes.  Flash-attention style: the probabilities are recomputed from the forward's log-sum-exp.
//
//   S^T  = K Q^T          P^T  = exp2(S^T * log2(e)/8 - lse2[q])            (rows = keys, columns = queries)
//   dP^T = V dO^T         dS^T = P^T * (dP^T - D[q]) / 8,   D[q] = sum_d dO[q,d] O[q,d]
//   dV   = P^T dO         dK   = dS^T Q          dQ = dS K
//
// Everything is kept in the TRANSPOSED orientation (TMEM lane = key): P^T and dS^T then sit in shared memory as K-major
// [key][query] tiles, which is what dV and dK want as their A operand, and dQ reads the same dS^T tile MN-major - no
// transposition anywhere.  One CTA per SM walks (sample, head) units:
//   warps 0-3 : D and lse2 of the unit, then thread = key row: S^T / dP^T out of TMEM -> P^T, dS^T (bf16) into shared memory;
//               later the epilogue: dV / dK / dQ accumulators -> bf16 -> coalesced rows of dqkv [B*T, 2304]
//   warp 4    : TMA producer - Q, K, V of the head out of the fused QKV activation, dO and O out of [B*T, 768]
//   warp 5    : MMA issuer
// T = 144 = 128 + 16.  Rows (keys) 128..143 are NOT given a second 128-row pass in the transposed orientation (one warp would
// do all its softmax work): their scores come from two small MMAs in the UNtransposed orientation, S[q, 128:144] = Q K_tail^T
// (query rows 0..127, and again with the query rows shifted by 16 so that lanes 112..127 hold queries 128..143), so every
// thread handles one query row x 16 keys and writes its 16 probabilities transposed (2-byte stores, contiguous per warp).
// The 16-row remainder of each OUTPUT is a second MMA pass over the A operand shifted by 16 rows (dV, dK: K-major rows
// [16,144), valid lanes 112..127) or over query blocks 2,3 of the MN-major dS^T tile (dQ: valid lanes 0..15).
// TMEM (512 columns): phase 1  S^T [0,144) dP^T [144,288) tails [288,352);  phase 2 (aliases phase 1, consumed by then)
// dV0 dV1 dK0 dK1 dQ0 dQ1, 64 columns each.
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace jp {

namespace {

constexpr int kBtThreads = 192;
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
__device__ __forceinline__ void sts_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 u;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w) : "r"(addr));
  return u;
}
__device__ __forceinline__ uint16_t bf16_bits(float v) {
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// tcgen05.mma from split descriptor words (see attention_tc.cu): low word = start address (+ LBO), high word shared
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);        // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo_k(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
}
template <bool ACC>
__device__ __forceinline__ void umma_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "n"(ACC ? 1 : 0), "r"(kDescHi)
      : "memory");
}

template <int T>
struct BtCfg {
  static_assert(T == 144, "the tcgen05 attention backward is laid out for 128 + 16 tokens");
  static constexpr int kTile = T * 128;                      // one of Q / K / V / dO / O: T rows x 64 bf16, 128-byte swizzled
  static constexpr int kQBlocks = (T + 63) / 64;             // 64-query blocks of the P^T / dS^T tiles
  static constexpr int kBlk = T * 128;                       // one block: T key rows x 128 B
  static constexpr int kOffQ = 0, kOffK = kTile, kOffV = 2 * kTile, kOffdO = 3 * kTile, kOffO = 4 * kTile;
  static constexpr int kOffdS = 5 * kTile;                   // dS^T first: dQ's second pass reads "blocks 2, 3" of it, block 3
  static constexpr int kOffP = kOffdS + kQBlocks * kBlk;     // being the first block of P^T (finite garbage in discarded rows)
  static constexpr int kOffStage = kOffP + kQBlocks * kBlk;  // 4 x 4 KB epilogue staging
  static constexpr int kOffL = kOffStage + 4 * 4096;         // lse2[T], D[T] fp32
  static constexpr int kBarOff = kOffL + 2 * ((T * 4 + 127) / 128 * 128);
  static constexpr int kSmemBytes = kBarOff + 128 + 1024;    // + alignment slack (the kernel adds 768 B of static shared memory)
  static constexpr int kInBytes = 5 * kTile;
  static_assert(kTile % 1024 == 0 && kBlk % 1024 == 0 && kOffP % 1024 == 0 && kOffStage % 1024 == 0, "swizzle atoms");
  static_assert(kSmemBytes + 1024 <= 227 * 1024, "shared memory");
  // TMEM columns
  static constexpr int kColS = 0, kColdP = T, kColSt0 = 2 * T, kColSt1 = 2 * T + 16, kColdPt0 = 2 * T + 32, kColdPt1 = 2 * T + 48;
  static constexpr int kColdV0 = 0, kColdV1 = 64, kColdK0 = 128, kColdK1 = 192, kColdQ0 = 256, kColdQ1 = 320;
  static_assert(2 * T + 64 <= 512, "TMEM columns");
};

// 32 accumulator rows of this warp (TMEM lane = row, 64 fp32 columns) -> bf16 -> global rows of `ld` elements.  The 32 x 128 B
// tile is transposed through a 4 KB staging tile (XOR-swizzled 16-byte chunks) so that every store instruction writes four
// full 128-byte rows.  Rows r with lo <= r < hi are live.
// colsum_s != 0: the live rows' bf16 values are also added, column by column, into 64 floats of shared memory (the bias
// gradient of the qkv Linear is the column sum of dqkv; folding it here saves a pass over the 85 MB the kernel has just written).
__device__ __forceinline__ void store_acc_rows(uint32_t t_row, uint32_t stage, __nv_bfloat16* dst_row0, long long ld, int lo, int hi,
                                               int lane, uint32_t colsum_s = 0) {
  uint32_t a[32], b[32];
  tmem_ld_32x32(t_row, a);
  tmem_ld_32x32(t_row + 32, b);
  tmem_ld_wait();
  const uint32_t mine = stage + lane * 128, sw = lane & 7;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sts_u4(mine + ((j ^ sw) << 4),
           make_uint4(pack_bf16(__uint_as_float(a[8 * j]), __uint_as_float(a[8 * j + 1])),
                      pack_bf16(__uint_as_float(a[8 * j + 2]), __uint_as_float(a[8 * j + 3])),
                      pack_bf16(__uint_as_float(a[8 * j + 4]), __uint_as_float(a[8 * j + 5])),
                      pack_bf16(__uint_as_float(a[8 * j + 6]), __uint_as_float(a[8 * j + 7]))));
    sts_u4(mine + (((4 + j) ^ sw) << 4),
           make_uint4(pack_bf16(__uint_as_float(b[8 * j]), __uint_as_float(b[8 * j + 1])),
                      pack_bf16(__uint_as_float(b[8 * j + 2]), __uint_as_float(b[8 * j + 3])),
                      pack_bf16(__uint_as_float(b[8 * j + 4]), __uint_as_float(b[8 * j + 5])),
                      pack_bf16(__uint_as_float(b[8 * j + 6]), __uint_as_float(b[8 * j + 7]))));
  }
  __syncwarp();
  const int sub = lane >> 3, ch = lane & 7;
  float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + sub;
    const uint4 u = lds_u4(stage + r * 128 + ((ch ^ (r & 7)) << 4));
    if (r >= lo && r < hi) {
      *reinterpret_cast<uint4*>(dst_row0 + static_cast<long long>(r) * ld + ch * 8) = u;
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { cs[2 * e] += __uint_as_float(w[e] << 16); cs[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u); }
    }
  }
  if (colsum_s != 0) {                   // lanes (sub, ch) -> sum over sub (4 lanes), then one shared-memory add per column
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 8);
      cs[e] += __shfl_xor_sync(0xffffffffu, cs[e], 16);
    }
    if (sub == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(colsum_s + static_cast<uint32_t>(ch * 8 + e) * 4u), "f"(cs[e]) : "memory");
    }
  }
  __syncwarp();
}

template <int T>
__global__ void __launch_bounds__(kBtThreads, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_o, const float* __restrict__ lse2, __nv_bfloat16* __restrict__ dqkv,
                        float* __restrict__ dbias, int num_units) {
  using Cfg = BtCfg<T>;
  extern __shared__ uint8_t att_bt_smem[];
  uint8_t* smem = att_bt_smem + ((1024u - (smem_u32(att_bt_smem) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* in_full = bars + 0;        // TMA: Q, K, V, dO, O of the unit landed
  uint64_t* s_full = bars + 1;         // MMA: S^T, dP^T and the tails are in TMEM
  uint64_t* p_full = bars + 2;         // math warps: P^T, dS^T are in shared memory, phase-1 TMEM columns consumed (4 arrivals)
  uint64_t* o_full = bars + 3;         // MMA: dV, dK, dQ are in TMEM; every operand tile of the unit has been read
  uint64_t* epi_done = bars + 4;       // math warps: accumulators have left TMEM (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);
  float* sL = reinterpret_cast<float*>(smem + Cfg::kOffL);
  float* sD = sL + (T * 4 + 127) / 128 * 32;
  __shared__ float s_colsum[3 * kHeadDim];     // dQ | dK | dV column sums of the current unit (dbias != null)
  if (threadIdx.x < 3 * kHeadDim) s_colsum[threadIdx.x] = 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(in_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(epi_done, 4);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tm_qkv); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_o); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV),
                 sdO = smem_u32(smem + Cfg::kOffdO), sO = smem_u32(smem + Cfg::kOffO), sdS = smem_u32(smem + Cfg::kOffdS),
                 sP = smem_u32(smem + Cfg::kOffP);

  if (warp == 4) {
    // ---------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int b = unit / kHeads, h = unit - b * kHeads;
        if (it > 0) mbar_wait_backoff(o_full, static_cast<uint32_t>((it - 1) & 1), 64);   // the previous unit's MMAs have read every tile
        mbar_expect_tx(in_full, Cfg::kInBytes);
        tma_load_2d(&tm_qkv, in_full, smem + Cfg::kOffK, kHidden + h * kHeadDim, b * T);
        tma_load_2d(&tm_qkv, in_full, smem + Cfg::kOffQ, h * kHeadDim, b * T);
        tma_load_2d(&tm_qkv, in_full, smem + Cfg::kOffV, 2 * kHidden + h * kHeadDim, b * T);
        tma_load_2d(&tm_do, in_full, smem + Cfg::kOffdO, h * kHeadDim, b * T);
        tma_load_2d(&tm_o, in_full, smem + Cfg::kOffO, h * kHeadDim, b * T);
        // the next unit's five tiles towards L2: its loads can only be issued when this unit's last MMA has read the buffers
        if (unit + static_cast<int>(gridDim.x) < num_units) {
          const int nu = unit + static_cast<int>(gridDim.x), nb = nu / kHeads, nh = nu - nb * kHeads;
          tma_prefetch_2d(&tm_qkv, kHidden + nh * kHeadDim, nb * T);
          tma_prefetch_2d(&tm_qkv, nh * kHeadDim, nb * T);
          tma_prefetch_2d(&tm_qkv, 2 * kHidden + nh * kHeadDim, nb * T);
          tma_prefetch_2d(&tm_do, nh * kHeadDim, nb * T);
          tma_prefetch_2d(&tm_o, nh * kHeadDim, nb * T);
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, T);                 // S^T / dP^T: M = 128 keys, N = T queries
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, 16);                // tails: M = 128 queries, N = 16 keys
      constexpr uint32_t idesc_kv = umma_idesc_bf16(128, kHeadDim, 0, 1);   // dV, dK: A K-major, B (dO / Q) MN-major
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, kHeadDim, 1, 1);    // dQ: A (dS^T) and B (K) both MN-major
      const uint32_t q_lo = desc_lo_k(sQ), k_lo = desc_lo_k(sK), v_lo = desc_lo_k(sV), do_lo = desc_lo_k(sdO);
      const uint32_t p_lo = desc_lo_k(sP), ds_lo = desc_lo_k(sdS);
      const uint32_t do_mn = desc_lo_mn(sdO, 8192), q_mn = desc_lo_mn(sQ, 8192), k_mn = desc_lo_mn(sK, 8192);
      const uint32_t ds_mn = desc_lo_mn(sdS, Cfg::kBlk);                    // next 64 queries: one block further
      constexpr uint32_t kBlkW = Cfg::kBlk / 16;                            // block stride in descriptor units (16 B)
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const uint32_t ph = static_cast<uint32_t>(it & 1);
        mbar_wait(in_full, ph);
        if (it > 0) mbar_wait(epi_done, static_cast<uint32_t>((it - 1) & 1));   // the previous unit's accumulators have left TMEM
        tc_fence_after();
        // ---- phase 1: scores and their gradient
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          if (k == 0) umma_lohi<false>(tmem_base + Cfg::kColS, k_lo, q_lo, idesc_s);
          else umma_lohi<true>(tmem_base + Cfg::kColS, k_lo + 2 * k, q_lo + 2 * k, idesc_s);
        }
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          if (k == 0) umma_lohi<false>(tmem_base + Cfg::kColdP, v_lo, do_lo, idesc_s);
          else umma_lohi<true>(tmem_base + Cfg::kColdP, v_lo + 2 * k, do_lo + 2 * k, idesc_s);
        }
        // tails, untransposed: [query rows] x keys 128..143; pass 0 = queries 0..127, pass 1 = queries 16..143 (row shift 16)
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const uint32_t sh = pass * 16 * 8;                                // 16 rows x 128 B in descriptor units
          const uint32_t cs = pass ? Cfg::kColSt1 : Cfg::kColSt0, cp = pass ? Cfg::kColdPt1 : Cfg::kColdPt0;
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
            if (k == 0) umma_lohi<false>(tmem_base + cs, q_lo + sh, k_lo + 128 * 8, idesc_t);
            else umma_lohi<true>(tmem_base + cs, q_lo + sh + 2 * k, k_lo + 128 * 8 + 2 * k, idesc_t);
          }
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
            if (k == 0) umma_lohi<false>(tmem_base + cp, do_lo + sh, v_lo + 128 * 8, idesc_t);
            else umma_lohi<true>(tmem_base + cp, do_lo + sh + 2 * k, v_lo + 128 * 8 + 2 * k, idesc_t);
          }
        }
        umma_commit(s_full);
        // ---- phase 2: the three gradients (contraction over all T queries / keys: T / 16 k-steps)
        mbar_wait(p_full, ph);
        tc_fence_after();
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {                              // pass 1: A rows shifted by 16 -> keys 16..143
          const uint32_t sh = pass * 16 * 8;
#pragma unroll
          for (int j = 0; j < T / 16; ++j) {
            const uint32_t a = p_lo + sh + (j >> 2) * kBlkW + (j & 3) * 2, bq = do_mn + j * 128;
            if (j == 0) umma_lohi<false>(tmem_base + (pass ? Cfg::kColdV1 : Cfg::kColdV0), a, bq, idesc_kv);
            else umma_lohi<true>(tmem_base + (pass ? Cfg::kColdV1 : Cfg::kColdV0), a, bq, idesc_kv);
          }
#pragma unroll
          for (int j = 0; j < T / 16; ++j) {
            const uint32_t a = ds_lo + sh + (j >> 2) * kBlkW + (j & 3) * 2, bq = q_mn + j * 128;
            if (j == 0) umma_lohi<false>(tmem_base + (pass ? Cfg::kColdK1 : Cfg::kColdK0), a, bq, idesc_kv);
            else umma_lohi<true>(tmem_base + (pass ? Cfg::kColdK1 : Cfg::kColdK0), a, bq, idesc_kv);
          }
        }
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {                              // pass 1: query blocks 2, 3 (queries 128..)
#pragma unroll
          for (int j = 0; j < T / 16; ++j) {
            const uint32_t a = ds_mn + pass * 2 * kBlkW + j * 128, bq = k_mn + j * 128;
            if (j == 0) umma_lohi<false>(tmem_base + (pass ? Cfg::kColdQ1 : Cfg::kColdQ0), a, bq, idesc_q);
            else umma_lohi<true>(tmem_base + (pass ? Cfg::kColdQ1 : Cfg::kColdQ0), a, bq, idesc_q);
          }
        }
        umma_commit(o_full);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- math + epilogue warps
    constexpr float sl2 = 0.125f * 1.4426950408889634f;       // head_dim^-0.5 * log2(e)
    const int tid = threadIdx.x;                              // 0..127 = TMEM lane
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t stage = smem_u32(smem + Cfg::kOffStage) + static_cast<uint32_t>(warp) * 4096u;
    int it = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const uint32_t ph = static_cast<uint32_t>(it & 1);
      const int b = unit / kHeads, h = unit - b * kHeads;
      const float* lrow = lse2 + (static_cast<long long>(b) * kHeads + h) * T;
      mbar_wait(in_full, ph);
      // ---- D[q] = sum_d dO[q, d] O[q, d] and lse2[q] for the unit's T queries (thread = query row; 16 threads take two)
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {
        const int q = tid + rep * 128;
        if (q < T) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t off = static_cast<uint32_t>(q) * 128u + static_cast<uint32_t>((j ^ (q & 7)) << 4);
            const uint4 a = lds_u4(sdO + off), c = lds_u4(sO + off);
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              acc = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(cw[e] << 16), acc);
              acc = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(cw[e] & 0xffff0000u), acc);
            }
          }
          sD[q] = acc;
          sL[q] = __ldg(lrow + q);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      mbar_wait(s_full, ph);
      tc_fence_after();
      // ---- main tile: thread = key row `tid`, columns = queries; 16 columns per step, the next step's loads in flight
      {
        uint32_t sa[16], da[16], sb[16], db[16];
        tmem_ld_32x16(t_lane + Cfg::kColS, sa);
        tmem_ld_32x16(t_lane + Cfg::kColdP, da);
        tmem_ld_wait();
        const uint32_t prow = sP + static_cast<uint32_t>(tid) * 128u, dsrow = sdS + static_cast<uint32_t>(tid) * 128u;
        const int sw = tid & 7;
#pragma unroll
        for (int c = 0; c < T / 16; ++c) {
          uint32_t (&s_cur)[16] = (c & 1) ? sb : sa;
          uint32_t (&d_cur)[16] = (c & 1) ? db : da;
          uint32_t (&s_nxt)[16] = (c & 1) ? sa : sb;
          uint32_t (&d_nxt)[16] = (c & 1) ? da : db;
          if (c + 1 < T / 16) {
            tmem_ld_32x16(t_lane + Cfg::kColS + 16 * (c + 1), s_nxt);
            tmem_ld_32x16(t_lane + Cfg::kColdP + 16 * (c + 1), d_nxt);
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {                        // 8 queries -> one 16-byte chunk of the row
            const int q0 = 16 * c + 8 * g;
            const float4 l0 = *reinterpret_cast<const float4*>(sL + q0), l1 = *reinterpret_cast<const float4*>(sL + q0 + 4);
            const float4 e0 = *reinterpret_cast<const float4*>(sD + q0), e1 = *reinterpret_cast<const float4*>(sD + q0 + 4);
            const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
            const float dv[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
            float p[8], ds[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              p[e] = ex2f(fmaf(__uint_as_float(s_cur[8 * g + e]), sl2, -lv[e]));
              ds[e] = p[e] * (__uint_as_float(d_cur[8 * g + e]) - dv[e]) * 0.125f;
            }
            const int chunk = q0 >> 3;                          // 16-byte chunk index along the row of T queries
            const uint32_t off = static_cast<uint32_t>(chunk >> 3) * Cfg::kBlk + static_cast<uint32_t>(((chunk & 7) ^ sw) << 4);
            sts_u4(prow + off, make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7])));
            sts_u4(dsrow + off, make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7])));
          }
          if (c + 1 < T / 16) tmem_ld_wait();
        }
      }
      // ---- tails: thread = query row, 16 keys (128..143); written transposed into rows 128..143 of P^T / dS^T
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int q = pass ? tid + 16 : tid;                    // pass 1 (rows shifted by 16): lanes 112..127 hold queries 128..143
        if (pass == 0 || warp == 3) {                            // warp-uniform: tcgen05.ld is a .sync.aligned instruction
          uint32_t st[16], dt[16];
          tmem_ld_32x16(t_lane + (pass ? Cfg::kColSt1 : Cfg::kColSt0), st);
          tmem_ld_32x16(t_lane + (pass ? Cfg::kColdPt1 : Cfg::kColdPt0), dt);
          tmem_ld_wait();
          if (pass == 0 || lane >= 16) {                         // pass 1: only lanes 112..127 hold new rows (queries 128..143)
            const float lq = sL[q], dq = sD[q];
            const uint32_t col = static_cast<uint32_t>(q >> 6) * Cfg::kBlk + static_cast<uint32_t>((q & 7) * 2);
            const int qc = (q & 63) >> 3;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float p = ex2f(fmaf(__uint_as_float(st[j]), sl2, -lq));
              const float ds = p * (__uint_as_float(dt[j]) - dq) * 0.125f;
              const uint32_t off = static_cast<uint32_t>(128 + j) * 128u + static_cast<uint32_t>((qc ^ (j & 7)) << 4) + col;
              sts_u16(sP + off, bf16_bits(p));
              sts_u16(sdS + off, bf16_bits(ds));
            }
          }
        }
      }
      fence_proxy_async_smem();                                 // generic-proxy stores -> visible to the tensor core's reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- epilogue: accumulators -> bf16 rows of dqkv (dQ | dK | dV column groups of the head)
      __nv_bfloat16* base = dqkv + static_cast<long long>(b) * T * kQkvCols + h * kHeadDim;
      mbar_wait(o_full, ph);
      tc_fence_after();
      __nv_bfloat16* row0 = base + static_cast<long long>(warp * 32) * kQkvCols;
      const uint32_t csq = dbias != nullptr ? smem_u32(s_colsum) : 0u;
      const uint32_t csk = csq ? csq + kHeadDim * 4 : 0u, csv = csq ? csq + 2 * kHeadDim * 4 : 0u;
      store_acc_rows(t_lane + Cfg::kColdV0, stage, row0 + 2 * kHidden, kQkvCols, 0, 32, lane, csv);
      store_acc_rows(t_lane + Cfg::kColdK0, stage, row0 + kHidden, kQkvCols, 0, 32, lane, csk);
      store_acc_rows(t_lane + Cfg::kColdQ0, stage, row0, kQkvCols, 0, 32, lane, csq);
      if (warp == 3) {        // shifted pass: lane 112 + i holds key 128 + i, i.e. row (96 + 16) + r for this warp's lane r >= 16
        __nv_bfloat16* r16 = base + static_cast<long long>(112) * kQkvCols;
        store_acc_rows(t_lane + Cfg::kColdV1, stage, r16 + 2 * kHidden, kQkvCols, 16, 32, lane, csv);
        store_acc_rows(t_lane + Cfg::kColdK1, stage, r16 + kHidden, kQkvCols, 16, 32, lane, csk);
      }
      if (warp == 0) {        // query blocks 2, 3: lane i holds query 128 + i
        store_acc_rows(t_lane + Cfg::kColdQ1, stage, base + static_cast<long long>(128) * kQkvCols, kQkvCols, 0, 16, lane, csq);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_done);
      if (dbias != nullptr) {
        // the unit's 3 x 64 column sums -> dbias[{q,k,v} * 768 + h * 64 + c]: six coalesced red.adds by warp 0, which also
        // re-zeroes the accumulators (the bar.sync at the top of the next unit orders that against the next adds)
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (warp == 0) {
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const int idx = k * 32 + lane;
            atomicAdd(dbias + (idx >> 6) * kHidden + h * kHeadDim + (idx & 63), s_colsum[idx]);
            s_colsum[idx] = 0.f;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// T = 256 (4x4 puzzles @256 px, BASELINE configs[3]): the score-sized tiles no longer fit shared memory next to the five
// operand tiles, so a unit is walked as 2 key blocks x 2 query blocks of 128 x 128:
//   per (kb, qb):  S^T = K[kb] Q[qb]^T, dP^T = V[kb] dO[qb]^T   -> TMEM columns [0,128), [128,256)
//                  math warps (thread = key row): P^T, dS^T (bf16, K-major [key][query]) -> shared memory
//                  dV[kb] += P^T dO[qb],  dK[kb] += dS^T Q[qb],  dQ[qb] += dS K[kb]      -> TMEM [256,320) [320,384) [384+64 qb, ..)
// dV / dK of a key block leave TMEM after its second query block, the two dQ tiles at the end of the unit.  P^T lives where O
// was (O is only needed for D[q] = sum_d dO O at the top of the unit).  The score MMAs of tile n+1 queue behind the gradient
// MMAs of tile n, so they are done by the time the math warps come back for them.
struct Bt256 {
  static constexpr int T = 256;
  static constexpr int kTile = T * 128;                       // 32 KB: one of Q / K / V / dO / O
  static constexpr int kBlk = 128 * 128;                      // 64 queries of the 128-key P^T / dS^T tile
  static constexpr int kOffQ = 0, kOffK = kTile, kOffV = 2 * kTile, kOffdO = 3 * kTile, kOffO = 4 * kTile;
  static constexpr int kOffP = kOffO;                         // aliases O
  static constexpr int kOffdS = 5 * kTile;
  static constexpr int kOffStage = kOffdS + 2 * kBlk;         // 4 x 4 KB epilogue staging
  static constexpr int kOffL = kOffStage + 4 * 4096;          // lse2[T], D[T]
  static constexpr int kBarOff = kOffL + 2 * T * 4;
  static constexpr int kSmemBytes = kBarOff + 128 + 1024;
  static constexpr int kInBytes = 5 * kTile;
  static constexpr int kColS = 0, kColdP = 128, kColdV = 256, kColdK = 320, kColdQ = 384;
  static_assert(kSmemBytes + 1024 <= 227 * 1024, "shared memory");
};

__global__ void __launch_bounds__(kBtThreads, 1)
attention_bwd_tc256_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                           const __grid_constant__ CUtensorMap tm_o, const float* __restrict__ lse2, __nv_bfloat16* __restrict__ dqkv,
                           float* __restrict__ dbias, int num_units) {
  using Cfg = Bt256;
  constexpr int T = Cfg::T;
  extern __shared__ uint8_t att_bt_smem[];
  uint8_t* smem = att_bt_smem + ((1024u - (smem_u32(att_bt_smem) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* in_full = bars + 0;        // TMA: the unit's five tiles landed                                  (1 / unit)
  uint64_t* s_full = bars + 1;         // MMA: S^T, dP^T of a tile are in TMEM                                (4 / unit)
  uint64_t* p_full = bars + 2;         // math: P^T, dS^T of a tile are in shared memory, S^T / dP^T consumed (4 / unit, 4 arrivals)
  uint64_t* g_done = bars + 3;         // MMA: the gradient MMAs of a tile have read P^T / dS^T               (4 / unit)
  uint64_t* kv_drained = bars + 4;     // math: dV / dK of a key block have left TMEM                         (2 / unit, 4 arrivals)
  uint64_t* epi_done = bars + 5;       // math: dQ has left TMEM                                              (1 / unit, 4 arrivals)
  uint64_t* unit_done = bars + 6;      // MMA: every MMA of the unit has read its operand tiles               (1 / unit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  float* sL = reinterpret_cast<float*>(smem + Cfg::kOffL);
  float* sD = sL + T;
  __shared__ float s_colsum[3 * kHeadDim];
  if (threadIdx.x < 3 * kHeadDim) s_colsum[threadIdx.x] = 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(in_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(g_done, 1); mbar_init(kv_drained, 4);
    mbar_init(epi_done, 4); mbar_init(unit_done, 1);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tm_qkv); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_o); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV),
                 sdO = smem_u32(smem + Cfg::kOffdO), sO = smem_u32(smem + Cfg::kOffO), sdS = smem_u32(smem + Cfg::kOffdS),
                 sP = smem_u32(smem + Cfg::kOffP);

  if (warp == 4) {
    // ---------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int b = unit / kHeads, h = unit - b * kHeads;
        if (it > 0) mbar_wait_backoff(unit_done, static_cast<uint32_t>((it - 1) & 1), 64);
        mbar_expect_tx(in_full, Cfg::kInBytes);
        tma_load_2d(&tm_qkv, in_full, smem + Cfg::kOffK, kHidden + h * kHeadDim, b * T);
        tma_load_2d(&tm_qkv, in_full, smem + Cfg::kOffQ, h * kHeadDim, b * T);
        tma_load_2d(&tm_qkv, in_full, smem + Cfg::kOffV, 2 * kHidden + h * kHeadDim, b * T);
        tma_load_2d(&tm_do, in_full, smem + Cfg::kOffdO, h * kHeadDim, b * T);
        tma_load_2d(&tm_o, in_full, smem + Cfg::kOffO, h * kHeadDim, b * T);
        if (unit + static_cast<int>(gridDim.x) < num_units) {       // the next unit's tiles towards L2
          const int nu = unit + static_cast<int>(gridDim.x), nb = nu / kHeads, nh = nu - nb * kHeads;
          tma_prefetch_2d(&tm_qkv, kHidden + nh * kHeadDim, nb * T);
          tma_prefetch_2d(&tm_qkv, nh * kHeadDim, nb * T);
          tma_prefetch_2d(&tm_qkv, 2 * kHidden + nh * kHeadDim, nb * T);
          tma_prefetch_2d(&tm_do, nh * kHeadDim, nb * T);
          tma_prefetch_2d(&tm_o, nh * kHeadDim, nb * T);
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);               // S^T / dP^T: M = 128 keys, N = 128 queries
      constexpr uint32_t idesc_kv = umma_idesc_bf16(128, kHeadDim, 0, 1);   // dV, dK: A K-major, B (dO / Q) MN-major
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, kHeadDim, 1, 1);    // dQ: A (dS^T) and B (K) both MN-major
      constexpr uint32_t kBlkW = Cfg::kBlk / 16;                            // P^T / dS^T block stride in descriptor units
      constexpr uint32_t kHalfW = 128 * 128 / 16;                           // 128 rows of an operand tile in descriptor units
      const uint32_t p_lo = desc_lo_k(sP), ds_lo = desc_lo_k(sdS), ds_mn = desc_lo_mn(sdS, Cfg::kBlk);
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        mbar_wait(in_full, static_cast<uint32_t>(it & 1));
        tc_fence_after();
#pragma unroll 1
        for (int n = 0; n < 4; ++n) {
          const int kb = n >> 1, qb = n & 1;
          const uint32_t g = static_cast<uint32_t>(it * 4 + n);
          const uint32_t k_lo = desc_lo_k(sK + kb * Cfg::kBlk), v_lo = desc_lo_k(sV + kb * Cfg::kBlk);
          const uint32_t q_lo = desc_lo_k(sQ + qb * Cfg::kBlk), do_lo = desc_lo_k(sdO + qb * Cfg::kBlk);
          // ---- scores and their gradient for tile n (the TMEM columns were released by p_full of tile n - 1)
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
            if (k == 0) umma_lohi<false>(tmem_base + Cfg::kColS, k_lo, q_lo, idesc_s);
            else umma_lohi<true>(tmem_base + Cfg::kColS, k_lo + 2 * k, q_lo + 2 * k, idesc_s);
          }
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
            if (k == 0) umma_lohi<false>(tmem_base + Cfg::kColdP, v_lo, do_lo, idesc_s);
            else umma_lohi<true>(tmem_base + Cfg::kColdP, v_lo + 2 * k, do_lo + 2 * k, idesc_s);
          }
          umma_commit(s_full);
          // ---- gradients of tile n
          mbar_wait(p_full, g & 1u);
          if (n == 0 && it > 0) {                        // dV / dK of the previous unit's second key block, and its dQ, have left TMEM
            mbar_wait(kv_drained, 1u);
            mbar_wait(epi_done, static_cast<uint32_t>((it - 1) & 1));
          }
          if (n == 2) mbar_wait(kv_drained, 0u);         // ... of this unit's first key block
          tc_fence_after();
          const uint32_t do_mn = desc_lo_mn(sdO + qb * Cfg::kBlk, 8192), q_mn = desc_lo_mn(sQ + qb * Cfg::kBlk, 8192);
          const uint32_t k_mn = desc_lo_mn(sK + kb * Cfg::kBlk, 8192);
          (void)kHalfW;
#pragma unroll
          for (int j = 0; j < 8; ++j) {                  // contraction over the tile's 128 queries
            const uint32_t a = p_lo + (j >> 2) * kBlkW + (j & 3) * 2;
            if (j == 0 && qb == 0) umma_lohi<false>(tmem_base + Cfg::kColdV, a, do_mn + j * 128, idesc_kv);
            else umma_lohi<true>(tmem_base + Cfg::kColdV, a, do_mn + j * 128, idesc_kv);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = ds_lo + (j >> 2) * kBlkW + (j & 3) * 2;
            if (j == 0 && qb == 0) umma_lohi<false>(tmem_base + Cfg::kColdK, a, q_mn + j * 128, idesc_kv);
            else umma_lohi<true>(tmem_base + Cfg::kColdK, a, q_mn + j * 128, idesc_kv);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {                  // contraction over the tile's 128 keys
            const uint32_t dq = tmem_base + Cfg::kColdQ + 64 * qb;
            if (j == 0 && kb == 0) umma_lohi<false>(dq, ds_mn + j * 128, k_mn + j * 128, idesc_q);
            else umma_lohi<true>(dq, ds_mn + j * 128, k_mn + j * 128, idesc_q);
          }
          umma_commit(g_done);
        }
        umma_commit(unit_done);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- math + epilogue warps
    constexpr float sl2 = 0.125f * 1.4426950408889634f;
    const int tid = threadIdx.x;                              // 0..127 = TMEM lane = key row inside the key block
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t stage = smem_u32(smem + Cfg::kOffStage) + static_cast<uint32_t>(warp) * 4096u;
    const uint32_t csq = dbias != nullptr ? smem_u32(s_colsum) : 0u;
    const uint32_t csk = csq ? csq + kHeadDim * 4 : 0u, csv = csq ? csq + 2 * kHeadDim * 4 : 0u;
    int it = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const int b = unit / kHeads, h = unit - b * kHeads;
      const float* lrow = lse2 + (static_cast<long long>(b) * kHeads + h) * T;
      __nv_bfloat16* base = dqkv + static_cast<long long>(b) * T * kQkvCols + h * kHeadDim;
      mbar_wait(in_full, static_cast<uint32_t>(it & 1));
#pragma unroll
      for (int rep = 0; rep < 2; ++rep) {                     // D[q] and lse2[q]: two query rows per thread
        const int q = tid + rep * 128;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t off = static_cast<uint32_t>(q) * 128u + static_cast<uint32_t>((j ^ (q & 7)) << 4);
          const uint4 a = lds_u4(sdO + off), c = lds_u4(sO + off);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(cw[e] << 16), acc);
            acc = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(cw[e] & 0xffff0000u), acc);
          }
        }
        sD[q] = acc;
        sL[q] = __ldg(lrow + q);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");          // D / lse complete; nobody reads O any more (P^T may overwrite it)
#pragma unroll 1
      for (int n = 0; n < 4; ++n) {
        const int kb = n >> 1, qb = n & 1;
        const uint32_t g = static_cast<uint32_t>(it * 4 + n);
        mbar_wait(s_full, g & 1u);
        if (g > 0) mbar_wait(g_done, (g - 1) & 1u);           // the previous tile's gradient MMAs have read P^T / dS^T
        tc_fence_after();
        {
          uint32_t sa[16], da[16], sb[16], db[16];
          tmem_ld_32x16(t_lane + Cfg::kColS, sa);
          tmem_ld_32x16(t_lane + Cfg::kColdP, da);
          tmem_ld_wait();
          const uint32_t prow = sP + static_cast<uint32_t>(tid) * 128u, dsrow = sdS + static_cast<uint32_t>(tid) * 128u;
          const int sw = tid & 7;
          const float* lq = sL + qb * 128;
          const float* dq = sD + qb * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t (&s_cur)[16] = (c & 1) ? sb : sa;
            uint32_t (&d_cur)[16] = (c & 1) ? db : da;
            uint32_t (&s_nxt)[16] = (c & 1) ? sa : sb;
            uint32_t (&d_nxt)[16] = (c & 1) ? da : db;
            if (c + 1 < 8) {
              tmem_ld_32x16(t_lane + Cfg::kColS + 16 * (c + 1), s_nxt);
              tmem_ld_32x16(t_lane + Cfg::kColdP + 16 * (c + 1), d_nxt);
            }
#pragma unroll
            for (int gg = 0; gg < 2; ++gg) {
              const int q0 = 16 * c + 8 * gg;
              const float4 l0 = *reinterpret_cast<const float4*>(lq + q0), l1 = *reinterpret_cast<const float4*>(lq + q0 + 4);
              const float4 e0 = *reinterpret_cast<const float4*>(dq + q0), e1 = *reinterpret_cast<const float4*>(dq + q0 + 4);
              const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
              const float dv[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
              float p[8], ds[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                p[e] = ex2f(fmaf(__uint_as_float(s_cur[8 * gg + e]), sl2, -lv[e]));
                ds[e] = p[e] * (__uint_as_float(d_cur[8 * gg + e]) - dv[e]) * 0.125f;
              }
              const int chunk = q0 >> 3;
              const uint32_t off = static_cast<uint32_t>(chunk >> 3) * Cfg::kBlk + static_cast<uint32_t>(((chunk & 7) ^ sw) << 4);
              sts_u4(prow + off, make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7])));
              sts_u4(dsrow + off, make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7])));
            }
            if (c + 1 < 8) tmem_ld_wait();
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        if (qb == 1) {                                        // the key block is complete: dV[kb], dK[kb] -> dqkv rows kb*128 + ...
          mbar_wait(g_done, g & 1u);
          tc_fence_after();
          __nv_bfloat16* row0 = base + static_cast<long long>(kb * 128 + warp * 32) * kQkvCols;
          store_acc_rows(t_lane + Cfg::kColdV, stage, row0 + 2 * kHidden, kQkvCols, 0, 32, lane, csv);
          store_acc_rows(t_lane + Cfg::kColdK, stage, row0 + kHidden, kQkvCols, 0, 32, lane, csk);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(kv_drained);
        }
      }
      // ---- dQ of both query blocks (g_done of the last tile was waited for above)
#pragma unroll
      for (int qb = 0; qb < 2; ++qb)
        store_acc_rows(t_lane + Cfg::kColdQ + 64 * qb, stage, base + static_cast<long long>(qb * 128 + warp * 32) * kQkvCols, kQkvCols,
                       0, 32, lane, csq);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_done);
      if (dbias != nullptr) {
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (warp == 0) {
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            const int idx = k * 32 + lane;
            atomicAdd(dbias + (idx >> 6) * kHidden + h * kHeadDim + (idx & 63), s_colsum[idx]);
            s_colsum[idx] = 0.f;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_bt256(const __nv_bfloat16* qkv, const __nv_bfloat16* o, const __nv_bfloat16* d_o, const float* lse2, __nv_bfloat16* dqkv,
                 float* dbias, int batch, cudaStream_t stream) {
  using Cfg = Bt256;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attention_bwd_tc256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "attention_bwd_tc256: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
                       cudaGetErrorString(cudaGetLastError()));
    configured = true;
  }
  CUtensorMap tm_qkv, tm_do, tm_o;
  const long long rows = static_cast<long long>(batch) * Cfg::T;
  int rc = make_tmap_bf16_kmajor(&tm_qkv, qkv, rows, kQkvCols, kQkvCols, Cfg::T);
  if (rc != kOk) return rc;
  rc = make_tmap_bf16_kmajor(&tm_do, d_o, rows, kHidden, kHidden, Cfg::T);
  if (rc != kOk) return rc;
  rc = make_tmap_bf16_kmajor(&tm_o, o, rows, kHidden, kHidden, Cfg::T);
  if (rc != kOk) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int units = batch * kHeads;
  attention_bwd_tc256_kernel<<<units < sms ? units : sms, kBtThreads, Cfg::kSmemBytes, stream>>>(tm_qkv, tm_do, tm_o, lse2, dqkv, dbias, units);
  return check_launch("attention_bwd_tc256_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// T = 324 (3x3 @288 px - the reference trainer's default image size): three key blocks x three query blocks of 128 (the last
// of each 68 live).  dV / dK of ONE key block plus all three dQ tiles plus the score tiles would need 576 TMEM columns, so the
// work item is (sample, head, key block): dV[kb], dK[kb] accumulate in TMEM over the item's three query blocks, while each
// tile's dQ contribution dS K[kb] is a fresh 128 x 64 accumulator (double buffered) that the math warps write into the key
// block's slice of an fp32 scratch; a closing pass adds the three slices and rounds them into dqkv's q columns.
// D[q] = sum_d dO O comes from a small pre-pass (O is not loaded here).  TMEM: S^T 0, dP^T 128, dV 256, dK 320, dQ parts 384 / 448.
struct Bt324 {
  static constexpr int T = 324, TB = 3;                       // tokens, 128-row blocks
  static constexpr int kBlk = 128 * 128;                      // 16 KB: 128 rows of an operand tile / 64 queries of P^T, dS^T
  static constexpr int kOffK = 0, kOffV = kBlk, kOffQ = 2 * kBlk, kOffdO = kOffQ + TB * kBlk;
  static constexpr int kOffP = kOffdO + TB * kBlk, kOffdS = kOffP + 2 * kBlk;
  static constexpr int kOffStage = kOffdS + 2 * kBlk;         // 4 x 4 KB epilogue staging
  static constexpr int kOffL = kOffStage + 4 * 4096;          // lse2[384], D[384]
  static constexpr int kBarOff = kOffL + 2 * TB * 128 * 4;
  static constexpr int kSmemBytes = kBarOff + 128 + 1024;
  static constexpr int kInBytes = (2 + 2 * TB) * kBlk;
  static constexpr int kColS = 0, kColdP = 128, kColdV = 256, kColdK = 320, kColdQ = 384;
  static_assert(kSmemBytes + 1024 <= 227 * 1024, "shared memory");
};

// D[(b, h), q] = sum_d dO[b T + q, 64 h + d] * O[b T + q, 64 h + d]; one warp per token row, 8 lanes per head
__global__ void __launch_bounds__(256)
attn_bwd_d_kernel(const __nv_bfloat16* __restrict__ d_o, const __nv_bfloat16* __restrict__ o, float* __restrict__ dsum, long long rows,
                  int tokens) {
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 3; ++j) {                               // 16-byte chunk lane + 32 j = columns 8 (lane + 32 j) ..: head (lane + 32 j) / 8
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(d_o + row * kHidden) + lane + 32 * j);
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(o + row * kHidden) + lane + 32 * j);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[j] = fmaf(__uint_as_float(aw[e] << 16), __uint_as_float(cw[e] << 16), acc[j]);
      acc[j] = fmaf(__uint_as_float(aw[e] & 0xffff0000u), __uint_as_float(cw[e] & 0xffff0000u), acc[j]);
    }
#pragma unroll
    for (int off = 4; off > 0; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
  }
  const long long b = row / tokens;
  const int q = static_cast<int>(row - b * tokens);
  if ((lane & 7) == 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int h = (lane >> 3) + 4 * j;
      dsum[(b * kHeads + h) * tokens + q] = acc[j];
    }
  }
}

// dqkv[row, 0:768] = bf16(sum over the key-block slices of dq32[slice][row, :])
__global__ void __launch_bounds__(256)
attn_bwd_dq_round_kernel(const float* __restrict__ dq32, long long slice_floats, int slices, __nv_bfloat16* __restrict__ dqkv, long long n4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = __ldcs(reinterpret_cast<const float4*>(dq32) + i);
  for (int s = 1; s < slices; ++s) {
    const float4 w = __ldcs(reinterpret_cast<const float4*>(dq32 + s * slice_floats) + i);
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  const long long row = i / (kHidden / 4);
  const int c4 = static_cast<int>(i - row * (kHidden / 4));
  uint2 u;
  u.x = pack_bf16(v.x, v.y); u.y = pack_bf16(v.z, v.w);
  reinterpret_cast<uint2*>(dqkv + row * kQkvCols)[c4] = u;
}

__global__ void __launch_bounds__(kBtThreads, 1)
attention_bwd_tc324_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                           const float* __restrict__ lse2, const float* __restrict__ dsum, __nv_bfloat16* __restrict__ dqkv,
                           float* __restrict__ dq32, long long slice_floats, float* __restrict__ dbias, int num_items) {
  using Cfg = Bt324;
  constexpr int T = Cfg::T, TB = Cfg::TB;
  extern __shared__ uint8_t att_bt_smem[];
  uint8_t* smem = att_bt_smem + ((1024u - (smem_u32(att_bt_smem) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* in_full = bars + 0;        // TMA: K[kb], V[kb], Q, dO of the item landed                         (1 / item)
  uint64_t* s_full = bars + 1;         // MMA: S^T, dP^T of a tile are in TMEM                                 (3 / item)
  uint64_t* p_full = bars + 2;         // math: P^T, dS^T of a tile are in shared memory (4 arrivals)          (3 / item)
  uint64_t* g_done = bars + 3;         // MMA: the gradient MMAs of a tile are done                            (3 / item)
  uint64_t* kv_drained = bars + 4;     // math: dV / dK of the item have left TMEM (4 arrivals)                (1 / item)
  uint64_t* dq_drained = bars + 5;     // [2] math: the dQ partial in buffer b has left TMEM (4 arrivals)
  uint64_t* item_done = bars + 7;      // MMA: every MMA of the item has read its operand tiles                (1 / item)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* sL = reinterpret_cast<float*>(smem + Cfg::kOffL);
  float* sD = sL + TB * 128;
  __shared__ float s_colsum[2 * kHeadDim];     // dK | dV column sums of the item

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 2 * kHeadDim) s_colsum[threadIdx.x] = 0.f;
  if (threadIdx.x == 0) {
    mbar_init(in_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(g_done, 1); mbar_init(kv_drained, 4);
    mbar_init(&dq_drained[0], 4); mbar_init(&dq_drained[1], 4); mbar_init(item_done, 1);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tm_qkv); tma_prefetch_desc(&tm_do); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV), sQ = smem_u32(smem + Cfg::kOffQ),
                 sdO = smem_u32(smem + Cfg::kOffdO), sP = smem_u32(smem + Cfg::kOffP), sdS = smem_u32(smem + Cfg::kOffdS);

  if (warp == 4) {
    // ---------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int unit = item / TB, kb = item - unit * TB;
        const int b = unit / kHeads, h = unit - b * kHeads;
        if (it > 0) mbar_wait_backoff(item_done, static_cast<uint32_t>((it - 1) & 1), 64);
        mbar_expect_tx(in_full, Cfg::kInBytes);
        // 128-row boxes; rows past the sample's 324 tokens belong to the next sample (finite values, masked below) or are
        // zero-filled past the end of the tensor
        tma_load_2d(&tm_qkv, in_full, smem + Cfg::kOffK, kHidden + h * kHeadDim, b * T + kb * 128);
        tma_load_2d(&tm_qkv, in_full, smem + Cfg::kOffV, 2 * kHidden + h * kHeadDim, b * T + kb * 128);
#pragma unroll
        for (int qb = 0; qb < TB; ++qb) {
          tma_load_2d(&tm_qkv, in_full, smem + Cfg::kOffQ + qb * Cfg::kBlk, h * kHeadDim, b * T + qb * 128);
          tma_load_2d(&tm_do, in_full, smem + Cfg::kOffdO + qb * Cfg::kBlk, h * kHeadDim, b * T + qb * 128);
        }
        if (item + static_cast<int>(gridDim.x) < num_items) {       // the next item's tiles towards L2
          const int ni = item + static_cast<int>(gridDim.x), nunit = ni / TB, nkb = ni - nunit * TB;
          const int nb = nunit / kHeads, nh = nunit - nb * kHeads;
          tma_prefetch_2d(&tm_qkv, kHidden + nh * kHeadDim, nb * T + nkb * 128);
          tma_prefetch_2d(&tm_qkv, 2 * kHidden + nh * kHeadDim, nb * T + nkb * 128);
#pragma unroll
          for (int qb = 0; qb < TB; ++qb) {
            tma_prefetch_2d(&tm_qkv, nh * kHeadDim, nb * T + qb * 128);
            tma_prefetch_2d(&tm_do, nh * kHeadDim, nb * T + qb * 128);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
      constexpr uint32_t idesc_kv = umma_idesc_bf16(128, kHeadDim, 0, 1);
      constexpr uint32_t idesc_q = umma_idesc_bf16(128, kHeadDim, 1, 1);
      constexpr uint32_t kBlkW = Cfg::kBlk / 16;
      const uint32_t p_lo = desc_lo_k(sP), ds_lo = desc_lo_k(sdS), ds_mn = desc_lo_mn(sdS, Cfg::kBlk);
      const uint32_t k_lo = desc_lo_k(sK), v_lo = desc_lo_k(sV), k_mn = desc_lo_mn(sK, 8192);
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        mbar_wait(in_full, static_cast<uint32_t>(it & 1));
        tc_fence_after();
#pragma unroll 1
        for (int qb = 0; qb < TB; ++qb) {
          const uint32_t g = static_cast<uint32_t>(it * TB + qb);
          const uint32_t q_lo = desc_lo_k(sQ + qb * Cfg::kBlk), do_lo = desc_lo_k(sdO + qb * Cfg::kBlk);
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
            if (k == 0) umma_lohi<false>(tmem_base + Cfg::kColS, k_lo, q_lo, idesc_s);
            else umma_lohi<true>(tmem_base + Cfg::kColS, k_lo + 2 * k, q_lo + 2 * k, idesc_s);
          }
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
            if (k == 0) umma_lohi<false>(tmem_base + Cfg::kColdP, v_lo, do_lo, idesc_s);
            else umma_lohi<true>(tmem_base + Cfg::kColdP, v_lo + 2 * k, do_lo + 2 * k, idesc_s);
          }
          umma_commit(s_full);
          mbar_wait(p_full, g & 1u);
          if (qb == 0 && it > 0) mbar_wait(kv_drained, static_cast<uint32_t>((it - 1) & 1));   // dV / dK of the previous item have left TMEM
          if (g >= 2) mbar_wait(&dq_drained[g & 1u], ((g >> 1) - 1u) & 1u);                      // ... and the dQ partial of tile g - 2
          tc_fence_after();
          const uint32_t do_mn = desc_lo_mn(sdO + qb * Cfg::kBlk, 8192), q_mn = desc_lo_mn(sQ + qb * Cfg::kBlk, 8192);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = p_lo + (j >> 2) * kBlkW + (j & 3) * 2;
            if (j == 0 && qb == 0) umma_lohi<false>(tmem_base + Cfg::kColdV, a, do_mn + j * 128, idesc_kv);
            else umma_lohi<true>(tmem_base + Cfg::kColdV, a, do_mn + j * 128, idesc_kv);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = ds_lo + (j >> 2) * kBlkW + (j & 3) * 2;
            if (j == 0 && qb == 0) umma_lohi<false>(tmem_base + Cfg::kColdK, a, q_mn + j * 128, idesc_kv);
            else umma_lohi<true>(tmem_base + Cfg::kColdK, a, q_mn + j * 128, idesc_kv);
          }
          const uint32_t dq = tmem_base + Cfg::kColdQ + 64 * (g & 1u);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j == 0) umma_lohi<false>(dq, ds_mn + j * 128, k_mn + j * 128, idesc_q);
            else umma_lohi<true>(dq, ds_mn + j * 128, k_mn + j * 128, idesc_q);
          }
          umma_commit(g_done);
        }
        umma_commit(item_done);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- math + epilogue warps
    constexpr float sl2 = 0.125f * 1.4426950408889634f;
    const int tid = threadIdx.x;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t stage = smem_u32(smem + Cfg::kOffStage) + static_cast<uint32_t>(warp) * 4096u;
    const uint32_t csk = dbias != nullptr ? smem_u32(s_colsum) : 0u, csv = csk ? csk + kHeadDim * 4 : 0u;
    // a drained dQ partial (this warp: 32 query rows x 64 fp32 columns) -> the item's key-block slice of the fp32 scratch,
    // transposed through the staging tile 32 columns at a time so that every store writes four full 128-byte row segments.
    // Plain stores: red.global.add into one buffer was the first form - 221 MB of L2 atomics per launch dominated the kernel.
    auto drain_dq = [&](uint32_t g, int b, int h, int qb, int kb) {
      uint32_t a[32], c[32];
      tmem_ld_32x32(t_lane + Cfg::kColdQ + 64 * (g & 1u), a);
      tmem_ld_32x32(t_lane + Cfg::kColdQ + 64 * (g & 1u) + 32, c);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dq_drained[g & 1u]);
      const int q_first = qb * 128 + warp * 32;
      const int live = T - q_first < 0 ? 0 : (T - q_first < 32 ? T - q_first : 32);
      float* dst0 = dq32 + static_cast<long long>(kb) * slice_floats + (static_cast<long long>(b) * T + q_first) * kHidden + h * kHeadDim;
      const uint32_t mine = stage + lane * 128, sw = lane & 7;
      const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const uint32_t (&r)[32] = half ? c : a;
#pragma unroll
        for (int j = 0; j < 8; ++j) sts_u4(mine + ((j ^ sw) << 4), make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]));
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + sub;
          const uint4 u = lds_u4(stage + rr * 128 + ((ch ^ (rr & 7)) << 4));
          if (rr < live) *reinterpret_cast<uint4*>(dst0 + static_cast<long long>(rr) * kHidden + half * 32 + ch * 4) = u;
        }
        __syncwarp();
      }
    };
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int unit = item / TB, kb = item - unit * TB;
      const int b = unit / kHeads, h = unit - b * kHeads;
      const float* lrow = lse2 + (static_cast<long long>(b) * kHeads + h) * T;
      const float* drow = dsum + (static_cast<long long>(b) * kHeads + h) * T;
      __nv_bfloat16* base = dqkv + static_cast<long long>(b) * T * kQkvCols + h * kHeadDim;
      const int key = kb * 128 + tid;
      const bool key_ok = key < T;
      // lse / D of the unit's queries; dead queries get lse = +inf (probability 0) and D = 0.  The previous item's last reads of
      // these arrays happened before its p_full arrivals, which every warp has passed.
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
      for (int rep = 0; rep < TB; ++rep) {
        const int q = tid + rep * 128;
        sL[q] = q < T ? __ldg(lrow + q) : INFINITY;
        sD[q] = q < T ? __ldg(drow + q) : 0.f;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll 1
      for (int qb = 0; qb < TB; ++qb) {
        const uint32_t g = static_cast<uint32_t>(it * TB + qb);
        mbar_wait(s_full, g & 1u);
        if (g > 0) mbar_wait(g_done, (g - 1) & 1u);           // P^T / dS^T are free again; the dQ partial of tile g - 1 is complete
        tc_fence_after();
        {
          uint32_t sa[16], da[16], sb[16], db[16];
          tmem_ld_32x16(t_lane + Cfg::kColS, sa);
          tmem_ld_32x16(t_lane + Cfg::kColdP, da);
          tmem_ld_wait();
          const uint32_t prow = sP + static_cast<uint32_t>(tid) * 128u, dsrow = sdS + static_cast<uint32_t>(tid) * 128u;
          const int sw = tid & 7;
          const float* lq = sL + qb * 128;
          const float* dq = sD + qb * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint32_t (&s_cur)[16] = (c & 1) ? sb : sa;
            uint32_t (&d_cur)[16] = (c & 1) ? db : da;
            uint32_t (&s_nxt)[16] = (c & 1) ? sa : sb;
            uint32_t (&d_nxt)[16] = (c & 1) ? da : db;
            if (c + 1 < 8) {
              tmem_ld_32x16(t_lane + Cfg::kColS + 16 * (c + 1), s_nxt);
              tmem_ld_32x16(t_lane + Cfg::kColdP + 16 * (c + 1), d_nxt);
            }
#pragma unroll
            for (int gg = 0; gg < 2; ++gg) {
              const int q0 = 16 * c + 8 * gg;
              const float4 l0 = *reinterpret_cast<const float4*>(lq + q0), l1 = *reinterpret_cast<const float4*>(lq + q0 + 4);
              const float4 e0 = *reinterpret_cast<const float4*>(dq + q0), e1 = *reinterpret_cast<const float4*>(dq + q0 + 4);
              const float lv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
              const float dv[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
              float p[8], ds[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {                   // dead key row or dead query column -> exactly 0 (garbage scores stay finite)
                p[e] = key_ok ? ex2f(fmaf(__uint_as_float(s_cur[8 * gg + e]), sl2, -lv[e])) : 0.f;
                ds[e] = (p[e] != 0.f) ? p[e] * (__uint_as_float(d_cur[8 * gg + e]) - dv[e]) * 0.125f : 0.f;
              }
              const int chunk = q0 >> 3;
              const uint32_t off = static_cast<uint32_t>(chunk >> 3) * Cfg::kBlk + static_cast<uint32_t>(((chunk & 7) ^ sw) << 4);
              sts_u4(prow + off, make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7])));
              sts_u4(dsrow + off, make_uint4(pack_bf16(ds[0], ds[1]), pack_bf16(ds[2], ds[3]), pack_bf16(ds[4], ds[5]), pack_bf16(ds[6], ds[7])));
            }
            if (c + 1 < 8) tmem_ld_wait();
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        // while the gradient MMAs of this tile run: the dQ partial of the previous tile (complete: g_done(g - 1) above)
        if (g > 0 && qb > 0) drain_dq(g - 1, b, h, qb - 1, kb);
      }
      // ---- end of the item: its last dQ partial, then dV[kb], dK[kb]
      const uint32_t g_last = static_cast<uint32_t>(it * TB + TB - 1);
      mbar_wait(g_done, g_last & 1u);
      tc_fence_after();
      drain_dq(g_last, b, h, TB - 1, kb);
      const int row_first = kb * 128 + warp * 32;
      const int live = T - row_first < 0 ? 0 : (T - row_first < 32 ? T - row_first : 32);
      __nv_bfloat16* row0 = base + static_cast<long long>(row_first) * kQkvCols;
      store_acc_rows(t_lane + Cfg::kColdV, stage, row0 + 2 * kHidden, kQkvCols, 0, live, lane, csv);
      store_acc_rows(t_lane + Cfg::kColdK, stage, row0 + kHidden, kQkvCols, 0, live, lane, csk);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(kv_drained);
      if (dbias != nullptr) {
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (warp == 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int idx = k * 32 + lane;                    // 0..63: dK columns, 64..127: dV columns
            atomicAdd(dbias + (1 + (idx >> 6)) * kHidden + h * kHeadDim + (idx & 63), s_colsum[idx]);
            s_colsum[idx] = 0.f;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int launch_bt324(const __nv_bfloat16* qkv, const __nv_bfloat16* o, const __nv_bfloat16* d_o, const float* lse2, __nv_bfloat16* dqkv,
                 float* dbias, float* scratch, int batch, cudaStream_t stream) {
  using Cfg = Bt324;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attention_bwd_tc324_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "attention_bwd_tc324: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
                       cudaGetErrorString(cudaGetLastError()));
    configured = true;
  }
  const long long rows = static_cast<long long>(batch) * Cfg::T;
  const size_t dq_floats = static_cast<size_t>(rows) * kHidden, d_floats = static_cast<size_t>(batch) * kHeads * Cfg::T;
  // scratch: 3 key-block slices of rows x 768 fp32 dQ partial sums (every live row of every slice is written: no memset) +
  // batch x 12 x 324 fp32 for D; the caller's buffer, or stream-ordered memory
  float* mine = nullptr;
  if (scratch == nullptr) {
    static bool pool_kept[64] = {};     // keep freed blocks in the device's stream-ordered pool: the next call reuses them
    int pdev = 0;
    cudaGetDevice(&pdev);
    if (pdev >= 0 && pdev < 64 && !pool_kept[pdev]) {
      cudaMemPool_t pool;
      unsigned long long keep = ~0ull;
      if (cudaDeviceGetDefaultMemPool(&pool, pdev) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      pool_kept[pdev] = true;
    }
    if (cudaMallocAsync(&mine, (Cfg::TB * dq_floats + d_floats) * sizeof(float), stream) != cudaSuccess)
      return set_error(kErrCuda, "attention_bwd_tc324: cudaMallocAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
    scratch = mine;
  }
  float* dq32 = scratch;
  float* dsum = scratch + Cfg::TB * dq_floats;
  int rc = kOk;
  do {
    attn_bwd_d_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(d_o, o, dsum, rows, Cfg::T);
    if ((rc = check_launch("attn_bwd_d_kernel")) != kOk) break;
    CUtensorMap tm_qkv, tm_do;
    if ((rc = make_tmap_bf16_kmajor(&tm_qkv, qkv, rows, kQkvCols, kQkvCols, 128)) != kOk) break;
    if ((rc = make_tmap_bf16_kmajor(&tm_do, d_o, rows, kHidden, kHidden, 128)) != kOk) break;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int items = batch * kHeads * Cfg::TB;
    attention_bwd_tc324_kernel<<<items < sms ? items : sms, kBtThreads, Cfg::kSmemBytes, stream>>>(tm_qkv, tm_do, lse2, dsum, dqkv, dq32,
                                                                                                  static_cast<long long>(dq_floats), dbias, items);
    if ((rc = check_launch("attention_bwd_tc324_kernel")) != kOk) break;
    const long long n4 = static_cast<long long>(dq_floats / 4);
    attn_bwd_dq_round_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(dq32, static_cast<long long>(dq_floats), Cfg::TB, dqkv, n4);
    if ((rc = check_launch("attn_bwd_dq_round_kernel")) != kOk) break;
    if (dbias != nullptr) rc = launch_colsum_bf16(dqkv, kQkvCols, rows, kHidden, dbias, stream);     // the q third of the qkv bias gradient
  } while (false);
  if (mine != nullptr) cudaFreeAsync(mine, stream);
  return rc;
}

template <int T>
int launch_bt(const __nv_bfloat16* qkv, const __nv_bfloat16* o, const __nv_bfloat16* d_o, const float* lse2, __nv_bfloat16* dqkv,
              float* dbias, int batch, cudaStream_t stream) {
  using Cfg = BtCfg<T>;
  static bool configured = false;
  auto kern = attention_bwd_tc_kernel<T>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "attention_bwd_tc: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
                       cudaGetErrorString(cudaGetLastError()));
    configured = true;
  }
  CUtensorMap tm_qkv, tm_do, tm_o;
  const long long rows = static_cast<long long>(batch) * T;
  int rc = make_tmap_bf16_kmajor(&tm_qkv, qkv, rows, kQkvCols, kQkvCols, T);
  if (rc != kOk) return rc;
  rc = make_tmap_bf16_kmajor(&tm_do, d_o, rows, kHidden, kHidden, T);
  if (rc != kOk) return rc;
  rc = make_tmap_bf16_kmajor(&tm_o, o, rows, kHidden, kHidden, T);
  if (rc != kOk) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int units = batch * kHeads;
  kern<<<units < sms ? units : sms, kBtThreads, Cfg::kSmemBytes, stream>>>(tm_qkv, tm_do, tm_o, lse2, dqkv, dbias, units);
  return check_launch("attention_bwd_tc_kernel");
}

}  // namespace

bool attention_bwd_tc_supported(int tokens) {
  static int legacy = -1;      // JPDVT_ATTN_BWD_LEGACY=1: the mma.sync backward for every size (A/B knob)
  if (legacy < 0) { const char* e = getenv("JPDVT_ATTN_BWD_LEGACY"); legacy = (e != nullptr && e[0] == '1') ? 1 : 0; }
  return !legacy && (tokens == 144 || tokens == 256 || tokens == 324);
}

int launch_attention_bwd_tc(const __nv_bfloat16* qkv, const __nv_bfloat16* o, const __nv_bfloat16* d_o, const float* lse2,
                            __nv_bfloat16* dqkv, float* dbias, int batch, int tokens, cudaStream_t stream) {
  if (batch <= 0) return kOk;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(o) & 15) || (reinterpret_cast<uintptr_t>(d_o) & 15) ||
      (reinterpret_cast<uintptr_t>(dqkv) & 15))
    return set_error(kErrBadArg, "attention_bwd_tc: pointers must be 16-byte aligned");
  switch (tokens) {
    case 144: return launch_bt<144>(qkv, o, d_o, lse2, dqkv, dbias, batch, stream);
    case 256: return launch_bt256(qkv, o, d_o, lse2, dqkv, dbias, batch, stream);
    case 324: return launch_bt324(qkv, o, d_o, lse2, dqkv, dbias, nullptr, batch, stream);
    default: return set_error(kErrUnsupported, "attention_bwd_tc: %d tokens not instantiated", tokens);
  }
}

}  // namespace jp
