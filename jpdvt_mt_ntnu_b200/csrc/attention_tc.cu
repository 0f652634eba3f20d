// tcgen05 attention for the JPDVT piece tokens: softmax(Q K^T / 8) V per (sample, head) with both contractions on the
// 5th-generation tensor cores and the scores / outputs living in TMEM.
//
// Replaces timm Attention.forward's F.scaled_dot_product_attention (image_model/models.py:108,120) for T = 144 (192 px)
// and T = 256 (256 px), inference and training forward (optionally writes the per-row log-sum-exp the backward reads);
// the mma.sync kernel in attention.cu keeps the other sizes.
//
// Kernels of this file, by token count (defaults; DESIGN.md section 4 has the measurements):
//   T = 144 : attention_rw_kernel    - main 128-row tile on tcgen05 with the probabilities kept in TMEM (A-from-TMEM MMAs), the
//                                      16-row remainder on one warp's mma.sync; attention_tc_kernel<144> / attention_hm_kernel /
//                                      attention_tc8_kernel are the earlier forms, kept as A/B opt-ins (JPDVT_ATTN_REM, JPDVT_ATTN_WARPS)
//   T = 256 : attention_qt_kernel    - work item = (unit, 128-query tile), P in TMEM, two CTAs per SM (attention_tc_kernel<256>: opt-in)
//   T = 324 : attention_ks_kernel    - work item = (unit, 128-query tile), keys in two blocks with the online-softmax rescale of O in
//                                      TMEM, two CTAs per SM (attention_tc_seq_kernel / attention_seq8_kernel: the earlier forms, opt-in)
// The description below is the first tcgen05 form (attention_tc_kernel), whose building blocks the others share.
//
// One CTA works on one (sample, head) unit at a time, several units per CTA (persistent grid):
//   warps 0-3 : softmax + output epilogue; warp w owns TMEM lanes [32w, 32w+32) = 32 query rows of a 128-row tile
//   warp 4    : TMA producer - Q, K, V of the head straight out of the fused QKV activation [B*T, 2304]
//               ({64 cols x T rows} boxes, 128-byte swizzle -> canonical K-major tiles)
//   warp 5    : MMA issuer   - S = Q K^T  (M = 128 query rows, N = T keys, K = 64)  -> TMEM columns [0, T)
//                              O = P V    (M = 128, N = 64, K = T; P is the bf16 probability tile the softmax warps
//                              wrote to shared memory in K-major swizzled form, V is read MN-major as loaded)
// T = 144 is 128 + 16 query rows.  The 16-row remainder is NOT given to one warp (one thread per row would leave half
// a warp idle and three SM sub-partitions waiting): its scores are produced by four small MMA groups, group j multiplying
// the Q rows shifted by 32 j against keys [32j, 32j+32) (the last group takes 48), so TMEM lanes [32j, 32j+16) - lane
// quadrant j - hold rows 128..143 x its own quarter of the keys.  All four softmax warps then work on the remainder, a
// quarter of the columns each, and combine the row maximum / row sum through shared memory.  Its P tile only has 16 live
// rows and is stored compactly (2 KB per 64-key block).
// TMEM per CTA: S [0,T) shared by both tiles in turn, O0 [T,T+64), O1 aliases S[0,64) -> 256 columns at T = 144, two
// CTAs per SM, so one CTA's loads / MMA latencies hide behind the other's softmax.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>

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

// tcgen05.mma from split descriptor words.  The issuing thread is alone in its warp: every instruction it spends building
// a 64-bit descriptor costs several cycles of exposed latency, and the MMAs of one unit are tiny, so the descriptors are
// formed as (precomputed low word + compile-time offset, constant high word) and the accumulate flag is compile-time.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);        // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo_k(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | ((8192u >> 4) << 16); }
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
struct TcCfg {
  static_assert(T % 16 == 0 && T >= 16 && T <= 256, "tokens must be a multiple of 16, at most 256");
  static constexpr int kTiles = T > 128 ? 2 : 1;
  static constexpr int kRem = T - 128;                       // live rows of tile 1
  static constexpr bool kSplit = kTiles == 2 && kRem <= 16;    // 16-row remainder: split its key columns over the four warps
  static_assert(!kSplit || T == 144, "the split remainder is laid out for 128 + 16 rows");
  static_assert(kTiles == 1 || kSplit || kRem == 128, "other remainders are not instantiated");
  static constexpr int kKBlocks = (T + 63) / 64;             // 64-key blocks of the P tile (128 B per row and block)
  static constexpr int kTileBytes = T * 128;                 // one of Q / K / V
  static constexpr int kPBytes = kKBlocks * 16384;           // P0: 128 rows x kKBlocks x 128 B
  // split remainder: its compact P tile (16 live rows x kKBlocks x 2 KB) sits right before P0, so the 128-row A operand the
  // MMA reads for it runs on into P0 (dead rows, read only) and tile 1 need not wait for the O0 MMAs to release P0.
  // dual (two full tiles, T = 256): both score tiles are computed up front into the two halves of TMEM and each tile has
  // its own P buffer, so the second tile's softmax starts the moment the first is done and O0's MMAs run underneath it.
  static constexpr bool kDual = kTiles == 2 && kRem > 16;
  static_assert(!kDual || 2 * T <= 512, "two score tiles must fit the 512 TMEM columns");
  static constexpr int kP1Bytes = kSplit ? kKBlocks * 2048 : (kDual ? kPBytes : 0);
  static constexpr int kOffQ = 0, kOffK = kTileBytes, kOffV = 2 * kTileBytes;
  static constexpr int kOffP1 = kSplit ? 3 * kTileBytes : 3 * kTileBytes + kPBytes;
  static constexpr int kOffP = kSplit ? kOffP1 + kP1Bytes : 3 * kTileBytes;
  // output staging (4 KB per softmax warp) lives in the P buffer of the tile being written out: it is idle once that
  // tile's O MMAs are done
  static constexpr int kEndBytes = (kOffP > kOffP1 ? kOffP + kPBytes : kOffP1 + kP1Bytes);
  static constexpr int kDataBytes = kEndBytes > 256 * 128 ? kEndBytes : 256 * 128;
  // TMEM columns: scores of tile 0 / tile 1, outputs of tile 0 / tile 1
  static constexpr int kColS1 = kDual ? 256 : 0, kColO0 = kDual ? 0 : T, kColO1 = kDual ? 256 : 0;
  static constexpr int kBarOff = kDataBytes;
  static constexpr int kXchOff = kBarOff + 128;              // [2][4][16] floats: row max / row sum of the split remainder
  static constexpr int kSmemBytes = kXchOff + 512 + 1024;    // + alignment slack
  static constexpr int kTmemCols = (!kDual && T + 64 <= 256) ? 256 : 512;
  static constexpr int kCtasPerSm = (T + 64 <= 256 && 2 * (kSmemBytes + 1024) <= 227 * 1024) ? 2 : 1;
  static_assert(kOffP % 1024 == 0 && kOffP1 % 1024 == 0 && kTileBytes % 1024 == 0, "operand tiles must stay 1024-byte aligned (swizzle atoms)");
  static_assert(!kSplit || T + 64 + 32 <= kTmemCols, "transposed remainder scores need 32 TMEM columns behind O0");
};

// One softmax pass of the thread's S row (TMEM lane = row, columns [0,T)): exact row maximum, then
// p = 2^((s - max) * log2(e) / 8) written as bf16 into the K-major swizzled P tile; returns sum(p) (fp32, unrounded p).
// PB: the 64-key blocks of the P row are announced one by one (blk_bar[0], blk_bar[1]: one arrival per warp) as soon as they are
// in shared memory, so the MMA warp can issue their P V k-steps while the later blocks are still being computed.
template <int T, int TV = T, bool PB = false>     // T columns are read (multiple of 16); columns >= TV are padding and get probability 0
__device__ __forceinline__ float softmax_row_to_p(uint32_t t_row, uint32_t p_row_addr, uint32_t blk_stride, int sw, bool store,
                                                  float& ms_out, uint64_t* blk_bar = nullptr, int lane = 0) {
  constexpr float sl2 = 0.125f * 1.4426950408889634f;       // head_dim^-0.5 * log2(e)
  constexpr int kFull = T / 32, kTail = T % 32;              // kTail is 0 or 16
  // chunk c = columns [32c, 32c+32) (the last one may be 16 wide); the load of chunk c+1 is in flight while chunk c is
  // processed (tcgen05.wait::ld after the math, not before it)
  constexpr int kChunks = kFull + (kTail ? 1 : 0);
  uint32_t ra[32], rb[32];
  auto load_chunk = [&](uint32_t (&r)[32], int c) {
    if (c < kFull) tmem_ld_32x32(t_row + c * 32, r);
    else tmem_ld_32x16(t_row + c * 32, reinterpret_cast<uint32_t (&)[16]>(r));
  };
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  load_chunk(ra, 0);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    uint32_t (&cur)[32] = (c & 1) ? rb : ra;
    uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
    if (c + 1 < kChunks) load_chunk(nxt, c + 1);
    const int n = (c < kFull) ? 32 : kTail;
#pragma unroll
    for (int j = 0; j < n; j += 8) {
      if (c * 32 + j + 8 <= TV) {
        m0 = fmaxf(m0, fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1])));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(cur[j + 2]), __uint_as_float(cur[j + 3])));
        m2 = fmaxf(m2, fmaxf(__uint_as_float(cur[j + 4]), __uint_as_float(cur[j + 5])));
        m3 = fmaxf(m3, fmaxf(__uint_as_float(cur[j + 6]), __uint_as_float(cur[j + 7])));
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (c * 32 + j + e < TV) m0 = fmaxf(m0, __uint_as_float(cur[j + e]));
      }
    }
    if (c + 1 < kChunks) tmem_ld_wait();
  }
  load_chunk(ra, 0);                                          // second pass: its first chunk is in flight during the reduction
  const float ms = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sl2;
  ms_out = ms;
  float sum = 0.f;
  uint64_t sum2 = f2_pack(0.f, 0.f);                           // packed partial row sums (even / odd keys)
  const uint64_t sl2p = f2_pack(sl2, sl2), nmsp = f2_pack(-ms, -ms);
  auto emit8 = [&](const uint32_t* r, int chunk) {            // 8 consecutive keys -> one 16-byte chunk of the P row
    float p[8];
    if (chunk * 8 + 8 <= TV) {                                 // all eight keys valid: scale / shift and the sum on packed pairs
      uint64_t e[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float a, b;
        f2_unpack(f2_fma(f2_pack(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), sl2p, nmsp), a, b);
        p[2 * j] = ex2f(a); p[2 * j + 1] = ex2f(b);
        e[j] = f2_pack(p[2 * j], p[2 * j + 1]);
      }
      sum2 = f2_add(sum2, f2_add(f2_add(e[0], e[1]), f2_add(e[2], e[3])));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) p[j] = (chunk * 8 + j < TV) ? ex2f(fmaf(__uint_as_float(r[j]), sl2, -ms)) : 0.f;
      sum += ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
    }
    const uint4 u = make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
    if (store) sts_u4(p_row_addr + static_cast<uint32_t>(chunk >> 3) * blk_stride + static_cast<uint32_t>(((chunk & 7) ^ sw) << 4), u);
  };
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    uint32_t (&cur)[32] = (c & 1) ? rb : ra;
    uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
    if (c + 1 < kChunks) load_chunk(nxt, c + 1);
    const int n = (c < kFull) ? 32 : kTail;
#pragma unroll
    for (int g = 0; g < n / 8; ++g) emit8(cur + 8 * g, c * 4 + g);
    if constexpr (PB) {
      if ((c & 1) == 1 && c + 1 < kChunks) {                  // keys [32 (c - 1), 32 (c + 1)) = one 64-key block of P is complete
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&blk_bar[c >> 1]);
      }
    }
    if (c + 1 < kChunks) tmem_ld_wait();
  }
  float s_even, s_odd;
  f2_unpack(sum2, s_even, s_odd);
  return sum + (s_even + s_odd);
}

__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// tcgen05.mma with the A operand in TENSOR MEMORY (K-major: lane = row, two bf16 per 32-bit column), B from shared memory
template <bool ACC>
__device__ __forceinline__ void umma_ts_lohi(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(idesc), "n"(ACC ? 1 : 0), "r"(kDescHi)
      : "memory");
}

// softmax_row_to_p with the probabilities going back into TENSOR MEMORY instead of shared memory: the bf16 pairs of keys
// [32c, 32c + 32) overwrite columns [16c, 16c + 16) of the thread's own score row - columns whose scores this thread has
// already read (chunk c is in registers, the load in flight is chunk c + 1 at columns >= 32c + 32) - and the P V MMAs read
// them as their A operand straight from there: no shared-memory tile, no store / operand-read traffic for P at all.
template <int T, int TV = T>      // T columns are read (multiple of 16); columns >= TV are padding and get probability 0
__device__ __forceinline__ float softmax_row_to_tmem(uint32_t t_row, float& ms_out) {
  constexpr float sl2 = 0.125f * 1.4426950408889634f;
  constexpr int kFull = T / 32, kTail = T % 32;
  static_assert(kTail == 0 || kTail == 16, "token count must be a multiple of 16");
  constexpr int kChunks = kFull + (kTail ? 1 : 0);
  uint32_t ra[32], rb[32];
  auto load_chunk = [&](uint32_t (&r)[32], int c) {
    if (c < kFull) tmem_ld_32x32(t_row + c * 32, r);
    else tmem_ld_32x16(t_row + c * 32, reinterpret_cast<uint32_t (&)[16]>(r));
  };
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  load_chunk(ra, 0);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    uint32_t (&cur)[32] = (c & 1) ? rb : ra;
    uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
    if (c + 1 < kChunks) load_chunk(nxt, c + 1);
    const int n = (c < kFull) ? 32 : kTail;
#pragma unroll
    for (int j = 0; j < n; j += 8) {
      if (c * 32 + j + 8 <= TV) {
        m0 = fmaxf(m0, fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1])));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(cur[j + 2]), __uint_as_float(cur[j + 3])));
        m2 = fmaxf(m2, fmaxf(__uint_as_float(cur[j + 4]), __uint_as_float(cur[j + 5])));
        m3 = fmaxf(m3, fmaxf(__uint_as_float(cur[j + 6]), __uint_as_float(cur[j + 7])));
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (c * 32 + j + e < TV) m0 = fmaxf(m0, __uint_as_float(cur[j + e]));
      }
    }
    if (c + 1 < kChunks) tmem_ld_wait();
  }
  load_chunk(ra, 0);
  const float ms = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * sl2;
  ms_out = ms;
  uint64_t sum2 = f2_pack(0.f, 0.f);
  const uint64_t sl2p = f2_pack(sl2, sl2), nmsp = f2_pack(-ms, -ms);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    uint32_t (&cur)[32] = (c & 1) ? rb : ra;
    uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
    if (c + 1 < kChunks) load_chunk(nxt, c + 1);
    const int n = (c < kFull) ? 32 : kTail;
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < n / 2; ++j) {
      float a, b;
      f2_unpack(f2_fma(f2_pack(__uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1])), sl2p, nmsp), a, b);
      float pa = ex2f(a), pb = ex2f(b);
      if (c * 32 + 2 * j >= TV) pa = 0.f;                      // padded keys (compile-time after unrolling)
      if (c * 32 + 2 * j + 1 >= TV) pb = 0.f;
      sum2 = f2_add(sum2, f2_pack(pa, pb));
      pk[j] = pack_bf16(pa, pb);
    }
    if (c < kFull) tmem_st_32x16(t_row + c * 16, pk);
    else tmem_st_32x8(t_row + c * 16, pk);
    if (c + 1 < kChunks) tmem_ld_wait();
  }
  tmem_st_wait();
  float s_even, s_odd;
  f2_unpack(sum2, s_even, s_odd);
  return s_even + s_odd;
}

// Split remainder: this warp's NCJ key columns of the 16 remainder rows (live in lanes 0..15 of the warp's TMEM quadrant).
// Row maximum and row sum are combined across the four warps through `xch`; p goes out as bf16 into the compact P tile
// (row = lane, 2 KB per 64-key block).  Returns the full row sum.
template <int NCJ>
__device__ __forceinline__ float softmax_rem_to_p(uint32_t t_addr, uint32_t p_row_addr, int chunk0, int sw, float* xch, int warp,
                                                  int lane, float& ms_out) {
  constexpr float sl2 = 0.125f * 1.4426950408889634f;
  static_assert(NCJ == 32 || NCJ == 48, "remainder column split is 32/32/32/48");
  uint32_t sreg[48];
  tmem_ld_32x32(t_addr, reinterpret_cast<uint32_t (&)[32]>(sreg[0]));
  if constexpr (NCJ == 48) tmem_ld_32x16(t_addr + 32, reinterpret_cast<uint32_t (&)[16]>(sreg[32]));
  tmem_ld_wait();
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < NCJ; j += 4) {
    m0 = fmaxf(m0, fmaxf(__uint_as_float(sreg[j]), __uint_as_float(sreg[j + 1])));
    m1 = fmaxf(m1, fmaxf(__uint_as_float(sreg[j + 2]), __uint_as_float(sreg[j + 3])));
  }
  const int l = lane & 15;
  if (lane < 16) xch[warp * 16 + lane] = fmaxf(m0, m1);
  asm volatile("bar.sync 1, 128;" ::: "memory");
  const float mx = fmaxf(fmaxf(xch[l], xch[16 + l]), fmaxf(xch[32 + l], xch[48 + l]));
  const float ms = mx * sl2;
  ms_out = ms;
  float sum = 0.f;
#pragma unroll
  for (int g = 0; g < NCJ / 8; ++g) {
    float p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = ex2f(fmaf(__uint_as_float(sreg[8 * g + j]), sl2, -ms));
    sum += ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
    const uint4 u = make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
    const int ch = chunk0 + g;
    if (lane < 16) sts_u4(p_row_addr + static_cast<uint32_t>(ch >> 3) * 2048u + static_cast<uint32_t>(((ch & 7) ^ sw) << 4), u);
  }
  if (lane < 16) xch[64 + warp * 16 + lane] = sum;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  return (xch[64 + l] + xch[80 + l]) + (xch[96 + l] + xch[112 + l]);
}

// Warp reduction of 16 per-lane values at once (value q of every lane -> one result per q): a butterfly that halves the value
// count while it doubles the lanes folded in - 8 + 4 + 2 + 1 exchanges, then one across the last lane bit - 16 shuffles instead
// of 16 x 5.  On return lanes 2q and 2q + 1 hold the reduction of value q over the 32 lanes.
template <bool MAX>
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
  auto op = [](float a, float b) { return MAX ? fmaxf(a, b) : a + b; };
  float w8[8], w4[4], w2[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = h16 ? v[i] : v[i + 8], keep = h16 ? v[i + 8] : v[i];
    w8[i] = op(keep, __shfl_xor_sync(0xffffffffu, send, 16));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = h8 ? w8[i] : w8[i + 4], keep = h8 ? w8[i + 4] : w8[i];
    w4[i] = op(keep, __shfl_xor_sync(0xffffffffu, send, 8));
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = h4 ? w4[i] : w4[i + 2], keep = h4 ? w4[i + 2] : w4[i];
    w2[i] = op(keep, __shfl_xor_sync(0xffffffffu, send, 4));
  }
  const float send = h2 ? w2[0] : w2[1], keep = h2 ? w2[1] : w2[0];
  const float w1 = op(keep, __shfl_xor_sync(0xffffffffu, send, 2));
  return op(w1, __shfl_xor_sync(0xffffffffu, w1, 1));       // value q = 8 h16 + 4 h8 + 2 h4 + h2 = lane >> 1
}

// Transposed remainder (T = 144, opt-in - measured slower than the split form below, kept as the A/B): the scores of query rows 128..143 are produced as S_rem^T = K Q_rem^T - TMEM lane =
// key, 16 columns = the remainder queries; a second group over the K rows shifted by 16 puts keys 128..143 into lanes 112..127 -
// in 32 columns of their own, so their MMAs are issued right behind the main score tile instead of after its softmax has
// released the columns (the untransposed form left every softmax warp waiting ~1,400 cycles per unit for them).  Each thread
// then holds one key x 16 queries: row maximum and row sum are warp shuffles + one exchange through `xch`; the probabilities
// go out as 2-byte stores into the compact K-major P tile (row = query, 2 KB per 64-key block), contiguous per warp and query.
// Returns the row sum / sets ms_out for query (lane & 15).
template <int T>
__device__ __forceinline__ float softmax_rem_t_to_p(const uint32_t (&a)[16], const uint32_t (&b)[16], uint32_t sP1, float* xch, int warp,
                                                    int lane, float& ms_out) {
  // a, b: the thread's 16 + 16 score columns, read out of TMEM by the caller BEFORE it hands P0 to the MMA warp (a tcgen05.ld
  // issued while the O0 MMAs run waits for them)
  constexpr float sl2 = 0.125f * 1.4426950408889634f;
  static_assert(T == 144, "transposed remainder: 128 + 16 tokens");
  const bool tail = warp == 3 && lane >= 16;                 // lanes 112..127 of the shifted group hold keys 128..143
  float m[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    m[q] = __uint_as_float(a[q]);
    if (tail) m[q] = fmaxf(m[q], __uint_as_float(b[q]));
  }
  const float mq = warp_reduce16<true>(m, lane);
  if ((lane & 1) == 0) xch[warp * 16 + (lane >> 1)] = mq;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  float ms[16];
#pragma unroll
  for (int q = 0; q < 16; q += 4) {
    const float4 x0 = *reinterpret_cast<const float4*>(xch + q), x1 = *reinterpret_cast<const float4*>(xch + 16 + q);
    const float4 x2 = *reinterpret_cast<const float4*>(xch + 32 + q), x3 = *reinterpret_cast<const float4*>(xch + 48 + q);
    ms[q] = fmaxf(fmaxf(x0.x, x1.x), fmaxf(x2.x, x3.x)) * sl2; ms[q + 1] = fmaxf(fmaxf(x0.y, x1.y), fmaxf(x2.y, x3.y)) * sl2;
    ms[q + 2] = fmaxf(fmaxf(x0.z, x1.z), fmaxf(x2.z, x3.z)) * sl2; ms[q + 3] = fmaxf(fmaxf(x0.w, x1.w), fmaxf(x2.w, x3.w)) * sl2;
  }
  const int key = warp * 32 + lane, key2 = 128 + (lane & 15);
  // element (q, key) of the compact P tile: block key >> 6, row q (128 B), 16-byte chunk ((key & 63) >> 3) ^ (q & 7)
  const uint32_t base1 = sP1 + static_cast<uint32_t>(key >> 6) * 2048u + static_cast<uint32_t>(key & 7) * 2u;
  const uint32_t base2 = sP1 + 2u * 2048u + static_cast<uint32_t>(key2 & 7) * 2u;
  const uint32_t c1 = static_cast<uint32_t>((key & 63) >> 3), c2 = static_cast<uint32_t>((key2 & 63) >> 3);
  float sm[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float p = ex2f(fmaf(__uint_as_float(a[q]), sl2, -ms[q]));
    sm[q] = p;
    const __nv_bfloat16 hb = __float2bfloat16_rn(p);
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(base1 + static_cast<uint32_t>(q) * 128u + ((c1 ^ static_cast<uint32_t>(q & 7)) << 4)),
                 "h"(*reinterpret_cast<const uint16_t*>(&hb)) : "memory");
    if (tail) {
      const float p2 = ex2f(fmaf(__uint_as_float(b[q]), sl2, -ms[q]));
      sm[q] += p2;
      const __nv_bfloat16 h2 = __float2bfloat16_rn(p2);
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(base2 + static_cast<uint32_t>(q) * 128u + ((c2 ^ static_cast<uint32_t>(q & 7)) << 4)),
                   "h"(*reinterpret_cast<const uint16_t*>(&h2)) : "memory");
    }
  }
  const float sq = warp_reduce16<false>(sm, lane);
  if ((lane & 1) == 0) xch[64 + warp * 16 + (lane >> 1)] = sq;
  const int l = lane & 15;
  ms_out = fmaxf(fmaxf(xch[l], xch[16 + l]), fmaxf(xch[32 + l], xch[48 + l])) * sl2;
  asm volatile("bar.sync 1, 128;" ::: "memory");
  return (xch[64 + l] + xch[80 + l]) + (xch[96 + l] + xch[112 + l]);
}

// O tile rows of this warp (TMEM lane = row, 64 fp32 columns) * inv -> bf16 -> global.  One thread per row would make every
// store instruction touch 32 different lines, so the 32 x 128 B tile is transposed through a 4 KB staging tile (XOR-swizzled
// 16-byte chunks, conflict free both ways) and leaves as 4 full 128-byte rows per instruction.
__device__ __forceinline__ void load_o_row(uint32_t t_row, uint32_t (&a)[32], uint32_t (&b)[32]) {
  tmem_ld_32x32(t_row, a);
  tmem_ld_32x32(t_row + 32, b);
  tmem_ld_wait();
}
__device__ __forceinline__ void store_o_rows(const uint32_t (&a)[32], const uint32_t (&b)[32], float inv, uint32_t stage,
                                             __nv_bfloat16* dst_row0, int live_rows, int lane) {
  const uint32_t mine = stage + lane * 128, sw = lane & 7;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sts_u4(mine + ((j ^ sw) << 4),
           make_uint4(pack_bf16(__uint_as_float(a[8 * j]) * inv, __uint_as_float(a[8 * j + 1]) * inv),
                      pack_bf16(__uint_as_float(a[8 * j + 2]) * inv, __uint_as_float(a[8 * j + 3]) * inv),
                      pack_bf16(__uint_as_float(a[8 * j + 4]) * inv, __uint_as_float(a[8 * j + 5]) * inv),
                      pack_bf16(__uint_as_float(a[8 * j + 6]) * inv, __uint_as_float(a[8 * j + 7]) * inv)));
    sts_u4(mine + (((4 + j) ^ sw) << 4),
           make_uint4(pack_bf16(__uint_as_float(b[8 * j]) * inv, __uint_as_float(b[8 * j + 1]) * inv),
                      pack_bf16(__uint_as_float(b[8 * j + 2]) * inv, __uint_as_float(b[8 * j + 3]) * inv),
                      pack_bf16(__uint_as_float(b[8 * j + 4]) * inv, __uint_as_float(b[8 * j + 5]) * inv),
                      pack_bf16(__uint_as_float(b[8 * j + 6]) * inv, __uint_as_float(b[8 * j + 7]) * inv)));
  }
  __syncwarp();
  const int sub = lane >> 3, ch = lane & 7;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = i * 4 + sub;
    uint4 u;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                 : "r"(stage + r * 128 + ((ch ^ (r & 7)) << 4)));
    if (r < live_rows) *reinterpret_cast<uint4*>(dst_row0 + static_cast<long long>(r) * kHidden + ch * 8) = u;
  }
  __syncwarp();
}

// TRACE (developer path, JPDVT_ATTN_TRACE=1): lane 0 of the MMA warp and of two softmax warps of CTA 0 log clock64() at the
// pipeline events of the first units into `trace` [role][unit][event]
constexpr int kTraceUnits = 6, kTraceEvents = 10, kTraceRoles = 4;   // roles: MMA warp, softmax warp 0, softmax warp 3, TMA warp

// RT: transposed remainder scores (T = 144), see softmax_rem_t_to_p.  R2 (T = 144, split remainder): the 16 small remainder
// MMAs are issued by TWO threads - groups 0, 1 by the MMA warp, groups 2, 3 by the TMA warp's lane - because issuing them
// (60-90 cycles of single-thread work each) is what the softmax warps wait for between the two tiles.
template <int T, bool TRACE, bool RT = false, bool R2 = false>
__global__ void __launch_bounds__(kTcThreads, TcCfg<T>::kCtasPerSm)
attention_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse2,
                    int num_units, int reverse, long long* __restrict__ trace, int l2_prefetch) {
  using Cfg = TcCfg<T>;
  auto mark = [&](int role, int it, int ev) {
    if constexpr (TRACE) {
      if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && it < kTraceUnits && role >= 0)
        trace[(role * kTraceUnits + it) * kTraceEvents + ev] = clock64();
    }
  };
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
  uint64_t* pb_full = bars + 10;       // [2] R2: the first / second 64-key block of P0 is in shared memory (4 arrivals)
  float* xch = reinterpret_cast<float*>(smem + Cfg::kXchOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(qk_full, 1); mbar_init(v_full, 1);
    for (int t = 0; t < 2; ++t) { mbar_init(&s_full[t], (t == 1 && R2) ? 2 : 1); mbar_init(&p_full[t], 4); mbar_init(&o_full[t], 1); }
    mbar_init(epi_done, 4);
    mbar_init(&pb_full[0], 4); mbar_init(&pb_full[1], 4);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  if (warp == 4 && lane == 0) tma_prefetch_desc(&tm_qkv);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();                  // the prologue above may overlap the previous kernel's tail (programmatic dependent launch)
  griddep_launch_dependents();
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV),
                 sP = smem_u32(smem + Cfg::kOffP), sP1 = smem_u32(smem + Cfg::kOffP1);
  constexpr int kLast = Cfg::kTiles - 1;

  if (warp == 4) {
    // ---------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int it = 0;
      [[maybe_unused]] auto issue_rem_half = [&](uint32_t ph_unit) {   // R2: quadrants 2, 3 <- rows 128..143 x keys [64, 96), [96, 144)
        constexpr uint32_t idesc_a = umma_idesc_bf16(128, 32), idesc_b = umma_idesc_bf16(128, T - 96);
        const uint32_t q_lo = desc_lo_k(sQ), k_lo = desc_lo_k(sK);
        mbar_wait(&p_full[0], ph_unit);                        // the softmax has consumed the main score tile
        tc_fence_after();
#pragma unroll
        for (int j = 2; j < 4; ++j) {
          const uint32_t idesc_j = j < 3 ? idesc_a : idesc_b;
          const uint32_t a0 = q_lo + (128 - 32 * j) * 8, b0 = k_lo + 32 * j * 8;
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
            if (k == 0) umma_lohi<false>(tmem_base + 32 * j, a0, b0, idesc_j);
            else umma_lohi<true>(tmem_base + 32 * j, a0 + 2 * k, b0 + 2 * k, idesc_j);
          }
        }
        umma_commit(&s_full[1]);
      };
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int uu = reverse ? num_units - 1 - unit : unit;   // sweep direction: see sweep_reverse() in common.cuh
        const int b = uu / kHeads, h = uu - b * kHeads;
        const uint32_t prev = static_cast<uint32_t>((it - 1) & 1);
        mark(3, it, 0);
        if constexpr (R2) {
          if (it > 0) issue_rem_half(prev);                   // groups 2, 3 of the previous unit's remainder scores
        }
        if (it > 0) mbar_wait(&s_full[kLast], prev);          // every score MMA of the previous unit has read Q, K
        mark(3, it, 1);
        mbar_expect_tx(qk_full, 2 * Cfg::kTileBytes);
        tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffQ, h * kHeadDim, b * T);
        tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffK, kHidden + h * kHeadDim, b * T);
        mark(3, it, 2);
        if (it > 0) mbar_wait(&o_full[kLast], prev);          // ... and every P V MMA has read V
        mark(3, it, 3);
        mbar_expect_tx(v_full, Cfg::kTileBytes);
        tma_load_2d(&tm_qkv, v_full, smem + Cfg::kOffV, 2 * kHidden + h * kHeadDim, b * T);
        // the next unit's Q / K / V towards L2 now: their loads are only issued when this unit's MMAs have released the
        // buffers, and that latency sits on the unit chain (2,860 cycles from HBM under load)
        if (l2_prefetch && unit + static_cast<int>(gridDim.x) < num_units) {
          const int nu = reverse ? num_units - 1 - (unit + static_cast<int>(gridDim.x)) : unit + static_cast<int>(gridDim.x);
          const int nb = nu / kHeads, nh = nu - nb * kHeads;
          tma_prefetch_2d(&tm_qkv, nh * kHeadDim, nb * T);
          tma_prefetch_2d(&tm_qkv, kHidden + nh * kHeadDim, nb * T);
          tma_prefetch_2d(&tm_qkv, 2 * kHidden + nh * kHeadDim, nb * T);
        }
      }
      if constexpr (R2) {
        if (it > 0) issue_rem_half(static_cast<uint32_t>((it - 1) & 1));   // the last unit's
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, T);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);   // B = V, MN-major (keys are the strided index)
      const uint32_t q_lo = desc_lo_k(sQ), k_lo = desc_lo_k(sK), p_lo = desc_lo_k(sP), v_lo = desc_lo_mn(sV);
      const uint32_t p1_lo = (Cfg::kSplit || Cfg::kDual) ? desc_lo_k(sP1) : p_lo;
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const uint32_t ph = static_cast<uint32_t>(it & 1);
        mark(0, it, 0);
        mbar_wait(qk_full, ph);
        mark(0, it, 1);
        if (it > 0) mbar_wait(epi_done, static_cast<uint32_t>((it - 1) & 1));   // O of the previous unit has left TMEM
        mark(0, it, 2);
        tc_fence_after();
        auto issue_o = [&](uint32_t d_col, uint32_t pa_lo, uint32_t blk_words, int j0 = 0, int j1 = T / 16) {   // blk_words = P block stride / 16 bytes
#pragma unroll
          for (int j = 0; j < T / 16; ++j) {
            if (j < j0 || j >= j1) continue;
            const uint32_t a = pa_lo + (j >> 2) * blk_words + (j & 3) * 2, bq = v_lo + j * 128;
            if (j == 0) umma_lohi<false>(tmem_base + d_col, a, bq, idesc_o);
            else umma_lohi<true>(tmem_base + d_col, a, bq, idesc_o);
          }
        };
        constexpr bool kPB = Cfg::kSplit && R2;                // P V k-steps of tile 0 issued per 64-key block of P0
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          if (k == 0) umma_lohi<false>(tmem_base, q_lo, k_lo, idesc_s);
          else umma_lohi<true>(tmem_base, q_lo + 2 * k, k_lo + 2 * k, idesc_s);
        }
        umma_commit(&s_full[0]);
        if constexpr (Cfg::kSplit && RT) {                     // remainder scores, transposed, into 32 columns of their own
          constexpr uint32_t idesc_r = umma_idesc_bf16(128, 16);
#pragma unroll
          for (int gq = 0; gq < 2; ++gq) {                     // keys 0..127, then keys 16..143 (lanes 112..127 = keys 128..143)
            const uint32_t a0 = k_lo + gq * 16 * 8, b0 = q_lo + 128 * 8;
#pragma unroll
            for (int k = 0; k < kHeadDim / 16; ++k) {
              if (k == 0) umma_lohi<false>(tmem_base + (T + 64) + 16 * gq, a0, b0, idesc_r);
              else umma_lohi<true>(tmem_base + (T + 64) + 16 * gq, a0 + 2 * k, b0 + 2 * k, idesc_r);
            }
          }
          umma_commit(&s_full[1]);
        }
        if constexpr (Cfg::kDual) {                            // second score tile right behind the first, into its own columns
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
            if (k == 0) umma_lohi<false>(tmem_base + Cfg::kColS1, q_lo + 128 * 8, k_lo, idesc_s);
            else umma_lohi<true>(tmem_base + Cfg::kColS1, q_lo + 128 * 8 + 2 * k, k_lo + 2 * k, idesc_s);
          }
          umma_commit(&s_full[1]);
        }
        mark(0, it, 3);
        if constexpr (kPB) {                                   // blocks 0 and 1 of P0 while the softmax is still on the later keys
          mbar_wait(v_full, ph);
          mbar_wait(&pb_full[0], ph);
          tc_fence_after();
          issue_o(Cfg::kColO0, p_lo, 1024, 0, 4);
          mbar_wait(&pb_full[1], ph);
          tc_fence_after();
          issue_o(Cfg::kColO0, p_lo, 1024, 4, 8);
        }
        mbar_wait(&p_full[0], ph);                            // softmax has consumed S and written P0
        mark(0, it, 4);
        tc_fence_after();
        if constexpr (Cfg::kTiles == 2 && !Cfg::kDual && !(Cfg::kSplit && RT)) {   // remainder scores first: the softmax warps wait for them
          if constexpr (Cfg::kSplit) {
#pragma unroll
            for (int j = 0; j < (R2 ? 2 : 4); ++j) {            // quadrant j <- rows 128..143 x keys [32j, 32j + n_j)
              constexpr uint32_t idesc_a = umma_idesc_bf16(128, 32), idesc_b = umma_idesc_bf16(128, T - 96);
              const uint32_t idesc_j = j < 3 ? idesc_a : idesc_b;
              const uint32_t a0 = q_lo + (128 - 32 * j) * 8, b0 = k_lo + 32 * j * 8;
#pragma unroll
              for (int k = 0; k < kHeadDim / 16; ++k) {
                if (k == 0) umma_lohi<false>(tmem_base + 32 * j, a0, b0, idesc_j);
                else umma_lohi<true>(tmem_base + 32 * j, a0 + 2 * k, b0 + 2 * k, idesc_j);
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < kHeadDim / 16; ++k) {
              if (k == 0) umma_lohi<false>(tmem_base, q_lo + 128 * 8, k_lo, idesc_s);
              else umma_lohi<true>(tmem_base, q_lo + 128 * 8 + 2 * k, k_lo + 2 * k, idesc_s);
            }
          }
          umma_commit(&s_full[1]);
        }
        mark(0, it, 5);
        if constexpr (!kPB) mbar_wait(v_full, ph);
        mark(0, it, 6);
        if constexpr (kPB) issue_o(Cfg::kColO0, p_lo, 1024, 8, T / 16);
        else issue_o(Cfg::kColO0, p_lo, 1024);
        umma_commit(&o_full[0]);
        mark(0, it, 7);
        if constexpr (Cfg::kTiles == 2) {
          mbar_wait(&p_full[1], ph);
          mark(0, it, 8);
          tc_fence_after();
          issue_o(Cfg::kColO1, p1_lo, Cfg::kSplit ? 128 : 1024);
          umma_commit(&o_full[1]);
          mark(0, it, 9);
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
      const int uu = reverse ? num_units - 1 - unit : unit;
      const int b = uu / kHeads, h = uu - b * kHeads;
      __nv_bfloat16* obase = out + static_cast<long long>(b) * T * kHidden + h * kHeadDim;
      const int role = (warp == 0) ? 1 : (warp == 3) ? 2 : -1;
      // ---- tile 0
      mark(role, it, 0);
      mbar_wait(&s_full[0], ph);
      mark(role, it, 1);
      tc_fence_after();
      float ms0, ms1 = 0.f;
      const float sum0 = softmax_row_to_p<T, T, Cfg::kSplit && R2>(t_lane, sP + r_tile * 128, 16384, r_tile & 7, true, ms0, pb_full, lane);
      [[maybe_unused]] uint32_t rem_a[16], rem_b[16];
      if constexpr (Cfg::kSplit && RT) {                      // transposed remainder scores: long since in TMEM, read them now
        mbar_wait(&s_full[1], ph);
        tc_fence_after();
        tmem_ld_32x16(t_lane + (T + 64), rem_a);
        tmem_ld_32x16(t_lane + (T + 64) + 16, rem_b);
        tmem_ld_wait();
      }
      fence_proxy_async_smem();                               // generic-proxy stores -> visible to the tensor core's reads
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[0]);
      mark(role, it, 2);
      // ---- tile 1 (the 16- or 128-row remainder)
      float sum1 = 1.f;
      if constexpr (Cfg::kTiles == 2) {
        mbar_wait(&s_full[1], ph);
        mark(role, it, 3);
        if constexpr (!Cfg::kSplit && !Cfg::kDual) mbar_wait(&o_full[0], ph);   // P1 would reuse P0: the O0 MMAs must have read it
        mark(role, it, 4);
        tc_fence_after();
        if constexpr (Cfg::kSplit && RT) {
          sum1 = softmax_rem_t_to_p<T>(rem_a, rem_b, sP1, xch, warp, lane, ms1);
        } else if constexpr (Cfg::kSplit) {
          if (warp < 3) sum1 = softmax_rem_to_p<32>(t_lane + 32 * warp, sP1 + lane * 128, 4 * warp, lane & 7, xch, warp, lane, ms1);
          else sum1 = softmax_rem_to_p<T - 96>(t_lane + 96, sP1 + lane * 128, 12, lane & 7, xch, warp, lane, ms1);
        } else {
          sum1 = softmax_row_to_p<T>(t_lane + Cfg::kColS1, (Cfg::kDual ? sP1 : sP) + r_tile * 128, 16384, r_tile & 7, true, ms1);
        }
        fence_proxy_async_smem();
      }
      if (lse2 != nullptr) {   // training: log2-domain log-sum-exp of the scaled scores, [B, 12, T] (read by the attention backward)
        float* lrow = lse2 + (static_cast<long long>(b) * kHeads + h) * T;
        if (r_tile < T) lrow[r_tile] = ms0 + log2f(sum0);
        if constexpr (Cfg::kTiles == 2) {
          if constexpr (Cfg::kSplit) { if (warp == 0 && lane < 16) lrow[128 + lane] = ms1 + log2f(sum1); }
          else if (128 + r_tile < T) lrow[128 + r_tile] = ms1 + log2f(sum1);
        }
      }
      // ---- outputs.  O0 is pulled into registers BEFORE the remainder's P tile is handed to the MMA warp: a tcgen05.ld
      // issued while the O1 MMAs run waits for them, so the other order parks every softmax warp behind the tensor pipe
      const uint32_t o_stage = sP + static_cast<uint32_t>(warp) * 4096u;                     // P0 is idle after o_full[0]
      const uint32_t o_stage1 = (Cfg::kDual ? sP1 : sP) + static_cast<uint32_t>(warp) * 4096u;   // dual: P1 after o_full[1]
      uint32_t oa[32], ob[32];
      mbar_wait(&o_full[0], ph);
      tc_fence_after();
      load_o_row(t_lane + Cfg::kColO0, oa, ob);
      if constexpr (Cfg::kTiles == 2) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[1]);
        mark(role, it, 5);
      }
      store_o_rows(oa, ob, 1.0f / sum0, o_stage, obase + static_cast<long long>(warp * 32) * kHidden,
                   T - warp * 32 < 32 ? T - warp * 32 : 32, lane);
      if constexpr (Cfg::kTiles == 2) {
        if (!Cfg::kSplit || warp == 0) {                       // split remainder: O1 rows 128..143 sit in lanes 0..15 of quadrant 0
          mark(role, it, 6);
          mbar_wait(&o_full[1], ph);
          mark(role, it, 7);
          tc_fence_after();
          load_o_row(t_lane + Cfg::kColO1, oa, ob);
          store_o_rows(oa, ob, 1.0f / sum1, o_stage1, obase + static_cast<long long>(Cfg::kSplit ? 128 : 128 + warp * 32) * kHidden,
                       Cfg::kSplit ? 16 : 32, lane);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_done);
      mark(role, it, 8);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// T = 144 with EIGHT softmax warps per CTA (opt-in, JPDVT_ATTN_WARPS=8 - built on the hypothesis that the four-warp kernel is
// bound by the latency of its math warps' instruction streams; parity-green and SLOWER, 57.8 vs 54.0 us at B = 256, so that is
// not the bound either: DESIGN.md section 4).  Warps w
// and w + 4 share a TMEM lane quadrant - the same 32 query rows - and split the COLUMNS: keys [0,80) / [80,144) of the score
// row (row maximum and row sum meet through shared memory), 16-key halves of each remainder group, columns [0,32) / [32,64)
// of the output rows.  Twice the warps per scheduler, half the work per warp.  Layout, barriers and MMA / TMA roles are the
// four-warp kernel's (split remainder); warps 0-7 softmax, warp 8 TMA, warp 9 MMA.
constexpr int kTc8Threads = 320;
struct Tc8Cfg {
  using Base = TcCfg<144>;
  static constexpr int T = 144;
  static constexpr int kXchOff = Base::kBarOff + 128;
  // floats: row max [2][128], row sum [2][128] of the main tile; row max [8][16], row sum [8][16] of the remainder
  static constexpr int kXchBytes = (4 * 128 + 2 * 8 * 16) * 4;
  static constexpr int kSmemBytes = kXchOff + kXchBytes + 1024;
  static_assert(2 * (kSmemBytes + 1024) <= 227 * 1024, "two CTAs per SM");
};

__global__ void __launch_bounds__(kTc8Threads, 2)
attention_tc8_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse2,
                     int num_units, int reverse) {
  using Cfg = TcCfg<144>;
  constexpr int T = 144;
  constexpr float sl2 = 0.125f * 1.4426950408889634f;
  extern __shared__ uint8_t att_tc_smem[];
  uint8_t* smem = att_tc_smem + ((1024u - (smem_u32(att_tc_smem) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* qk_full = bars + 0;
  uint64_t* v_full = bars + 1;
  uint64_t* s_full = bars + 2;         // [2]
  uint64_t* p_full = bars + 4;         // [2] 8 arrivals
  uint64_t* o_full = bars + 6;         // [2]
  uint64_t* epi_done = bars + 8;       // 8 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* xmax = reinterpret_cast<float*>(smem + Tc8Cfg::kXchOff);   // [2][128]
  float* xsum = xmax + 256;                                        // [2][128]
  float* rmax = xsum + 256;                                        // [8][16]
  float* rsum = rmax + 128;                                        // [8][16]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(qk_full, 1); mbar_init(v_full, 1);
    for (int t = 0; t < 2; ++t) { mbar_init(&s_full[t], 1); mbar_init(&p_full[t], 8); mbar_init(&o_full[t], 1); }
    mbar_init(epi_done, 8);
    fence_mbar_init();
  }
  if (warp == 9) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  if (warp == 8 && lane == 0) tma_prefetch_desc(&tm_qkv);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch_dependents();
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV),
                 sP = smem_u32(smem + Cfg::kOffP), sP1 = smem_u32(smem + Cfg::kOffP1);

  if (warp == 8) {
    // ---------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int uu = reverse ? num_units - 1 - unit : unit;
        const int b = uu / kHeads, h = uu - b * kHeads;
        const uint32_t prev = static_cast<uint32_t>((it - 1) & 1);
        if (it > 0) mbar_wait(&s_full[1], prev);              // every score MMA of the previous unit has read Q, K
        mbar_expect_tx(qk_full, 2 * Cfg::kTileBytes);
        tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffQ, h * kHeadDim, b * T);
        tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffK, kHidden + h * kHeadDim, b * T);
        if (it > 0) mbar_wait(&o_full[1], prev);              // ... and every P V MMA has read V
        mbar_expect_tx(v_full, Cfg::kTileBytes);
        tma_load_2d(&tm_qkv, v_full, smem + Cfg::kOffV, 2 * kHidden + h * kHeadDim, b * T);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ---------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, T);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);
      constexpr uint32_t idesc_a = umma_idesc_bf16(128, 32), idesc_b = umma_idesc_bf16(128, T - 96);
      const uint32_t q_lo = desc_lo_k(sQ), k_lo = desc_lo_k(sK), p_lo = desc_lo_k(sP), v_lo = desc_lo_mn(sV), p1_lo = desc_lo_k(sP1);
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const uint32_t ph = static_cast<uint32_t>(it & 1);
        mbar_wait(qk_full, ph);
        if (it > 0) mbar_wait(epi_done, static_cast<uint32_t>((it - 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          if (k == 0) umma_lohi<false>(tmem_base, q_lo, k_lo, idesc_s);
          else umma_lohi<true>(tmem_base, q_lo + 2 * k, k_lo + 2 * k, idesc_s);
        }
        umma_commit(&s_full[0]);
        mbar_wait(&p_full[0], ph);                            // the main score tile is consumed, P0 is in shared memory
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 4; ++j) {                         // quadrant j <- rows 128..143 x keys [32j, 32j + n_j)
          const uint32_t idesc_j = j < 3 ? idesc_a : idesc_b;
          const uint32_t a0 = q_lo + (128 - 32 * j) * 8, b0 = k_lo + 32 * j * 8;
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
            if (k == 0) umma_lohi<false>(tmem_base + 32 * j, a0, b0, idesc_j);
            else umma_lohi<true>(tmem_base + 32 * j, a0 + 2 * k, b0 + 2 * k, idesc_j);
          }
        }
        umma_commit(&s_full[1]);
        mbar_wait(v_full, ph);
#pragma unroll
        for (int j = 0; j < T / 16; ++j) {
          const uint32_t a = p_lo + (j >> 2) * 1024 + (j & 3) * 2, bq = v_lo + j * 128;
          if (j == 0) umma_lohi<false>(tmem_base + Cfg::kColO0, a, bq, idesc_o);
          else umma_lohi<true>(tmem_base + Cfg::kColO0, a, bq, idesc_o);
        }
        umma_commit(&o_full[0]);
        mbar_wait(&p_full[1], ph);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < T / 16; ++j) {
          const uint32_t a = p1_lo + (j >> 2) * 128 + (j & 3) * 2, bq = v_lo + j * 128;
          if (j == 0) umma_lohi<false>(tmem_base + Cfg::kColO1, a, bq, idesc_o);
          else umma_lohi<true>(tmem_base + Cfg::kColO1, a, bq, idesc_o);
        }
        umma_commit(&o_full[1]);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- softmax + epilogue
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int r_tile = quad * 32 + lane;
    const int nch = half ? 4 : 5, cbase = half ? 80 : 0;      // this warp's 16-column chunks of the score row
    const int pair_bar = 3 + quad;                            // named barrier of the two warps that share a lane quadrant
    const int sw = r_tile & 7;
    const uint32_t p_row = sP + r_tile * 128;
    int it = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const uint32_t ph = static_cast<uint32_t>(it & 1);
      const int uu = reverse ? num_units - 1 - unit : unit;
      const int b = uu / kHeads, h = uu - b * kHeads;
      __nv_bfloat16* obase = out + static_cast<long long>(b) * T * kHidden + h * kHeadDim;
      // ---- main tile: row maximum over this warp's columns, exchanged with the sibling warp
      mbar_wait(&s_full[0], ph);
      tc_fence_after();
      uint32_t ra[16], rb[16];
      float m0 = -INFINITY, m1 = -INFINITY;
      tmem_ld_32x16(t_lane + cbase, ra);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        if (c < nch) {
          uint32_t (&cur)[16] = (c & 1) ? rb : ra;
          uint32_t (&nxt)[16] = (c & 1) ? ra : rb;
          if (c + 1 < nch) tmem_ld_32x16(t_lane + cbase + 16 * (c + 1), nxt);
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(cur[j + 2]), __uint_as_float(cur[j + 3])));
          }
          if (c + 1 < nch) tmem_ld_wait();
        }
      }
      tmem_ld_32x16(t_lane + cbase, ra);                      // second pass: its first chunk is in flight during the exchange
      xmax[half * 128 + r_tile] = fmaxf(m0, m1);
      asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
      const float ms0 = fmaxf(xmax[r_tile], xmax[128 + r_tile]) * sl2;
      // ---- probabilities of this warp's columns -> P0 (bf16, K-major swizzled), partial row sum
      uint64_t sum2 = f2_pack(0.f, 0.f);
      const uint64_t sl2p = f2_pack(sl2, sl2), nmsp = f2_pack(-ms0, -ms0);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 5; ++c) {
        if (c < nch) {
          uint32_t (&cur)[16] = (c & 1) ? rb : ra;
          uint32_t (&nxt)[16] = (c & 1) ? ra : rb;
          if (c + 1 < nch) tmem_ld_32x16(t_lane + cbase + 16 * (c + 1), nxt);
#pragma unroll
          for (int g = 0; g < 2; ++g) {                       // 8 keys -> one 16-byte chunk of the P row
            float p[8];
            uint64_t e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float a, bb;
              f2_unpack(f2_fma(f2_pack(__uint_as_float(cur[8 * g + 2 * j]), __uint_as_float(cur[8 * g + 2 * j + 1])), sl2p, nmsp), a, bb);
              p[2 * j] = ex2f(a); p[2 * j + 1] = ex2f(bb);
              e[j] = f2_pack(p[2 * j], p[2 * j + 1]);
            }
            sum2 = f2_add(sum2, f2_add(f2_add(e[0], e[1]), f2_add(e[2], e[3])));
            const int chunk = (cbase >> 3) + 2 * c + g;
            sts_u4(p_row + static_cast<uint32_t>(chunk >> 3) * 16384u + static_cast<uint32_t>(((chunk & 7) ^ sw) << 4),
                   make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7])));
          }
          if (c + 1 < nch) tmem_ld_wait();
        }
      }
      {
        float s_even, s_odd;
        f2_unpack(sum2, s_even, s_odd);
        xsum[half * 128 + r_tile] = s_even + s_odd;            // read after the remainder's CTA-wide barriers below
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[0]);
      // ---- remainder rows 128..143 (lanes 0..15 of every quadrant): group `quad` = keys [32 quad, 32 quad + n), split in two
      mbar_wait(&s_full[1], ph);
      tc_fence_after();
      const int rc0 = 32 * quad + 16 * half;                  // first key of this warp's slice; quad 3 / half 1 takes 32 keys
      const bool wide = (quad == 3 && half == 1);
      uint32_t sa[16], sb2[16];
      tmem_ld_32x16(t_lane + rc0, sa);
      if (wide) tmem_ld_32x16(t_lane + rc0 + 16, sb2);
      tmem_ld_wait();
      float rm0 = -INFINITY, rm1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        rm0 = fmaxf(rm0, __uint_as_float(sa[j])); rm1 = fmaxf(rm1, __uint_as_float(sa[j + 1]));
        if (wide) { rm0 = fmaxf(rm0, __uint_as_float(sb2[j])); rm1 = fmaxf(rm1, __uint_as_float(sb2[j + 1])); }
      }
      const int l = lane & 15;
      if (lane < 16) rmax[warp * 16 + lane] = fmaxf(rm0, rm1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float mx = rmax[l];
#pragma unroll
      for (int w = 1; w < 8; ++w) mx = fmaxf(mx, rmax[w * 16 + l]);
      const float ms1 = mx * sl2;
      float rs = 0.f;
      {
        const uint32_t p1_row = sP1 + lane * 128;
        const int swr = lane & 7;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          float p[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) p[j] = ex2f(fmaf(__uint_as_float(sa[8 * g + j]), sl2, -ms1));
          rs += ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
          const int ch = (rc0 >> 3) + g;
          if (lane < 16) sts_u4(p1_row + static_cast<uint32_t>(ch >> 3) * 2048u + static_cast<uint32_t>(((ch & 7) ^ swr) << 4),
                                make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7])));
        }
        if (wide) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float p[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) p[j] = ex2f(fmaf(__uint_as_float(sb2[8 * g + j]), sl2, -ms1));
            rs += ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
            const int ch = (rc0 >> 3) + 2 + g;
            if (lane < 16) sts_u4(p1_row + static_cast<uint32_t>(ch >> 3) * 2048u + static_cast<uint32_t>(((ch & 7) ^ swr) << 4),
                                  make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7])));
          }
        }
      }
      if (lane < 16) rsum[warp * 16 + lane] = rs;
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float sum1 = rsum[l];
#pragma unroll
      for (int w = 1; w < 8; ++w) sum1 += rsum[w * 16 + l];
      const float sum0 = xsum[r_tile] + xsum[128 + r_tile];     // both halves were written before the barriers above
      if (lse2 != nullptr) {
        float* lrow = lse2 + (static_cast<long long>(b) * kHeads + h) * T;
        if (half == 0) lrow[r_tile] = ms0 + log2f(sum0);
        if (warp == 0 && lane < 16) lrow[128 + lane] = ms1 + log2f(sum1);
      }
      // ---- outputs: this warp's 32 columns of its 32 rows.  O0 into registers BEFORE P1 is handed over (a tcgen05.ld issued
      // while the O1 MMAs run waits for them)
      uint32_t oa[32];
      mbar_wait(&o_full[0], ph);
      tc_fence_after();
      tmem_ld_32x32(t_lane + Cfg::kColO0 + 32 * half, oa);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[1]);
      auto store_half = [&](const uint32_t (&o)[32], float inv, __nv_bfloat16* dst_row0, int live_rows) {
        // 32 rows x 32 bf16 columns (64 B per row) through a 2 KB staging tile in the idle P0 buffer: 16-byte chunk c of row r
        // sits at r * 64 + ((c ^ (r >> 1)) & 3) * 16; every store instruction then writes 8 rows x 64 contiguous bytes
        const uint32_t stg = sP + static_cast<uint32_t>(warp) * 2048u;
        const uint32_t mine = stg + lane * 64;
        const uint32_t swz = static_cast<uint32_t>(lane >> 1) & 3u;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          sts_u4(mine + ((static_cast<uint32_t>(j) ^ swz) << 4),
                 make_uint4(pack_bf16(__uint_as_float(o[8 * j]) * inv, __uint_as_float(o[8 * j + 1]) * inv),
                            pack_bf16(__uint_as_float(o[8 * j + 2]) * inv, __uint_as_float(o[8 * j + 3]) * inv),
                            pack_bf16(__uint_as_float(o[8 * j + 4]) * inv, __uint_as_float(o[8 * j + 5]) * inv),
                            pack_bf16(__uint_as_float(o[8 * j + 6]) * inv, __uint_as_float(o[8 * j + 7]) * inv)));
        __syncwarp();
        const int sub = lane >> 2, ch = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = i * 8 + sub;
          uint4 u;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                       : "r"(stg + r * 64 + ((static_cast<uint32_t>(ch) ^ (static_cast<uint32_t>(r >> 1) & 3u)) << 4)));
          if (r < live_rows) *reinterpret_cast<uint4*>(dst_row0 + static_cast<long long>(r) * kHidden + 32 * half + ch * 8) = u;
        }
        __syncwarp();
      };
      store_half(oa, 1.0f / sum0, obase + static_cast<long long>(quad * 32) * kHidden, T - quad * 32 < 32 ? T - quad * 32 : 32);
      if (quad == 0) {                                         // O1 rows 128..143 sit in lanes 0..15 of quadrant 0: warps 0 and 4
        mbar_wait(&o_full[1], ph);
        tc_fence_after();
        tmem_ld_32x32(t_lane + Cfg::kColO1 + 32 * half, oa);
        tmem_ld_wait();
        store_half(oa, 1.0f / sum1, obase + static_cast<long long>(128) * kHidden, 16);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_done);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int launch_tc8(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int batch, cudaStream_t stream) {
  constexpr int T = 144;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attention_tc8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc8Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "attention_tc8: cudaFuncSetAttribute(smem=%d) failed: %s", Tc8Cfg::kSmemBytes,
                       cudaGetErrorString(cudaGetLastError()));
    cudaFuncSetAttribute(attention_tc8_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
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
  const int grid = units < 2 * sms ? units : 2 * sms;
  if (launch_pdl(attention_tc8_kernel, dim3(grid), dim3(kTc8Threads), Tc8Cfg::kSmemBytes, stream, tm, out, lse2, units, sweep_reverse()) != cudaSuccess)
    return set_error(kErrCuda, "attention_tc8_kernel: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return check_launch("attention_tc8_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// Longer sequences (T = 324 @288 px): keys padded to TP = 336 (a multiple of 16, columns >= T masked to probability 0), the
// score tile is 128 x TP fp32 = TP TMEM columns (two MMAs per k-step, N <= 256 each), so one score tile at a time: the
// ceil(T / 128) query tiles of a unit run one after the other, O in its own 64 columns, one CTA per SM.
template <int TP, int TV>
struct SeqCfg {
  static_assert(TP % 16 == 0 && TV <= TP && TP - TV < 16 && TP > 256 && TP + 64 <= 512, "padded key count");
  static constexpr int kTiles = (TV + 127) / 128;
  static constexpr int kN1 = (TP / 2 + 15) / 16 * 16, kN2 = TP - kN1;     // the two halves of the key range
  static constexpr int kKBlocks = (TP + 63) / 64;
  static constexpr int kTileBytes = TP * 128;                  // Q / K / V each (rows >= TV: whatever follows in memory, masked)
  static constexpr int kBoxRows = TP / 2;                      // TMA boxes are at most 256 rows: two loads per matrix
  static constexpr int kOffQ = 0, kOffK = kTileBytes, kOffV = 2 * kTileBytes, kOffP = 3 * kTileBytes;
  static constexpr int kPBytes = kKBlocks * 16384;
  static constexpr int kBarOff = kOffP + kPBytes;
  static constexpr int kSmemBytes = kBarOff + 128 + 1024;
  static_assert(kOffP % 1024 == 0 && kBoxRows <= 256 && kBoxRows % 8 == 0 && kN2 % 16 == 0 && kN1 <= 256, "layout");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory");
};

// PT: the probabilities go back into tensor memory over the consumed score columns (softmax_row_to_tmem) and the P V MMAs read
// them from there; the output of tile n - 1 is drained after the softmax of tile n (as in attention_rw_kernel)
template <int TP, int TV, bool PT = false>
__global__ void __launch_bounds__(kTcThreads, 1)
attention_tc_seq_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse2,
                        int num_units, int reverse) {
  using Cfg = SeqCfg<TP, TV>;
  constexpr int kTiles = Cfg::kTiles;
  extern __shared__ uint8_t att_tc_smem[];
  uint8_t* smem = att_tc_smem + ((1024u - (smem_u32(att_tc_smem) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* qk_full = bars + 0;        // TMA: Q and K landed                     (once per unit)
  uint64_t* v_full = bars + 1;         // TMA: V landed                           (once per unit)
  uint64_t* s_full = bars + 2;         // MMA: scores of the current tile ready   (once per tile)
  uint64_t* p_full = bars + 3;         // softmax warps: P written, S consumed    (once per tile, 4 arrivals)
  uint64_t* o_full = bars + 4;         // MMA: O of the current tile ready        (once per tile)
  uint64_t* o_read = bars + 5;         // softmax warps: O read out of TMEM       (once per tile, 4 arrivals)
  uint64_t* qk_free = bars + 6;        // MMA: every score MMA of the unit has read Q, K   (once per unit - the per-tile barriers
  uint64_t* v_free = bars + 7;         // MMA: every P V MMA of the unit has read V          cannot tell tile 0 from tile 2 by parity)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1); mbar_init(o_read, 4);
    mbar_init(qk_free, 1); mbar_init(v_free, 1);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  if (warp == 4 && lane == 0) tma_prefetch_desc(&tm_qkv);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV),
                 sP = smem_u32(smem + Cfg::kOffP);

  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int uu = reverse ? num_units - 1 - unit : unit;   // sweep direction: see sweep_reverse() in common.cuh
        const int b = uu / kHeads, h = uu - b * kHeads;
        if (it > 0) mbar_wait_backoff(qk_free, static_cast<uint32_t>((it - 1) & 1), 100);   // previous unit's score MMAs are done with Q, K
        mbar_expect_tx(qk_full, 2 * Cfg::kTileBytes);
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffQ + part * Cfg::kBoxRows * 128, h * kHeadDim, b * TV + part * Cfg::kBoxRows);
          tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffK + part * Cfg::kBoxRows * 128, kHidden + h * kHeadDim, b * TV + part * Cfg::kBoxRows);
        }
        if (it > 0) mbar_wait_backoff(v_free, static_cast<uint32_t>((it - 1) & 1), 100);    // ... and its P V MMAs with V
        mbar_expect_tx(v_full, Cfg::kTileBytes);
#pragma unroll
        for (int part = 0; part < 2; ++part)
          tma_load_2d(&tm_qkv, v_full, smem + Cfg::kOffV + part * Cfg::kBoxRows * 128, 2 * kHidden + h * kHeadDim, b * TV + part * Cfg::kBoxRows);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s1 = umma_idesc_bf16(128, Cfg::kN1), idesc_s2 = umma_idesc_bf16(128, Cfg::kN2);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);
      const uint32_t q_lo = desc_lo_k(sQ), k_lo = desc_lo_k(sK), p_lo = desc_lo_k(sP), v_lo = desc_lo_mn(sV);
      int it = 0;
      long long tile_no = 0;            // running tile counter: parity of the per-tile barriers
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        mbar_wait(qk_full, static_cast<uint32_t>(it & 1));
        for (int t = 0; t < kTiles; ++t, ++tile_no) {
          const uint32_t ph = static_cast<uint32_t>(tile_no & 1);
          if (tile_no > 0) mbar_wait(p_full, ph ^ 1);            // the previous tile's softmax has consumed the score columns
          if constexpr (PT) { if (tile_no > 0) mbar_wait(o_full, ph ^ 1); }   // ... and its P V MMAs the probabilities written over them
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {                // S[:, 0:N1) and S[:, N1:TP) of query rows [128 t, 128 t + 128)
            const uint32_t a = q_lo + t * 128 * 8 + 2 * k;
            if (k == 0) { umma_lohi<false>(tmem_base, a, k_lo, idesc_s1); umma_lohi<false>(tmem_base + Cfg::kN1, a, k_lo + Cfg::kN1 * 8, idesc_s2); }
            else { umma_lohi<true>(tmem_base, a, k_lo + 2 * k, idesc_s1); umma_lohi<true>(tmem_base + Cfg::kN1, a, k_lo + Cfg::kN1 * 8 + 2 * k, idesc_s2); }
          }
          umma_commit(s_full);
          if (t == kTiles - 1) umma_commit(qk_free);
          mbar_wait(p_full, ph);                                   // P of this tile is in shared memory
          if (t == 0) mbar_wait(v_full, static_cast<uint32_t>(it & 1));
          if (tile_no > 0) mbar_wait(o_read, ph ^ 1);            // the previous tile's O has left TMEM
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < TP / 16; ++j) {
            const uint32_t bq = v_lo + j * 128;
            if constexpr (PT) {
              if (j == 0) umma_ts_lohi<false>(tmem_base + TP, tmem_base + 8 * j, bq, idesc_o);
              else umma_ts_lohi<true>(tmem_base + TP, tmem_base + 8 * j, bq, idesc_o);
            } else {
              const uint32_t a = p_lo + (j >> 2) * 1024 + (j & 3) * 2;
              if (j == 0) umma_lohi<false>(tmem_base + TP, a, bq, idesc_o);
              else umma_lohi<true>(tmem_base + TP, a, bq, idesc_o);
            }
          }
          umma_commit(o_full);
          if (t == kTiles - 1) umma_commit(v_free);
        }
      }
    }
    __syncwarp();
  } else {
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int r_tile = warp * 32 + lane;
    long long tile_no = 0;
    // PT: the previous tile's output is drained after this tile's softmax
    [[maybe_unused]] __nv_bfloat16* prev_dst = nullptr;
    [[maybe_unused]] float prev_inv = 0.f;
    [[maybe_unused]] int prev_live = 0;
    [[maybe_unused]] auto drain = [&](uint32_t ph_prev) {
      mbar_wait(o_full, ph_prev);
      tc_fence_after();
      uint32_t oa[32], ob[32];
      load_o_row(t_lane + TP, oa, ob);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_read);
      store_o_rows(oa, ob, prev_inv, sP + static_cast<uint32_t>(warp) * 4096u, prev_dst, prev_live, lane);
    };
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int uu = reverse ? num_units - 1 - unit : unit;
      const int b = uu / kHeads, h = uu - b * kHeads;
      __nv_bfloat16* obase = out + static_cast<long long>(b) * TV * kHidden + h * kHeadDim;
      for (int t = 0; t < kTiles; ++t, ++tile_no) {
        const uint32_t ph = static_cast<uint32_t>(tile_no & 1);
        mbar_wait(s_full, ph);
        if constexpr (PT) {
          tc_fence_after();
          float ms;
          const float sum = softmax_row_to_tmem<TP, TV>(t_lane, ms);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(p_full);
          const int row = t * 128 + r_tile;
          if (lse2 != nullptr && row < TV) lse2[(static_cast<long long>(b) * kHeads + h) * TV + row] = ms + log2f(sum);
          if (tile_no > 0) drain(ph ^ 1);
          const int live = TV - (t * 128 + warp * 32);
          prev_dst = obase + static_cast<long long>(t * 128 + warp * 32) * kHidden;
          prev_inv = 1.0f / sum;
          prev_live = live < 0 ? 0 : (live < 32 ? live : 32);
          continue;
        }
        if (tile_no > 0) mbar_wait(o_full, ph ^ 1);               // the previous tile's P V MMAs have read the P buffer
        tc_fence_after();
        float ms;
        const float sum = softmax_row_to_p<TP, TV>(t_lane, sP + r_tile * 128, 16384, r_tile & 7, true, ms);
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        const int row = t * 128 + r_tile;
        if (lse2 != nullptr && row < TV) lse2[(static_cast<long long>(b) * kHeads + h) * TV + row] = ms + log2f(sum);
        mbar_wait(o_full, ph);
        tc_fence_after();
        uint32_t oa[32], ob[32];
        load_o_row(t_lane + TP, oa, ob);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_read);
        // staging in the P buffer: idle between this tile's O MMAs and the next tile's softmax stores (this warp's own rows)
        const int live = TV - (t * 128 + warp * 32);
        store_o_rows(oa, ob, 1.0f / sum, sP + static_cast<uint32_t>(warp) * 4096u,
                     obase + static_cast<long long>(t * 128 + warp * 32) * kHidden, live < 0 ? 0 : (live < 32 ? live : 32), lane);
      }
    }
    if constexpr (PT) { if (tile_no > 0) drain(static_cast<uint32_t>((tile_no - 1) & 1)); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// T = 324, eight softmax warps (default; JPDVT_ATTN_SEQ_WARPS=4 restores the kernel above): with 336 score columns one SM can
// hold one chain only, and that chain is the softmax of a 336-column row per thread (two passes, ~5,000 cycles per tile).  Here
// two threads share a row: warps 0-3 take keys [0, 176), warps 4-7 keys [176, 336) of the same 128 rows (same TMEM lane
// quadrant: warp index mod 4), exchange the row maximum and the row sum through shared memory (one 64-thread named barrier per
// quadrant each), and write their probabilities to tensor memory - the first half over the score columns it has itself read
// ([0, 88)), the second half into free columns behind O ([400, 480)), because its natural place lies inside the columns the
// first warp may still be reading.  The P V MMAs take k-steps 0..10 from the first region and 11..20 from the second.  The
// output tile is drained by warps 0-3 as before (after the next tile's softmax).
constexpr int kSeq8Threads = 320;                           // 8 softmax warps + TMA warp + MMA warp

// columns [0, NC) of the score row at t_src: maximum over the valid ones
template <int NC, int NV>
__device__ __forceinline__ float row_max_tmem(uint32_t t_src) {
  static_assert(NC % 16 == 0 && NV <= NC, "column count");
  constexpr int kFull = NC / 32, kTail = NC % 32, kChunks = kFull + (kTail ? 1 : 0);
  uint32_t ra[32], rb[32];
  auto load_chunk = [&](uint32_t (&r)[32], int c) {
    if (c < kFull) tmem_ld_32x32(t_src + c * 32, r);
    else tmem_ld_32x16(t_src + c * 32, reinterpret_cast<uint32_t (&)[16]>(r));
  };
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  load_chunk(ra, 0);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    uint32_t (&cur)[32] = (c & 1) ? rb : ra;
    uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
    if (c + 1 < kChunks) load_chunk(nxt, c + 1);
    const int n = (c < kFull) ? 32 : kTail;
#pragma unroll
    for (int j = 0; j < n; j += 8) {
      if (c * 32 + j + 8 <= NV) {
        m0 = fmaxf(m0, fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1])));
        m1 = fmaxf(m1, fmaxf(__uint_as_float(cur[j + 2]), __uint_as_float(cur[j + 3])));
        m2 = fmaxf(m2, fmaxf(__uint_as_float(cur[j + 4]), __uint_as_float(cur[j + 5])));
        m3 = fmaxf(m3, fmaxf(__uint_as_float(cur[j + 6]), __uint_as_float(cur[j + 7])));
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (c * 32 + j + e < NV) m0 = fmaxf(m0, __uint_as_float(cur[j + e]));
      }
    }
    if (c + 1 < kChunks) tmem_ld_wait();
  }
  return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}
// p = 2^(s * log2(e)/8 - ms) of columns [0, NC) at t_src as bf16 pairs into columns [0, NC / 2) at t_dst (t_dst may be t_src: the
// pairs of chunk c land on columns the thread has already read); columns >= NV get probability 0; returns the sum
template <int NC, int NV>
__device__ __forceinline__ float row_exp_to_tmem(uint32_t t_src, uint32_t t_dst, float ms) {
  constexpr float sl2 = 0.125f * 1.4426950408889634f;
  constexpr int kFull = NC / 32, kTail = NC % 32, kChunks = kFull + (kTail ? 1 : 0);
  uint32_t ra[32], rb[32];
  auto load_chunk = [&](uint32_t (&r)[32], int c) {
    if (c < kFull) tmem_ld_32x32(t_src + c * 32, r);
    else tmem_ld_32x16(t_src + c * 32, reinterpret_cast<uint32_t (&)[16]>(r));
  };
  uint64_t sum2 = f2_pack(0.f, 0.f);
  const uint64_t sl2p = f2_pack(sl2, sl2), nmsp = f2_pack(-ms, -ms);
  load_chunk(ra, 0);
  tmem_ld_wait();
#pragma unroll
  for (int c = 0; c < kChunks; ++c) {
    uint32_t (&cur)[32] = (c & 1) ? rb : ra;
    uint32_t (&nxt)[32] = (c & 1) ? ra : rb;
    if (c + 1 < kChunks) load_chunk(nxt, c + 1);
    const int n = (c < kFull) ? 32 : kTail;
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < n / 2; ++j) {
      float a, b;
      f2_unpack(f2_fma(f2_pack(__uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1])), sl2p, nmsp), a, b);
      float pa = ex2f(a), pb = ex2f(b);
      if (c * 32 + 2 * j >= NV) pa = 0.f;
      if (c * 32 + 2 * j + 1 >= NV) pb = 0.f;
      sum2 = f2_add(sum2, f2_pack(pa, pb));
      pk[j] = pack_bf16(pa, pb);
    }
    if (c < kFull) tmem_st_32x16(t_dst + c * 16, pk);
    else tmem_st_32x8(t_dst + c * 16, pk);
    if (c + 1 < kChunks) tmem_ld_wait();
  }
  tmem_st_wait();
  float s_even, s_odd;
  f2_unpack(sum2, s_even, s_odd);
  return s_even + s_odd;
}

template <int TP, int TV>
__global__ void __launch_bounds__(kSeq8Threads, 1)
attention_seq8_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse2,
                      int num_units, int reverse) {
  using Cfg = SeqCfg<TP, TV>;
  constexpr int kTiles = Cfg::kTiles;
  constexpr int kNA = 176, kNB = TP - kNA;                    // key split between the two warps of a row (whole 16-key steps)
  constexpr int kColO = TP, kColPB = TP + 64;                 // O behind the scores, the second half's probabilities behind O
  static_assert(TP == 336 && kNB % 16 == 0 && TV > kNA && kColPB + kNB / 2 <= 512, "laid out for 324 tokens padded to 336");
  constexpr float sl2 = 0.125f * 1.4426950408889634f;
  extern __shared__ uint8_t att_tc_smem[];
  uint8_t* smem = att_tc_smem + ((1024u - (smem_u32(att_tc_smem) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* qk_full = bars + 0;
  uint64_t* v_full = bars + 1;
  uint64_t* s_full = bars + 2;         // MMA: scores of the current tile ready   (once per tile)
  uint64_t* p_full = bars + 3;         // softmax warps: P written, S consumed    (once per tile, 8 arrivals)
  uint64_t* o_full = bars + 4;         // MMA: O of the current tile ready        (once per tile)
  uint64_t* o_read = bars + 5;         // warps 0-3: O read out of TMEM           (once per tile, 4 arrivals)
  uint64_t* qk_free = bars + 6;        // MMA: every score MMA of the unit has read Q, K
  uint64_t* v_free = bars + 7;         // MMA: every P V MMA of the unit has read V
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* xch = reinterpret_cast<float*>(smem + Cfg::kOffP);   // [2][2][128] row maximum / row sum per half (the P tile's slot is free)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 8); mbar_init(o_full, 1); mbar_init(o_read, 4);
    mbar_init(qk_free, 1); mbar_init(v_free, 1);
    fence_mbar_init();
  }
  if (warp == 9) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  if (warp == 8 && lane == 0) tma_prefetch_desc(&tm_qkv);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV),
                 sStage = smem_u32(smem + Cfg::kOffP) + 4096u;      // output staging behind the exchange words

  if (warp == 8) {
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int uu = reverse ? num_units - 1 - unit : unit;
        const int b = uu / kHeads, h = uu - b * kHeads;
        if (it > 0) mbar_wait_backoff(qk_free, static_cast<uint32_t>((it - 1) & 1), 100);
        mbar_expect_tx(qk_full, 2 * Cfg::kTileBytes);
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffQ + part * Cfg::kBoxRows * 128, h * kHeadDim, b * TV + part * Cfg::kBoxRows);
          tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffK + part * Cfg::kBoxRows * 128, kHidden + h * kHeadDim, b * TV + part * Cfg::kBoxRows);
        }
        if (it > 0) mbar_wait_backoff(v_free, static_cast<uint32_t>((it - 1) & 1), 100);
        mbar_expect_tx(v_full, Cfg::kTileBytes);
#pragma unroll
        for (int part = 0; part < 2; ++part)
          tma_load_2d(&tm_qkv, v_full, smem + Cfg::kOffV + part * Cfg::kBoxRows * 128, 2 * kHidden + h * kHeadDim, b * TV + part * Cfg::kBoxRows);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    if (lane == 0) {
      constexpr uint32_t idesc_s1 = umma_idesc_bf16(128, Cfg::kN1), idesc_s2 = umma_idesc_bf16(128, Cfg::kN2);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);
      const uint32_t q_lo = desc_lo_k(sQ), k_lo = desc_lo_k(sK), v_lo = desc_lo_mn(sV);
      int it = 0;
      long long tile_no = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        mbar_wait(qk_full, static_cast<uint32_t>(it & 1));
        for (int t = 0; t < kTiles; ++t, ++tile_no) {
          const uint32_t ph = static_cast<uint32_t>(tile_no & 1);
          if (tile_no > 0) { mbar_wait(p_full, ph ^ 1); mbar_wait(o_full, ph ^ 1); }   // scores consumed, P read by the previous P V
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
            const uint32_t a = q_lo + t * 128 * 8 + 2 * k;
            if (k == 0) { umma_lohi<false>(tmem_base, a, k_lo, idesc_s1); umma_lohi<false>(tmem_base + Cfg::kN1, a, k_lo + Cfg::kN1 * 8, idesc_s2); }
            else { umma_lohi<true>(tmem_base, a, k_lo + 2 * k, idesc_s1); umma_lohi<true>(tmem_base + Cfg::kN1, a, k_lo + Cfg::kN1 * 8 + 2 * k, idesc_s2); }
          }
          umma_commit(s_full);
          if (t == kTiles - 1) umma_commit(qk_free);
          mbar_wait(p_full, ph);
          if (t == 0) mbar_wait(v_full, static_cast<uint32_t>(it & 1));
          if (tile_no > 0) mbar_wait(o_read, ph ^ 1);            // the previous tile's O has left TMEM
          tc_fence_after();
#pragma unroll
          for (int j = 0; j < TP / 16; ++j) {                    // key step j: first half's pairs at [8j, 8j + 8), second half's behind O
            const uint32_t a = (j < kNA / 16) ? tmem_base + 8 * j : tmem_base + kColPB + 8 * (j - kNA / 16);
            if (j == 0) umma_ts_lohi<false>(tmem_base + kColO, a, v_lo + j * 128, idesc_o);
            else umma_ts_lohi<true>(tmem_base + kColO, a, v_lo + j * 128, idesc_o);
          }
          umma_commit(o_full);
          if (t == kTiles - 1) umma_commit(v_free);
        }
      }
    }
    __syncwarp();
  } else {
    const int quad = warp & 3, half = warp >> 2;                // TMEM lane quadrant / which half of the keys
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int r_tile = quad * 32 + lane;
    float* xmax = xch + half * 128 + r_tile;                    // mine; the partner's is xch[(1 - half) * 128 + r_tile]
    float* xsum = xch + 256 + half * 128 + r_tile;
    const float* pmax = xch + (1 - half) * 128 + r_tile;
    const float* psum = xch + 256 + (1 - half) * 128 + r_tile;
    long long tile_no = 0;
    __nv_bfloat16* prev_dst = nullptr;
    float prev_inv = 0.f;
    int prev_live = 0;
    auto drain = [&](uint32_t ph_prev) {                        // warps 0-3 only
      mbar_wait(o_full, ph_prev);
      tc_fence_after();
      uint32_t oa[32], ob[32];
      load_o_row(t_lane + kColO, oa, ob);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_read);
      store_o_rows(oa, ob, prev_inv, sStage + static_cast<uint32_t>(quad) * 4096u, prev_dst, prev_live, lane);
    };
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const int uu = reverse ? num_units - 1 - unit : unit;
      const int b = uu / kHeads, h = uu - b * kHeads;
      __nv_bfloat16* obase = out + static_cast<long long>(b) * TV * kHidden + h * kHeadDim;
      for (int t = 0; t < kTiles; ++t, ++tile_no) {
        const uint32_t ph = static_cast<uint32_t>(tile_no & 1);
        mbar_wait(s_full, ph);
        tc_fence_after();
        const float mx = half == 0 ? row_max_tmem<kNA, kNA>(t_lane) : row_max_tmem<kNB, TV - kNA>(t_lane + kNA);
        *xmax = mx;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        const float ms = fmaxf(mx, *pmax) * sl2;
        const float part = half == 0 ? row_exp_to_tmem<kNA, kNA>(t_lane, t_lane, ms)
                                     : row_exp_to_tmem<kNB, TV - kNA>(t_lane + kNA, t_lane + kColPB, ms);
        *xsum = part;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        if (half == 0) {
          const float sum = part + *psum;
          const int row = t * 128 + r_tile;
          if (lse2 != nullptr && row < TV) lse2[(static_cast<long long>(b) * kHeads + h) * TV + row] = ms + log2f(sum);
          if (tile_no > 0) drain(ph ^ 1);
          const int live = TV - (t * 128 + quad * 32);
          prev_dst = obase + static_cast<long long>(t * 128 + quad * 32) * kHidden;
          prev_inv = 1.0f / sum;
          prev_live = live < 0 ? 0 : (live < 32 ? live : 32);
        }
      }
    }
    if (half == 0 && tile_no > 0) drain(static_cast<uint32_t>((tile_no - 1) & 1));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int TP, int TV>
int launch_seq8(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int batch, cudaStream_t stream) {
  using Cfg = SeqCfg<TP, TV>;
  static bool configured = false;
  auto kern = attention_seq8_kernel<TP, TV>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "attention_seq8: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
                       cudaGetErrorString(cudaGetLastError()));
    configured = true;
  }
  CUtensorMap tm;
  int rc = make_tmap_bf16_kmajor(&tm, qkv, static_cast<long long>(batch) * TV, kQkvCols, kQkvCols, Cfg::kBoxRows);
  if (rc != kOk) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int units = batch * kHeads;
  kern<<<units < sms ? units : sms, kSeq8Threads, Cfg::kSmemBytes, stream>>>(tm, out, lse2, units, sweep_reverse());
  return check_launch("attention_seq8_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// T = 324, key-split form (default; JPDVT_ATTN_SEQ_SPLIT=0 restores the sequential-tile kernel): two chains per SM for the size whose score row (336 columns) does not
// fit a 256-column CTA.  Work item = (unit, 128-query tile); the keys go through in two blocks - [0, 176) and [176, 336) - with
// the online-softmax rule between them: block 0 gives m0, P0 = 2^(S0 - m0), O = P0 V0; block 1 gives m1 = max(m0, rowmax S1),
// P1 = 2^(S1 - m1), and the accumulator is rescaled by a = 2^(m0 - m1) IN tensor memory (tcgen05.ld, scale, tcgen05.st - one
// 64-column round trip per row) before O += P1 V1; the row sum follows l = a l0 + l1.  TMEM per CTA: scores / probabilities of
// the current block in columns [0, 176), O in [176, 240); shared memory: Q tile 16 KB + K 42 KB + V 42 KB (the output staging
// tile re-uses V once the second P V has read it) - two CTAs per SM.  Every per-block barrier completes exactly twice per item,
// so block 0 always waits on parity 0 and block 1 on parity 1.
struct KsCfg {
  static constexpr int TV = 324, TP = 336, kNA = 176, kNB = TP - kNA, kVB = TV - kNA;     // valid keys of block 1
  static constexpr int kQBytes = 128 * 128, kKvBytes = TP * 128, kBoxRows = TP / 2;
  static constexpr int kOffQ = 0, kOffK = kQBytes, kOffV = kQBytes + kKvBytes;
  static constexpr int kBarOff = kOffV + kKvBytes;
  static constexpr int kSmemBytes = kBarOff + 128 + 1024;
  static constexpr int kColO = kNA;
  static_assert(kOffK % 1024 == 0 && kOffV % 1024 == 0 && kNA % 16 == 0 && kNB % 16 == 0 && kColO + kHeadDim <= 256, "layout");
  static_assert(2 * (kSmemBytes + 1024) <= 227 * 1024, "two CTAs per SM");
};

__global__ void __launch_bounds__(kTcThreads, 2)
attention_ks_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    __nv_bfloat16* __restrict__ out, float* __restrict__ lse2, int num_items, int reverse) {
  using Cfg = KsCfg;
  constexpr int TV = Cfg::TV, kQt = 3;
  constexpr float sl2 = 0.125f * 1.4426950408889634f;
  extern __shared__ uint8_t att_tc_smem[];
  uint8_t* smem = att_tc_smem + ((1024u - (smem_u32(att_tc_smem) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* qk_full = bars + 0;        // TMA: Q tile and K landed                              (once per item)
  uint64_t* v_full = bars + 1;         // TMA: V landed                                         (once per item)
  uint64_t* s_full = bars + 2;         // MMA: scores of the current key block are in TMEM      (twice per item)
  uint64_t* p_full = bars + 3;         // softmax warps: P of the block written (block 1: O rescaled too), 4 arrivals (twice per item)
  uint64_t* o_full = bars + 4;         // MMA: the block's P V MMAs are done                    (twice per item)
  uint64_t* o_read = bars + 5;         // softmax warps: O has left TMEM, 4 arrivals            (once per item)
  uint64_t* qk_free = bars + 6;        // MMA: both score blocks have read Q, K                 (once per item)
  uint64_t* v_free = bars + 7;         // MMA: both P V blocks have read V                      (once per item)
  uint64_t* stage_free = bars + 8;     // softmax warps: the staging tile inside V is done, 4 arrivals (once per item)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
    mbar_init(o_read, 4); mbar_init(qk_free, 1); mbar_init(v_free, 1); mbar_init(stage_free, 4);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch_dependents();
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV);
  auto item_bhq = [&](int item, int& b, int& h, int& qt) {
    const int ii = reverse ? num_items - 1 - item : item;
    const int uu = ii / kQt;
    qt = ii - uu * kQt;
    b = uu / kHeads; h = uu - b * kHeads;
  };

  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        int b, h, qt;
        item_bhq(item, b, h, qt);
        const uint32_t prev = static_cast<uint32_t>((it - 1) & 1);
        if (it > 0) mbar_wait(qk_free, prev);
        mbar_expect_tx(qk_full, Cfg::kQBytes + Cfg::kKvBytes);
        tma_load_2d(&tm_q, qk_full, smem + Cfg::kOffQ, h * kHeadDim, b * TV + qt * 128);
#pragma unroll
        for (int part = 0; part < 2; ++part)
          tma_load_2d(&tm_kv, qk_full, smem + Cfg::kOffK + part * Cfg::kBoxRows * 128, kHidden + h * kHeadDim, b * TV + part * Cfg::kBoxRows);
        if (it > 0) { mbar_wait(v_free, prev); mbar_wait(stage_free, prev); }
        mbar_expect_tx(v_full, Cfg::kKvBytes);
#pragma unroll
        for (int part = 0; part < 2; ++part)
          tma_load_2d(&tm_kv, v_full, smem + Cfg::kOffV + part * Cfg::kBoxRows * 128, 2 * kHidden + h * kHeadDim, b * TV + part * Cfg::kBoxRows);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s0 = umma_idesc_bf16(128, Cfg::kNA), idesc_s1 = umma_idesc_bf16(128, Cfg::kNB);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);
      const uint32_t q_lo = desc_lo_k(sQ), k_lo = desc_lo_k(sK), v_lo = desc_lo_mn(sV);
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const uint32_t ph = static_cast<uint32_t>(it & 1);
        mbar_wait(qk_full, ph);
        if (it > 0) mbar_wait(o_full, 1);                       // the previous item's second P V has read its probabilities
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {                 // block 0: keys [0, 176)
          if (k == 0) umma_lohi<false>(tmem_base, q_lo, k_lo, idesc_s0);
          else umma_lohi<true>(tmem_base, q_lo + 2 * k, k_lo + 2 * k, idesc_s0);
        }
        umma_commit(s_full);
        mbar_wait(p_full, 0);
        mbar_wait(v_full, ph);
        if (it > 0) mbar_wait(o_read, static_cast<uint32_t>((it - 1) & 1));   // the previous item's O has left TMEM
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < Cfg::kNA / 16; ++j) {
          if (j == 0) umma_ts_lohi<false>(tmem_base + Cfg::kColO, tmem_base + 8 * j, v_lo + j * 128, idesc_o);
          else umma_ts_lohi<true>(tmem_base + Cfg::kColO, tmem_base + 8 * j, v_lo + j * 128, idesc_o);
        }
        umma_commit(o_full);
        mbar_wait(o_full, 0);                                   // P0 has been read: the second score block may overwrite it
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {                 // block 1: keys [176, 336)
          if (k == 0) umma_lohi<false>(tmem_base, q_lo, k_lo + Cfg::kNA * 8, idesc_s1);
          else umma_lohi<true>(tmem_base, q_lo + 2 * k, k_lo + Cfg::kNA * 8 + 2 * k, idesc_s1);
        }
        umma_commit(s_full);
        umma_commit(qk_free);
        mbar_wait(p_full, 1);                                   // P1 written and O rescaled
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < Cfg::kNB / 16; ++j)
          umma_ts_lohi<true>(tmem_base + Cfg::kColO, tmem_base + 8 * j, v_lo + (Cfg::kNA / 16 + j) * 128, idesc_o);
        umma_commit(o_full);
        umma_commit(v_free);
      }
    }
    __syncwarp();
  } else {
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int r_tile = warp * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t ph = static_cast<uint32_t>(it & 1);
      int b, h, qt;
      item_bhq(item, b, h, qt);
      // ---- block 0
      mbar_wait(s_full, 0);
      tc_fence_after();
      const float ms0 = row_max_tmem<Cfg::kNA, Cfg::kNA>(t_lane) * sl2;
      const float l0 = row_exp_to_tmem<Cfg::kNA, Cfg::kNA>(t_lane, t_lane, ms0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      // ---- block 1
      mbar_wait(s_full, 1);
      tc_fence_after();
      const float ms1 = fmaxf(ms0, row_max_tmem<Cfg::kNB, Cfg::kVB>(t_lane) * sl2);
      const float alpha = ex2f(ms0 - ms1);
      const float l1 = row_exp_to_tmem<Cfg::kNB, Cfg::kVB>(t_lane, t_lane, ms1);
      const float lsum = fmaf(l0, alpha, l1);
      if (!__all_sync(0xffffffffu, alpha == 1.0f)) {            // rescale the accumulator of block 0 (its P V is done: the second
        uint32_t oa[32], ob[32];                                //  score block is only issued behind it)
        load_o_row(t_lane + Cfg::kColO, oa, ob);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          oa[j] = __float_as_uint(__uint_as_float(oa[j]) * alpha);
          ob[j] = __float_as_uint(__uint_as_float(ob[j]) * alpha);
        }
        tmem_st_32x16(t_lane + Cfg::kColO, reinterpret_cast<const uint32_t (&)[16]>(oa[0]));
        tmem_st_32x16(t_lane + Cfg::kColO + 16, reinterpret_cast<const uint32_t (&)[16]>(oa[16]));
        tmem_st_32x16(t_lane + Cfg::kColO + 32, reinterpret_cast<const uint32_t (&)[16]>(ob[0]));
        tmem_st_32x16(t_lane + Cfg::kColO + 48, reinterpret_cast<const uint32_t (&)[16]>(ob[16]));
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      const int row = qt * 128 + r_tile;
      if (lse2 != nullptr && row < TV) lse2[(static_cast<long long>(b) * kHeads + h) * TV + row] = ms1 + log2f(lsum);
      // ---- output
      uint32_t oa[32], ob[32];
      mbar_wait(o_full, 1);
      tc_fence_after();
      load_o_row(t_lane + Cfg::kColO, oa, ob);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_read);
      const int live = TV - (qt * 128 + warp * 32);
      store_o_rows(oa, ob, 1.0f / lsum, sV + static_cast<uint32_t>(warp) * 4096u,
                   out + (static_cast<long long>(b) * TV + qt * 128 + warp * 32) * kHidden + h * kHeadDim,
                   live < 0 ? 0 : (live < 32 ? live : 32), lane);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(stage_free);
      (void)ph;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

int launch_ks(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int batch, cudaStream_t stream) {
  using Cfg = KsCfg;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(attention_ks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "attention_ks: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
                       cudaGetErrorString(cudaGetLastError()));
    cudaFuncSetAttribute(attention_ks_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  CUtensorMap tq, tkv;
  const long long rows = static_cast<long long>(batch) * Cfg::TV;
  int rc = make_tmap_bf16_kmajor(&tq, qkv, rows, kQkvCols, kQkvCols, 128);
  if (rc != kOk) return rc;
  rc = make_tmap_bf16_kmajor(&tkv, qkv, rows, kQkvCols, kQkvCols, Cfg::kBoxRows);
  if (rc != kOk) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = batch * kHeads * 3;
  const int slots = sms * 2;
  const int grid = items < slots ? items : slots;
  if (launch_pdl(attention_ks_kernel, dim3(grid), dim3(kTcThreads), Cfg::kSmemBytes, stream, tq, tkv, out, lse2, items, sweep_reverse()) != cudaSuccess)
    return set_error(kErrCuda, "attention_ks_kernel: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return check_launch("attention_ks_kernel");
}

template <int TP, int TV>
int launch_tc_seq(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int batch, cudaStream_t stream) {
  using Cfg = SeqCfg<TP, TV>;
  static bool configured = false;
  static int pt = -1;               // JPDVT_ATTN_SEQ_TMEM=0: probabilities through the shared-memory P tile (the first form) instead of TMEM
  if (pt < 0) { const char* e = getenv("JPDVT_ATTN_SEQ_TMEM"); pt = (e != nullptr && e[0] == '0') ? 0 : 1; }
  auto kern = pt ? attention_tc_seq_kernel<TP, TV, true> : attention_tc_seq_kernel<TP, TV, false>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "attention_tc_seq: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
                       cudaGetErrorString(cudaGetLastError()));
    configured = true;
  }
  CUtensorMap tm;
  int rc = make_tmap_bf16_kmajor(&tm, qkv, static_cast<long long>(batch) * TV, kQkvCols, kQkvCols, Cfg::kBoxRows);
  if (rc != kOk) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int units = batch * kHeads;
  kern<<<units < sms ? units : sms, kTcThreads, Cfg::kSmemBytes, stream>>>(tm, out, lse2, units, sweep_reverse());
  return check_launch("attention_tc_seq_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// T = 144, hybrid form (JPDVT_ATTN_REM=hybrid): the 128-row main tile on tcgen05 as above, the 16-row remainder on the softmax
// warps themselves with mma.sync (registers), so that NO tcgen05 instruction is spent on it.
//
// Why: a tcgen05.mma costs the same with 16 live rows as with 128, so the remainder doubled the tensor-pipe time and the
// operand traffic of a unit (16 score MMAs + 9 P V MMAs at M = 128 for 16 rows), its scores aliased the main tile's TMEM
// columns (issued only after the main softmax), its output aliased them again (so the next unit's scores had to wait for this
// unit's epilogue), and Q / K stayed busy until those late MMAs had read them (the next unit's load latency sat on the chain).
// Here a softmax warp computes S_rem[16, its keys] = Q_rem K^T with mma.sync.m16n8k16 straight from the TMA-swizzled Q / K tiles
// (ldmatrix) while the main score MMAs run, the four warps exchange row maximum / row sum through shared memory and leave the
// bf16 probabilities in a 4.8 KB row-major tile; after the main softmax each warp multiplies that tile with ITS 16 output
// columns of V (ldmatrix.trans) while the tensor core runs the main P V, and writes its 16 x 16 outputs directly.  Consequences
// for the pipeline: Q / K are released ~400 cycles into a unit (the next unit's loads hide behind the main softmax), the
// TMEM score columns are free as soon as the main softmax has read them, so the MMA warp issues the NEXT unit's scores right
// behind this unit's P V - they are done before the softmax warps come back from the epilogue.
constexpr int kHmPRowBytes = 304;                           // 144 keys x 2 B + 16 B: conflict-free ldmatrix rows / 4-byte stores
static_assert(16 * kHmPRowBytes <= TcCfg<144>::kP1Bytes, "the remainder's probability tile reuses the compact P1 slot");

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk `ch` of row `row` in a TMA SWIZZLE_128B tile of 128-byte rows (1024-byte aligned base)
__device__ __forceinline__ uint32_t swz128(int row, int ch) { return static_cast<uint32_t>(row * 128 + ((ch ^ (row & 7)) << 4)); }

// PT: the main tile's probabilities live in tensor memory (softmax_row_to_tmem, A-from-TMEM MMAs) instead of a shared-memory tile
template <bool PT, bool TRACE = false>
__global__ void __launch_bounds__(kTcThreads, 2)
attention_hm_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse2,
                    int num_units, int reverse, long long* __restrict__ trace = nullptr) {
  constexpr int T = 144;
  using Cfg = TcCfg<T>;
  auto mark = [&](int role, int it, int ev) {                 // developer path (JPDVT_ATTN_TRACE=1): clocks of CTA 0
    if constexpr (TRACE) {
      if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && it < kTraceUnits && role >= 0)
        trace[(role * kTraceUnits + it) * kTraceEvents + ev] = clock64();
    }
  };
  extern __shared__ uint8_t att_tc_smem[];
  uint8_t* smem = att_tc_smem + ((1024u - (smem_u32(att_tc_smem) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* qk_full = bars + 0;        // TMA: Q and K landed
  uint64_t* v_full = bars + 1;         // TMA: V landed
  uint64_t* s_full = bars + 2;         // MMA: main scores are in TMEM (and the score MMAs have read Q, K)
  uint64_t* p_full = bars + 3;         // softmax warps: P0 is in shared memory, the score columns are free (4 arrivals)
  uint64_t* o_full = bars + 4;         // MMA: O0 is in TMEM (and the P V MMAs have read V, P0)
  uint64_t* epi_done = bars + 5;       // softmax warps: O0 has left TMEM (4 arrivals)
  uint64_t* qk_read = bars + 6;        // softmax warps: the remainder's ldmatrix reads of Q, K are done (4 arrivals)
  uint64_t* v_read = bars + 7;         // softmax warps: the remainder's ldmatrix reads of V are done (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  float* xch = reinterpret_cast<float*>(smem + Cfg::kXchOff);   // [2][4][16]: row maximum / row sum of the remainder per warp

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
    mbar_init(epi_done, 4); mbar_init(qk_read, 4); mbar_init(v_read, 4);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  if (warp == 4 && lane == 0) tma_prefetch_desc(&tm_qkv);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch_dependents();
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV),
                 sP = smem_u32(smem + Cfg::kOffP), sPr = smem_u32(smem + Cfg::kOffP1);

  if (warp == 4) {
    // ---------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const int uu = reverse ? num_units - 1 - unit : unit;
        const int b = uu / kHeads, h = uu - b * kHeads;
        const uint32_t prev = static_cast<uint32_t>((it - 1) & 1);
        mark(3, it, 0);
        if (it > 0) { mbar_wait(s_full, prev); mbar_wait(qk_read, prev); }     // tensor core and ldmatrix are done with Q, K
        mark(3, it, 1);
        mbar_expect_tx(qk_full, 2 * Cfg::kTileBytes);
        tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffQ, h * kHeadDim, b * T);
        tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffK, kHidden + h * kHeadDim, b * T);
        mark(3, it, 2);
        if (it > 0) { mbar_wait(o_full, prev); mbar_wait(v_read, prev); }      // ... and with V
        mark(3, it, 3);
        mbar_expect_tx(v_full, Cfg::kTileBytes);
        tma_load_2d(&tm_qkv, v_full, smem + Cfg::kOffV, 2 * kHidden + h * kHeadDim, b * T);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, T);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);   // B = V, MN-major
      const uint32_t q_lo = desc_lo_k(sQ), k_lo = desc_lo_k(sK), p_lo = desc_lo_k(sP), v_lo = desc_lo_mn(sV);
      auto issue_s = [&]() {
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          if (k == 0) umma_lohi<false>(tmem_base, q_lo, k_lo, idesc_s);
          else umma_lohi<true>(tmem_base, q_lo + 2 * k, k_lo + 2 * k, idesc_s);
        }
        umma_commit(s_full);
      };
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const uint32_t ph = static_cast<uint32_t>(it & 1);
        if (it == 0) {
          mbar_wait(qk_full, 0);
          tc_fence_after();
          issue_s();
        }
        mark(0, it, 0);
        mbar_wait(p_full, ph);                                  // P0 written, score columns read
        mark(0, it, 1);
        mbar_wait(v_full, ph);
        if (it > 0) mbar_wait(epi_done, static_cast<uint32_t>((it - 1) & 1));   // O0 of the previous unit has left TMEM
        mark(0, it, 2);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < T / 16; ++j) {
          const uint32_t bq = v_lo + j * 128;
          if constexpr (PT) {                                  // A = P out of tensor memory: key step j = columns [8j, 8j + 8)
            if (j == 0) umma_ts_lohi<false>(tmem_base + Cfg::kColO0, tmem_base + 8 * j, bq, idesc_o);
            else umma_ts_lohi<true>(tmem_base + Cfg::kColO0, tmem_base + 8 * j, bq, idesc_o);
          } else {
            const uint32_t a = p_lo + (j >> 2) * 1024 + (j & 3) * 2;
            if (j == 0) umma_lohi<false>(tmem_base + Cfg::kColO0, a, bq, idesc_o);
            else umma_lohi<true>(tmem_base + Cfg::kColO0, a, bq, idesc_o);
          }
        }
        umma_commit(o_full);
        mark(0, it, 3);
        if (unit + static_cast<int>(gridDim.x) < num_units) {  // the next unit's scores, behind this unit's P V on the tensor pipe
          mbar_wait(qk_full, ph ^ 1u);
          mark(0, it, 4);
          if constexpr (PT) mbar_wait(o_full, ph);             // the scores overwrite the columns the P V MMAs read P from
          mark(0, it, 5);
          tc_fence_after();
          issue_s();
          mark(0, it, 6);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- softmax + remainder + epilogue
    constexpr float sl2 = 0.125f * 1.4426950408889634f;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int r_tile = warp * 32 + lane;
    const int g = lane >> 2, tq = lane & 3, li = lane >> 3, lr = lane & 7;
    // remainder scores: warp 0 takes keys [0, 48), warps 1..3 32 keys each - whole 16-key steps for the P V contraction
    const int key0 = (warp == 0) ? 0 : 16 + 32 * warp;
    const int npairs = (warp == 0) ? 3 : 2;                    // pairs of 8-key tiles
    int it = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const uint32_t ph = static_cast<uint32_t>(it & 1);
      const int uu = reverse ? num_units - 1 - unit : unit;
      const int b = uu / kHeads, h = uu - b * kHeads;
      __nv_bfloat16* obase = out + static_cast<long long>(b) * T * kHidden + h * kHeadDim;

      // ---- remainder scores S_rem[16, my keys] = Q[128:144] K[my keys]^T, fp32 in registers
      const int role = (warp == 0) ? 1 : (warp == 3) ? 2 : -1;
      mark(role, it, 0);
      mbar_wait(qk_full, ph);
      mark(role, it, 1);
      float sr[3][2][4];
#pragma unroll
      for (int pp = 0; pp < 3; ++pp)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) { sr[pp][nt][0] = sr[pp][nt][1] = sr[pp][nt][2] = sr[pp][nt][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t qa[4];
        ldsm_x4(qa, sQ + swz128(128 + (li & 1) * 8 + lr, ks * 2 + (li >> 1)));
#pragma unroll
        for (int pp = 0; pp < 3; ++pp) {
          if (pp < npairs) {
            uint32_t kf[4];
            ldsm_x4(kf, sK + swz128(key0 + 16 * pp + (li >> 1) * 8 + lr, ks * 2 + (li & 1)));
            mma_bf16(sr[pp][0], qa, kf[0], kf[1]);
            mma_bf16(sr[pp][1], qa, kf[2], kf[3]);
          }
        }
      }
      float mlo = -INFINITY, mhi = -INFINITY;                   // rows 128 + g and 128 + g + 8
#pragma unroll
      for (int pp = 0; pp < 3; ++pp) {
        if (pp < npairs) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            mlo = fmaxf(mlo, fmaxf(sr[pp][nt][0], sr[pp][nt][1]));
            mhi = fmaxf(mhi, fmaxf(sr[pp][nt][2], sr[pp][nt][3]));
          }
        }
      }
      mlo = fmaxf(mlo, __shfl_xor_sync(0xffffffffu, mlo, 1)); mlo = fmaxf(mlo, __shfl_xor_sync(0xffffffffu, mlo, 2));
      mhi = fmaxf(mhi, __shfl_xor_sync(0xffffffffu, mhi, 1)); mhi = fmaxf(mhi, __shfl_xor_sync(0xffffffffu, mhi, 2));
      if (tq == 0) { xch[warp * 16 + g] = mlo; xch[warp * 16 + g + 8] = mhi; }
      __syncwarp();
      if (lane == 0) mbar_arrive(qk_read);                      // every ldmatrix of Q / K has delivered (the maxima depend on them)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float ms_lo = fmaxf(fmaxf(xch[g], xch[16 + g]), fmaxf(xch[32 + g], xch[48 + g])) * sl2;
      const float ms_hi = fmaxf(fmaxf(xch[g + 8], xch[24 + g]), fmaxf(xch[40 + g], xch[56 + g])) * sl2;
      float slo = 0.f, shi = 0.f;
#pragma unroll
      for (int pp = 0; pp < 3; ++pp) {
        if (pp < npairs) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const float p0 = ex2f(fmaf(sr[pp][nt][0], sl2, -ms_lo)), p1 = ex2f(fmaf(sr[pp][nt][1], sl2, -ms_lo));
            const float p2 = ex2f(fmaf(sr[pp][nt][2], sl2, -ms_hi)), p3 = ex2f(fmaf(sr[pp][nt][3], sl2, -ms_hi));
            slo += p0 + p1; shi += p2 + p3;
            const uint32_t col = static_cast<uint32_t>(key0 + 16 * pp + 8 * nt + 2 * tq) * 2u;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(sPr + g * kHmPRowBytes + col), "r"(pack_bf16(p0, p1)) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(sPr + (g + 8) * kHmPRowBytes + col), "r"(pack_bf16(p2, p3)) : "memory");
          }
        }
      }
      slo += __shfl_xor_sync(0xffffffffu, slo, 1); slo += __shfl_xor_sync(0xffffffffu, slo, 2);
      shi += __shfl_xor_sync(0xffffffffu, shi, 1); shi += __shfl_xor_sync(0xffffffffu, shi, 2);
      if (tq == 0) { xch[64 + warp * 16 + g] = slo; xch[64 + warp * 16 + g + 8] = shi; }
      asm volatile("bar.sync 1, 128;" ::: "memory");           // probabilities and row sums of all four warps are in place
      mark(role, it, 2);

      // ---- main tile: S out of TMEM, P0 into shared memory
      mbar_wait(s_full, ph);
      mark(role, it, 3);
      tc_fence_after();
      float ms0, sum0;
      if constexpr (PT) {
        sum0 = softmax_row_to_tmem<T>(t_lane, ms0);
      } else {
        sum0 = softmax_row_to_p<T, T, false>(t_lane, sP + r_tile * 128, 16384, r_tile & 7, true, ms0);
        fence_proxy_async_smem();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      mark(role, it, 4);

      // ---- remainder outputs: O_rem[16, my 16 columns] = P_rem[16, 144] V[144, my columns] while the tensor core runs the main P V
      mbar_wait(v_full, ph);
      mark(role, it, 5);
      float orr[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { orr[nt][0] = orr[nt][1] = orr[nt][2] = orr[nt][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < T / 16; ++ks) {
        uint32_t pa[4], vf[4];
        ldsm_x4(pa, sPr + static_cast<uint32_t>(((li & 1) * 8 + lr) * kHmPRowBytes + (ks * 2 + (li >> 1)) * 16));
        ldsm_x4_trans(vf, sV + swz128(16 * ks + (li & 1) * 8 + lr, 2 * warp + (li >> 1)));
        mma_bf16(orr[0], pa, vf[0], vf[1]);
        mma_bf16(orr[1], pa, vf[2], vf[3]);
      }
      const float sum_lo = (xch[64 + g] + xch[80 + g]) + (xch[96 + g] + xch[112 + g]);
      const float sum_hi = (xch[72 + g] + xch[88 + g]) + (xch[104 + g] + xch[120 + g]);
      const float inv_lo = 1.0f / sum_lo, inv_hi = 1.0f / sum_hi;
      __syncwarp();
      if (lane == 0) mbar_arrive(v_read);                       // (the accumulators feed the stores below: the ldmatrix reads are done)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = 16 * warp + 8 * nt + 2 * tq;
        *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(128 + g) * kHidden + col) = pack_bf16(orr[nt][0] * inv_lo, orr[nt][1] * inv_lo);
        *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(136 + g) * kHidden + col) = pack_bf16(orr[nt][2] * inv_hi, orr[nt][3] * inv_hi);
      }
      if (lse2 != nullptr) {   // training: log2-domain log-sum-exp of the scaled scores, [B, 12, T]
        float* lrow = lse2 + (static_cast<long long>(b) * kHeads + h) * T;
        lrow[r_tile] = ms0 + log2f(sum0);
        if (warp == 0 && tq == 0) { lrow[128 + g] = ms_lo + log2f(sum_lo); lrow[136 + g] = ms_hi + log2f(sum_hi); }
      }

      // ---- main output
      uint32_t oa[32], ob[32];
      mark(role, it, 6);
      mbar_wait(o_full, ph);
      mark(role, it, 7);
      tc_fence_after();
      load_o_row(t_lane + Cfg::kColO0, oa, ob);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_done);
      mark(role, it, 8);
      store_o_rows(oa, ob, 1.0f / sum0, sP + static_cast<uint32_t>(warp) * 4096u, obase + static_cast<long long>(warp * 32) * kHidden, 32, lane);
      mark(role, it, 9);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// T = 144, remainder-warp form (JPDVT_ATTN_REM=warp): as the hybrid kernel with the probabilities in tensor memory, but the
// 16-row remainder is the job of ONE warp of its own - the former TMA warp, idle but for three bulk loads per unit - which
// runs it flash-attention-2 style entirely in registers (scores of all 144 keys as mma.sync accumulators, exact softmax with
// quad shuffles, the probabilities re-used as A fragments, 16 x 64 outputs written from the accumulators): no shared-memory
// tile, no barrier with the softmax warps.  The trace of the hybrid kernel showed why: the four softmax warps spent 1,500 +
// 1,100 of a unit's 6,400 cycles on the remainder (two named barriers, dependent mma.sync chains), i.e. the remainder sat on
// the unit chain although nothing of the main tile depends on it.  Roles:
//   warps 0-3 : main tile only - S out of TMEM, P back into TMEM, O0 out of TMEM -> global
//   warp 4    : remainder rows 128..143 (all 32 lanes) + the Q / K loads of the next unit (lane 0)
//   warp 5    : MMA issuer + the V load of the next unit (lane 0)
template <bool TRACE>
__global__ void __launch_bounds__(kTcThreads, 2)
attention_rw_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse2,
                    int num_units, int reverse, long long* __restrict__ trace, int sms) {
  constexpr int T = 144;
  using Cfg = TcCfg<T>;
  auto mark = [&](int role, int it, int ev) {                 // developer path (JPDVT_ATTN_TRACE=1): clocks of CTA 0
    if constexpr (TRACE) {
      if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && it < kTraceUnits && role >= 0)
        trace[(role * kTraceUnits + it) * kTraceEvents + ev] = clock64();
    }
  };
  extern __shared__ uint8_t att_tc_smem[];
  uint8_t* smem = att_tc_smem + ((1024u - (smem_u32(att_tc_smem) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* qk_full = bars + 0;        // TMA: Q and K landed
  uint64_t* v_full = bars + 1;         // TMA: V landed
  uint64_t* s_full = bars + 2;         // MMA: main scores are in TMEM (and the score MMAs have read Q, K)
  uint64_t* p_full = bars + 3;         // softmax warps: P is in TMEM, the scores are consumed (4 arrivals)
  uint64_t* o_full = bars + 4;         // MMA: O0 is in TMEM (and the P V MMAs have read V and P)
  uint64_t* epi_done = bars + 5;       // softmax warps: O0 has left TMEM (4 arrivals)
  uint64_t* v_read = bars + 6;         // remainder warp: its ldmatrix reads of V are done (1 arrival)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // The two CTAs of an SM would both put their remainder warp on the same SM sub-partition (warp index mod 4), where their
  // mma.sync streams share one legacy tensor pipe (30 cycles per HMMA in the trace); the second CTA of an SM (blocks are dealt
  // breadth first) therefore swaps the two single-warp roles, so that the remainder warps sit on sub-partitions 0 and 1.
  const bool swap_roles = ((static_cast<int>(blockIdx.x) / sms) & 1) != 0;
  const int rem_warp = swap_roles ? 5 : 4, mma_warp = swap_roles ? 4 : 5;
  if (threadIdx.x == 0) {
    mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
    mbar_init(epi_done, 4); mbar_init(v_read, 1);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  if (warp == 4 && lane == 0) tma_prefetch_desc(&tm_qkv);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch_dependents();
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV),
                 sP = smem_u32(smem + Cfg::kOffP);
  auto unit_bh = [&](int unit, int& b, int& h) {
    const int uu = reverse ? num_units - 1 - unit : unit;
    b = uu / kHeads; h = uu - b * kHeads;
  };

  if (warp == rem_warp) {
    // ---------------------------------------------------------------------------------------------- remainder rows + Q / K loads
    constexpr float sl2 = 0.125f * 1.4426950408889634f;
    const int g = lane >> 2, tq = lane & 3, li = lane >> 3, lr = lane & 7;
    auto load_qk = [&](int unit) {
      int b, h;
      unit_bh(unit, b, h);
      mbar_expect_tx(qk_full, 2 * Cfg::kTileBytes);
      tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffQ, h * kHeadDim, b * T);
      tma_load_2d(&tm_qkv, qk_full, smem + Cfg::kOffK, kHidden + h * kHeadDim, b * T);
    };
    if (lane == 0 && static_cast<int>(blockIdx.x) < num_units) load_qk(blockIdx.x);
    __syncwarp();
    int it = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const uint32_t ph = static_cast<uint32_t>(it & 1);
      int b, h;
      unit_bh(unit, b, h);
      __nv_bfloat16* obase = out + static_cast<long long>(b) * T * kHidden + h * kHeadDim;
      mark(3, it, 0);
      mbar_wait(qk_full, ph);
      mark(3, it, 1);
      // ---- S_rem[16, 144] = Q[128:144] K^T: 18 accumulator tiles of 8 keys
      float sr[18][4];
#pragma unroll
      for (int j = 0; j < 18; ++j) { sr[j][0] = sr[j][1] = sr[j][2] = sr[j][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t qa[4];
        ldsm_x4(qa, sQ + swz128(128 + (li & 1) * 8 + lr, ks * 2 + (li >> 1)));
#pragma unroll
        for (int pp = 0; pp < 9; ++pp) {
          uint32_t kf[4];
          ldsm_x4(kf, sK + swz128(16 * pp + (li >> 1) * 8 + lr, ks * 2 + (li & 1)));
          mma_bf16(sr[2 * pp], qa, kf[0], kf[1]);
          mma_bf16(sr[2 * pp + 1], qa, kf[2], kf[3]);
        }
      }
      float mlo = -INFINITY, mhi = -INFINITY;                   // rows 128 + g and 136 + g
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        mlo = fmaxf(mlo, fmaxf(sr[j][0], sr[j][1]));
        mhi = fmaxf(mhi, fmaxf(sr[j][2], sr[j][3]));
      }
      mlo = fmaxf(mlo, __shfl_xor_sync(0xffffffffu, mlo, 1)); mlo = fmaxf(mlo, __shfl_xor_sync(0xffffffffu, mlo, 2));
      mhi = fmaxf(mhi, __shfl_xor_sync(0xffffffffu, mhi, 1)); mhi = fmaxf(mhi, __shfl_xor_sync(0xffffffffu, mhi, 2));
      // every ldmatrix of Q / K has delivered (the maxima depend on all of them): once the main score MMAs have read the tiles
      // too, the next unit's Q / K may land - a whole main softmax ahead of their use
      if (lane == 0 && unit + static_cast<int>(gridDim.x) < num_units) {
        mbar_wait(s_full, ph);
        load_qk(unit + gridDim.x);
      }
      __syncwarp();
      mark(3, it, 2);
      const float ms_lo = mlo * sl2, ms_hi = mhi * sl2;
      float slo = 0.f, shi = 0.f;
      uint32_t pf[9][4];                                        // the probabilities as A fragments of the P V contraction
#pragma unroll
      for (int j = 0; j < 18; ++j) {
        const float p0 = ex2f(fmaf(sr[j][0], sl2, -ms_lo)), p1 = ex2f(fmaf(sr[j][1], sl2, -ms_lo));
        const float p2 = ex2f(fmaf(sr[j][2], sl2, -ms_hi)), p3 = ex2f(fmaf(sr[j][3], sl2, -ms_hi));
        slo += p0 + p1; shi += p2 + p3;
        pf[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p0, p1);
        pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p2, p3);
      }
      slo += __shfl_xor_sync(0xffffffffu, slo, 1); slo += __shfl_xor_sync(0xffffffffu, slo, 2);
      shi += __shfl_xor_sync(0xffffffffu, shi, 1); shi += __shfl_xor_sync(0xffffffffu, shi, 2);
      // ---- O_rem[16, 64] = P_rem V
      mbar_wait(v_full, ph);
      mark(3, it, 3);
      float orr[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) { orr[j][0] = orr[j][1] = orr[j][2] = orr[j][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 9; ++ks) {
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t vf[4];
          ldsm_x4_trans(vf, sV + swz128(16 * ks + (li & 1) * 8 + lr, dp * 2 + (li >> 1)));
          mma_bf16(orr[2 * dp], pf[ks], vf[0], vf[1]);
          mma_bf16(orr[2 * dp + 1], pf[ks], vf[2], vf[3]);
        }
      }
      const float inv_lo = 1.0f / slo, inv_hi = 1.0f / shi;
#pragma unroll
      for (int j = 0; j < 8; ++j) {                            // the accumulators feed the stores: every ldmatrix of V has delivered
        const int col = 8 * j + 2 * tq;
        *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(128 + g) * kHidden + col) = pack_bf16(orr[j][0] * inv_lo, orr[j][1] * inv_lo);
        *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(136 + g) * kHidden + col) = pack_bf16(orr[j][2] * inv_hi, orr[j][3] * inv_hi);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(v_read);
      if (lse2 != nullptr && tq == 0) {
        float* lrow = lse2 + (static_cast<long long>(b) * kHeads + h) * T;
        lrow[128 + g] = ms_lo + log2f(slo);
        lrow[136 + g] = ms_hi + log2f(shi);
      }
      mark(3, it, 4);
    }
  } else if (warp == mma_warp) {
    // ---------------------------------------------------------------------------------------------- MMA issuer + V loads
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, T);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);   // B = V, MN-major
      const uint32_t q_lo = desc_lo_k(sQ), k_lo = desc_lo_k(sK), v_lo = desc_lo_mn(sV);
      auto issue_s = [&]() {
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          if (k == 0) umma_lohi<false>(tmem_base, q_lo, k_lo, idesc_s);
          else umma_lohi<true>(tmem_base, q_lo + 2 * k, k_lo + 2 * k, idesc_s);
        }
        umma_commit(s_full);
      };
      auto load_v = [&](int unit) {
        int b, h;
        unit_bh(unit, b, h);
        mbar_expect_tx(v_full, Cfg::kTileBytes);
        tma_load_2d(&tm_qkv, v_full, smem + Cfg::kOffV, 2 * kHidden + h * kHeadDim, b * T);
      };
      int it = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        const uint32_t ph = static_cast<uint32_t>(it & 1);
        if (it == 0) {
          load_v(unit);
          mbar_wait(qk_full, 0);
          tc_fence_after();
          issue_s();
        }
        mark(0, it, 0);
        mbar_wait(p_full, ph);                                  // P written, score columns read
        mark(0, it, 1);
        mbar_wait(v_full, ph);
        if (it > 0) mbar_wait(epi_done, static_cast<uint32_t>((it - 1) & 1));   // O0 of the previous unit has left TMEM
        mark(0, it, 2);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < T / 16; ++j) {                      // A = P out of tensor memory: key step j = columns [8j, 8j + 8)
          if (j == 0) umma_ts_lohi<false>(tmem_base + Cfg::kColO0, tmem_base + 8 * j, v_lo + j * 128, idesc_o);
          else umma_ts_lohi<true>(tmem_base + Cfg::kColO0, tmem_base + 8 * j, v_lo + j * 128, idesc_o);
        }
        umma_commit(o_full);
        mark(0, it, 3);
        if (unit + static_cast<int>(gridDim.x) < num_units) {
          mbar_wait(o_full, ph);                                // P V has read V and P: V may be replaced, the scores may overwrite P
          mbar_wait(v_read, ph);                                // ... and the remainder warp has read V
          mark(0, it, 4);
          load_v(unit + gridDim.x);
          mbar_wait(qk_full, ph ^ 1u);
          mark(0, it, 5);
          tc_fence_after();
          issue_s();                                            // the next unit's scores: done before the softmax warps are back
          mark(0, it, 6);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- main tile: softmax + epilogue
    // Software pipeline: the output of unit n - 1 is drained AFTER the softmax of unit n, so the warps do not sit out the P V
    // latency (p_full -> nine MMAs -> o_full, ~800 cycles in the trace) between a unit's softmax and its own epilogue.  The
    // P V MMAs of unit n wait for that drain (epi_done), the scores of unit n + 1 for the P V MMAs of unit n (o_full).
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int r_tile = warp * 32 + lane;
    const int role = (warp == 0) ? 1 : (warp == 3) ? 2 : -1;
    auto drain = [&](int it_prev, int unit_prev, float inv) {
      int b, h;
      unit_bh(unit_prev, b, h);
      __nv_bfloat16* obase = out + static_cast<long long>(b) * T * kHidden + h * kHeadDim;
      uint32_t oa[32], ob[32];
      mbar_wait(o_full, static_cast<uint32_t>(it_prev & 1));
      mark(role, it_prev, 3);
      tc_fence_after();
      load_o_row(t_lane + Cfg::kColO0, oa, ob);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_done);
      mark(role, it_prev, 4);
      store_o_rows(oa, ob, inv, sP + static_cast<uint32_t>(warp) * 4096u, obase + static_cast<long long>(warp * 32) * kHidden, 32, lane);
      mark(role, it_prev, 5);
    };
    int it = 0, unit_prev = -1;
    float inv_prev = 0.f;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const uint32_t ph = static_cast<uint32_t>(it & 1);
      int b, h;
      unit_bh(unit, b, h);
      mark(role, it, 0);
      mbar_wait(s_full, ph);
      mark(role, it, 1);
      tc_fence_after();
      float ms0;
      const float sum0 = softmax_row_to_tmem<T>(t_lane, ms0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      mark(role, it, 2);
      if (lse2 != nullptr) lse2[(static_cast<long long>(b) * kHeads + h) * T + r_tile] = ms0 + log2f(sum0);
      if (it > 0) drain(it - 1, unit_prev, inv_prev);
      unit_prev = unit;
      inv_prev = 1.0f / sum0;
    }
    if (it > 0) drain(it - 1, unit_prev, inv_prev);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// T = 256 (4x4 @256 px), query-tile form (default; JPDVT_ATTN_QT=0 restores the dual-tile kernel above): what the T = 144
// rewrite taught, applied to the size without a remainder.  The dual-tile kernel holds a whole (sample, head) unit per CTA -
// two 128 x 256 score tiles in 512 TMEM columns, two 64 KB probability tiles in shared memory - so an SM runs ONE chain.
// Here the work item is (unit, 128-query tile): Q tile 16 KB + K 32 KB + V 32 KB, the scores in columns [0, 256), the bf16
// probabilities written back over columns [0, 128) (tcgen05.st; A-from-TMEM MMAs), the output accumulator in the score tile's
// dead columns [128, 192) - 256 TMEM columns and < 100 KB of shared memory, i.e. two independent chains per SM, and no
// shared-memory traffic for P at all.  K / V are fetched once per item (twice per unit); the two items of a unit run on
// neighbouring CTAs at the same time, so the second fetch is an L2 hit.
template <int T>
struct QtCfg {
  static_assert(T == 256, "instantiated for 256 tokens");
  static constexpr int kQBytes = 128 * 128, kKvBytes = T * 128;
  static constexpr int kOffQ = 0, kOffK = kQBytes, kOffV = kQBytes + kKvBytes, kOffStage = kQBytes + 2 * kKvBytes;
  static constexpr int kBarOff = kOffStage + 4 * 4096;
  static constexpr int kSmemBytes = kBarOff + 128 + 1024;
  static constexpr int kColP = 0, kColO = T / 2;             // P: T / 2 columns of bf16 pairs; O behind them, inside the score tile
  static_assert(kColO + kHeadDim <= T && T <= 256, "TMEM columns");
  static_assert(2 * (kSmemBytes + 1024) <= 227 * 1024, "two CTAs per SM");
};

template <int T>
__global__ void __launch_bounds__(kTcThreads, 2)
attention_qt_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                    __nv_bfloat16* __restrict__ out, float* __restrict__ lse2, int num_items, int reverse) {
  using Cfg = QtCfg<T>;
  constexpr int kQt = T / 128;                                // query tiles per unit
  extern __shared__ uint8_t att_tc_smem[];
  uint8_t* smem = att_tc_smem + ((1024u - (smem_u32(att_tc_smem) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* qk_full = bars + 0;        // TMA: Q tile and K landed
  uint64_t* v_full = bars + 1;         // TMA: V landed
  uint64_t* s_full = bars + 2;         // MMA: scores are in TMEM (and the score MMAs have read Q, K)
  uint64_t* p_full = bars + 3;         // softmax warps: P is in TMEM, the scores are consumed (4 arrivals)
  uint64_t* o_full = bars + 4;         // MMA: O is in TMEM (and the P V MMAs have read V and P)
  uint64_t* epi_done = bars + 5;       // softmax warps: O has left TMEM (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
    mbar_init(epi_done, 4);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  if (warp == 4 && lane == 0) { tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_kv); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_wait();
  griddep_launch_dependents();
  const uint32_t sQ = smem_u32(smem + Cfg::kOffQ), sK = smem_u32(smem + Cfg::kOffK), sV = smem_u32(smem + Cfg::kOffV),
                 sStage = smem_u32(smem + Cfg::kOffStage);
  auto item_bhq = [&](int item, int& b, int& h, int& qt) {
    const int ii = reverse ? num_items - 1 - item : item;
    const int uu = ii / kQt;
    qt = ii - uu * kQt;
    b = uu / kHeads; h = uu - b * kHeads;
  };

  if (warp == 4) {
    // ---------------------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        int b, h, qt;
        item_bhq(item, b, h, qt);
        const uint32_t prev = static_cast<uint32_t>((it - 1) & 1);
        if (it > 0) mbar_wait(s_full, prev);                  // the score MMAs of the previous item have read Q, K
        mbar_expect_tx(qk_full, Cfg::kQBytes + Cfg::kKvBytes);
        tma_load_2d(&tm_q, qk_full, smem + Cfg::kOffQ, h * kHeadDim, b * T + qt * 128);
        tma_load_2d(&tm_kv, qk_full, smem + Cfg::kOffK, kHidden + h * kHeadDim, b * T);
        if (it > 0) mbar_wait(o_full, prev);                  // ... and its P V MMAs have read V
        mbar_expect_tx(v_full, Cfg::kKvBytes);
        tma_load_2d(&tm_kv, v_full, smem + Cfg::kOffV, 2 * kHidden + h * kHeadDim, b * T);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ---------------------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, T);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kHeadDim, 0, 1);   // B = V, MN-major
      const uint32_t q_lo = desc_lo_k(sQ), k_lo = desc_lo_k(sK), v_lo = desc_lo_mn(sV);
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const uint32_t ph = static_cast<uint32_t>(it & 1);
        mbar_wait(qk_full, ph);
        if (it > 0) mbar_wait(epi_done, static_cast<uint32_t>((it - 1) & 1));   // O of the previous item (inside the score tile) has left TMEM
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          if (k == 0) umma_lohi<false>(tmem_base, q_lo, k_lo, idesc_s);
          else umma_lohi<true>(tmem_base, q_lo + 2 * k, k_lo + 2 * k, idesc_s);
        }
        umma_commit(s_full);
        mbar_wait(p_full, ph);                                  // P written, score columns read
        mbar_wait(v_full, ph);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < T / 16; ++j) {                      // A = P out of tensor memory: key step j = columns [8j, 8j + 8)
          if (j == 0) umma_ts_lohi<false>(tmem_base + Cfg::kColO, tmem_base + Cfg::kColP + 8 * j, v_lo + j * 128, idesc_o);
          else umma_ts_lohi<true>(tmem_base + Cfg::kColO, tmem_base + Cfg::kColP + 8 * j, v_lo + j * 128, idesc_o);
        }
        umma_commit(o_full);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- softmax + epilogue
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int r_tile = warp * 32 + lane;
    int it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const uint32_t ph = static_cast<uint32_t>(it & 1);
      int b, h, qt;
      item_bhq(item, b, h, qt);
      __nv_bfloat16* obase = out + (static_cast<long long>(b) * T + qt * 128) * kHidden + h * kHeadDim;
      mbar_wait(s_full, ph);
      tc_fence_after();
      float ms0;
      const float sum0 = softmax_row_to_tmem<T>(t_lane, ms0);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (lse2 != nullptr) lse2[(static_cast<long long>(b) * kHeads + h) * T + qt * 128 + r_tile] = ms0 + log2f(sum0);
      uint32_t oa[32], ob[32];
      mbar_wait(o_full, ph);
      tc_fence_after();
      load_o_row(t_lane + Cfg::kColO, oa, ob);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(epi_done);
      store_o_rows(oa, ob, 1.0f / sum0, sStage + static_cast<uint32_t>(warp) * 4096u, obase + static_cast<long long>(warp * 32) * kHidden, 32, lane);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

template <int T>
int launch_qt(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int batch, cudaStream_t stream) {
  using Cfg = QtCfg<T>;
  static bool configured = false;
  auto kern = attention_qt_kernel<T>;
  if (!configured) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "attention_qt: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
                       cudaGetErrorString(cudaGetLastError()));
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  CUtensorMap tq, tkv;
  const long long rows = static_cast<long long>(batch) * T;
  int rc = make_tmap_bf16_kmajor(&tq, qkv, rows, kQkvCols, kQkvCols, 128);
  if (rc != kOk) return rc;
  rc = make_tmap_bf16_kmajor(&tkv, qkv, rows, kQkvCols, kQkvCols, T);
  if (rc != kOk) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = batch * kHeads * (T / 128);
  const int slots = sms * 2;
  const int grid = items < slots ? items : slots;
  if (launch_pdl(kern, dim3(grid), dim3(kTcThreads), Cfg::kSmemBytes, stream, tq, tkv, out, lse2, items, sweep_reverse()) != cudaSuccess)
    return set_error(kErrCuda, "attention_qt_kernel: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return check_launch("attention_qt_kernel");
}

int launch_hm(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int batch, int mode, cudaStream_t stream) {
  using Cfg = TcCfg<144>;
  static bool configured = false;
  const bool p_in_tmem = mode == 2;
  auto kern = p_in_tmem ? attention_hm_kernel<true, false> : attention_hm_kernel<false, false>;
  if (!configured) {
    for (auto k : {attention_rw_kernel<false>, attention_rw_kernel<true>}) {
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    for (auto k : {attention_hm_kernel<true, false>, attention_hm_kernel<false, false>}) {
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
        return set_error(kErrCuda, "attention_hm: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
                         cudaGetErrorString(cudaGetLastError()));
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    configured = true;
  }
  CUtensorMap tm;
  int rc = make_tmap_bf16_kmajor(&tm, qkv, static_cast<long long>(batch) * 144, kQkvCols, kQkvCols, 144);
  if (rc != kOk) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int units = batch * kHeads;
  const int slots = sms * 2;
  const int grid = units < slots ? units : slots;
  static int trace_mode = -1;
  if (trace_mode < 0) { const char* e = getenv("JPDVT_ATTN_TRACE"); trace_mode = (e != nullptr && e[0] == '1') ? 1 : 0; }
  if (trace_mode) {   // developer path: synchronous, prints the event clocks of CTA 0 (relative to its first event)
    auto kt = p_in_tmem ? attention_hm_kernel<true, true> : attention_hm_kernel<false, true>;
    cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    cudaFuncSetAttribute(kt, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    constexpr int n = kTraceRoles * kTraceUnits * kTraceEvents;
    long long* d = nullptr;
    cudaMalloc(&d, n * sizeof(long long));
    cudaMemsetAsync(d, 0, n * sizeof(long long), stream);
    if (mode == 3) attention_rw_kernel<true><<<grid, kTcThreads, Cfg::kSmemBytes, stream>>>(tm, out, lse2, units, 0, d, sms);
    else kt<<<grid, kTcThreads, Cfg::kSmemBytes, stream>>>(tm, out, lse2, units, 0, d);
    long long hbuf[n];
    cudaMemcpyAsync(hbuf, d, sizeof(hbuf), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    cudaFree(d);
    long long t0 = 0;
    for (int i = 0; i < n; ++i) if (hbuf[i] != 0 && (t0 == 0 || hbuf[i] < t0)) t0 = hbuf[i];
    const char* names[kTraceRoles] = {"mma  ", "sm_w0", "sm_w3", "tma  "};
    for (int r = 0; r < kTraceRoles; ++r)
      for (int u = 0; u < kTraceUnits; ++u) {
        fprintf(stderr, "trace %s unit %d:", names[r], u);
        for (int e = 0; e < kTraceEvents; ++e) {
          const long long v = hbuf[(r * kTraceUnits + u) * kTraceEvents + e];
          fprintf(stderr, " %7lld", v ? v - t0 : -1LL);
        }
        fprintf(stderr, "\n");
      }
    return check_launch("attention_hm_kernel<trace>");
  }
  if (mode == 3) {
    if (launch_pdl(attention_rw_kernel<false>, dim3(grid), dim3(kTcThreads), Cfg::kSmemBytes, stream, tm, out, lse2, units, sweep_reverse(),
                   static_cast<long long*>(nullptr), sms) != cudaSuccess)
      return set_error(kErrCuda, "attention_rw_kernel: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return check_launch("attention_rw_kernel");
  }
  if (launch_pdl(kern, dim3(grid), dim3(kTcThreads), Cfg::kSmemBytes, stream, tm, out, lse2, units, sweep_reverse(), static_cast<long long*>(nullptr)) != cudaSuccess)
    return set_error(kErrCuda, "attention_hm_kernel: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return check_launch("attention_hm_kernel");
}

template <int T>
int launch_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int batch, cudaStream_t stream) {
  using Cfg = TcCfg<T>;
  static bool configured = false;
  static int trace_mode = -1;
  if (trace_mode < 0) { const char* e = getenv("JPDVT_ATTN_TRACE"); trace_mode = (e != nullptr && e[0] == '1') ? 1 : 0; }
  static int rem_t = -1;            // JPDVT_ATTN_REM=transposed: remainder scores as K Q_rem^T in columns of their own (A/B knob;
                                    // measured slower: 58.9 vs 54.1 us at B = 256 - DESIGN.md section 4); default: the split remainder
  if (rem_t < 0) { const char* e = getenv("JPDVT_ATTN_REM"); rem_t = (e != nullptr && e[0] == 't') ? 1 : 0; }
  static int rem_2 = -1;            // JPDVT_ATTN_REM=pipelined: remainder MMAs issued by two threads + P V k-steps per 64-key block of
                                    // P0 (A/B knob; shorter unit chain in the trace, same kernel time: DESIGN.md section 4)
  if (rem_2 < 0) { const char* e = getenv("JPDVT_ATTN_REM"); rem_2 = (e != nullptr && e[0] == 'p') ? 1 : 0; }
  constexpr bool kCanRT = TcCfg<T>::kSplit;
  auto kern = (kCanRT && rem_t) ? attention_tc_kernel<T, false, kCanRT, false>
              : (kCanRT && rem_2) ? attention_tc_kernel<T, false, false, kCanRT> : attention_tc_kernel<T, false, false, false>;
  auto kern_trace = (kCanRT && rem_t) ? attention_tc_kernel<T, true, kCanRT, false>
                    : (kCanRT && rem_2) ? attention_tc_kernel<T, true, false, kCanRT> : attention_tc_kernel<T, true, false, false>;
  if (!configured) {
    for (auto k : {kern, kern_trace}) {
      if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
        return set_error(kErrCuda, "attention_tc: cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes,
                         cudaGetErrorString(cudaGetLastError()));
      cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
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
  static int l2pf = -1;             // JPDVT_ATTN_L2PF=1: TMA L2 prefetch of the next unit's tiles (A/B knob; measured neutral: with two
                                    // CTAs per SM the other CTA already covers the load latency - 54.9 vs 54.2 us at B = 256)
  if (l2pf < 0) { const char* e = getenv("JPDVT_ATTN_L2PF"); l2pf = (e != nullptr && e[0] == '1') ? 1 : 0; }
  static int cta_cap = -1;          // JPDVT_ATTN_CTAS=1: one CTA per SM (the experiment behind DESIGN.md section 4: 76 vs 54 us)
  if (cta_cap < 0) { const char* e = getenv("JPDVT_ATTN_CTAS"); cta_cap = (e != nullptr && e[0] == '1') ? 1 : 0; }
  const int slots = sms * (cta_cap ? 1 : Cfg::kCtasPerSm);
  const int grid = units < slots ? units : slots;
  if (trace_mode) {   // developer path: synchronous, prints the event clocks of CTA 0 (relative to its first event)
    constexpr int n = kTraceRoles * kTraceUnits * kTraceEvents;
    long long* d = nullptr;
    cudaMalloc(&d, n * sizeof(long long));
    cudaMemsetAsync(d, 0, n * sizeof(long long), stream);
    kern_trace<<<grid, kTcThreads, Cfg::kSmemBytes, stream>>>(tm, out, lse2, units, 0, d, l2pf);
    long long h[n];
    cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    cudaFree(d);
    long long t0 = 0;
    for (int i = 0; i < n; ++i) if (h[i] != 0 && (t0 == 0 || h[i] < t0)) t0 = h[i];
    const char* names[kTraceRoles] = {"mma  ", "sm_w0", "sm_w3", "tma  "};
    for (int r = 0; r < kTraceRoles; ++r)
      for (int u = 0; u < kTraceUnits; ++u) {
        fprintf(stderr, "trace %s unit %d:", names[r], u);
        for (int e = 0; e < kTraceEvents; ++e) {
          const long long v = h[(r * kTraceUnits + u) * kTraceEvents + e];
          fprintf(stderr, " %7lld", v ? v - t0 : -1LL);
        }
        fprintf(stderr, "\n");
      }
    return check_launch("attention_tc_kernel<trace>");
  }
  if (launch_pdl(kern, dim3(grid), dim3(kTcThreads), Cfg::kSmemBytes, stream, tm, out, lse2, units, sweep_reverse(), static_cast<long long*>(nullptr), l2pf) != cudaSuccess)
    return set_error(kErrCuda, "attention_tc_kernel: launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  return check_launch("attention_tc_kernel");
}

}  // namespace

bool attention_tc_supported(int tokens) { return tokens == 144 || tokens == 256 || tokens == 324; }

int launch_attention_tc(const __nv_bfloat16* qkv, __nv_bfloat16* out, float* lse2, int batch, int tokens, cudaStream_t stream) {
  if (batch <= 0) return kOk;
  if ((reinterpret_cast<uintptr_t>(qkv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return set_error(kErrBadArg, "attention_tc: pointers must be 16-byte aligned");
  switch (tokens) {
    case 144: {
      static int warps8 = -1;       // JPDVT_ATTN_WARPS=8: the eight-softmax-warp kernel (A/B knob; measured slower: 57.8 vs 54.0 us at
                                    // B = 256 - DESIGN.md section 4); default: four
      if (warps8 < 0) { const char* e = getenv("JPDVT_ATTN_WARPS"); warps8 = (e != nullptr && e[0] == '8') ? 1 : 0; }
      static int hybrid = -1;       // default (JPDVT_ATTN_REM unset or =warp): attention_rw_kernel, 42 us at B = 256; =split: the round-1 kernel
                                    // (54 us; every remainder MMA on tcgen05); =transposed / =pipelined: its variants (launch_tc);
                                    // JPDVT_ATTN_REM=hybrid: main tile on tcgen05, the 16-row remainder on mma.sync (attention_hm_kernel);
                                    // JPDVT_ATTN_REM=hybrid-tmem: the same with the main tile's probabilities kept in tensor memory; JPDVT_ATTN_REM=warp: the
                                    // remainder as one warp's register-resident job (attention_rw_kernel)
      if (hybrid < 0) { const char* e = getenv("JPDVT_ATTN_REM"); hybrid = (e == nullptr || e[0] == 'w') ? 3 : (e[0] != 'h' ? 0 : (strstr(e, "tmem") != nullptr ? 2 : 1)); }
      if (hybrid) return launch_hm(qkv, out, lse2, batch, hybrid, stream);
      return warps8 ? launch_tc8(qkv, out, lse2, batch, stream) : launch_tc<144>(qkv, out, lse2, batch, stream);
    }
    case 256: {
      static int qt = -1;           // JPDVT_ATTN_QT=0: the dual-tile kernel (one unit per CTA, P tiles in shared memory) instead of the
                                    // query-tile kernel (attention_qt_kernel: two chains per SM, P in tensor memory)
      if (qt < 0) { const char* e = getenv("JPDVT_ATTN_QT"); qt = (e != nullptr && e[0] == '0') ? 0 : 1; }
      return qt ? launch_qt<256>(qkv, out, lse2, batch, stream) : launch_tc<256>(qkv, out, lse2, batch, stream);
    }
    case 324: {
      static int w8 = -1;           // JPDVT_ATTN_SEQ_WARPS=8: two threads per score row (attention_seq8_kernel); default: four softmax warps
      if (w8 < 0) { const char* e = getenv("JPDVT_ATTN_SEQ_WARPS"); w8 = (e != nullptr && e[0] == '8') ? 1 : 0; }
      static int ks = -1;           // default: key-split form, two chains per SM (attention_ks_kernel, 96.6 us at B = 128);
                                    // JPDVT_ATTN_SEQ_SPLIT=0: the sequential-tile kernel (121.5 us)
      if (ks < 0) { const char* e = getenv("JPDVT_ATTN_SEQ_SPLIT"); ks = (e != nullptr && e[0] == '0') ? 0 : 1; }
      if (ks) return launch_ks(qkv, out, lse2, batch, stream);
      return w8 ? launch_seq8<336, 324>(qkv, out, lse2, batch, stream) : launch_tc_seq<336, 324>(qkv, out, lse2, batch, stream);
    }
    default: return set_error(kErrUnsupported, "attention_tc: %d tokens not instantiated", tokens);
  }
}

}  // namespace jp
