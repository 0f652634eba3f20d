// Persistent warp-specialised tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue).
//
//   warp 0 : TMA producer   (cp.async.bulk.tensor 2-D, 128B swizzle, mbarrier complete_tx)
//   warp 1 : MMA issuer     (one lane issues tcgen05.mma 128 x BN x 16, accumulators in TMEM, double buffered)
//   warps 2-5 : epilogue    (tcgen05.ld 32 lanes x 32 columns per warp, fused bias / GELU / gated residual / ...)
//
// Replaces the cuBLAS calls behind nn.Linear in the reference denoiser (image_model/models.py:108-121,132,176-179,
// timm Attention.qkv/proj, Mlp.fc1/fc2, PatchEmbed.proj) - see SURVEY.md 2.1.
// A and W are both K-major bf16, so both operands use the canonical K-major SWIZZLE_128B smem layout that a
// {64 x rows} TMA box produces.  fp32 accumulation.
#include <cstdio>

#include "common.cuh"
#include "ptx.cuh"

namespace jp {

constexpr int BM = 128;
constexpr int BK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;   // two accumulator buffers (power of two for BN in {64,128,256})
  static constexpr int kBarBytes = (2 * kStages + 4) * 8 + 16;
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // +1024: manual 1 KiB alignment
};

__device__ __forceinline__ void st_bf16x8(__nv_bfloat16* dst, const float* v) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}

// One 32-column chunk of one accumulator row.  `acc` holds fp32 accumulators for columns [n0, n0+32).
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, float (&acc)[32], long long row, int n0, bool row_ok,
                                               const float* xt_row) {
  // bias (uniform across the warp -> broadcast loads served by L1)
  const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float4 b = __ldg(b4 + j);
    acc[4 * j + 0] += b.x; acc[4 * j + 1] += b.y; acc[4 * j + 2] += b.z; acc[4 * j + 3] += b.w;
  }
  if (!row_ok) return;

  if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = gelu_tanh(acc[j]);
  }
  if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_BIAS_GELU_BF16 || EPI == EPI_BIAS_BF16_F32) {
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ldo + n0;
#pragma unroll
    for (int j = 0; j < 4; ++j) st_bf16x8(o + 8 * j, &acc[8 * j]);
    if constexpr (EPI == EPI_BIAS_BF16_F32) {
      if (p.out2 != nullptr) {
        float4* o2 = reinterpret_cast<float4*>(p.out2 + row * p.ldo + n0);
#pragma unroll
        for (int j = 0; j < 8; ++j) o2[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
      }
    }
  }
  if constexpr (EPI == EPI_BIAS_F32) {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * p.ldo + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
  }
  if constexpr (EPI == EPI_GATE_RESID_F32) {
    const long long sample = row / p.tokens;
    const float4* g4 = reinterpret_cast<const float4*>(p.gate + sample * p.gate_stride + n0);
    const float4* r4 = reinterpret_cast<const float4*>(p.resid + row * p.ldo + n0);
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * p.ldo + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 g = __ldg(g4 + j);
      float4 r = r4[j];
      r.x = fmaf(g.x, acc[4 * j + 0], r.x); r.y = fmaf(g.y, acc[4 * j + 1], r.y);
      r.z = fmaf(g.z, acc[4 * j + 2], r.z); r.w = fmaf(g.w, acc[4 * j + 3], r.w);
      o[j] = r;
    }
  }
  if constexpr (EPI == EPI_PATCH_EMBED_F32) {
    const int tok = static_cast<int>(row % p.tokens);
    const float4* pos4 = reinterpret_cast<const float4*>(p.pos + static_cast<long long>(tok) * p.N + n0);
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * p.ldo + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 v = __ldg(pos4 + j);
      v.x += acc[4 * j + 0]; v.y += acc[4 * j + 1]; v.z += acc[4 * j + 2]; v.w += acc[4 * j + 3];
#pragma unroll
      for (int d = 0; d < kLatent; ++d) {
        float4 w = __ldg(reinterpret_cast<const float4*>(p.w_in_t + static_cast<long long>(d) * p.N + n0) + j);
        v.x = fmaf(xt_row[d], w.x, v.x); v.y = fmaf(xt_row[d], w.y, v.y);
        v.z = fmaf(xt_row[d], w.z, v.z); v.w = fmaf(xt_row[d], w.w, v.w);
      }
      o[j] = v;
    }
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;   // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;             // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = p.N / BN;
  const int num_tiles = num_m * num_n;
  const int num_kb = p.K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int i = 0; i < Cfg::kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / num_n, n_blk = tile % num_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(&tma_a, &full_bar[stage], sa, kb * BK, m_blk * BM);
          tma_load_2d(&tma_b, &full_bar[stage], sb, kb * BK, n_blk * BN);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = umma_desc_k_sw128(a_addr + k * UMMA_K * 2);
            const uint64_t db = umma_desc_k_sw128(b_addr + k * UMMA_K * 2);
            umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // smem slot reusable once these MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);        // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;             // TMEM lane quadrant this warp may access
    int acc = 0; uint32_t acc_phase = 0;
    __shared__ float s_w2[kLatent * 64];
    __shared__ float s_b2[kLatent];
    if constexpr (EPI == EPI_HEAD) {
      for (int i = threadIdx.x - 64; i < kLatent * 64; i += 128) s_w2[i] = p.w2[i];
      if (threadIdx.x - 64 < kLatent) s_b2[threadIdx.x - 64] = p.b2[threadIdx.x - 64];
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / num_n, n_blk = tile % num_n;
      const long long row = static_cast<long long>(m_blk) * BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      float xt_row[kLatent];
      if constexpr (EPI == EPI_PATCH_EMBED_F32) {
        if (row_ok) {
          const float4* x4 = reinterpret_cast<const float4*>(p.xt + row * kLatent);
          float4 u = x4[0], v = x4[1];
          xt_row[0] = u.x; xt_row[1] = u.y; xt_row[2] = u.z; xt_row[3] = u.w;
          xt_row[4] = v.x; xt_row[5] = v.y; xt_row[6] = v.z; xt_row[7] = v.w;
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN);
      if constexpr (EPI == EPI_HEAD) {
        static_assert(EPI != EPI_HEAD || BN == 64, "head epilogue needs the whole 64-wide row");
        uint32_t r0[32], r1[32];
        tmem_ld_32x32(t_row, r0);
        tmem_ld_32x32(t_row + 32, r1);
        tmem_ld_wait();
        float h[64];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          h[j] = silu(__uint_as_float(r0[j]) + __ldg(p.bias + j));
          h[32 + j] = silu(__uint_as_float(r1[j]) + __ldg(p.bias + 32 + j));
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
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + row * p.ldo);
          dst[0] = make_float4(o[0], o[1], o[2], o[3]);
          dst[1] = make_float4(o[4], o[5], o[6], o[7]);
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + c * 32, r);
          tmem_ld_wait();
          float accv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) accv[j] = __uint_as_float(r[j]);
          epilogue_chunk<EPI>(p, accv, row, n_blk * BN + c * 32, row_ok, xt_row);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
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

template <int BN, int EPI>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  auto kern = gemm_kernel<BN, EPI>;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess)
      return set_error(kErrCuda, "cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::kSmemBytes, cudaGetErrorString(cudaGetLastError()));
    attr_set = true;
  }
  const int tiles = ((p.M + BM - 1) / BM) * (p.N / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  kern<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tb, p);
  return check_launch("gemm_kernel");
}

int launch_gemm(int epi, const __nv_bfloat16* a, long long lda, const __nv_bfloat16* w, long long ldw, const GemmParams& p,
                cudaStream_t stream) {
  if (p.M <= 0) return kOk;
  if (p.K % BK != 0 || p.K <= 0) return set_error(kErrBadArg, "gemm: K=%d must be a positive multiple of %d", p.K, BK);
  if ((lda % 8) != 0 || (ldw % 8) != 0) return set_error(kErrBadArg, "gemm: leading dimensions must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(w) & 15))
    return set_error(kErrBadArg, "gemm: operand pointers must be 16-byte aligned");
  const int bn = (epi == EPI_HEAD) ? 64 : ((p.N % 256 == 0) ? 256 : 128);
  if (p.N % bn != 0) return set_error(kErrBadArg, "gemm: N=%d is not a multiple of the %d-wide tile", p.N, bn);
  if (epi == EPI_HEAD && p.N != 64) return set_error(kErrBadArg, "gemm: head epilogue requires N == 64");
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_kmajor(&ta, a, p.M, p.K, lda, BM);
  if (rc != kOk) return rc;
  rc = make_tmap_bf16_kmajor(&tb, w, p.N, p.K, ldw, bn);
  if (rc != kOk) return rc;
#define JP_CASE(E)                                                             \
  case E:                                                                      \
    return bn == 256 ? launch_cfg<256, E>(ta, tb, p, stream) : launch_cfg<128, E>(ta, tb, p, stream);
  switch (epi) {
    JP_CASE(EPI_BIAS_BF16)
    JP_CASE(EPI_BIAS_GELU_BF16)
    JP_CASE(EPI_GATE_RESID_F32)
    JP_CASE(EPI_PATCH_EMBED_F32)
    JP_CASE(EPI_BIAS_F32)
    JP_CASE(EPI_BIAS_BF16_F32)
    case EPI_HEAD:
      return launch_cfg<64, EPI_HEAD>(ta, tb, p, stream);
    default:
      return set_error(kErrBadArg, "gemm: unknown epilogue %d", epi);
  }
#undef JP_CASE
}

}  // namespace jp
