// Data-parallel optimizer step over NVLink / NVSwitch peer memory: gradient reduce-scatter + AdamW + EMA + all-gather of
// the bf16 operand copy as ONE kernel per rank, no NCCL call on the data path.
//
// Replaces, for world > 1, what DistributedDataParallel's bucketed all-reduce + torch.optim.AdamW.step() + update_ema()
// do in the reference trainer (image_model/train_JPDVT.py:231, 370-372, 36-46).  Every rank maps every other rank's flat
// fp32 gradient buffer and flat bf16 operand buffer (torch symmetric memory: CUDA VMM handles + an NVSwitch multicast
// object; host plumbing in jpdvt_mt_ntnu_b200/peer.py).  Rank r owns the contiguous slice [shard_begin, shard_end) of
// the flat parameter index space:
//
//   barrier A   "every rank's backward has finished writing its gradients"      (flags in peer memory, system scope)
//   for each group of 8 parameters of my slice:
//       g    = sum over ranks of grads[rank][i]        multimem.ld_reduce (the switch adds, one 16-byte response per
//                                                      request) or plain peer loads in rank order
//       p, m, v, ema  updated locally  (fp32 master state exists ONLY on the owner: 1/world of the optimizer traffic;
//                                       the few fp32 values the kernels read directly - biases, timestep MLP, head -
//                                       are replicated like the bf16 copy)
//       bf16(p) -> every rank's operand buffer         multimem.st (one store, the switch replicates) or peer stores
//   barrier B   "my stores have landed everywhere and I no longer read anyone's gradients"
//
// so the step costs one pass over 1/world of the optimizer state plus (world-1)/world of 4 + 2 bytes per parameter over
// the links, instead of an all-reduce (2 x 4 bytes per parameter over the links) followed by the full 38 B/param pass.
#include <cstdlib>

#include "../../include/jpdvt_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace jp {

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add_f32x4(const float* mc_addr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_addr)
               : "memory");
  return r;
}
__device__ __forceinline__ void multimem_st_b32x4(void* mc_addr, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}

// Wait until flag[q] has reached `epoch` (wrap-safe) for every peer q != rank; one polling thread per peer.
// Returns false on timeout (a peer died or never launched its step): the caller records it and carries on, so a broken
// job fails loudly on the host instead of hanging the GPU.
__device__ __forceinline__ bool wait_flags(const uint32_t* flags, int world, int rank, uint32_t epoch, unsigned long long timeout_ns) {
  bool ok = true;
  const int q = threadIdx.x;
  if (q < world && q != rank) {
    const unsigned long long t0 = global_timer_ns();
    while (static_cast<int32_t>(ld_acquire_sys(flags + q) - epoch) < 0) {
      __nanosleep(64);
      if (global_timer_ns() - t0 > timeout_ns) { ok = false; break; }
    }
  }
  return ok;
}

struct AdamHyper {
  float grad_scale, lr, beta1, beta2, eps, weight_decay, step_size, inv_sqrt_bc2, ema_decay;
  const long long* step_dev;      // != null: step count on the device, bias corrections derived in the kernel (graph replay)
};

// step-dependent scalars of one launch: from the host (parameters) or from the device-side counters
struct StepScalars { uint32_t epoch; float step_size, inv_sqrt_bc2; };
__device__ __forceinline__ StepScalars step_scalars(const jpdvt_peer_step& px, const AdamHyper& h) {
  __shared__ StepScalars sc;
  if (threadIdx.x == 0) {
    sc.epoch = px.epoch_dev != nullptr ? *reinterpret_cast<volatile uint32_t*>(px.epoch_dev) + 1u : px.epoch;
    sc.step_size = h.step_size; sc.inv_sqrt_bc2 = h.inv_sqrt_bc2;
    if (h.step_dev != nullptr) {
      const double st = static_cast<double>(*h.step_dev);
      sc.step_size = static_cast<float>(static_cast<double>(h.lr) / (1.0 - pow(static_cast<double>(h.beta1), st)));
      sc.inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(1.0 - pow(static_cast<double>(h.beta2), st)));
    }
  }
  __syncthreads();
  return sc;
}

template <bool MC>
__global__ void __launch_bounds__(256)
peer_adamw_ema_kernel(const jpdvt_peer_step px, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                      float* __restrict__ ema, const AdamHyper h) {
  const int world = px.world, rank = px.rank;
  const unsigned long long timeout_ns = static_cast<unsigned long long>(px.timeout_ms) * 1000000ull;
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(px.signals[rank]);
  __shared__ int s_last;
  const StepScalars sc = step_scalars(px, h);
  const uint32_t epoch = sc.epoch;

  // ---- barrier A: the gradients of every rank are final -------------------------------------------------------------
  if (blockIdx.x == 0 && threadIdx.x < world && static_cast<int>(threadIdx.x) != rank) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(px.signals[threadIdx.x]) + rank, epoch);
  }
  if (!wait_flags(my_flags, world, rank, epoch, timeout_ns)) atomicExch(px.status, 1);
  __syncthreads();

  // ---- my slice: reduce, update, broadcast ---------------------------------------------------------------------------
  const long long g0 = px.shard_begin / 8, g1 = px.shard_end / 8;          // groups of 8 parameters
  const float wd_mul = 1.0f - h.lr * h.weight_decay;
  for (long long g = g0 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; g < g1;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long i = g * 8;
    float gr[8];
    if constexpr (MC) {
      const float4 a = multimem_ld_reduce_add_f32x4(px.grads_mc + i), b = multimem_ld_reduce_add_f32x4(px.grads_mc + i + 4);
      gr[0] = a.x; gr[1] = a.y; gr[2] = a.z; gr[3] = a.w; gr[4] = b.x; gr[5] = b.y; gr[6] = b.z; gr[7] = b.w;
    } else {
      float4 a[JPDVT_MAX_PEERS], b[JPDVT_MAX_PEERS];
#pragma unroll
      for (int q = 0; q < JPDVT_MAX_PEERS; ++q) {
        if (q < world) {                                                    // every load in flight before the first add
          a[q] = __ldcg(reinterpret_cast<const float4*>(px.grads[q] + i));
          b[q] = __ldcg(reinterpret_cast<const float4*>(px.grads[q] + i + 4));
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) gr[k] = 0.f;
#pragma unroll
      for (int q = 0; q < JPDVT_MAX_PEERS; ++q) {                           // rank order: every rank would add in the same order
        if (q < world) {
          gr[0] += a[q].x; gr[1] += a[q].y; gr[2] += a[q].z; gr[3] += a[q].w;
          gr[4] += b[q].x; gr[5] += b[q].y; gr[6] += b[q].z; gr[7] += b[q].w;
        }
      }
    }
    float4 p0 = reinterpret_cast<const float4*>(p + i)[0], p1 = reinterpret_cast<const float4*>(p + i)[1];
    float4 m0 = reinterpret_cast<const float4*>(m + i)[0], m1 = reinterpret_cast<const float4*>(m + i)[1];
    float4 v0 = reinterpret_cast<const float4*>(v + i)[0], v1 = reinterpret_cast<const float4*>(v + i)[1];
    float pa[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float ma[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    float va[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {                                           // identical arithmetic to adamw_ema_kernel (optim.cu)
      const float gk = gr[k] * h.grad_scale;
      pa[k] *= wd_mul;
      ma[k] = h.beta1 * ma[k] + (1.0f - h.beta1) * gk;
      va[k] = h.beta2 * va[k] + (1.0f - h.beta2) * gk * gk;
      const float denom = sqrtf(va[k]) * sc.inv_sqrt_bc2 + h.eps;
      pa[k] -= sc.step_size * (ma[k] / denom);
    }
    reinterpret_cast<float4*>(p + i)[0] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(p + i)[1] = make_float4(pa[4], pa[5], pa[6], pa[7]);
    reinterpret_cast<float4*>(m + i)[0] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(m + i)[1] = make_float4(ma[4], ma[5], ma[6], ma[7]);
    reinterpret_cast<float4*>(v + i)[0] = make_float4(va[0], va[1], va[2], va[3]);
    reinterpret_cast<float4*>(v + i)[1] = make_float4(va[4], va[5], va[6], va[7]);
    if (ema != nullptr) {
      float4 e0 = reinterpret_cast<const float4*>(ema + i)[0], e1 = reinterpret_cast<const float4*>(ema + i)[1];
      const float d = h.ema_decay, c = 1.0f - h.ema_decay;
      e0.x = d * e0.x + c * pa[0]; e0.y = d * e0.y + c * pa[1]; e0.z = d * e0.z + c * pa[2]; e0.w = d * e0.w + c * pa[3];
      e1.x = d * e1.x + c * pa[4]; e1.y = d * e1.y + c * pa[5]; e1.z = d * e1.z + c * pa[6]; e1.w = d * e1.w + c * pa[7];
      reinterpret_cast<float4*>(ema + i)[0] = e0;
      reinterpret_cast<float4*>(ema + i)[1] = e1;
    }
    bool rep = false;                       // does the group touch a range whose fp32 values every rank reads (biases, ...)?
    for (int k = 0; k < px.n_f32_ranges; ++k) rep = rep || (i < px.f32_ranges[2 * k + 1] && i + 8 > px.f32_ranges[2 * k]);
    if (rep) {
      const uint4 q0 = make_uint4(__float_as_uint(pa[0]), __float_as_uint(pa[1]), __float_as_uint(pa[2]), __float_as_uint(pa[3]));
      const uint4 q1 = make_uint4(__float_as_uint(pa[4]), __float_as_uint(pa[5]), __float_as_uint(pa[6]), __float_as_uint(pa[7]));
      if constexpr (MC) {
        multimem_st_b32x4(px.params_mc + i, q0);
        multimem_st_b32x4(px.params_mc + i + 4, q1);
      } else {
#pragma unroll
        for (int q = 0; q < JPDVT_MAX_PEERS; ++q) {
          if (q < world && q != rank) {
            *reinterpret_cast<uint4*>(px.params[q] + i) = q0;
            *reinterpret_cast<uint4*>(px.params[q] + i + 4) = q1;
          }
        }
      }
    }
    uint4 w;
    w.x = pack_bf16(pa[0], pa[1]); w.y = pack_bf16(pa[2], pa[3]); w.z = pack_bf16(pa[4], pa[5]); w.w = pack_bf16(pa[6], pa[7]);
    if constexpr (MC) {
      multimem_st_b32x4(px.weights_mc + i, w);
    } else {
#pragma unroll
      for (int q = 0; q < JPDVT_MAX_PEERS; ++q)
        if (q < world) *reinterpret_cast<uint4*>(px.weights_bf16[q] + i) = w;
    }
  }

  // ---- barrier B: the last CTA of this rank tells every peer and waits for theirs ------------------------------------
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(px.local_sync, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
    __threadfence();
  }
  __syncthreads();
  if (s_last) {
    if (threadIdx.x == 0) *px.local_sync = 0u;                              // ready for the next call (stream ordered)
    if (threadIdx.x < world && static_cast<int>(threadIdx.x) != rank) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(px.signals[threadIdx.x]) + JPDVT_MAX_PEERS + rank, epoch);
    }
    if (!wait_flags(my_flags + JPDVT_MAX_PEERS, world, rank, epoch, timeout_ns)) atomicExch(px.status, 2);
    __syncthreads();
    if (threadIdx.x == 0 && px.epoch_dev != nullptr) *px.epoch_dev = epoch;      // every CTA of this launch has read the old value
  }
}

// ---- bulk variant -----------------------------------------------------------------------------------------------------
// Per-thread peer loads keep at most ~16 KB per SM in flight (the LSU's outstanding-miss budget), which at the ~10 us loaded
// latency of an NVLink read caps the kernel near 230 GB/s of gradient traffic (measured, 2 GPUs: 1.15 ms).  Here the remote
// gradient tiles are pulled by the bulk-copy engine instead (cp.async.bulk global -> shared, mbarrier completion): a CTA
// keeps kStages x (world - 1) x 8 KB in flight, the threads only ever read shared memory, and the bf16 results leave the
// same way (cp.async.bulk shared -> every rank's operand buffer), so no thread waits on a link.
constexpr int kTile = 2048;                      // parameters per tile: 256 threads x 8
constexpr int kTileBytes = kTile * 4;
constexpr int kBulkStages = 4;

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}

__global__ void __launch_bounds__(256)
peer_adamw_ema_bulk_kernel(const jpdvt_peer_step px, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                           float* __restrict__ ema, const AdamHyper h, const int stages) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int world = px.world, rank = px.rank, nrem = world - 1;
  const unsigned long long timeout_ns = static_cast<unsigned long long>(px.timeout_ms) * 1000000ull;
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(px.signals[rank]);
  // [stages][nrem][kTileBytes] remote gradient tiles | [2][kTile * 2] bf16 output tiles | full barriers
  uint8_t* in_buf = smem;
  uint8_t* out_buf = smem + static_cast<size_t>(stages) * nrem * kTileBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(out_buf + 2 * kTile * 2);
  __shared__ int s_last;
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < stages; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
  }
  // ---- barrier A ---------------------------------------------------------------------------------------------------
  const StepScalars sc = step_scalars(px, h);
  const uint32_t epoch = sc.epoch;
  if (blockIdx.x == 0 && tid < world && tid != rank) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(px.signals[tid]) + rank, epoch);
  }
  if (!wait_flags(my_flags, world, rank, epoch, timeout_ns)) atomicExch(px.status, 1);
  __syncthreads();

  const long long t0 = px.shard_begin / kTile, t1 = px.shard_end / kTile;   // tiles of my slice
  const long long first = t0 + blockIdx.x, stride = gridDim.x;
  auto issue = [&](long long tile, int stage) {                              // thread 0: pull the remote tiles of `tile`
    mbar_expect_tx(&full[stage], static_cast<uint32_t>(nrem * kTileBytes));
    int slot = 0;
    for (int q = 0; q < world; ++q) {
      if (q == rank) continue;
      bulk_load(in_buf + (static_cast<size_t>(stage) * nrem + slot) * kTileBytes, px.grads[q] + tile * kTile, kTileBytes, &full[stage]);
      ++slot;
    }
  };
  if (tid == 0) {
    asm volatile("fence.proxy.async;" ::: "memory");       // the acquire above orders generic accesses; the bulk engine is another proxy
    for (int k = 0; k < stages - 1; ++k)
      if (first + k * stride < t1) issue(first + k * stride, k);
  }
  const float wd_mul = 1.0f - h.lr * h.weight_decay;
  long long it = 0;
  for (long long tile = first; tile < t1; tile += stride, ++it) {
    const int stage = static_cast<int>(it % stages);
    if (tid == 0) {                                                          // refill the stage consumed in the previous iteration
      const long long nxt = tile + static_cast<long long>(stages - 1) * stride;
      if (nxt < t1) issue(nxt, static_cast<int>((it + stages - 1) % stages));
    }
    const long long i = tile * kTile + tid * 8;
    // local operands first: their latency overlaps the wait for the remote tile
    const float4 ga = __ldcs(reinterpret_cast<const float4*>(px.grads[rank] + i)), gb = __ldcs(reinterpret_cast<const float4*>(px.grads[rank] + i) + 1);
    const float4 p0 = reinterpret_cast<const float4*>(p + i)[0], p1 = reinterpret_cast<const float4*>(p + i)[1];
    const float4 m0 = reinterpret_cast<const float4*>(m + i)[0], m1 = reinterpret_cast<const float4*>(m + i)[1];
    const float4 v0 = reinterpret_cast<const float4*>(v + i)[0], v1 = reinterpret_cast<const float4*>(v + i)[1];
    float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
    if (ema != nullptr) { e0 = reinterpret_cast<const float4*>(ema + i)[0]; e1 = reinterpret_cast<const float4*>(ema + i)[1]; }
    mbar_wait(&full[stage], static_cast<uint32_t>((it / stages) & 1));
    // sum in rank order (the local contribution takes its place in the sequence), so every owner adds identically
    float gr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int slot = 0;
    for (int q = 0; q < world; ++q) {
      float4 a, b;
      if (q == rank) { a = ga; b = gb; }
      else {
        const float4* src = reinterpret_cast<const float4*>(in_buf + (static_cast<size_t>(stage) * nrem + slot) * kTileBytes) + tid * 2;
        a = src[0]; b = src[1];
        ++slot;
      }
      gr[0] += a.x; gr[1] += a.y; gr[2] += a.z; gr[3] += a.w; gr[4] += b.x; gr[5] += b.y; gr[6] += b.z; gr[7] += b.w;
    }
    float pa[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float ma[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    float va[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {                                            // identical arithmetic to adamw_ema_kernel (optim.cu)
      const float gk = gr[k] * h.grad_scale;
      pa[k] *= wd_mul;
      ma[k] = h.beta1 * ma[k] + (1.0f - h.beta1) * gk;
      va[k] = h.beta2 * va[k] + (1.0f - h.beta2) * gk * gk;
      const float denom = sqrtf(va[k]) * sc.inv_sqrt_bc2 + h.eps;
      pa[k] -= sc.step_size * (ma[k] / denom);
    }
    reinterpret_cast<float4*>(p + i)[0] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(p + i)[1] = make_float4(pa[4], pa[5], pa[6], pa[7]);
    reinterpret_cast<float4*>(m + i)[0] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(m + i)[1] = make_float4(ma[4], ma[5], ma[6], ma[7]);
    reinterpret_cast<float4*>(v + i)[0] = make_float4(va[0], va[1], va[2], va[3]);
    reinterpret_cast<float4*>(v + i)[1] = make_float4(va[4], va[5], va[6], va[7]);
    if (ema != nullptr) {
      const float d = h.ema_decay, c = 1.0f - h.ema_decay;
      e0.x = d * e0.x + c * pa[0]; e0.y = d * e0.y + c * pa[1]; e0.z = d * e0.z + c * pa[2]; e0.w = d * e0.w + c * pa[3];
      e1.x = d * e1.x + c * pa[4]; e1.y = d * e1.y + c * pa[5]; e1.z = d * e1.z + c * pa[6]; e1.w = d * e1.w + c * pa[7];
      reinterpret_cast<float4*>(ema + i)[0] = e0;
      reinterpret_cast<float4*>(ema + i)[1] = e1;
    }
    bool rep = false;
    for (int k = 0; k < px.n_f32_ranges; ++k) rep = rep || (i < px.f32_ranges[2 * k + 1] && i + 8 > px.f32_ranges[2 * k]);
    if (rep) {                                                               // rare (biases, timestep MLP, head): direct peer stores
      const float4 q0 = make_float4(pa[0], pa[1], pa[2], pa[3]), q1 = make_float4(pa[4], pa[5], pa[6], pa[7]);
      for (int q = 0; q < world; ++q) {
        if (q != rank) {
          reinterpret_cast<float4*>(px.params[q] + i)[0] = q0;
          reinterpret_cast<float4*>(px.params[q] + i)[1] = q1;
        }
      }
    }
    // bf16 tile -> shared -> every rank's operand buffer by bulk stores (two output tiles: the stores of tile it - 2 must
    // have READ their buffer before it is refilled)
    uint8_t* ob = out_buf + (it & 1) * (kTile * 2);
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncthreads();
    uint4 w;
    w.x = pack_bf16(pa[0], pa[1]); w.y = pack_bf16(pa[2], pa[3]); w.z = pack_bf16(pa[4], pa[5]); w.w = pack_bf16(pa[6], pa[7]);
    reinterpret_cast<uint4*>(ob)[tid] = w;
    fence_proxy_async_smem();
    __syncthreads();               // also: every thread is done reading in_buf[stage], so iteration it + 1 may refill it
    if (tid == 0) {
      for (int q = 0; q < world; ++q) bulk_store(px.weights_bf16[q] + tile * kTile, ob, kTile * 2);
      tma_store_commit();
    }
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // my bulk stores are complete (visible) ...

  // ---- barrier B -----------------------------------------------------------------------------------------------------
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(px.local_sync, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
    __threadfence();
  }
  __syncthreads();
  if (s_last) {
    if (tid == 0) *px.local_sync = 0u;
    if (tid < world && tid != rank) {
      __threadfence_system();
      st_release_sys(reinterpret_cast<uint32_t*>(px.signals[tid]) + JPDVT_MAX_PEERS + rank, epoch);
    }
    if (!wait_flags(my_flags + JPDVT_MAX_PEERS, world, rank, epoch, timeout_ns)) atomicExch(px.status, 2);
    __syncthreads();
    if (threadIdx.x == 0 && px.epoch_dev != nullptr) *px.epoch_dev = epoch;      // every CTA of this launch has read the old value
  }
}

}  // namespace jp

using namespace jp;

extern "C" {

static int peer_step_impl(const jpdvt_peer_step* px, float* p, float* m, float* v, float* ema_or_null, int64_t step,
                          const int64_t* step_dev, float grad_scale, float lr, float beta1, float beta2, float eps,
                          float weight_decay, float ema_decay, void* stream);

int jpdvt_adamw_ema_peer(const jpdvt_peer_step* px, float* p, float* m, float* v, float* ema_or_null, int64_t step,
                         float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay, float ema_decay,
                         void* stream) {
  if (step < 1) return set_error(kErrBadArg, "adamw_ema_peer: step counts from 1");
  return peer_step_impl(px, p, m, v, ema_or_null, step, nullptr, grad_scale, lr, beta1, beta2, eps, weight_decay, ema_decay, stream);
}

int jpdvt_adamw_ema_peer_dev(const jpdvt_peer_step* px, float* p, float* m, float* v, float* ema_or_null, const int64_t* step_dev,
                             float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay,
                             float ema_decay, void* stream) {
  if (!step_dev || !px || !px->epoch_dev) return set_error(kErrBadArg, "adamw_ema_peer_dev: device step / epoch counters required");
  return peer_step_impl(px, p, m, v, ema_or_null, 1, step_dev, grad_scale, lr, beta1, beta2, eps, weight_decay, ema_decay, stream);
}

static int peer_step_impl(const jpdvt_peer_step* px, float* p, float* m, float* v, float* ema_or_null, int64_t step,
                          const int64_t* step_dev, float grad_scale, float lr, float beta1, float beta2, float eps,
                          float weight_decay, float ema_decay, void* stream) {
  if (!px || !p || !m || !v) return set_error(kErrBadArg, "adamw_ema_peer: null pointer");
  if (px->world < 2 || px->world > JPDVT_MAX_PEERS || px->rank < 0 || px->rank >= px->world)
    return set_error(kErrBadArg, "adamw_ema_peer: world=%d rank=%d (2..%d ranks)", px->world, px->rank, JPDVT_MAX_PEERS);
  if ((px->shard_begin & 7) || (px->shard_end & 7) || px->shard_end < px->shard_begin)
    return set_error(kErrBadArg, "adamw_ema_peer: the slice [%lld, %lld) must be multiples of 8 parameters",
                     static_cast<long long>(px->shard_begin), static_cast<long long>(px->shard_end));
  if (!px->local_sync || !px->status) return set_error(kErrBadArg, "adamw_ema_peer: null sync / status word");
  const bool mc = px->grads_mc != nullptr && px->weights_mc != nullptr && px->params_mc != nullptr;
  if (px->n_f32_ranges < 0 || px->n_f32_ranges > JPDVT_MAX_F32_RANGES)
    return set_error(kErrBadArg, "adamw_ema_peer: n_f32_ranges=%d (0..%d)", px->n_f32_ranges, JPDVT_MAX_F32_RANGES);
  for (int q = 0; q < px->world; ++q) {
    if (!px->signals[q] || (!mc && (!px->grads[q] || !px->weights_bf16[q] || !px->params[q])))
      return set_error(kErrBadArg, "adamw_ema_peer: rank %d's buffers are not mapped", q);
  }
  AdamHyper h;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), static_cast<double>(step));
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), static_cast<double>(step));
  h.grad_scale = grad_scale; h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay;
  h.step_size = static_cast<float>(lr / bc1);
  h.inv_sqrt_bc2 = static_cast<float>(1.0 / sqrt(bc2));
  h.ema_decay = ema_decay;
  h.step_dev = reinterpret_cast<const long long*>(step_dev);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long groups = (px->shard_end - px->shard_begin) / 8;
  long long blocks = (groups + 255) / 256;
  if (blocks > 8LL * sms) blocks = 8LL * sms;
  if (blocks < 1) blocks = 1;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static int variant = -1;          // JPDVT_PEER_VARIANT=thread: per-thread peer loads / stores instead of bulk copies (A/B knob)
  if (variant < 0) { const char* e = getenv("JPDVT_PEER_VARIANT"); variant = (e != nullptr && e[0] == 't') ? 0 : 1; }
  if (!mc && variant == 1 && (px->shard_begin % kTile) == 0 && (px->shard_end % kTile) == 0) {
    const int nrem = px->world - 1;
    int stages = kBulkStages;
    auto smem_for = [&](int s) { return static_cast<size_t>(s) * nrem * kTileBytes + 2 * kTile * 2 + 8 * kBulkStages + 128; };
    while (stages > 2 && smem_for(stages) > 200 * 1024) --stages;
    const size_t smem = smem_for(stages);
    static size_t smem_set = 0;
    if (smem > smem_set) {
      if (cudaFuncSetAttribute(peer_adamw_ema_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
        return set_error(kErrCuda, "adamw_ema_peer: cudaFuncSetAttribute(smem=%zu) failed: %s", smem, cudaGetErrorString(cudaGetLastError()));
      smem_set = smem;
    }
    int per_sm = static_cast<int>((220 * 1024) / smem);
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    const long long tiles = (px->shard_end - px->shard_begin) / kTile;
    long long grid = static_cast<long long>(sms) * per_sm;
    if (grid > tiles) grid = tiles > 0 ? tiles : 1;
    peer_adamw_ema_bulk_kernel<<<static_cast<unsigned>(grid), 256, smem, st>>>(*px, p, m, v, ema_or_null, h, stages);
    return check_launch("peer_adamw_ema_bulk_kernel");
  }
  if (mc) peer_adamw_ema_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(*px, p, m, v, ema_or_null, h);
  else peer_adamw_ema_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(*px, p, m, v, ema_or_null, h);
  return check_launch("peer_adamw_ema_kernel");
}

}  // extern "C"
