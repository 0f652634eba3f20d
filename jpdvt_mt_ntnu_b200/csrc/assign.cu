// Position-to-grid assignment: one warp per puzzle, integer output, bit-exact against the reference's
// numpy/sklearn snippet on identical score matrices.
//
//   image_model/inference.py:294-301  per-slot mean of the predicted 8-d embedding
//   image_model/inference.py:304      sklearn pairwise_distances(metric='manhattan') -> fp64 L1 scores (sequential over d)
//   image_model/inference.py:113-125  find_permutation: column j takes argmin over rows; a taken row is overwritten with
//                                     the sentinel (1e9; 2024 in sample.py:103 / train_JPDVT.py:556) in the REMAINING
//                                     columns (it is not removed), ties -> lowest row, NaN wins (numpy argmin)
//   image_model/inference.py:306      pred = argsort(order)
#include "common.cuh"

namespace jp {

constexpr int kMaxSlots = 32;       // G*G <= 32 (3x3, 4x4, 5x5)
constexpr int kAssignWarps = 4;

// Greedy column scan.  Lane i (< n) owns row i.  `col(i, j)` yields the live score of row i in column j.
template <typename ColFn>
__device__ __forceinline__ void greedy_and_rank(int n, double sentinel, ColFn col, int* order_out, int* pred_out) {
  const int lane = threadIdx.x & 31;
  bool taken = false;
  int my_order = 0;   // lane j keeps order[j]
  for (int j = 0; j < n; ++j) {
    double v = INFINITY;
    if (lane < n) v = taken ? sentinel : col(lane, j);
    const unsigned nan_mask = __ballot_sync(0xffffffffu, lane < n && v != v);
    int best;
    if (nan_mask != 0u) {
      best = __ffs(nan_mask) - 1;                 // numpy argmin returns the first NaN
    } else {
      int idx = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
      }
      best = idx;                                  // first minimum; lanes >= n hold +inf with larger indices
    }
    if (lane == best) taken = true;
    if (lane == j) my_order = best;
  }
  // pred = argsort(order) (stable; for a permutation pred[order[j]] = j)
  int rank = 0;
  for (int i = 0; i < n; ++i) {
    const int oi = __shfl_sync(0xffffffffu, my_order, i);
    if (lane < n && (oi < my_order || (oi == my_order && i < lane))) ++rank;
  }
  if (lane < n) {
    order_out[lane] = my_order;
    pred_out[rank] = lane;
  }
}

__global__ void __launch_bounds__(kAssignWarps * 32)
assign_scores_kernel(const double* __restrict__ scores, int batch, int n, double sentinel, int* __restrict__ order,
                     int* __restrict__ pred) {
  const int puzzle = blockIdx.x * kAssignWarps + (threadIdx.x >> 5);
  if (puzzle >= batch) return;
  const double* sc = scores + static_cast<long long>(puzzle) * n * n;
  greedy_and_rank(n, sentinel, [&](int i, int j) { return sc[i * n + j]; }, order + static_cast<long long>(puzzle) * n,
                  pred + static_cast<long long>(puzzle) * n);
}

// latents [B, T, 8] fp32, token order (p1 h1 p2 w1); canon [n, 8] fp32 (get_2d_sincos_pos_embed(8, G) cast to fp32).
__global__ void __launch_bounds__(kAssignWarps * 32)
assign_latents_kernel(const float* __restrict__ latents, const float* __restrict__ canon, int batch, int grid, int tok,
                      double sentinel, int* __restrict__ order, int* __restrict__ pred, double* __restrict__ scores_out) {
  __shared__ double s_scores[kAssignWarps][kMaxSlots * kMaxSlots];
  __shared__ float s_feat[kAssignWarps][kMaxSlots * kLatent];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int puzzle = blockIdx.x * kAssignWarps + w;
  if (puzzle >= batch) return;
  const int n = grid * grid;
  const int T = n * tok * tok;
  const float* lat = latents + static_cast<long long>(puzzle) * T * kLatent;
  // per-slot mean: 8 lanes-worth of work per slot -> (slot, d) pairs strided over the warp, sequential fp32 sum
  for (int e = lane; e < n * kLatent; e += 32) {
    const int slot = e / kLatent, d = e % kLatent;
    const int p1 = slot / grid, p2 = slot % grid;
    float acc = 0.f;
    for (int h1 = 0; h1 < tok; ++h1)
      for (int w1 = 0; w1 < tok; ++w1) {
        const int token = ((p1 * tok + h1) * grid + p2) * tok + w1;
        acc += lat[token * kLatent + d];
      }
    s_feat[w][e] = acc / static_cast<float>(tok * tok);
  }
  __syncwarp();
  for (int e = lane; e < n * n; e += 32) {
    const int i = e / n, j = e % n;
    double acc = 0.0;
#pragma unroll
    for (int d = 0; d < kLatent; ++d)
      acc += fabs(static_cast<double>(s_feat[w][i * kLatent + d]) - static_cast<double>(canon[j * kLatent + d]));
    s_scores[w][e] = acc;
    if (scores_out != nullptr) scores_out[static_cast<long long>(puzzle) * n * n + e] = acc;
  }
  __syncwarp();
  const double* sc = s_scores[w];
  greedy_and_rank(n, sentinel, [&](int i, int j) { return sc[i * n + j]; }, order + static_cast<long long>(puzzle) * n,
                  pred + static_cast<long long>(puzzle) * n);
}

int launch_assign_scores(const double* scores, int batch, int n, double sentinel, int* order, int* pred,
                         cudaStream_t stream) {
  if (batch <= 0) return kOk;
  if (n <= 0 || n > kMaxSlots) return set_error(kErrUnsupported, "assign: %d slots per puzzle (supported: 1..%d)", n, kMaxSlots);
  assign_scores_kernel<<<(batch + kAssignWarps - 1) / kAssignWarps, kAssignWarps * 32, 0, stream>>>(scores, batch, n, sentinel,
                                                                                                     order, pred);
  return check_launch("assign_scores_kernel");
}

int launch_assign_latents(const float* latents, const float* canon, int batch, int grid, int tok, double sentinel,
                          int* order, int* pred, double* scores_out, cudaStream_t stream) {
  if (batch <= 0) return kOk;
  const int n = grid * grid;
  if (grid <= 0 || tok <= 0 || n > kMaxSlots)
    return set_error(kErrUnsupported, "assign: grid %d / tokens-per-side %d not supported (G*G <= %d)", grid, tok, kMaxSlots);
  assign_latents_kernel<<<(batch + kAssignWarps - 1) / kAssignWarps, kAssignWarps * 32, 0, stream>>>(
      latents, canon, batch, grid, tok, sentinel, order, pred, scores_out);
  return check_launch("assign_latents_kernel");
}

}  // namespace jp
