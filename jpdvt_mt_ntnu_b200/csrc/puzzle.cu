// Puzzle plumbing either side of the sampling loop, batched on the device:
//
//   scramble / mask    image_model/inference_ddp.py:382-395 (= inference.py:266-278; batched in inferencetexmet.py:318-338)
//                      einops 'b c (g1 h1) (g2 w1) -> b c (g1 g2) h1 w1', pieces[:, :, indices], inverse rearrange: slot i of
//                      the output holds input piece indices[i]; masked-puzzle inference (C5) zeroes selected slots
//   reconstruct        inference_ddp.py:449-455: reconstructed[pred[i]] = scrambled piece i, i.e. cell j shows slot order[j]
//                      - the same gather with indices = order
//   scoring            inference_ddp.py:431-447: puzzle_correct = (pred == indices).all(), patch_matches = (pred == indices).sum()
//
// Pure data movement and integer compares: results are bit-exact against the reference snippets.
#include "common.cuh"

namespace jp {

// dst[b, c, Y, X] = keep[b, slot(Y, X)] ? src[b, c, piece-local offset inside piece perm[b, slot]] : 0
// VEC = 4: one float4 (4 pixels of a row) per thread, valid when the piece width is a multiple of 4.
template <int VEC>
__global__ void gather_pieces_kernel(const float* __restrict__ src, float* __restrict__ dst, const int* __restrict__ perm,
                                     const unsigned char* __restrict__ keep, long long total, int channels, int size, int grid) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int wv = size / VEC;                       // vectors per image row
  const int xv = static_cast<int>(idx % wv);
  const int y = static_cast<int>((idx / wv) % size);
  const long long bc = idx / (static_cast<long long>(wv) * size);
  const long long b = bc / channels;
  const int piece = size / grid;
  const int x = xv * VEC;
  const int slot = (y / piece) * grid + x / piece;
  const int n = grid * grid;
  float* out = dst + (bc * size + y) * static_cast<long long>(size) + x;
  if (keep != nullptr && keep[b * n + slot] == 0) {
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(out) = make_float4(0.f, 0.f, 0.f, 0.f);
    else *out = 0.f;
    return;
  }
  const int from = perm[b * n + slot];
  const int sy = (from / grid) * piece + y % piece, sx = (from % grid) * piece + x % piece;
  const float* in = src + (bc * size + sy) * static_cast<long long>(size) + sx;
  if constexpr (VEC == 4) *reinterpret_cast<float4*>(out) = __ldg(reinterpret_cast<const float4*>(in));
  else *out = __ldg(in);
}

int launch_gather_pieces(const float* src, float* dst, const int* perm, const unsigned char* keep, int batch, int channels,
                         int size, int grid, cudaStream_t stream) {
  if (batch <= 0) return kOk;
  if (grid <= 0 || size <= 0 || size % grid != 0) return set_error(kErrBadArg, "gather_pieces: image size %d is not a multiple of the grid %d", size, grid);
  if (channels <= 0) return set_error(kErrBadArg, "gather_pieces: channels must be positive");
  if (src == dst) return set_error(kErrBadArg, "gather_pieces: in-place permutation is not supported");
  const int piece = size / grid;
  const bool vec = (piece % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const long long total = static_cast<long long>(batch) * channels * size * (vec ? size / 4 : size);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  if (vec) gather_pieces_kernel<4><<<blocks, 256, 0, stream>>>(src, dst, perm, keep, total, channels, size, grid);
  else gather_pieces_kernel<1><<<blocks, 256, 0, stream>>>(src, dst, perm, keep, total, channels, size, grid);
  return check_launch("gather_pieces_kernel");
}

// Crop-gap erosion of the training loader (image_model/train_JPDVT.py:345-349): every piece of a G x G puzzle is centre-cropped
// from in_piece to out_piece pixels and the crops are re-tiled, so neighbouring pieces no longer share a boundary:
// dst[b, c, gy*out + y, gx*out + x] = src[b, c, gy*in + off + y, gx*in + off + x]   (off = torchvision CenterCrop's offset)
__global__ void crop_pieces_kernel(const float* __restrict__ src, float* __restrict__ dst, long long total, int grid, int in_piece,
                                   int out_piece, int off) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int so = grid * out_piece, si = grid * in_piece;
  const int X = static_cast<int>(idx % so);
  const int Y = static_cast<int>((idx / so) % so);
  const long long bc = idx / (static_cast<long long>(so) * so);
  const int gy = Y / out_piece, y = Y - gy * out_piece, gx = X / out_piece, x = X - gx * out_piece;
  dst[idx] = __ldg(src + (bc * si + gy * in_piece + off + y) * static_cast<long long>(si) + gx * in_piece + off + x);
}

int launch_crop_pieces(const float* src, float* dst, int batch, int channels, int grid, int in_piece, int out_piece, int off,
                       cudaStream_t stream) {
  if (batch <= 0) return kOk;
  if (grid <= 0 || in_piece <= 0 || out_piece <= 0 || out_piece > in_piece || off < 0 || off + out_piece > in_piece)
    return set_error(kErrBadArg, "crop_pieces: cannot crop %d-pixel pieces to %d at offset %d", in_piece, out_piece, off);
  const long long total = static_cast<long long>(batch) * channels * grid * out_piece * grid * out_piece;
  crop_pieces_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(src, dst, total, grid, in_piece, out_piece, off);
  return check_launch("crop_pieces_kernel");
}

// One warp per puzzle: matches[b] = #{i : pred[b,i] == truth[b,i]}, correct[b] = (matches[b] == n); totals[0..2] +=
// (puzzles correct, pieces correct, puzzles) - the three counters the reference all-reduces (inference_ddp.py:485-490).
__global__ void score_placements_kernel(const int* __restrict__ pred, const int* __restrict__ truth, int batch, int n,
                                        int* __restrict__ correct, int* __restrict__ matches, long long* __restrict__ totals) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= batch) return;
  int m = 0;
  for (int i = lane; i < n; i += 32) m += pred[b * n + i] == truth[b * n + i] ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m += __shfl_xor_sync(0xffffffffu, m, o);
  if (lane == 0) {
    matches[b] = m;
    correct[b] = (m == n) ? 1 : 0;
    if (totals != nullptr) {
      atomicAdd(reinterpret_cast<unsigned long long*>(totals), static_cast<unsigned long long>(m == n ? 1 : 0));
      atomicAdd(reinterpret_cast<unsigned long long*>(totals) + 1, static_cast<unsigned long long>(m));
      atomicAdd(reinterpret_cast<unsigned long long*>(totals) + 2, 1ull);
    }
  }
}

int launch_score_placements(const int* pred, const int* truth, int batch, int n, int* correct, int* matches, long long* totals,
                            cudaStream_t stream) {
  if (batch <= 0) return kOk;
  if (n <= 0) return set_error(kErrBadArg, "score_placements: n must be positive");
  const int warps = 4;
  score_placements_kernel<<<(batch + warps - 1) / warps, warps * 32, 0, stream>>>(pred, truth, batch, n, correct, matches, totals);
  return check_launch("score_placements_kernel");
}

}  // namespace jp
