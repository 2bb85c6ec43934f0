// On-device greedy reduction of an NMS suppression bit-matrix (64 columns per word).
// Replaces the host loop of paddle3d/ops/iou3d_nms/iou3d_nms.cpp:115-135 (and its copies at
// :177-197 and centerpoint_postprocess/postprocess.cu:234-245), which costs the reference a
// blocking D2H copy of the matrix per call.  One CTA walks the boxes 64 at a time:
//   - warp 0 resolves the 64 boxes of the diagonal tile with the tile's words held in registers
//     (2 per lane, read through shuffles): strictly sequential, but register-resident;
//   - the whole CTA then ORs the rows of the boxes that survived into the running `removed`
//     bitset (shared memory) for all later column words.
// Only words with column-block >= row-block are ever read, so the mask kernel may skip the
// lower triangle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace p3d {

// s_removed: shared memory, col_blocks words.  s_misc: shared memory, >= 2 words.
// keep[] receives the kept box indices in order; returns (to every thread) the number kept.
// row_stride = words per mask row.  blockDim.x must be a multiple of 32 and >= 64.
__device__ inline int nms_greedy_cta(const unsigned long long *__restrict__ mask, int n, int row_stride,
                                     int32_t *__restrict__ keep, unsigned long long *s_removed,
                                     unsigned long long *s_misc) {
  const int col_blocks = (n + 63) / 64;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int j = tid; j < col_blocks; j += blockDim.x) s_removed[j] = 0ull;
  int kept_total = 0;
  __syncthreads();
  for (int nb = 0; nb < col_blocks; ++nb) {
    const int base = nb * 64;
    const int valid = min(64, n - base);
    if (wid == 0) {
      unsigned long long d0 = 0ull, d1 = 0ull;
      if (lane < valid) d0 = mask[static_cast<size_t>(base + lane) * row_stride + nb];
      if (lane + 32 < valid) d1 = mask[static_cast<size_t>(base + lane + 32) * row_stride + nb];
      unsigned long long dead = s_removed[nb];
      unsigned long long kept = 0ull;
      for (int t = 0; t < valid; ++t) {
        const unsigned long long row = __shfl_sync(0xffffffffu, t < 32 ? d0 : d1, t & 31);
        if (!((dead >> t) & 1ull)) {
          kept |= 1ull << t;
          dead |= row;
        }
      }
      if (lane == 0) s_misc[0] = kept;
    }
    __syncthreads();
    const unsigned long long kept = s_misc[0];
    if (tid < 64 && ((kept >> tid) & 1ull))
      keep[kept_total + __popcll(kept & ((1ull << tid) - 1ull))] = base + tid;
    const int rest = col_blocks - nb - 1;
    for (int p = tid; p < rest * 64; p += blockDim.x) {
      const int t = p / rest, j = nb + 1 + (p - t * rest);
      if ((kept >> t) & 1ull) {
        const unsigned long long w = mask[static_cast<size_t>(base + t) * row_stride + j];
        if (w) atomicOr(&s_removed[j], w);
      }
    }
    kept_total += __popcll(kept);
    __syncthreads();
  }
  return kept_total;
}

}  // namespace p3d
