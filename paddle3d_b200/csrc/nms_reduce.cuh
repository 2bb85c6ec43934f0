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

#include "box_geom.cuh"

namespace p3d {

// s_removed: shared memory, col_blocks words.  s_misc: shared memory, >= 2 words.
// keep[] receives the kept box indices in order; returns (to every thread) the number kept.
// row_stride = words per mask row.  blockDim.x must be a multiple of 32 and >= 64.
__device__ inline int nms_greedy_cta(const unsigned long long *__restrict__ mask, int n, int row_stride,
                                     int32_t *__restrict__ keep, unsigned long long *s_removed,
                                     unsigned long long *s_misc, int limit = 0x7fffffff) {
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
    // a caller that consumes only the first `limit` kept boxes (in score order) does not need the later tiles:
    // the first `limit` entries of keep[] are final once kept_total reaches it (uniform exit: same value in all threads)
    if (kept_total >= limit) break;
  }
  return kept_total;
}

// One 64 x 64 tile of the rotated-IoU suppression matrix, 64 threads.
// The reference evaluates the full polygon intersection for every pair (iou3d_nms_kernel.cu:340-354).  Here a
// cheap exact reject runs first — two boxes whose centres are farther apart than the sum of their half
// diagonals (+ slack covering the 1e-2 inside-margin) cannot intersect, so their IoU is exactly 0 and the
// `> thr` test is false for any thr >= 0 — and the surviving pairs are compacted in shared memory and shared
// out evenly over the threads, so the expensive path runs without warp divergence.  Results are bit-identical.
// s_row / s_col: 64 x 7 floats each; s_pairs: 4096 uint16; s_bits: 64 words; s_cnt: 3 ints.
__device__ inline unsigned long long nms_rotated_tile(const float *s_row, const float *s_col, int rows, int cols,
                                                       bool diag, float thr, unsigned short *s_pairs,
                                                       unsigned long long *s_bits, int *s_cnt) {
  const int tid = threadIdx.x, lane = tid & 31;
  s_bits[tid] = 0ull;
  unsigned long long cand = 0ull;
  if (tid < rows) {
    const float *a = s_row + tid * 7;
    const float ra = 0.5f * sqrtf(a[3] * a[3] + a[4] * a[4]);
    for (int j = diag ? tid + 1 : 0; j < cols; ++j) {
      const float *b = s_col + j * 7;
      const float dx = a[0] - b[0], dy = a[1] - b[1];
      const float R = ra + 0.5f * sqrtf(b[3] * b[3] + b[4] * b[4]) + 0.1f;
      if (thr < 0.f || dx * dx + dy * dy <= R * R * 1.001f) cand |= 1ull << j;
    }
  }
  // exclusive offsets of each thread's candidate run (2 warps)
  const int mine = __popcll(cand);
  int inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) s_cnt[tid >> 5] = inc;
  __syncthreads();
  int off = inc - mine + ((tid >> 5) ? s_cnt[0] : 0);
  const int total = s_cnt[0] + s_cnt[1];
  while (cand) {
    const int j = __ffsll(static_cast<long long>(cand)) - 1;
    cand &= cand - 1;
    s_pairs[off++] = static_cast<unsigned short>((tid << 6) | j);
  }
  __syncthreads();
  for (int p = tid; p < total; p += 64) {
    const int r = s_pairs[p] >> 6, j = s_pairs[p] & 63;
    if (geom::iou_rotated(s_row + r * 7, s_col + j * 7) > thr) atomicOr(&s_bits[r], 1ull << j);
  }
  __syncthreads();
  return s_bits[tid];
}

}  // namespace p3d
