// Dense BEV canvas writer for sm_100a: PointPillarsScatter and sparse to_dense(+transpose), gather-formulated.
//
// The reference does zeros + paddle.scatter + transpose (pillar_scatter.py:71-92) or
// to_dense + transpose + reshape (sparse_resnet.py:202-206): three passes over the dense tensor.
// Here the dense output [batch, C, D, ny, nx] is written exactly once:
//   S1 memset      cell -> row map <- -1                                  (4 B per cell)
//   S2 scat_map    map[cell(row)] = max(row)   (later rows win on duplicates, as scatter-overwrite)
//   S3 scat_write  one CTA per tile of 128 consecutive cells: occupied rows are staged channel-major
//                  in shared memory (coalesced 128 B row reads), then every channel row of the tile is
//                  written with one float4 per lane (512 B per warp store), zeros where the map is empty.
// Algorithmic bytes: 4*n*C (features) + 16*n (coords) + 4*batch*C*D*ny*nx (canvas).
#include "common.cuh"

namespace p3d {
namespace {

constexpr int kTile = 128;         // cells per CTA
constexpr int kChunk = 64;         // channels staged per pass
constexpr int kStride = kTile + 4; // smem row stride (floats): keeps float4 alignment, spreads banks

__global__ void __launch_bounds__(256) scat_map_kernel(const int32_t *__restrict__ coords,
                                                       const int32_t *__restrict__ n_dev, int n_cap, int batch, int D,
                                                       int ny, int nx, int use_z, int32_t *__restrict__ map) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = n_dev ? min(n_dev[0], n_cap) : n_cap;
  if (i >= n) return;
  const int4 c = *reinterpret_cast<const int4 *>(coords + static_cast<size_t>(i) * 4);
  const int z = use_z ? c.y : 0;
  if (c.x < 0 || c.x >= batch || z < 0 || z >= D || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) return;
  atomicMax(&map[((static_cast<size_t>(c.x) * D + z) * ny + c.z) * nx + c.w], i);
}

template <bool kVec>
__global__ void __launch_bounds__(256) scat_write_kernel(const float *__restrict__ feats,
                                                         const int32_t *__restrict__ map, int C, long long S,
                                                         float *__restrict__ out) {
  __shared__ int s_row[kTile];
  __shared__ __align__(16) float s_val[kChunk * kStride];
  const int b = blockIdx.y;
  const long long cell0 = static_cast<long long>(blockIdx.x) * kTile;
  const int ncell = static_cast<int>(min(static_cast<long long>(kTile), S - cell0));
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  int any = 0;
  if (tid < kTile) {
    const int r = tid < ncell ? map[static_cast<size_t>(b) * S + cell0 + tid] : -1;
    s_row[tid] = r;
    any = r >= 0;
  }
  const int occupied = __syncthreads_or(any);
  for (int c0 = 0; c0 < C; c0 += kChunk) {
    const int cc = min(kChunk, C - c0);
    if (occupied) {
      // stage: warp w takes cells w, w+8, ...; lanes sweep the channels of that row (coalesced)
      for (int cell = wid; cell < ncell; cell += 8) {
        const int r = s_row[cell];
        if (r < 0) continue;
        const float *src = feats + static_cast<size_t>(r) * C + c0;
        for (int c = lane; c < cc; c += 32) s_val[c * kStride + cell] = __ldg(src + c);
      }
      __syncthreads();
    }
    // emit: one warp per channel row of the tile
    for (int c = wid; c < cc; c += 8) {
      float *dst = out + (static_cast<size_t>(b) * C + c0 + c) * S + cell0;
      if (kVec) {
        const int k = lane * 4;
        if (k < ncell) {  // S % 4 == 0 => ncell % 4 == 0
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (occupied) {
            const float4 sv = *reinterpret_cast<const float4 *>(&s_val[c * kStride + k]);
            v.x = s_row[k] >= 0 ? sv.x : 0.f;
            v.y = s_row[k + 1] >= 0 ? sv.y : 0.f;
            v.z = s_row[k + 2] >= 0 ? sv.z : 0.f;
            v.w = s_row[k + 3] >= 0 ? sv.w : 0.f;
          }
          __stcs(reinterpret_cast<float4 *>(dst + k), v);
        }
      } else {
        for (int k = lane; k < ncell; k += 32) dst[k] = (occupied && s_row[k] >= 0) ? s_val[c * kStride + k] : 0.f;
      }
    }
    if (occupied) __syncthreads();
  }
}

// Sparse rows (fp16-pair H16 rows of C channels, C % 32 == 0) -> pixel H16 image [batch, ny, nx][D * C channels]: the
// fp16-pair form of SparseResNet3D's to_dense + transpose + reshape (sparse_resnet.py:202-206, channel = z * C + c) that
// the dense RPN consumes directly (no fp32 NCHW tensor, no layout conversion pass).  One warp copies one row (4 * C bytes).
__global__ void __launch_bounds__(256) rows_to_pixel_h16_kernel(const uint4 *__restrict__ rows, const int32_t *__restrict__ coords,
                                                                const int32_t *__restrict__ n_dev, int n_cap, int C, int batch,
                                                                int D, int ny, int nx, uint4 *__restrict__ out) {
  const int n = n_dev ? min(n_dev[0], n_cap) : n_cap;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n) return;
  const int4 c = *reinterpret_cast<const int4 *>(coords + static_cast<size_t>(r) * 4);
  if (c.x < 0 || c.x >= batch || c.y < 0 || c.y >= D || c.z < 0 || c.z >= ny || c.w < 0 || c.w >= nx) return;
  const int q_row = C / 4;  // uint4 per row (4 * C bytes)
  uint4 *dst = out + ((static_cast<size_t>(c.x) * ny + c.z) * nx + c.w) * (static_cast<size_t>(D) * q_row) + static_cast<size_t>(c.y) * q_row;
  const uint4 *src = rows + static_cast<size_t>(r) * q_row;
  for (int q = lane; q < q_row; q += 32) dst[q] = __ldg(src + q);
}

}  // namespace
}  // namespace p3d

using namespace p3d;

extern "C" int p3d_sparse_rows_to_pixel_h16(const void *rows_h16, const int32_t *coords, const int32_t *n_dev, int n_cap, int C,
                                            int batch, int D, int ny, int nx, void *out_pixel_h16, p3d_stream_t stream) {
  if (n_cap < 0 || C < 32 || C % 32 || batch < 1 || D < 1 || ny < 1 || nx < 1 || !out_pixel_h16 || (n_cap && (!rows_h16 || !coords)))
    return P3D_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(rows_h16) & 15) || (reinterpret_cast<uintptr_t>(coords) & 15) ||
      (reinterpret_cast<uintptr_t>(out_pixel_h16) & 15))
    return P3D_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bytes = static_cast<size_t>(batch) * ny * nx * D * C * 4;
  P3D_CUDA_CHECK(cudaMemsetAsync(out_pixel_h16, 0, bytes, st));  // (hi, lo') = (0, 0) is the value 0
  if (n_cap > 0) {
    rows_to_pixel_h16_kernel<<<div_up(static_cast<long long>(n_cap) * 32, 256), 256, 0, st>>>(
        static_cast<const uint4 *>(rows_h16), coords, n_dev, n_cap, C, batch, D, ny, nx, static_cast<uint4 *>(out_pixel_h16));
    P3D_LAUNCH_CHECK();
  }
  return P3D_OK;
}

extern "C" size_t p3d_scatter_dense_workspace_bytes(int batch, int D, int ny, int nx) {
  if (batch < 1 || D < 1 || ny < 1 || nx < 1) return 0;
  return align_up(static_cast<size_t>(batch) * D * ny * nx * sizeof(int32_t));
}

extern "C" int p3d_scatter_dense(const float *feats, const int32_t *coords, const int32_t *n_dev, int n_cap, int C,
                                 int batch, int D, int ny, int nx, int use_z, float *out, void *workspace,
                                 size_t workspace_bytes, p3d_stream_t stream) {
  if (n_cap < 0 || C < 1 || batch < 1 || D < 1 || ny < 1 || nx < 1 || !out || !workspace ||
      (n_cap && (!feats || !coords)))
    return P3D_ERR_INVALID_ARG;
  if (batch > 65535) return P3D_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(coords) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return P3D_ERR_INVALID_ARG;
  const size_t need = p3d_scatter_dense_workspace_bytes(batch, D, ny, nx);
  if (workspace_bytes < need) return P3D_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t *map = static_cast<int32_t *>(workspace);
  const long long S = static_cast<long long>(D) * ny * nx;
  P3D_CUDA_CHECK(cudaMemsetAsync(map, 0xff, static_cast<size_t>(batch) * S * sizeof(int32_t), st));
  if (n_cap > 0) {
    scat_map_kernel<<<div_up(n_cap, 256), 256, 0, st>>>(coords, n_dev, n_cap, batch, D, ny, nx, use_z, map);
    P3D_LAUNCH_CHECK();
  }
  dim3 grid(div_up(S, kTile), batch);
  if (S % 4 == 0)
    scat_write_kernel<true><<<grid, 256, 0, st>>>(feats, map, C, S, out);
  else
    scat_write_kernel<false><<<grid, 256, 0, st>>>(feats, map, C, S, out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}
