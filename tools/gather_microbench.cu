// Micro-benchmark: how fast can one SM gather random 64 B / 128 B row pieces from an L2-resident table into shared
// memory, for the lane mappings the sparse-conv producers can use?  (Round-1 finding: the gather is the largest
// per-use cost of the tcgen05 sparse conv; this isolates it from barriers / MMA.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/gather_mb tools/gather_microbench.cu && ./gpurun_out/gather_mb
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

constexpr int kRows = 128;       // rows per "use" (one UMMA M tile)
constexpr int kThreads = 128;    // producer threads

__device__ __forceinline__ void cp16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ uint32_t s32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// mode 0: LDG.128, lane = row (32 rows per instruction), 4 chunks per lane sequentially (64 B per row), ST.128
// mode 1: LDG.128, 4 lanes per row (8 rows per instruction, 64 B contiguous per row), ST.128
// mode 2: cp.async 16 B, 4 lanes per row (8 rows per instruction, 64 B per row)
// mode 3: cp.async 16 B, 8 lanes per row (4 rows per instruction, 128 B per row)  [2x the bytes per use]
// mode 4: LDG.128, 8 lanes per row (4 rows per instruction, 128 B per row), ST.128  [2x the bytes per use]
template <int MODE>
__global__ void __launch_bounds__(kThreads) gather_kernel(const float *__restrict__ table, int row_floats,
                                                          const int *__restrict__ idx, int uses, long long *cycles,
                                                          float *sink) {
  extern __shared__ __align__(128) uint8_t smem[];  // ring of 4 x 16 KB
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int *my_idx = idx + static_cast<size_t>(blockIdx.x) * uses * kRows;
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int u = 0; u < uses; ++u) {
    uint8_t *st = smem + (u & 3) * 16384;
    const int *ix = my_idx + u * kRows;
    if (MODE == 0) {
      float4 v[4];
      const int row = tid;
      const float4 *p = reinterpret_cast<const float4 *>(table + static_cast<size_t>(ix[row]) * row_floats);
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = __ldg(p + c);
#pragma unroll
      for (int c = 0; c < 4; ++c) *reinterpret_cast<float4 *>(st + c * 2048 + (row >> 3) * 128 + (row & 7) * 16) = v[c];
    } else if (MODE == 1) {
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = wid * 32 + i * 8 + (lane >> 2);
        v[i] = __ldg(reinterpret_cast<const float4 *>(table + static_cast<size_t>(ix[row]) * row_floats) + (lane & 3));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4 *>(st + (wid * 4 + i) * 512 + (lane & 3) * 128 + (lane >> 2) * 16) = v[i];
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = wid * 32 + i * 8 + (lane >> 2);
        cp16(s32(st + (wid * 4 + i) * 512 + (lane & 3) * 128 + (lane >> 2) * 16),
             reinterpret_cast<const float4 *>(table + static_cast<size_t>(ix[row]) * row_floats) + (lane & 3));
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 3;" ::: "memory");
    } else if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = wid * 32 + i * 4 + (lane >> 3);
        // SWIZZLE_128B-style placement: 128 B per row, 16-byte chunk index XOR (row & 7)
        cp16(s32(st + row * 128 + (((lane & 7) ^ (row & 7)) * 16)),
             reinterpret_cast<const float4 *>(table + static_cast<size_t>(ix[row]) * row_floats) + (lane & 7));
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 3;" ::: "memory");
    } else {
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = wid * 32 + i * 4 + (lane >> 3);
        v[i] = __ldg(reinterpret_cast<const float4 *>(table + static_cast<size_t>(ix[row]) * row_floats) + (lane & 7));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = wid * 32 + i * 4 + (lane >> 3);
        *reinterpret_cast<float4 *>(st + row * 128 + (((lane & 7) ^ (row & 7)) * 16)) = v[i];
      }
    }
  }
  if (MODE == 2 || MODE == 3) asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const long long t1 = clock64();
  acc += reinterpret_cast<float *>(smem)[tid];
  if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
}

template <int MODE>
void run(const char *name, const float *table, int row_floats, const int *idx, int uses, int ctas_per_sm, int bytes_per_row) {
  const int grid = 148 * ctas_per_sm;
  long long *cyc;
  float *sink;
  cudaMalloc(&cyc, grid * sizeof(long long));
  cudaMalloc(&sink, 4);
  auto k = gather_kernel<MODE>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int rep = 0; rep < 2; ++rep) k<<<grid, kThreads, 65536>>>(table, row_floats, idx, uses, cyc, sink);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  k<<<grid, kThreads, 65536>>>(table, row_floats, idx, uses, cyc, sink);
  cudaEventRecord(b);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (auto c : h) mean += c;
  mean /= grid;
  const double per_use = mean / uses;                        // cycles per 128-row use, per CTA
  const double bytes = static_cast<double>(kRows) * bytes_per_row;
  printf("{\"mode\": \"%s\", \"ctas_per_sm\": %d, \"cycles_per_use_per_cta\": %.0f, \"bytes_per_clk_per_sm\": %.1f, "
         "\"chip_TBps\": %.2f, \"kernel_us\": %.1f}\n",
         name, ctas_per_sm, per_use, bytes * ctas_per_sm / per_use,
         bytes * uses * grid / (ms * 1e-3) / 1e12, ms * 1e3);
  cudaFree(cyc);
  cudaFree(sink);
}

int main() {
  const int table_rows = 65536, row_floats = 64;  // 16 MB table of 256-byte rows (64 fp32 channels), L2 resident
  const int uses = 256;
  float *table;
  int *idx;
  cudaMalloc(&table, static_cast<size_t>(table_rows) * row_floats * 4);
  cudaMemset(table, 0, static_cast<size_t>(table_rows) * row_floats * 4);
  const size_t n_idx = static_cast<size_t>(148) * 4 * uses * kRows;
  std::vector<int> h(n_idx);
  srand(1);
  for (auto &x : h) x = rand() % table_rows;
  cudaMalloc(&idx, n_idx * 4);
  cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
  for (int c = 1; c <= 3; ++c) {
    run<0>("ldg_lane_per_row_64B", table, row_floats, idx, uses, c, 64);
    run<1>("ldg_4lanes_per_row_64B", table, row_floats, idx, uses, c, 64);
    run<2>("cpasync_4lanes_per_row_64B", table, row_floats, idx, uses, c, 64);
    run<3>("cpasync_8lanes_per_row_128B", table, row_floats, idx, uses, c, 128);
    run<4>("ldg_8lanes_per_row_128B", table, row_floats, idx, uses, c, 128);
  }
  return 0;
}
