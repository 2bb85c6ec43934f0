// Dense 2-D convolution on tcgen05 for the RPN / neck / CenterHead row (reference:
// backbones/second_backbone.py:72-120, necks/second_fpn.py:99-160, detection/centerpoint/center_head.py:43-220).
// Parity-green on a B200 (tests/test_gpu_dense.py) at the end of round 1; performance not measured yet, not on the
// default bench path.
//
// The image is kept as "pixel split rows" [B*H*W][2][C] (NHWC with the tf32 hi half of all channels, then the lo half —
// the row format of the sparse-conv layers with row = pixel), so a 3x3 tap of 32 input channels for a 16 x 8 pixel
// tile is one 4-D TMA box {32 ch, 16 x, 8 y, 1 b} at the shifted coordinate: zero padding is the tensor map's
// out-of-bounds fill, a stride-2 conv is the map's element stride, and the box lands in shared memory as the
// SWIZZLE_128B K-major operand tile (128 rows x 128 B) that the sparse kernel builds with 64 cp.async per thread.
// Everything after the load is the sparse kernel's pipeline: 3xTF32 with acc[0,2N) += A_hi x [B_hi | B_lo],
// acc[2N,3N) += A_lo x B_hi in TMEM, weights of the use by cp.async.bulk, BN/bias/ReLU epilogue writing split rows
// (at a column offset: channel concat for free) or fp32 NCHW planes (the head outputs centerpoint_postprocess reads).
//
//   work item   (batch, 16 x 8 output tile, N tile of the output channels[, tap of a k = s transposed conv])
//   warps 0-3   epilogue        warp 4   MMA issue        warp 5   TMA (rows + weights), one lane
#include <cuda.h>

#include "p3d_b200.h"
#include "tc_common.cuh"

namespace p3d {
namespace dc {

using namespace tc;

constexpr int kTW = 16, kTH = 8;  // output tile: 16 x 8 pixels = the 128 rows of one UMMA
static_assert(kTW * kTH == kM, "tile must have 128 pixels");

struct Params {
  int B, H, W, Cin;          // input image (pixel split rows [B*H*W][2][Cin])
  int taps, kw, stride, pad;  // conv geometry (taps = kh * kw); transposed conv: taps = up * up, stride = pad = 0 here
  int up;                    // 1 = convolution; > 1 = transposed conv with kernel = stride = up
  int oH, oW;                // extent of the tiled grid (conv: output image; transposed: input image)
  int tiles_x, tiles_y, n_ntiles;
  int cout;                  // valid output channels (N tiles are zero-padded above it)
  int out_H, out_W;          // output image
  int out_C, out_c0;         // split-row output: channels per row and first column written by this layer
  int relu;
};

template <int N>
struct Cfg {
  static constexpr int KC = 32;
  static constexpr int A_TILE = KC * kM * 4;
  static constexpr int A_STAGE = 2 * A_TILE;
  static constexpr int B_STAGE = 2 * KC * N * 4;
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int MIN_CTAS = (N <= 64) ? 2 : 1;
  static constexpr int BUDGET = (N <= 64) ? 100 * 1024 : 192 * 1024;
  static constexpr int S_RAW = BUDGET / STAGE;
  static constexpr int STAGES = S_RAW > 8 ? 8 : (S_RAW < 2 ? 2 : S_RAW);
  static constexpr int TMEM_COLS = (3 * N <= 64) ? 64 : (3 * N <= 128) ? 128 : (3 * N <= 256) ? 256 : 512;
  static constexpr uint32_t IDESC2 = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>((2 * N) >> 3) << 17) |
                                     (static_cast<uint32_t>(kM >> 4) << 24);
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
                                    (static_cast<uint32_t>(kM >> 4) << 24);
  static_assert(N % 16 == 0 && N >= 16 && N <= 128, "N tile: 16 .. 128 in steps of 16");
  static_assert(MIN_CTAS * TMEM_COLS <= 512, "TMEM over-subscribed");
  static_assert(STAGE % 1024 == 0, "SWIZZLE_128B tiles need 1024-byte alignment");
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3ffffu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ void mma_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_tile4d(uint32_t dst, const CUtensorMap *map, int c, int x, int y, int b, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c), "r"(x), "r"(y), "r"(b), "r"(bar)
      : "memory");
}

template <int N>
__global__ void __launch_bounds__(kThreads, Cfg<N>::MIN_CTAS)
    dense_conv_kernel(const __grid_constant__ CUtensorMap in_map, const Params p, const float *__restrict__ packed_w,
                      const float *__restrict__ scale, const float *__restrict__ shift, float *__restrict__ out_split,
                      float *__restrict__ out_nchw) {
  using C = Cfg<N>;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int up2 = p.up * p.up;
  const long long n_work = static_cast<long long>(p.B) * p.tiles_y * p.tiles_x * p.n_ntiles * (p.up > 1 ? up2 : 1);
  if (static_cast<long long>(blockIdx.x) >= n_work) return;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) unsigned long long s_bar[8 + 8 + 1];  // full[8] empty[8] tmem_full
  constexpr int kF = 0, kE = 8, kTF = 16;
  __shared__ uint32_t s_tmem_base;

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  if (tid == 128) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(smem_u32(&s_bar[kF + s]), 1);  // the TMA lane's expect_tx arrival (rows + weights)
      mbar_init(smem_u32(&s_bar[kE + s]), 1);  // tcgen05.commit
    }
    mbar_init(smem_u32(&s_bar[kTF]), 1);
    fence_mbar_init();
  }
  if (wid == 4) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)),
                 "r"(static_cast<uint32_t>(C::TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  const uint32_t ring = smem_u32(smem);
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the input image is the previous layer's output

  const int G = p.Cin / C::KC;
  const int n_uses = (p.up > 1 ? 1 : p.taps) * G;
  uint32_t ph = (wid == 4) ? 0u : 0xffffffffu;  // expected parity per stage (bit s), as in the sparse kernel
  int item_it = 0;
  for (long long w = blockIdx.x; w < n_work; w += gridDim.x, ++item_it) {
    // decode: N tile fastest (neighbouring CTAs share the image tile in L2), then tap (transposed conv), x, y, batch
    long long q = w;
    const int nt = static_cast<int>(q % p.n_ntiles);
    q /= p.n_ntiles;
    int tap0 = 0;
    if (p.up > 1) {
      tap0 = static_cast<int>(q % up2);
      q /= up2;
    }
    const int tx0 = static_cast<int>(q % p.tiles_x) * kTW;
    q /= p.tiles_x;
    const int ty0 = static_cast<int>(q % p.tiles_y) * kTH;
    const int b = static_cast<int>(q / p.tiles_y);
    const float *w_tile = packed_w + static_cast<size_t>(nt) * p.taps * p.Cin * (2 * N);

    if (wid < 4) {
      // ---------------------------------------------------------------- epilogue
      mbar_wait(smem_u32(&s_bar[kTF]), static_cast<uint32_t>(item_it & 1));
      tc_fence_after();
      const int m = tid;  // tile pixel of this thread = TMEM lane
      const int iy = ty0 + m / kTW, ix = tx0 + m % kTW;
      const bool live = iy < p.oH && ix < p.oW;
      const int Y = iy * p.up + (p.up > 1 ? tap0 / p.up : 0), X = ix * p.up + (p.up > 1 ? tap0 % p.up : 0);
      const size_t orow = (static_cast<size_t>(b) * p.out_H + Y) * p.out_W + X;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wid * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t a[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
            "%15}, [%16];"
            : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]),
              "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
            : "r"(taddr + static_cast<uint32_t>(c0)));
#pragma unroll
        for (int acc = 1; acc < 3; ++acc) {
          uint32_t t[16];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
              "%15}, [%16];"
              : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
                "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
              : "r"(taddr + static_cast<uint32_t>(acc * N + c0)));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) a[j] = __float_as_uint(__uint_as_float(a[j]) + __uint_as_float(t[j]));
        }
        const int ch0 = nt * N + c0;  // first output channel of this chunk
        if (live && ch0 < p.cout) {
          float o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v = __uint_as_float(a[j]);
            if (ch0 + j < p.cout) {
              if (scale) v = v * __ldg(scale + ch0 + j);
              if (shift) v = v + __ldg(shift + ch0 + j);
            }
            if (p.relu) v = fmaxf(v, 0.f);
            o[j] = v;
          }
          if (out_split) {  // channel counts of split-row layers are multiples of 16: whole chunks
            float h[16], l[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split_tf32(o[j], h[j], l[j]);
            float *oh = out_split + orow * (2 * static_cast<size_t>(p.out_C)) + p.out_c0 + ch0;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              *reinterpret_cast<float4 *>(oh + j) = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
              *reinterpret_cast<float4 *>(oh + p.out_C + j) = make_float4(l[j], l[j + 1], l[j + 2], l[j + 3]);
            }
          }
          if (out_nchw) {  // fp32 planes [B, cout, out_H, out_W]: consecutive threads = consecutive x
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (ch0 + j < p.cout)
                out_nchw[((static_cast<size_t>(b) * p.cout + ch0 + j) * p.out_H + Y) * p.out_W + X] = o[j];
          }
        }
      }
      tc_fence_before();
    } else if (wid == 4) {
      // ---------------------------------------------------------------- MMA issuer (whole warp runs the stream)
      int s = 0;
      for (int u = 0; u < n_uses; ++u) {
        mbar_wait(smem_u32(&s_bar[kF + s]), (ph >> s) & 1u);
        ph ^= 1u << s;
        tc_fence_after();
        const uint32_t a_hi = ring + static_cast<uint32_t>(s * C::STAGE), a_lo = a_hi + C::A_TILE;
        const uint32_t b_all = a_hi + C::A_STAGE;
#pragma unroll
        for (int j = 0; j < C::KC / 8; ++j) {
          const uint32_t bo = static_cast<uint32_t>(2 * j) * (2 * N * 16);
          const uint64_t db = smem_desc(b_all + bo, 2 * N * 16, 128);  // rows 0..N-1 = hi, N..2N-1 = lo
          const uint32_t first = (u | j) ? 1u : 0u;
          mma_elect(tmem_base, desc_sw128(a_hi + j * 32), db, C::IDESC2, first);          // A_hi x [B_hi | B_lo]
          mma_elect(tmem_base + 2 * N, desc_sw128(a_lo + j * 32), db, C::IDESC, first);   // A_lo x B_hi
        }
        commit_elect(smem_u32(&s_bar[kE + s]));
        if (u == n_uses - 1) commit_elect(smem_u32(&s_bar[kTF]));
        s = (s + 1 == C::STAGES) ? 0 : s + 1;
      }
      tc_fence_before();
    } else if (lane == 0) {
      // ---------------------------------------------------------------- TMA: image tile (hi, lo) + weights per use
      int s = 0;
      const int t_begin = p.up > 1 ? tap0 : 0, t_end = p.up > 1 ? tap0 + 1 : p.taps;
      for (int t = t_begin; t < t_end; ++t) {
        const int dy = p.up > 1 ? 0 : t / p.kw, dx = p.up > 1 ? 0 : t % p.kw;
        const int x = tx0 * p.stride - p.pad + dx, y = ty0 * p.stride - p.pad + dy;  // may be negative: zero fill
        for (int g = 0; g < G; ++g) {
          mbar_wait(smem_u32(&s_bar[kE + s]), (ph >> s) & 1u);
          ph ^= 1u << s;
          const uint32_t st = ring + static_cast<uint32_t>(s * C::STAGE), bar = smem_u32(&s_bar[kF + s]);
          mbar_arrive_expect_tx(bar, static_cast<uint32_t>(C::STAGE));
          tma_tile4d(st, &in_map, g * C::KC, x, y, b, bar);                       // hi half of the 32 channels
          tma_tile4d(st + C::A_TILE, &in_map, p.Cin + g * C::KC, x, y, b, bar);   // lo half
          bulk_g2s(st + C::A_STAGE, w_tile + (static_cast<size_t>(t) * p.Cin + g * C::KC) * (2 * N),
                   static_cast<uint32_t>(C::B_STAGE), bar);
          s = (s + 1 == C::STAGES) ? 0 : s + 1;
        }
      }
    }
    __syncthreads();  // item drained (epilogue has read TMEM) before the next item's first MMA overwrites it
  }
  tc_fence_before();
  __syncthreads();
  if (wid == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(C::TMEM_COLS))
                 : "memory");
  }
}

// fp32 NCHW image -> pixel split rows [B*H*W][2][C] (32 x 32 tile transpose through shared memory)
__global__ void __launch_bounds__(256) nchw_to_pixel_split_kernel(const float *__restrict__ in, int C, long long HW,
                                                                  float *__restrict__ out) {
  __shared__ float s_t[32][33];
  const int b = blockIdx.z;
  const long long p0 = static_cast<long long>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k;
    const long long px = p0 + tx;
    s_t[k][tx] = (c < C && px < HW) ? in[(static_cast<size_t>(b) * C + c) * HW + px] : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const long long px = p0 + k;
    const int c = c0 + tx;
    if (c < C && px < HW) {
      float h, l;
      split_tf32(s_t[tx][k], h, l);
      float *row = out + (static_cast<size_t>(b) * HW + px) * (2 * static_cast<size_t>(C));
      row[c] = h;
      row[C + c] = l;
    }
  }
}

inline int make_image_map(const float *img, int B, int H, int W, int Cin, int stride, CUtensorMap *map) {
  using Encode = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Encode encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn)
      return P3D_ERR_UNSUPPORTED;
    encode = reinterpret_cast<Encode>(fn);
  }
  const cuuint64_t row = static_cast<cuuint64_t>(2 * Cin) * sizeof(float);
  const cuuint64_t gdim[4] = {static_cast<cuuint64_t>(2 * Cin), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                              static_cast<cuuint64_t>(B)};
  const cuuint64_t gstride[3] = {row, row * W, row * W * H};
  // with an element stride s the box extent is s x the number of elements loaded (cuda.h, cuTensorMapEncodeTiled)
  const cuuint32_t box[4] = {32u, static_cast<cuuint32_t>(kTW * stride), static_cast<cuuint32_t>(kTH * stride), 1u};
  const cuuint32_t estride[4] = {1u, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1u};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(img), gdim, gstride, box, estride,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? P3D_OK : P3D_ERR_INVALID_ARG;
}

template <int N>
int launch(const CUtensorMap &map, const Params &p, const float *packed, const float *scale, const float *shift,
           float *out_split, float *out_nchw, cudaStream_t st) {
  using C = Cfg<N>;
  const size_t smem = static_cast<size_t>(C::STAGES) * C::STAGE + 1024;
  auto kern = dense_conv_kernel<N>;
  P3D_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const long long work = static_cast<long long>(p.B) * p.tiles_y * p.tiles_x * p.n_ntiles * (p.up > 1 ? p.up * p.up : 1);
  const long long slots = static_cast<long long>(kNumSMs) * C::MIN_CTAS;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned int>(work < slots ? work : slots));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  P3D_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, map, p, packed, scale, shift, out_split, out_nchw));
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

}  // namespace dc
}  // namespace p3d

using namespace p3d;

extern "C" int p3d_nchw_to_pixel_split(const float *in, int B, int C, int H, int W, float *out_split,
                                       p3d_stream_t stream) {
  if (!in || !out_split || B < 1 || C < 1 || H < 1 || W < 1) return P3D_ERR_INVALID_ARG;
  const long long hw = static_cast<long long>(H) * W;
  dim3 grid(static_cast<unsigned int>((hw + 31) / 32), static_cast<unsigned int>((C + 31) / 32), B);
  dc::nchw_to_pixel_split_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, C, hw, out_split);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

extern "C" size_t p3d_dense_conv2d_packed_weight_bytes(int taps, int Cin, int Cout, int n_tile) {
  if (taps < 1 || Cin < 32 || Cin % 32 || Cout < 1 || (n_tile != 16 && n_tile != 64 && n_tile != 128)) return 0;
  const size_t tiles = static_cast<size_t>((Cout + n_tile - 1) / n_tile);
  return align_up(tiles * taps * Cin * (2 * static_cast<size_t>(n_tile)) * sizeof(float));
}

extern "C" int p3d_dense_conv2d_split(const float *in_split, int B, int H, int W, int Cin, const float *packed_weight,
                                      int Cout, int n_tile, int kh, int kw, int stride, int pad, int up,
                                      const float *scale, const float *shift, int relu, float *out_split, int out_C,
                                      int out_c0, float *out_nchw, p3d_stream_t stream) {
  if (!in_split || !packed_weight || (!out_split && !out_nchw) || B < 1 || H < 1 || W < 1 || Cout < 1)
    return P3D_ERR_INVALID_ARG;
  if (Cin < 32 || Cin % 32 || (n_tile != 16 && n_tile != 64 && n_tile != 128)) return P3D_ERR_UNSUPPORTED;
  if (up < 1 || (up > 1 && (kh != up || kw != up || stride != up || pad != 0))) return P3D_ERR_UNSUPPORTED;
  if (up == 1 && (kh < 1 || kw < 1 || kh * kw > 32 || stride < 1 || stride > 2 || pad < 0)) return P3D_ERR_UNSUPPORTED;
  if (out_split && (Cout % 16 || out_C % 4 || out_c0 % 4 || out_c0 + Cout > out_C)) return P3D_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(in_split) & 15) || (reinterpret_cast<uintptr_t>(packed_weight) & 15) ||
      (reinterpret_cast<uintptr_t>(out_split) & 15))
    return P3D_ERR_INVALID_ARG;
  dc::Params p;
  p.B = B;
  p.H = H;
  p.W = W;
  p.Cin = Cin;
  p.taps = kh * kw;
  p.kw = kw;
  p.up = up;
  p.stride = up > 1 ? 1 : stride;
  p.pad = up > 1 ? 0 : pad;
  p.oH = up > 1 ? H : (H + 2 * pad - kh) / stride + 1;
  p.oW = up > 1 ? W : (W + 2 * pad - kw) / stride + 1;
  if (p.oH < 1 || p.oW < 1) return P3D_ERR_INVALID_ARG;
  p.out_H = up > 1 ? H * up : p.oH;
  p.out_W = up > 1 ? W * up : p.oW;
  p.tiles_x = (p.oW + dc::kTW - 1) / dc::kTW;
  p.tiles_y = (p.oH + dc::kTH - 1) / dc::kTH;
  p.n_ntiles = (Cout + n_tile - 1) / n_tile;
  p.cout = Cout;
  p.out_C = out_C;
  p.out_c0 = out_c0;
  p.relu = relu;
  CUtensorMap map;
  const int rc = dc::make_image_map(in_split, B, H, W, Cin, p.stride, &map);
  if (rc != P3D_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_tile == 16) return dc::launch<16>(map, p, packed_weight, scale, shift, out_split, out_nchw, st);
  if (n_tile == 64) return dc::launch<64>(map, p, packed_weight, scale, shift, out_split, out_nchw, st);
  return dc::launch<128>(map, p, packed_weight, scale, shift, out_split, out_nchw, st);
}
