// Sparse-conv gather-GEMM for the NARROW layers (Cin, Cout <= 32) on fp16 hi/lo' pair rows: register gather + warp MMA.
//
// Why a second kernel next to sparse_conv_f16.cu (tcgen05): on the 16- and 32-channel levels the tcgen05 pipeline is
// bound by what it costs to stage a gathered row in shared memory, not by the tensor pipe (profiles/r02_f16_sweep.md:
// a missing neighbour costs a full cp.async slot, 55 % of the slots are missing neighbours, one full/empty mbarrier
// round trip per 2-4 taps) - a level-0 layer takes 44 us for 0.38 GFLOP.  Here nothing is staged:
//
//   * one warp owns 16 output rows; per tap each lane loads ITS OWN A-fragment words of the two rows it serves straight
//     from L2 into registers (8 B for Cin = 16, 16 B for Cin = 32, per half), predicated off for a missing neighbour
//     (no request, no bytes), and a tap no row of the warp has is skipped altogether;
//   * the k dimension of mma.sync.m16n8k16 is permuted so that a lane's fragment (k-slots 2t, 2t+1, 2t+8, 2t+9) is
//     4 CONTIGUOUS channels of the H16 row -> one vector load per row half, no shuffles; the weight image is packed
//     with the same permutation, in fragment order, and sits in shared memory for the whole kernel (27-110 KB);
//   * loads run D taps ahead of the MMAs in a register ring; no mbarrier, no TMEM, no block-level synchronisation
//     after the weight load: warps stride over the 16-row tiles independently;
//   * products exactly as the tcgen05 kernel: acc_m += A_hi x B_hi, acc_l += A_hi x B_lo' + A_lo' x B_hi (fp32
//     accumulators), out = acc_m + acc_l * 2^-11, then the same fused BN / bias / residual / ReLU epilogue and the same
//     H16 / fp32 outputs (drop-in for p3d_sparse_conv_f16 on these shapes).
//
// The tensor pipe is irrelevant at these widths (N = 16/32): the roofline of these layers is the L2 gather
// (pairs x 4 Cin bytes) - see DESIGN.md section 3.
#include <cuda_fp16.h>

#include "common.cuh"

namespace p3d {
namespace wm {

constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;

struct Params {
  const uint8_t *in;         // H16 rows [n_in][4 * CIN bytes]
  const int32_t *nbr;        // [n_cap][K]
  const int32_t *n_out_dev;  // device row count (or null: n_cap)
  long long n_cap;
  int K;
  const uint4 *packed_w;     // [K][CIN / 16][COUT / 8][32 lanes] x (b0_hi, b1_hi, b0_lo, b1_lo)
  const float *scale, *shift;
  const uint8_t *residual;   // H16 rows [n][4 * COUT bytes] or null
  int relu;
  float *out_f32;            // [n][COUT] or null
  uint8_t *out_h16;          // [n][4 * COUT bytes] or null
  int32_t *status;           // bit 0: fp16 range overflow while producing out_h16
};

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void split_h16(float x, __half &hi, __half &lo, bool &ovf) {
  if (fabsf(x) > 65504.0f) {
    ovf = true;
    x = copysignf(65504.0f, x);
  }
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * kLoScale);
}
__device__ __forceinline__ float merge_h16(__half hi, __half lo) { return fmaf(__half2float(lo), kLoInv, __half2float(hi)); }

// A-fragment words of one tap for the two rows (g, g + 8) a lane serves: [row][half: hi, lo'][CIN / 8 words]
template <int CIN>
struct Frag {
  static constexpr int W = CIN / 8;  // 32-bit words per row half: 2 (8 B) or 4 (16 B)
  uint32_t w[2][2][W];
  unsigned any;  // ballot: rows of the warp that have this tap (0 = skip the tap)
};

template <int CIN>
__device__ __forceinline__ void load_frag(Frag<CIN> &f, const uint8_t *__restrict__ in, const int32_t *s_idx, int K,
                                          int tap, int g, int t) {
  const int i0 = s_idx[g * K + tap], i1 = s_idx[(g + 8) * K + tap];
  f.any = __ballot_sync(0xffffffffu, (i0 >= 0) || (i1 >= 0));
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int idx = r ? i1 : i0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (CIN == 16) {
        uint2 v = make_uint2(0u, 0u);
        if (idx >= 0) v = __ldg(reinterpret_cast<const uint2 *>(in + static_cast<size_t>(idx) * 64 + h * 32 + t * 8));
        f.w[r][h][0] = v.x;
        f.w[r][h][1] = v.y;
      } else {
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (idx >= 0) v = __ldg(reinterpret_cast<const uint4 *>(in + static_cast<size_t>(idx) * 128 + h * 64 + t * 16));
        f.w[r][h][0] = v.x;
        f.w[r][h][1] = v.y;
        f.w[r][h][2] = v.z;
        f.w[r][h][3] = v.w;
      }
    }
  }
}

template <int CIN, int COUT, int WARPS, int MINB, int D>
__global__ void __launch_bounds__(WARPS * 32, MINB) conv_wm_kernel(const Params p) {
  constexpr int KS = CIN / 16, NT = COUT / 8;
  constexpr int KCO = (COUT >= 32) ? 32 : 16;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint4 *s_w = reinterpret_cast<uint4 *>(smem_raw);  // [K][KS][NT][32]
  const int K = p.K;
  const int w_vec = K * KS * NT * 32;
  int32_t *s_idx_all = reinterpret_cast<int32_t *>(smem_raw + static_cast<size_t>(w_vec) * 16);  // [WARPS][16 * K]
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  for (int i = tid; i < w_vec; i += WARPS * 32) s_w[i] = __ldg(p.packed_w + i);  // static parameters
  __shared__ float s_sc[COUT], s_sh[COUT];
  const int g = lane >> 2, t = lane & 3;
  if (tid < COUT) {
    s_sc[tid] = p.scale ? __ldg(p.scale + tid) : 1.0f;
    s_sh[tid] = p.shift ? __ldg(p.shift + tid) : 0.0f;
  }
  __syncthreads();
  asm volatile("griddepcontrol.wait;" ::: "memory");  // rows / neighbour map / residual belong to earlier kernels
  const long long n = p.n_out_dev ? min(static_cast<long long>(p.n_out_dev[0]), p.n_cap) : p.n_cap;
  const long long n_tiles = (n + 15) / 16;
  int32_t *s_idx = s_idx_all + wid * 16 * K;
  bool ovf = false;
  for (long long tile = static_cast<long long>(blockIdx.x) * WARPS + wid; tile < n_tiles;
       tile += static_cast<long long>(gridDim.x) * WARPS) {
    const long long row0 = tile * 16;
    const int rows = static_cast<int>(min(16ll, n - row0));
    // this warp's 16 x K neighbour indices, coalesced; rows beyond n read as "missing"
    __syncwarp();
    for (int i = lane; i < 16 * K; i += 32) s_idx[i] = (i < rows * K) ? __ldg(p.nbr + row0 * K + i) : -1;
    __syncwarp();
    float am[NT][4], al[NT][4];
#pragma unroll
    for (int nn = 0; nn < NT; ++nn)
#pragma unroll
      for (int j = 0; j < 4; ++j) am[nn][j] = al[nn][j] = 0.f;
    Frag<CIN> buf[D];
#pragma unroll
    for (int d = 0; d < D; ++d)
      if (d < K) load_frag<CIN>(buf[d], p.in, s_idx, K, d, g, t);
    for (int tap0 = 0; tap0 < K; tap0 += D) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int tap = tap0 + d;
        if (tap < K) {
          if (buf[d].any) {
            const uint4 *wt = s_w + static_cast<size_t>(tap) * (KS * NT * 32) + lane;
#pragma unroll
            for (int s = 0; s < KS; ++s) {
              uint32_t ah[4], alo[4];
              ah[0] = buf[d].w[0][0][2 * s];
              ah[1] = buf[d].w[1][0][2 * s];
              ah[2] = buf[d].w[0][0][2 * s + 1];
              ah[3] = buf[d].w[1][0][2 * s + 1];
              alo[0] = buf[d].w[0][1][2 * s];
              alo[1] = buf[d].w[1][1][2 * s];
              alo[2] = buf[d].w[0][1][2 * s + 1];
              alo[3] = buf[d].w[1][1][2 * s + 1];
#pragma unroll
              for (int nn = 0; nn < NT; ++nn) {
                const uint4 b = wt[(s * NT + nn) * 32];
                mma16816(am[nn], ah, b.x, b.y);
                mma16816(al[nn], ah, b.z, b.w);
                mma16816(al[nn], alo, b.x, b.y);
              }
            }
          }
          if (tap + D < K) load_frag<CIN>(buf[d], p.in, s_idx, K, tap + D, g, t);
        }
      }
    }
    // epilogue: lane holds rows g, g + 8, columns 8 nn + 2t, + 1
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int lr = g + 8 * r;
      if (lr >= rows) continue;
      const size_t orow = static_cast<size_t>(row0 + lr);
#pragma unroll
      for (int nn = 0; nn < NT; ++nn) {
        const int c = nn * 8 + 2 * t;
        float v0 = fmaf(al[nn][2 * r], kLoInv, am[nn][2 * r]);
        float v1 = fmaf(al[nn][2 * r + 1], kLoInv, am[nn][2 * r + 1]);
        v0 = fmaf(v0, s_sc[c], s_sh[c]);
        v1 = fmaf(v1, s_sc[c + 1], s_sh[c + 1]);
        const size_t hoff = orow * (4 * COUT) + (c / KCO) * (4 * KCO) + (c % KCO) * 2;
        if (p.residual) {
          const __half2 rh = *reinterpret_cast<const __half2 *>(p.residual + hoff);
          const __half2 rl = *reinterpret_cast<const __half2 *>(p.residual + hoff + 2 * KCO);
          v0 = v0 + merge_h16(__low2half(rh), __low2half(rl));
          v1 = v1 + merge_h16(__high2half(rh), __high2half(rl));
        }
        if (p.relu) {
          v0 = fmaxf(v0, 0.f);
          v1 = fmaxf(v1, 0.f);
        }
        if (p.out_f32) *reinterpret_cast<float2 *>(p.out_f32 + orow * COUT + c) = make_float2(v0, v1);
        if (p.out_h16) {
          __half h0, l0, h1, l1;
          split_h16(v0, h0, l0, ovf);
          split_h16(v1, h1, l1, ovf);
          *reinterpret_cast<__half2 *>(p.out_h16 + hoff) = __halves2half2(h0, h1);
          *reinterpret_cast<__half2 *>(p.out_h16 + hoff + 2 * KCO) = __halves2half2(l0, l1);
        }
      }
    }
  }
  if (ovf && p.status) atomicOr(p.status, 1);
}

// fp32 [K][Cin][Cout] -> fragment-order image [K][Cin / 16][Cout / 8][32 lanes][b0_hi, b1_hi, b0_lo, b1_lo]:
// lane (g = lane / 4, t = lane % 4) of k-step s, n-tile nn holds output channel 8 nn + g and the input channels
// (Cin / 4) t + 4 s + {0, 1, 2, 3} (the k permutation of the header comment)
__global__ void __launch_bounds__(256) pack_weights_wm_kernel(const float *__restrict__ w, int K, int Cin, int Cout,
                                                              uint32_t *__restrict__ packed, int32_t *status) {
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int KS = Cin / 16, NT = Cout / 8;
  const long long total = static_cast<long long>(K) * KS * NT * 32;
  if (q >= total) return;
  const int lane = static_cast<int>(q % 32);
  const int nn = static_cast<int>((q / 32) % NT);
  const int s = static_cast<int>((q / (32 * NT)) % KS);
  const int tap = static_cast<int>(q / (32ll * NT * KS));
  const int g = lane >> 2, t = lane & 3;
  const int co = nn * 8 + g;
  const int ch0 = (Cin / 4) * t + 4 * s;
  __half hi[4], lo[4];
  bool ovf = false;
  for (int j = 0; j < 4; ++j) split_h16(w[(static_cast<size_t>(tap) * Cin + ch0 + j) * Cout + co], hi[j], lo[j], ovf);
  const __half2 b0h = __halves2half2(hi[0], hi[1]), b1h = __halves2half2(hi[2], hi[3]);
  const __half2 b0l = __halves2half2(lo[0], lo[1]), b1l = __halves2half2(lo[2], lo[3]);
  packed[q * 4 + 0] = *reinterpret_cast<const uint32_t *>(&b0h);
  packed[q * 4 + 1] = *reinterpret_cast<const uint32_t *>(&b1h);
  packed[q * 4 + 2] = *reinterpret_cast<const uint32_t *>(&b0l);
  packed[q * 4 + 3] = *reinterpret_cast<const uint32_t *>(&b1l);
  if (ovf && status) atomicOr(status, 1);
}

template <int CIN, int COUT, int WARPS, int MINB, int D>
int launch(const Params &p, cudaStream_t st) {
  constexpr int KS = CIN / 16, NT = COUT / 8;
  const size_t smem = static_cast<size_t>(p.K) * KS * NT * 512 + static_cast<size_t>(WARPS) * 16 * p.K * sizeof(int32_t);
  if (smem > 227 * 1024) return P3D_ERR_UNSUPPORTED;
  auto kern = conv_wm_kernel<CIN, COUT, WARPS, MINB, D>;
  P3D_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  int per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  if (per_sm > MINB) per_sm = MINB;
  if (per_sm < 1) per_sm = 1;
  const long long blocks_needed = (p.n_cap + 16 * WARPS - 1) / (16 * WARPS);
  long long grid = static_cast<long long>(kNumSMs) * per_sm;
  if (grid > blocks_needed) grid = blocks_needed;
  if (grid < 1) grid = 1;
  static const bool pdl = !(getenv("P3D_PDL") && atoi(getenv("P3D_PDL")) == 0);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned int>(grid));
  cfg.blockDim = dim3(WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  P3D_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, p));
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

inline bool supported(int K, int Cin, int Cout) {
  if (K < 1 || K > 32) return false;
  return (Cin == 16 && (Cout == 16 || Cout == 32)) || (Cin == 32 && Cout == 32);
}

}  // namespace wm
}  // namespace p3d

using namespace p3d;

extern "C" size_t p3d_sparse_conv_wm_packed_weight_bytes(int K, int Cin, int Cout) {
  if (!wm::supported(K, Cin, Cout)) return 0;
  return align_up(static_cast<size_t>(K) * (Cin / 16) * (Cout / 8) * 512);
}

extern "C" int p3d_sparse_conv_wm_pack_weights(const float *weight, int K, int Cin, int Cout, void *packed,
                                               int32_t *status_dev, p3d_stream_t stream) {
  if (!weight || !packed) return P3D_ERR_INVALID_ARG;
  if (!wm::supported(K, Cin, Cout)) return P3D_ERR_UNSUPPORTED;
  const long long total = static_cast<long long>(K) * (Cin / 16) * (Cout / 8) * 32;
  wm::pack_weights_wm_kernel<<<div_up(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      weight, K, Cin, Cout, static_cast<uint32_t *>(packed), status_dev);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

extern "C" int p3d_sparse_conv_wm(const void *in_h16, const int32_t *nbr, const int32_t *n_out_dev, int64_t n_out_cap,
                                  int K, int Cin, int Cout, const void *packed_weight, const float *scale,
                                  const float *shift, const void *residual_h16, int relu, float *out_f32, void *out_h16,
                                  int32_t *status_dev, p3d_stream_t stream) {
  if (n_out_cap < 0 || !packed_weight || (!out_f32 && !out_h16) || (n_out_cap && (!in_h16 || !nbr)))
    return P3D_ERR_INVALID_ARG;
  if (!wm::supported(K, Cin, Cout)) return P3D_ERR_UNSUPPORTED;
  if (n_out_cap == 0) return P3D_OK;
  if ((reinterpret_cast<uintptr_t>(in_h16) & 15) || (reinterpret_cast<uintptr_t>(out_f32) & 15) ||
      (reinterpret_cast<uintptr_t>(out_h16) & 15) || (reinterpret_cast<uintptr_t>(packed_weight) & 15) ||
      (reinterpret_cast<uintptr_t>(residual_h16) & 15) || (reinterpret_cast<uintptr_t>(nbr) & 3))
    return P3D_ERR_INVALID_ARG;
  wm::Params p;
  p.in = static_cast<const uint8_t *>(in_h16);
  p.nbr = nbr;
  p.n_out_dev = n_out_dev;
  p.n_cap = n_out_cap;
  p.K = K;
  p.packed_w = static_cast<const uint4 *>(packed_weight);
  p.scale = scale;
  p.shift = shift;
  p.residual = static_cast<const uint8_t *>(residual_h16);
  p.relu = relu;
  p.out_f32 = out_f32;
  p.out_h16 = static_cast<uint8_t *>(out_h16);
  p.status = status_dev;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (Cin == 16 && Cout == 16) return wm::launch<16, 16, 8, 3, 4>(p, st);
  if (Cin == 16 && Cout == 32) return wm::launch<16, 32, 16, 1, 4>(p, st);
  if (Cin == 32 && Cout == 32) return wm::launch<32, 32, 16, 1, 3>(p, st);
  return P3D_ERR_UNSUPPORTED;
}
