// Sparse-conv gather-GEMM on Blackwell tensor cores: tcgen05.mma kind::tf32, accumulators in TMEM,
// 3xTF32 split (a = hi + lo, b = hi + lo; D += a_hi*b_hi + a_hi*b_lo + a_lo*b_hi) for fp32-level accuracy.
//
// One CTA owns 128 consecutive output rows (UMMA M = 128, N = Cout, fp32 accumulate in Cout TMEM columns)
// and walks the K dimension = (active taps) x (Cin in chunks of KC <= 32 channels) through a ring of
// shared-memory stages:
//
//   warps 0-3  producers: thread r gathers row nbr[row0 + r][tap] of the input (KC floats, 16-byte
//              loads), splits every value into tf32 hi / lo and stores both into the stage's A tiles in the
//              canonical no-swizzle K-major core-matrix layout (8 rows x 16 B per core matrix) — a missing
//              neighbour is a zero row.  Thread 0 also fires ONE cp.async.bulk per stage that drops the
//              pre-packed weight slice [hi | lo] for (tap, chunk) into the B tiles (TMA engine, mbarrier
//              complete_tx).  Generic-proxy stores are made visible to the tensor core with
//              fence.proxy.async before the mbarrier arrive.
//   warp 4     one elected lane issues 3 x KC/8 tcgen05.mma per stage and tcgen05.commit's the stage back to
//              the producers; after the last stage it commits the accumulator to the epilogue barrier.
//   warps 0-3  epilogue: tcgen05.ld 32 lanes x 16 columns at a time, fused BN scale/shift (+bias), residual,
//              ReLU, 64 B per thread row stores.
//
// Taps that no row of the tile uses are skipped for the whole tile (block-uniform bitmask).
// Weights are packed once per layer by p3d_sparse_conv_pack_weights into exactly the shared-memory image:
//   packed[tap][chunk g][hi|lo][KC/4 k-chunks][Cout rows][4 floats].
#include "tc_common.cuh"

namespace p3d {
namespace tc {

template <int CIN, int COUT>
struct Cfg {
  static constexpr int KC = kc_of(CIN);          // channels per stage
  static constexpr int G = CIN / KC;             // stages per tap
  static constexpr int CH = KC / 4;              // 16-byte k-chunks per stage
  static constexpr int A_TILE = KC * kM * 4;     // bytes, one of hi / lo
  static constexpr int B_TILE = KC * COUT * 4;   // bytes, one of hi / lo
  static constexpr int A_STAGE = 2 * A_TILE;      // hi + lo
  static constexpr int B_STAGE = 2 * B_TILE;      // hi + lo (one cp.async.bulk)
  // two independent rings (<= 96 KB together: two CTAs per SM): SA slots of gathered rows, SB slots of weights.
  // The weight ring is deeper so that the TMA latency never sits on the gather -> MMA critical path.
  // Narrow layers (Cout <= 32) have 300-530 row tiles: three CTAs per SM (444 slots) instead of two keeps the
  // 309-tile 32-channel layers in ONE wave and adds a third MMA-issuing warp per SM.
  static constexpr int MIN_CTAS = (COUT <= 32) ? 3 : 2;
  static constexpr int SA = (COUT <= 32) ? 3 : ((COUT <= 64) ? 4 : 3);
  static constexpr int BUDGET = (COUT <= 32) ? 56 * 1024 : 96 * 1024;  // 3 x (56 K ring + 14 K map + 3 K) <= 228 KB per SM
  static constexpr int SB_RAW = (BUDGET - SA * A_STAGE) / B_STAGE;
  static constexpr int SB = SB_RAW > 8 ? 8 : SB_RAW;
  static constexpr int RING_BYTES = SA * A_STAGE + SB * B_STAGE;
  // Independent TMEM accumulators: back-to-back tcgen05.mma into ONE accumulator serialise on its dependency
  // latency (~200 cycles each, measured), so the three 3xTF32 products go to separate column ranges and are summed
  // in the epilogue.  Cout = 128 keeps two (256 columns) so that two CTAs still fit the 512 TMEM columns of an SM.
  static constexpr int NACC = (COUT <= 64) ? 3 : 2;
  static constexpr int TMEM_COLS = (NACC * COUT <= 32) ? 32 : (NACC * COUT <= 64) ? 64 : (NACC * COUT <= 128) ? 128
                                   : (NACC * COUT <= 256) ? 256 : 512;
  static constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(COUT >> 3) << 17) |
                                    (static_cast<uint32_t>(kM >> 4) << 24);
  static_assert(CIN % 16 == 0 && COUT % 16 == 0 && COUT <= 256, "tensor-core path needs 16-channel multiples");
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kThreads, Cfg<CIN, COUT>::MIN_CTAS)
    gather_gemm_tf32x3_kernel(const float *__restrict__ in, const int32_t *__restrict__ nbr,
                              const int32_t *__restrict__ n_out_dev, long long n_cap, int K, int splits,
                              const float *__restrict__ packed_w, const float *__restrict__ scale,
                              const float *__restrict__ shift, const float *__restrict__ residual, int relu,
                              float *__restrict__ out_base) {
  using C = Cfg<CIN, COUT>;
  // Persistent CTAs: the grid is sized for the SMs (not for the row CAPACITY, which would launch thousands of empty
  // CTAs) and every CTA walks the work items  w = blockIdx.x, +gridDim.x, ...  with  w = tile * splits + split.
  // split-K over taps: split s owns the taps t == s (mod splits) and writes raw partial sums to slab s of `out_base`
  // (the host passes no epilogue parameters then; rows_finalize_kernel adds the slabs in a fixed order).
  const long long n = n_out_dev ? min(static_cast<long long>(n_out_dev[0]), n_cap) : n_cap;
  const long long n_work = ((n + kM - 1) / kM) * splits;
  if (static_cast<long long>(blockIdx.x) >= n_work) return;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *a_base = smem;                                                  // SA x A_STAGE
  uint8_t *b_base = smem + C::SA * C::A_STAGE;                             // SB x B_STAGE
  int32_t *s_nbr = reinterpret_cast<int32_t *>(smem + C::RING_BYTES);      // [kM][K]
  // barriers: a_full[4] a_empty[4] b_full[8] b_empty[8] tmem_full
  __shared__ __align__(8) unsigned long long s_bar[4 + 4 + 8 + 8 + 1];
  constexpr int kAF = 0, kAE = 4, kBF = 8, kBE = 16, kTF = 24;
  __shared__ uint32_t s_tmem_base;
  __shared__ uint32_t s_active;                                             // bit t: some row uses tap t (K <= 32)

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  if (tid == kProducers) {  // first lane of the MMA warp
    for (int s = 0; s < C::SA; ++s) {
      mbar_init(smem_u32(&s_bar[kAF + s]), 32);        // the 32 lanes of the slot's producer warp
      mbar_init(smem_u32(&s_bar[kAE + s]), 1);         // tcgen05.commit
    }
    for (int s = 0; s < C::SB; ++s) {
      mbar_init(smem_u32(&s_bar[kBF + s]), 1);         // arrive.expect_tx + bulk-copy bytes
      mbar_init(smem_u32(&s_bar[kBE + s]), 1);         // tcgen05.commit
    }
    mbar_init(smem_u32(&s_bar[kTF]), 1);
    fence_mbar_init();
  }
  if (wid == 4) {  // TMEM allocation is warp-collective
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

  int use_base = 0;  // pipeline uses consumed by earlier work items of this CTA: the ring phases keep running
  int item_it = 0;
  for (long long w = blockIdx.x; w < n_work; w += gridDim.x, ++item_it) {
    const long long tile = w / splits;
    const int split = static_cast<int>(w - tile * splits);
    const long long row0 = tile * kM;
    const int rows = static_cast<int>(min(static_cast<long long>(kM), n - row0));
    float *out = out_base + static_cast<size_t>(split) * static_cast<size_t>(n_cap) * COUT;
    if (tid == 0) s_active = 0u;
    __syncthreads();  // previous item drained: epilogue finished reading TMEM, s_nbr free
    {
      uint32_t mine = 0u;
      for (int q = tid; q < kM * K; q += kThreads) {
        const int v = (q < rows * K) ? __ldg(nbr + row0 * K + q) : -1;
        s_nbr[q] = v;
        if (v >= 0) mine |= 1u << (q % K);
      }
      mine = __reduce_or_sync(0xffffffffu, mine);
      if (lane == 0 && mine) atomicOr(&s_active, mine);
    }
    __syncthreads();
    uint32_t tap_mask = 0xffffffffu;
    if (splits > 1) {
      tap_mask = 0u;
      for (int t = split; t < K; t += splits) tap_mask |= 1u << t;
    }
    const uint32_t active = s_active & tap_mask;
    const int n_stage_uses = __popc(active) * C::G;

    if (wid < 4) {
      // ---------------------------------------------------------------- producers
      // Warp w (< SA) owns ring slot w: it gathers all 128 rows of its (tap, chunk) uses — lane l takes rows
      // l, l+32, l+64, l+96, i.e. 4 x CH independent 16-byte loads in flight per lane — so SA stages' worth of L2
      // gathers are in flight per CTA.  With one warp per slot a warp can never run a full mbarrier phase ahead of
      // the tensor core, so the parity waits cannot alias.
      const int r = tid;
      int use = use_base;
      for (int t = 0; t < K; ++t) {
        if (!((active >> t) & 1u)) continue;
        for (int g = 0; g < C::G; ++g, ++use) {
          const int s = use % C::SA;
          if (s != wid) continue;
          const uint32_t ph = static_cast<uint32_t>((use / C::SA) & 1);
          float4 v[4][C::CH];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int src = s_nbr[(lane + 32 * q) * K + t];
            if (src >= 0) {
              const float4 *p = reinterpret_cast<const float4 *>(in + static_cast<size_t>(src) * CIN + g * C::KC);
#pragma unroll
              for (int c = 0; c < C::CH; ++c) v[q][c] = __ldg(p + c);
            } else {
#pragma unroll
              for (int c = 0; c < C::CH; ++c) v[q][c] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          mbar_wait(smem_u32(&s_bar[kAE + s]), ph ^ 1u);  // slot free (first pass returns at once)
          uint8_t *st = a_base + s * C::A_STAGE;
          // A-tile layout: k-chunk c at c*2048 (LBO), 8-row groups 128 B apart (SBO): the 32 lanes of one store
          // instruction (32 consecutive rows, same chunk) cover 512 contiguous bytes — no bank conflicts.
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int row = lane + 32 * q;
            const uint32_t a_off = static_cast<uint32_t>((row >> 3) * 128 + (row & 7) * 16);
#pragma unroll
            for (int c = 0; c < C::CH; ++c) {
              float4 h, l;
              split_tf32(v[q][c].x, h.x, l.x);
              split_tf32(v[q][c].y, h.y, l.y);
              split_tf32(v[q][c].z, h.z, l.z);
              split_tf32(v[q][c].w, h.w, l.w);
              *reinterpret_cast<float4 *>(st + c * (kM * 16) + a_off) = h;
              *reinterpret_cast<float4 *>(st + C::A_TILE + c * (kM * 16) + a_off) = l;
            }
          }
          fence_proxy_async();
          mbar_arrive(smem_u32(&s_bar[kAF + s]));
        }
      }
      // ---------------------------------------------------------------- epilogue
      mbar_wait(smem_u32(&s_bar[kTF]), static_cast<uint32_t>(item_it & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(wid * 32) << 16);
      const bool live = r < rows;
      float *orow = out + (row0 + r) * COUT;
      const float *rrow = residual ? residual + (row0 + r) * COUT : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < COUT; c0 += 16) {
        uint32_t a[16];
        if (n_stage_uses > 0) {
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
              "%15}, [%16];"
              : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
                "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
              : "r"(taddr + static_cast<uint32_t>(c0)));
#pragma unroll
          for (int acc = 1; acc < C::NACC; ++acc) {
            uint32_t b[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
                "%15}, [%16];"
                : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
                  "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
                : "r"(taddr + static_cast<uint32_t>(acc * COUT + c0)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) a[j] = __float_as_uint(__uint_as_float(a[j]) + __uint_as_float(b[j]));
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) a[j] = 0u;
        }
        if (live) {
          float o[16], res[16];
          if (rrow) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 rv = __ldg(reinterpret_cast<const float4 *>(rrow + c0 + j));
              res[j] = rv.x;
              res[j + 1] = rv.y;
              res[j + 2] = rv.z;
              res[j + 3] = rv.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float v = __uint_as_float(a[j]);
            if (scale) v = v * __ldg(scale + c0 + j);
            if (shift) v = v + __ldg(shift + c0 + j);
            if (rrow) v = v + res[j];
            if (relu) v = fmaxf(v, 0.f);
            o[j] = v;
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4 *>(orow + c0 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
        }
      }
      tc_fence_before();
    } else if (wid == 4) {
      // ---------------------------------------------------------------- MMA issuer (warp 4)
      if (n_stage_uses == 0) {
        if (lane == 0) mbar_arrive(smem_u32(&s_bar[kTF]));
      } else {
        for (int u = 0; u < n_stage_uses; ++u) {
          const int use = use_base + u;
          const int sa = use % C::SA, sb = use % C::SB;
          mbar_wait(smem_u32(&s_bar[kBF + sb]), static_cast<uint32_t>((use / C::SB) & 1));
          mbar_wait(smem_u32(&s_bar[kAF + sa]), static_cast<uint32_t>((use / C::SA) & 1));
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_hi = smem_u32(a_base + sa * C::A_STAGE), a_lo = a_hi + C::A_TILE;
            const uint32_t b_hi = smem_u32(b_base + sb * C::B_STAGE), b_lo = b_hi + COUT * 16;  // lo rows follow the hi rows
#pragma unroll
            for (int j = 0; j < C::KC / 8; ++j) {
              const uint32_t ao = static_cast<uint32_t>(2 * j) * (kM * 16), bo = static_cast<uint32_t>(2 * j) * (2 * COUT * 16);
              const uint64_t dah = smem_desc(a_hi + ao, kM * 16, 128), dal = smem_desc(a_lo + ao, kM * 16, 128);
              const uint64_t dbh = smem_desc(b_hi + bo, 2 * COUT * 16, 128), dbl = smem_desc(b_lo + bo, 2 * COUT * 16, 128);
              const uint32_t first = (u | j) ? 1u : 0u;
              umma_tf32(tmem_base, dal, dbh, C::IDESC, first);                               // acc 0
              umma_tf32(tmem_base + (C::NACC == 3 ? COUT : 0), dah, dbl, C::IDESC, C::NACC == 3 ? first : 1u);  // acc 1 (or 0)
              umma_tf32(tmem_base + (C::NACC - 1) * COUT, dah, dbh, C::IDESC, first);         // last acc
            }
            umma_commit(smem_u32(&s_bar[kAE + sa]));                             // gathered rows consumed
            umma_commit(smem_u32(&s_bar[kBE + sb]));                             // weights consumed
            if (u == n_stage_uses - 1) umma_commit(smem_u32(&s_bar[kTF]));      // accumulator ready
          }
          __syncwarp();
        }
      }
      tc_fence_before();
    } else {
      // ---------------------------------------------------------------- weight TMA (warp 5, one lane)
      if (lane == 0) {
        int use = use_base;
        for (int t = 0; t < K; ++t) {
          if (!((active >> t) & 1u)) continue;
          for (int g = 0; g < C::G; ++g, ++use) {
            const int sb = use % C::SB;
            mbar_wait(smem_u32(&s_bar[kBE + sb]), static_cast<uint32_t>((use / C::SB) & 1) ^ 1u);
            mbar_arrive_expect_tx(smem_u32(&s_bar[kBF + sb]), static_cast<uint32_t>(C::B_STAGE));
            bulk_g2s(smem_u32(b_base + sb * C::B_STAGE), packed_w + (static_cast<size_t>(t) * C::G + g) * (2 * C::KC * COUT),
                     static_cast<uint32_t>(C::B_STAGE), smem_u32(&s_bar[kBF + sb]));
          }
        }
      }
    }
    use_base += n_stage_uses;
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

// packed[tap][g][c][hl * Cout + n][j] = split(W[tap][g*KC + 4c + j][n]): per (tap, 16-channel chunk) a K-major
// operand of 2*Cout rows (tf32 hi rows, then lo rows), so [B_hi | B_lo] is ONE UMMA operand with N = 2*Cout.
__global__ void __launch_bounds__(256) pack_weights_kernel(const float *__restrict__ w, int K, int Cin, int Cout,
                                                           float *__restrict__ packed) {
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(K) * Cin * Cout;
  if (q >= total) return;
  const int KC = kc_of(Cin), G = Cin / KC;
  const int n = static_cast<int>(q % Cout);
  const int ci = static_cast<int>((q / Cout) % Cin);
  const int t = static_cast<int>(q / (static_cast<long long>(Cout) * Cin));
  const int g = ci / KC, c = (ci % KC) / 4, j = ci & 3;
  float hi, lo;
  split_tf32(w[q], hi, lo);
  const size_t stage = static_cast<size_t>(2) * KC * Cout;  // floats per (tap, g)
  const size_t base = (static_cast<size_t>(t) * G + g) * stage + static_cast<size_t>(c) * (2 * Cout) * 4;
  packed[base + static_cast<size_t>(n) * 4 + j] = hi;
  packed[base + static_cast<size_t>(Cout + n) * 4 + j] = lo;
}

template <int CIN, int COUT>
int launch(const float *in, const int32_t *nbr, const int32_t *n_out_dev, int64_t n_cap, int K, const float *packed,
           const float *scale, const float *shift, const float *residual, int relu, float *out, cudaStream_t st,
           int splits = 1) {
  using C = Cfg<CIN, COUT>;
  const size_t smem = static_cast<size_t>(C::RING_BYTES) + static_cast<size_t>(kM) * K * sizeof(int32_t) + 1024;
  if (smem > 227 * 1024) return P3D_ERR_UNSUPPORTED;
  auto kern = gather_gemm_tf32x3_kernel<CIN, COUT>;
  P3D_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const long long work = ((n_cap + kM - 1) / kM) * splits;
  const long long slots = static_cast<long long>(kNumSMs) * C::MIN_CTAS;
  kern<<<static_cast<unsigned int>(work < slots ? work : slots), kThreads, smem, st>>>(in, nbr, n_out_dev, n_cap, K, splits, packed, scale, shift, residual, relu,
                                                  out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

// out[r, c] = act((sum_s partial[s][r][c]) * scale + shift (+ residual)), slabs added in index order
__global__ void __launch_bounds__(256) rows_finalize_kernel(const float *__restrict__ partial, int splits,
                                                            const int32_t *__restrict__ n_dev, long long n_cap, int C,
                                                            const float *__restrict__ scale, const float *__restrict__ shift,
                                                            const float *__restrict__ residual, int relu,
                                                            float *__restrict__ out) {
  const long long n = n_dev ? min(static_cast<long long>(n_dev[0]), n_cap) : n_cap;
  const int c4 = C / 4;
  const size_t slab = static_cast<size_t>(n_cap) * C / 4;
  // persistent grid-stride loop over float4s: the grid is sized for the SMs, the trip count follows the device row count
  for (long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; q < n * c4;
       q += static_cast<long long>(gridDim.x) * blockDim.x) {
  const int c = static_cast<int>(q % c4) * 4;
  const float4 *p = reinterpret_cast<const float4 *>(partial) + q;
  float4 a = __ldg(p);
  for (int s = 1; s < splits; ++s) {
    const float4 b = __ldg(p + s * slab);
    a.x += b.x;
    a.y += b.y;
    a.z += b.z;
    a.w += b.w;
  }
  float v[4] = {a.x, a.y, a.z, a.w};
  float r[4] = {0.f, 0.f, 0.f, 0.f};
  if (residual) {
    const float4 rv = __ldg(reinterpret_cast<const float4 *>(residual) + q);
    r[0] = rv.x;
    r[1] = rv.y;
    r[2] = rv.z;
    r[3] = rv.w;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (scale) v[j] = v[j] * __ldg(scale + c + j);
    if (shift) v[j] = v[j] + __ldg(shift + c + j);
    if (residual) v[j] = v[j] + r[j];
    if (relu) v[j] = fmaxf(v[j], 0.f);
  }
  reinterpret_cast<float4 *>(out)[q] = make_float4(v[0], v[1], v[2], v[3]);
  }
}


}  // namespace tc
}  // namespace p3d

using namespace p3d;

extern "C" size_t p3d_sparse_conv_packed_weight_bytes(int K, int Cin, int Cout) {
  if (K < 1 || Cin < 16 || Cout < 16 || Cin % 16 || Cout % 16) return 0;
  return align_up(static_cast<size_t>(K) * Cin * Cout * 2 * sizeof(float));
}

extern "C" int p3d_sparse_conv_pack_weights(const float *weight, int K, int Cin, int Cout, float *packed,
                                            p3d_stream_t stream) {
  if (!weight || !packed || K < 1) return P3D_ERR_INVALID_ARG;
  if (Cin < 16 || Cout < 16 || Cin % 16 || Cout % 16 || false) return P3D_ERR_UNSUPPORTED;
  const long long total = static_cast<long long>(K) * Cin * Cout;
  tc::pack_weights_kernel<<<div_up(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(weight, K, Cin, Cout,
                                                                                            packed);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

extern "C" size_t p3d_sparse_conv_splitk_workspace_bytes(int64_t n_out_cap, int Cin, int Cout) {
  (void)Cin;
  if (n_out_cap <= 0 || Cout < 16) return 0;
  const int s = tc::splits_for(Cout);
  return s > 1 ? align_up(static_cast<size_t>(s) * static_cast<size_t>(n_out_cap) * Cout * sizeof(float)) : 0;
}

extern "C" int p3d_sparse_conv_gather_gemm_tf32x3_ws(const float *in, const int32_t *nbr, const int32_t *n_out_dev,
                                                     int64_t n_out_cap, int K, int Cin, int Cout, const float *weight,
                                                     const float *scale, const float *shift, const float *residual,
                                                     int relu, float *out, void *workspace, size_t workspace_bytes,
                                                     p3d_stream_t stream) {
  if (n_out_cap < 0 || K < 1 || K > 32 || !weight || (n_out_cap && (!in || !nbr || !out))) return P3D_ERR_INVALID_ARG;
  if (n_out_cap == 0) return P3D_OK;
  if ((reinterpret_cast<uintptr_t>(in) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (reinterpret_cast<uintptr_t>(weight) & 15) || (reinterpret_cast<uintptr_t>(residual) & 15) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15))
    return P3D_ERR_INVALID_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Wide layers have few 128-row tiles (52 - 130 at the C3 sizes): split the taps over 2 - 3 CTAs per tile so that all
  // SMs work on them; needs the scratch slab(s) of p3d_sparse_conv_splitk_workspace_bytes.
  int splits = tc::splits_for(Cout);
  if (splits > K) splits = K;
  const size_t need = static_cast<size_t>(splits) * static_cast<size_t>(n_out_cap) * Cout * sizeof(float);
  const bool split = splits > 1 && workspace && workspace_bytes >= need;
  float *conv_out = split ? static_cast<float *>(workspace) : out;
  const float *k_scale = split ? nullptr : scale, *k_shift = split ? nullptr : shift, *k_res = split ? nullptr : residual;
  const int k_relu = split ? 0 : relu, k_splits = split ? splits : 1;
  int rc = P3D_ERR_UNSUPPORTED;
#define P3D_TC_CASE(CI, CO)                                                                                         \
  if (Cin == CI && Cout == CO)                                                                                      \
    rc = tc::launch<CI, CO>(in, nbr, n_out_dev, n_out_cap, K, weight, k_scale, k_shift, k_res, k_relu, conv_out, st, \
                            k_splits);
  P3D_TC_CASE(16, 16)
  P3D_TC_CASE(16, 32)
  P3D_TC_CASE(32, 32)
  P3D_TC_CASE(32, 64)
  P3D_TC_CASE(64, 64)
  P3D_TC_CASE(64, 128)
  P3D_TC_CASE(128, 128)
#undef P3D_TC_CASE
  if (rc != P3D_OK || !split) return rc;
  const long long fin_blocks = (n_out_cap * (Cout / 4) + 255) / 256;
  tc::rows_finalize_kernel<<<static_cast<unsigned int>(fin_blocks < kNumSMs * 8 ? fin_blocks : kNumSMs * 8), 256, 0, st>>>(conv_out, splits, n_out_dev, n_out_cap, Cout,
                                                                                scale, shift, residual, relu, out);
  P3D_LAUNCH_CHECK();
  return P3D_OK;
}

extern "C" int p3d_sparse_conv_gather_gemm_tf32x3(const float *in, const int32_t *nbr, const int32_t *n_out_dev,
                                                  int64_t n_out_cap, int K, int Cin, int Cout, const float *weight,
                                                  const float *scale, const float *shift, const float *residual,
                                                  int relu, float *out, p3d_stream_t stream) {
  return p3d_sparse_conv_gather_gemm_tf32x3_ws(in, nbr, n_out_dev, n_out_cap, K, Cin, Cout, weight, scale, shift, residual,
                                               relu, out, nullptr, 0, stream);
}
