// mlcv_score: K-fold (or leave-one-out) cross-validation likelihood for a whole
// grid of bandwidth factors in one pass over the pair space.
// Replaces the joblib loop over _kde_cv_score, python/kdsource/kde.py:106-153 and
// :208-211: for every test point i and grid factor g
//     S_ig = sum_{j not in fold(i)} c_j * 2^(nk_g * |x_i - x_j|^2 / h_j^2)
// with nk_g = -log2(e) / (2 grid_g^2) and c_j = w_j h_j^-D / c_ref, followed by
//     partial[block][g] = sum_{i in block} coefw_i * (ln S_ig - lwtr_i).
// One pair distance feeds G exponentials, so the kernel is MUFU.EX2 (SFU) bound:
// per kernel evaluation 1 FMUL + 1 FFMA (packed as f32x2) + 1 MUFU.
//
// The fold structure enters only as a per-test-point excluded index interval
// [excl_lo, excl_hi) of sources (contiguous KFold without shuffle, kde.py:134),
// so leave-one-out is the interval [i, i+1).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "kdsb200.h"

namespace kdsb {

#ifndef KDSB_MLCV_POLY_DEFAULT
#define KDSB_MLCV_POLY_DEFAULT 4
#endif

constexpr int CV_THREADS = 256;
constexpr int CV_STAGES = 3;

template <int GT>
struct NkParams {
  float nk[GT];
};

// deterministic block sum of v over the CTA for each g; S2 is [GT][CV_THREADS]
template <int GT>
__device__ __forceinline__ void block_reduce_store(double* S2, double* __restrict__ out_row) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int g = warp; g < GT; g += CV_THREADS / 32) {
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < CV_THREADS / 32; e++) s += S2[g * CV_THREADS + e * 32 + lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0) out_row[g] = s;
  }
}

// Which grid slots take the FMA-pipe exponential (the rest use MUFU.EX2).  PM is a
// small pattern id: 0 none, 2 every 2nd, 3 every 3rd, 4 every 4th, 38 three of eight.
template <int PM>
__device__ __forceinline__ constexpr bool use_poly(int g) {
  return PM == 2 ? (g % 2 == 1)
       : PM == 3 ? (g % 3 == 2)
       : PM == 4 ? (g % 4 == 3)
       : PM == 38 ? (g % 8 == 2 || g % 8 == 5 || g % 8 == 7)
       : false;
}

template <int D, int GT, int PM>
__global__ void __launch_bounds__(CV_THREADS, 2)
    mlcv_pairs_kernel(const float* __restrict__ packed, uint32_t n_tiles, uint32_t tiles_per_split,
                      const float* __restrict__ qT, uint64_t M, uint64_t Mpad,
                      const int32_t* __restrict__ excl_lo, const int32_t* __restrict__ excl_hi,
                      const double* __restrict__ coefw, const double* __restrict__ lwtr,
                      NkParams<GT> P, int fused, double* __restrict__ sbuf,
                      double* __restrict__ partial) {
  constexpr int ROWS = D + 2;
  constexpr int TILE_FLOATS = ROWS * TS;
  constexpr uint32_t TILE_BYTES = TILE_FLOATS * sizeof(float);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  double* S2 = reinterpret_cast<double*>(smem_raw + (size_t)CV_STAGES * TILE_BYTES);
  __shared__ __align__(8) uint64_t full[CV_STAGES];

  const int tid = threadIdx.x;
  const uint32_t t0 = blockIdx.y * tiles_per_split;
  const uint32_t t1 = min(n_tiles, t0 + tiles_per_split);
  const uint32_t nt_all = t1 > t0 ? t1 - t0 : 0;

  if (tid == 0) {
    for (int s = 0; s < CV_STAGES; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const uint64_t m = (uint64_t)blockIdx.x * CV_THREADS + tid;
  float q[D];
#pragma unroll
  for (int k = 0; k < D; k++) q[k] = qT[(uint64_t)k * Mpad + m];
  const int lo = excl_lo[m], hi = excl_hi[m];
#pragma unroll
  for (int g = 0; g < GT; g++) S2[g * CV_THREADS + tid] = 0.0;

  // Source tiles that lie inside the excluded interval of EVERY row of this CTA
  // (the rows' own fold, for all CTAs that do not straddle a fold boundary)
  // contribute nothing and are skipped: [skip_a, skip_b) in tile units.  Padding
  // rows (m >= M, empty interval) do not take part in the intersection.
  __shared__ int cta_lo, cta_hi;
  if (tid == 0) {
    cta_lo = 0;
    cta_hi = 0x7fffffff;
  }
  __syncthreads();
  if (m < M) {
    atomicMax(&cta_lo, lo);
    atomicMin(&cta_hi, hi);
  }
  __syncthreads();
  uint32_t skip_a = 0, skip_b = 0;
  if (cta_hi > cta_lo) {
    const uint32_t a = ((uint32_t)cta_lo + TS - 1) / TS, b = (uint32_t)cta_hi / TS;
    const uint32_t a2 = max(a, t0), b2 = min(b, t1);
    if (b2 > a2) {
      skip_a = a2 - t0;
      skip_b = b2 - t0;
    }
  }
  const uint32_t nskip = skip_b - skip_a;
  const uint32_t nt = nt_all - nskip;
  // logical tile index (0..nt) -> physical tile of this split
  auto phys = [&](uint32_t l) { return l < skip_a ? l : l + nskip; };

  if (tid == 0) {
    for (uint32_t p = 0; p < (uint32_t)(CV_STAGES - 1) && p < nt; p++) {
      mbar_expect_tx(&full[p], TILE_BYTES);
      tma_load_1d(tiles + p * TILE_FLOATS, packed + (uint64_t)(t0 + phys(p)) * TILE_FLOATS,
                  TILE_BYTES, &full[p]);
    }
  }

  for (uint32_t it = 0; it < nt; it++) {
    const uint32_t s = it % CV_STAGES;
    const uint32_t parity = (it / CV_STAGES) & 1u;
    if (tid == 0 && it + CV_STAGES - 1 < nt) {
      const uint32_t ps = (it + CV_STAGES - 1) % CV_STAGES;
      mbar_expect_tx(&full[ps], TILE_BYTES);
      tma_load_1d(tiles + ps * TILE_FLOATS,
                  packed + (uint64_t)(t0 + phys(it + CV_STAGES - 1)) * TILE_FLOATS, TILE_BYTES,
                  &full[ps]);
    }
    mbar_wait(&full[s], parity);
    const float* tile = tiles + s * TILE_FLOATS;
    // excluded interval relative to this tile, clamped to [0, TS]
    const int tbase = (int)((t0 + phys(it)) * TS);
    const int xlo = max(lo - tbase, 0), xhi = min(hi - tbase, TS);

    f2_t acc[GT];
#pragma unroll
    for (int g = 0; g < GT; g++) acc[g] = 0ull;

#pragma unroll 1
    for (int j = 0; j < TS; j += 2) {
      float2 r[ROWS];
#pragma unroll
      for (int k = 0; k < ROWS; k++) r[k] = *reinterpret_cast<const float2*>(tile + k * TS + j);
      const f2_t a2 = f2_pack(r[D].x, r[D].y);
      f2_t u2 = 0ull;
#pragma unroll
      for (int k = 0; k < D; k++) {
        const f2_t d2 = f2_fma(f2_pack(q[k], q[k]), a2, f2_pack(r[k].x, r[k].y));
        u2 = f2_fma(d2, d2, u2);
      }
      const float c0 = (j >= xlo && j < xhi) ? 0.f : r[D + 1].x;
      const float c1 = (j + 1 >= xlo && j + 1 < xhi) ? 0.f : r[D + 1].y;
      const f2_t c2 = f2_pack(c0, c1);
#pragma unroll
      for (int g = 0; g < GT; g++) {
        const f2_t arg = f2_mul(u2, f2_pack(P.nk[g], P.nk[g]));
        if (use_poly<PM>(g)) {
          acc[g] = f2_fma(c2, ex2_poly2(arg), acc[g]);
        } else {
          float x0, x1;
          f2_unpack(arg, x0, x1);
          acc[g] = f2_fma(c2, f2_pack(ex2_approx(x0), ex2_approx(x1)), acc[g]);
        }
      }
    }
#pragma unroll
    for (int g = 0; g < GT; g++) {
      float a, b;
      f2_unpack(acc[g], a, b);
      S2[g * CV_THREADS + tid] += (double)a + (double)b;
    }
    __syncthreads();
  }

  if (!fused) {
    double* dst = sbuf + ((uint64_t)blockIdx.y * Mpad + m) * GT;
#pragma unroll
    for (int g = 0; g < GT; g++) dst[g] = S2[g * CV_THREADS + tid];
    return;
  }
  const bool live = m < M;
  const double cw = live ? coefw[m] : 0.0, lw = live ? lwtr[m] : 0.0;
#pragma unroll
  for (int g = 0; g < GT; g++) {
    const double Sg = S2[g * CV_THREADS + tid];
    S2[g * CV_THREADS + tid] = live ? cw * (log(Sg) - lw) : 0.0;
  }
  __syncthreads();
  block_reduce_store<GT>(S2, partial + (uint64_t)blockIdx.x * GT);
}

template <int GT>
__global__ void __launch_bounds__(CV_THREADS)
    mlcv_finalize_kernel(const double* __restrict__ sbuf, int nsplit, uint64_t M, uint64_t Mpad,
                         const double* __restrict__ coefw, const double* __restrict__ lwtr,
                         double* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char fin_raw[];
  double* S2 = reinterpret_cast<double*>(fin_raw);
  const int tid = threadIdx.x;
  const uint64_t m = (uint64_t)blockIdx.x * CV_THREADS + tid;
  const bool live = m < M;
  const double cw = live ? coefw[m] : 0.0, lw = live ? lwtr[m] : 0.0;
  for (int g = 0; g < GT; g++) {
    double s = 0.0;
    for (int p = 0; p < nsplit; p++) s += sbuf[((uint64_t)p * Mpad + m) * GT + g];
    S2[g * CV_THREADS + tid] = live ? cw * (log(s) - lw) : 0.0;
  }
  __syncthreads();
  block_reduce_store<GT>(S2, partial + (uint64_t)blockIdx.x * GT);
}

static int poly_mode() {
  // KDSB200_MLCV_POLY selects the SFU/FMA split of the exponentials (tuning knob)
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("KDSB200_MLCV_POLY");
    mode = e ? atoi(e) : KDSB_MLCV_POLY_DEFAULT;
  }
  return mode;
}

template <int D, int GT, int PM>
static int launch_mlcv(const float* packed, uint32_t ntl, uint32_t tps, int nsplit,
                       const float* qT, uint64_t M, uint64_t Mpad, const int32_t* lo,
                       const int32_t* hi, const double* coefw, const double* lwtr,
                       const float* h_nk, int G, double* sbuf, double* partial, cudaStream_t st) {
  NkParams<GT> P;
  for (int g = 0; g < GT; g++) P.nk[g] = g < G ? h_nk[g] : -1e30f;  // unused slots -> 2^-inf = 0
  const size_t smem = (size_t)CV_STAGES * (D + 2) * TS * sizeof(float) +
                      (size_t)GT * CV_THREADS * sizeof(double);
  KDSB_CUDA(cudaFuncSetAttribute(mlcv_pairs_kernel<D, GT, PM>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(Mpad / CV_THREADS), (unsigned)nsplit);
  const int fused = nsplit == 1;
  mlcv_pairs_kernel<D, GT, PM><<<grid, CV_THREADS, smem, st>>>(
      packed, ntl, tps, qT, M, Mpad, lo, hi, coefw, lwtr, P, fused, sbuf, partial);
  KDSB_CUDA(cudaGetLastError());
  if (!fused) {
    const size_t fsm = (size_t)GT * CV_THREADS * sizeof(double);
    KDSB_CUDA(cudaFuncSetAttribute(mlcv_finalize_kernel<GT>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
    mlcv_finalize_kernel<GT><<<(unsigned)(Mpad / CV_THREADS), CV_THREADS, fsm, st>>>(
        sbuf, nsplit, M, Mpad, coefw, lwtr, partial);
    KDSB_CUDA(cudaGetLastError());
  }
  return 0;
}

#define KDSB_MLCV_ARGS packed, ntl, tps, nsplit, qT, M, Mpad, lo, hi, coefw, lwtr, h_nk, G, sbuf, partial, st

template <int D>
static int dispatch_gt(int GT, const float* packed, uint32_t ntl, uint32_t tps, int nsplit,
                       const float* qT, uint64_t M, uint64_t Mpad, const int32_t* lo,
                       const int32_t* hi, const double* coefw, const double* lwtr,
                       const float* h_nk, int G, double* sbuf, double* partial, cudaStream_t st) {
  const int pm = poly_mode();
  switch (GT) {
    case 8:
      return launch_mlcv<D, 8, 0>(KDSB_MLCV_ARGS);
    case 10:
      return launch_mlcv<D, 10, 0>(KDSB_MLCV_ARGS);
    case 16:
      return pm ? launch_mlcv<D, 16, KDSB_MLCV_POLY_DEFAULT>(KDSB_MLCV_ARGS)
                : launch_mlcv<D, 16, 0>(KDSB_MLCV_ARGS);
    case 32:
      if (D == 6) {  // tuning variants are only built for the benchmark shape
        if (pm == 2) return launch_mlcv<6, 32, 2>(KDSB_MLCV_ARGS);
        if (pm == 3) return launch_mlcv<6, 32, 3>(KDSB_MLCV_ARGS);
        if (pm == 4) return launch_mlcv<6, 32, 4>(KDSB_MLCV_ARGS);
        if (pm == 38) return launch_mlcv<6, 32, 38>(KDSB_MLCV_ARGS);
      }
      return pm ? launch_mlcv<D, 32, KDSB_MLCV_POLY_DEFAULT>(KDSB_MLCV_ARGS)
                : launch_mlcv<D, 32, 0>(KDSB_MLCV_ARGS);
  }
  set_error("mlcv: unsupported grid tile %d", GT);
  return 1;
}

}  // namespace kdsb

using namespace kdsb;

extern "C" {

int kdsb200_mlcv_gtile(int G) {
  if (G <= 8) return 8;
  if (G <= 10) return 10;
  if (G <= 16) return 16;
  return 32;
}

int kdsb200_mlcv_poly_slots(int G) {
  // grid slots (of kdsb200_mlcv_gtile(G)) whose exponential runs on the FMA pipe
  const int GT = kdsb200_mlcv_gtile(G);
  int pm = poly_mode();
  if (GT < 16 || pm == 0) return 0;
  if (GT == 16) pm = KDSB_MLCV_POLY_DEFAULT;  // only the default pattern is built for 16
  if (pm != 2 && pm != 3 && pm != 4 && pm != 38) pm = KDSB_MLCV_POLY_DEFAULT;
  int n = 0;
  for (int g = 0; g < GT; g++)
    n += pm == 2 ? (g % 2 == 1) : pm == 3 ? (g % 3 == 2) : pm == 4 ? (g % 4 == 3)
                                : (g % 8 == 2 || g % 8 == 5 || g % 8 == 7);
  return n;
}

uint64_t kdsb200_mlcv_mpad(uint64_t M) { return (M + CV_THREADS - 1) / CV_THREADS * CV_THREADS; }

int kdsb200_mlcv_nsplit(uint64_t N, uint64_t M) {
  // Source-dimension split that fills whole waves of 2 CTAs/SM best.  A CTA of the
  // headline shape runs for hundreds of milliseconds, so a ragged last wave costs
  // several percent (3907 row tiles = 13.2 waves -> 14); splitting the sources three
  // ways gives 39.6 -> 40.  The split costs one extra pass over nsplit*M*G doubles.
  const uint64_t qt = kdsb200_mlcv_mpad(M) / CV_THREADS;
  const uint64_t ntl = n_tiles_for(N);
  if (!qt || !ntl) return 1;
  const uint64_t slots = (uint64_t)sm_count() * 2;
  uint64_t lo = (3 * slots + qt - 1) / qt;  // at least ~3 waves when the problem allows
  if (lo < 1) lo = 1;
  if (lo > ntl) lo = ntl;
  uint64_t hi = lo + 7;
  if (hi > ntl) hi = ntl;
  if (hi > 64) hi = lo > 64 ? lo : 64;
  uint64_t best = lo;
  double best_eff = -1.0;
  for (uint64_t ns = lo; ns <= hi; ns++) {
    const uint64_t tps = (ntl + ns - 1) / ns;
    const uint64_t real = (ntl + tps - 1) / tps;  // splits actually launched
    const uint64_t ctas = real * qt;
    const uint64_t waves = (ctas + slots - 1) / slots;
    // useful tile-visits over the capacity of the waves the grid occupies
    const double eff = (double)(qt * ntl) / ((double)waves * slots * tps);
    if (eff > best_eff + 0.004) {  // prefer fewer splits unless the gain is real
      best_eff = eff;
      best = ns;
    }
  }
  return (int)best;
}

int kdsb200_mlcv_scores_f32(const float* d_packed, uint64_t N, int D, const float* d_qT,
                            uint64_t M, uint64_t Mpad, const int32_t* d_excl_lo,
                            const int32_t* d_excl_hi, const double* d_coefw,
                            const double* d_lwtr, const float* h_nk, int G, int nsplit,
                            double* d_sbuf, double* d_partial, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) return 0;
  if (G < 1 || G > 32) {
    set_error("mlcv: G=%d out of range 1..32 per call", G);
    return 1;
  }
  if (Mpad % CV_THREADS != 0 || Mpad < M) {
    set_error("mlcv: Mpad must be a multiple of %d and >= M", CV_THREADS);
    return 1;
  }
  if (N >= (1ull << 31)) {
    set_error("mlcv: N too large for int32 fold intervals");
    return 1;
  }
  const uint32_t ntl = (uint32_t)n_tiles_for(N);
  if (nsplit < 1) nsplit = 1;
  if ((uint32_t)nsplit > ntl && ntl) nsplit = (int)ntl;
  const uint32_t tps = ntl ? (ntl + nsplit - 1) / nsplit : 0;
  if (ntl) nsplit = (int)((ntl + tps - 1) / tps);
  if (nsplit > 1 && !d_sbuf) {
    set_error("mlcv: sbuf required for nsplit=%d", nsplit);
    return 1;
  }
  const int GT = kdsb200_mlcv_gtile(G);
  switch (D) {
#define KDSB_CASE(d)                                                                          \
  case d:                                                                                     \
    return dispatch_gt<d>(GT, d_packed, ntl, tps, nsplit, d_qT, M, Mpad, d_excl_lo, d_excl_hi, \
                          d_coefw, d_lwtr, h_nk, G, d_sbuf, d_partial, st);
    KDSB_CASE(1)
    KDSB_CASE(2)
    KDSB_CASE(3)
    KDSB_CASE(4)
    KDSB_CASE(5)
    KDSB_CASE(6)
    KDSB_CASE(7)
    KDSB_CASE(8)
#undef KDSB_CASE
  }
  set_error("mlcv: D=%d out of range 1..8", D);
  return 1;
}

}  // extern "C"
