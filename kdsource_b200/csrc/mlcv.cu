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
// ex2.approx.ftz returns 0 below 2^-126, so a test point whose nearest training source
// is 13 < r/h < 38 bandwidths away gets a sum that is too small or zero, where fp64
// (the reference) still has a finite logarithm.  The coefficients are scaled so that
// the largest is 1 and at most 2^31 terms of < 2^-126 can have been lost, hence a sum
// above 2^-70 is accurate to 2^-25; anything below is flagged (bit g of flagmask[row])
// and recomputed in fp64 by mlcv_rows_f64_kernel.
constexpr double CV_FLAG_BELOW = 8.470329472543003e-22;  // 2^-70

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
                      NkParams<GT> P, int G, int fused, double* __restrict__ sbuf,
                      double* __restrict__ partial, uint32_t* __restrict__ flagmask) {
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
  uint32_t mask = 0;
#pragma unroll
  for (int g = 0; g < GT; g++) {
    const double Sg = S2[g * CV_THREADS + tid];
    const bool bad = flagmask && live && g < G && !(Sg >= CV_FLAG_BELOW);
    if (bad) mask |= 1u << g;
    S2[g * CV_THREADS + tid] = (live && !bad) ? cw * (log(Sg) - lw) : 0.0;
  }
  if (flagmask) flagmask[m] = mask;
  __syncthreads();
  block_reduce_store<GT>(S2, partial + (uint64_t)blockIdx.x * GT);
}

template <int GT>
__global__ void __launch_bounds__(CV_THREADS)
    mlcv_finalize_kernel(const double* __restrict__ sbuf, int nsplit, uint64_t M, uint64_t Mpad,
                         const double* __restrict__ coefw, const double* __restrict__ lwtr, int G,
                         double* __restrict__ partial, uint32_t* __restrict__ flagmask) {
  extern __shared__ __align__(16) unsigned char fin_raw[];
  double* S2 = reinterpret_cast<double*>(fin_raw);
  const int tid = threadIdx.x;
  const uint64_t m = (uint64_t)blockIdx.x * CV_THREADS + tid;
  const bool live = m < M;
  const double cw = live ? coefw[m] : 0.0, lw = live ? lwtr[m] : 0.0;
  uint32_t mask = 0;
  for (int g = 0; g < GT; g++) {
    double s = 0.0;
    for (int p = 0; p < nsplit; p++) s += sbuf[((uint64_t)p * Mpad + m) * GT + g];
    const bool bad = flagmask && live && g < G && !(s >= CV_FLAG_BELOW);
    if (bad) mask |= 1u << g;
    S2[g * CV_THREADS + tid] = (live && !bad) ? cw * (log(s) - lw) : 0.0;
  }
  if (flagmask) flagmask[m] = mask;
  __syncthreads();
  block_reduce_store<GT>(S2, partial + (uint64_t)blockIdx.x * GT);
}

// ---------------------------------------------------------------------------
// fp64 rows: the 1e-12 mode of the search, and the recomputation of the rows the fp32
// kernel flagged.  One warp per test row (lanes stride over the sources, G fp64
// accumulators per lane, fixed-order shuffle reduction), 8 rows per CTA sharing the
// source tiles staged in shared memory.  rowmask == NULL: every row, every g;
// otherwise only the (row, g) pairs whose bit is set contribute (the others add 0).
// ---------------------------------------------------------------------------
constexpr int R64_WARPS = 8;
constexpr int R64_THREADS = R64_WARPS * 32;
constexpr int R64_TILE = 128;

template <int GT>
struct GridParams {
  double inv_g2[GT];  // 1 / grid[g]^2
};

// Rows flagged by the fp32 kernel are usually a handful of isolated tail points, and one
// warp walking the whole list for each of them would take ~0.1 ms per 1000 sources.  The
// row blocks that hold flagged rows therefore get a slot each (ordered, so the result does
// not depend on scheduling); the first R64_CAP of them are recomputed by R64_NSPLIT CTAs
// per block, each over a slice of the sources, and combined in a fixed order.
constexpr int R64_CAP = 256;
constexpr int R64_NSPLIT = 64;

// slots[b] = rank of row block b among the blocks with a flagged row (-1: none);
// slot_block[slot] = b for slot < R64_CAP; *count = number of such blocks.  One CTA.
__global__ void __launch_bounds__(256)
    mlcv_assign_slots_kernel(const uint32_t* __restrict__ rowmask, uint64_t M, int G,
                             uint64_t nblocks, int32_t* __restrict__ slots,
                             int32_t* __restrict__ slot_block, int32_t* __restrict__ count) {
  __shared__ uint32_t wcnt[8];
  __shared__ uint32_t run;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gm = G < 32 ? (1u << G) - 1u : 0xffffffffu;
  if (threadIdx.x == 0) run = 0;
  __syncthreads();
  for (uint64_t base = 0; base < nblocks; base += 256) {
    const uint64_t b = base + threadIdx.x;
    bool any = false;
    if (b < nblocks)
      for (int r = 0; r < R64_WARPS; r++) {
        const uint64_t m = b * R64_WARPS + r;
        if (m < M && (rowmask[m] & gm)) any = true;
      }
    const uint32_t bal = __ballot_sync(0xffffffffu, any);
    if (lane == 0) wcnt[warp] = __popc(bal);
    __syncthreads();
    uint32_t off = run;
    for (int w = 0; w < warp; w++) off += wcnt[w];
    if (b < nblocks) {
      const int32_t slot = any ? (int32_t)(off + __popc(bal & ((1u << lane) - 1u))) : -1;
      slots[b] = slot;
      if (slot >= 0 && slot < R64_CAP) slot_block[slot] = (int32_t)b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
      for (int w = 0; w < 8; w++) tot += wcnt[w];
      run += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = (int32_t)run;
}

// MODE 0: every row block b = blockIdx.x over the whole list (the fp64 mode, and in the
//         recomputation the blocks beyond R64_CAP; blocks with a slot below the cap are
//         left to MODE 1; blocks without flagged rows write zeros).
// MODE 1: blockIdx.x = slot, blockIdx.y = source slice; raw partial sums go to sfix.
template <int D, int GT, int MODE>
__global__ void __launch_bounds__(R64_THREADS)
    mlcv_rows_f64_kernel(const double* __restrict__ data, const double* __restrict__ coef,
                         const double* __restrict__ ih2, uint64_t N, uint64_t row0, uint64_t M,
                         const int32_t* __restrict__ excl_lo, const int32_t* __restrict__ excl_hi,
                         const double* __restrict__ coefw, const double* __restrict__ lwtr,
                         GridParams<GT> P, int G, double log_shift,
                         const uint32_t* __restrict__ rowmask, const int32_t* __restrict__ slots,
                         const int32_t* __restrict__ slot_block,
                         const int32_t* __restrict__ slot_count, double* __restrict__ sfix,
                         double* __restrict__ partial) {
  __shared__ double sx[D * R64_TILE];
  __shared__ double sc[R64_TILE];
  __shared__ double sh[R64_TILE];
  __shared__ double contrib[R64_WARPS][GT];
  __shared__ int any_live;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint64_t blk = blockIdx.x;
  uint64_t j0 = 0, j1 = N;
  if (MODE == 1) {
    const int cnt = min(*slot_count, R64_CAP);
    if ((int)blockIdx.x >= cnt) return;
    blk = (uint64_t)slot_block[blockIdx.x];
    const uint64_t per = (N + R64_NSPLIT - 1) / R64_NSPLIT;
    j0 = min(N, (uint64_t)blockIdx.y * per);
    j1 = min(N, j0 + per);
  }
  const uint64_t m = blk * R64_WARPS + warp;
  uint32_t mask = 0;
  if (m < M) mask = rowmask ? rowmask[m] : 0xffffffffu;
  if (G < 32) mask &= (1u << G) - 1u;
  if (MODE == 0) {
    if (threadIdx.x == 0) any_live = 0;
    __syncthreads();
    if (mask && lane == 0) any_live = 1;
    __syncthreads();
    if (!any_live) {  // CTA-uniform: nothing flagged among these rows
      if (threadIdx.x < GT) partial[blk * GT + threadIdx.x] = 0.0;
      return;
    }
    if (slots && slots[blk] < R64_CAP) return;  // recomputed by the sliced launch
  }
  double q[D];
#pragma unroll
  for (int k = 0; k < D; k++) q[k] = mask ? data[(row0 + m) * D + k] : 0.0;
  const int64_t lo = mask ? excl_lo[m] : 0, hi = mask ? excl_hi[m] : 0;
  double acc[GT];
#pragma unroll
  for (int g = 0; g < GT; g++) acc[g] = 0.0;

  for (uint64_t base = j0; base < j1; base += R64_TILE) {
    const int cnt = (int)min((uint64_t)R64_TILE, j1 - base);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * D; e += R64_THREADS)
      sx[(e % D) * R64_TILE + e / D] = data[base * D + e];
    for (int e = threadIdx.x; e < cnt; e += R64_THREADS) {
      sc[e] = coef[base + e];
      sh[e] = ih2[base + e];
    }
    __syncthreads();
    if (!mask) continue;
    for (int j = lane; j < cnt; j += 32) {
      const int64_t gj = (int64_t)(base + j);
      if (gj >= lo && gj < hi) continue;  // the row's own fold
      double r2 = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) {
        const double t = q[k] - sx[k * R64_TILE + j];
        r2 += t * t;
      }
      const double u = r2 * sh[j], c = sc[j];
#pragma unroll
      for (int g = 0; g < GT; g++)  // only the grid points this row needs (warp-uniform)
        if ((mask >> g) & 1u) acc[g] += c * exp(-u * P.inv_g2[g]);
    }
  }
  const double cw = mask ? coefw[m] : 0.0, lw = mask ? lwtr[m] : 0.0;
#pragma unroll
  for (int g = 0; g < GT; g++) {
    double v = acc[g];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) {
      if (MODE == 1)
        sfix[(((uint64_t)blockIdx.x * R64_NSPLIT + blockIdx.y) * R64_WARPS + warp) * GT + g] = v;
      else
        contrib[warp][g] = ((mask >> g) & 1u) ? cw * (log(v) + log_shift - lw) : 0.0;
    }
  }
  if (MODE == 1) return;
  __syncthreads();
  if (threadIdx.x < GT) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < R64_WARPS; w++) t += contrib[w][threadIdx.x];
    partial[blk * GT + threadIdx.x] = t;
  }
}

// sums the source slices of the sliced launch in slice order and writes the row blocks'
// contributions; grid = R64_CAP CTAs of R64_THREADS threads
template <int GT>
__global__ void __launch_bounds__(R64_THREADS)
    mlcv_combine_slices_kernel(const double* __restrict__ sfix, const int32_t* __restrict__ slot_block,
                               const int32_t* __restrict__ slot_count,
                               const uint32_t* __restrict__ rowmask, uint64_t M, int G,
                               const double* __restrict__ coefw, const double* __restrict__ lwtr,
                               double log_shift, double* __restrict__ partial) {
  __shared__ double contrib[R64_WARPS][GT];
  const int cnt = min(*slot_count, R64_CAP);
  if ((int)blockIdx.x >= cnt) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t blk = (uint64_t)slot_block[blockIdx.x];
  const uint64_t m = blk * R64_WARPS + warp;
  uint32_t mask = m < M ? rowmask[m] : 0u;
  if (G < 32) mask &= (1u << G) - 1u;
  const double cw = mask ? coefw[m] : 0.0, lw = mask ? lwtr[m] : 0.0;
  for (int g = lane; g < GT; g += 32) {
    double S = 0.0;
    for (int sl = 0; sl < R64_NSPLIT; sl++)
      S += sfix[(((uint64_t)blockIdx.x * R64_NSPLIT + sl) * R64_WARPS + warp) * GT + g];
    contrib[warp][g] = ((mask >> g) & 1u) ? cw * (log(S) + log_shift - lw) : 0.0;
  }
  __syncthreads();
  if (threadIdx.x < GT) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < R64_WARPS; w++) t += contrib[w][threadIdx.x];
    partial[blk * GT + threadIdx.x] = t;
  }
}

// coef_j = w_j * w_scale * h_j^-D, ih2_j = 1 / (2 h_j^2)
__global__ void mlcv_coef_f64_kernel(const double* __restrict__ w, const double* __restrict__ bw,
                                     double bw_scalar, uint64_t N, int D, double w_scale,
                                     double* __restrict__ coef, double* __restrict__ ih2) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double h = bw ? bw[i] : bw_scalar;
  coef[i] = (w ? w[i] : 1.0) * w_scale * exp(-D * log(h));
  ih2[i] = 0.5 / (h * h);
}

// lo/hi: fold interval of each test row; coefw = w_i * cwf[f] (cwf[f] = 1/(n_f K));
// lwtr = lwt[f] (log of the training weight of fold f).  Padding rows get an empty
// interval and zero coefficient.
__global__ void mlcv_rows_setup_kernel(uint64_t N, uint64_t K, uint64_t row0, uint64_t M,
                                       uint64_t Mpad, const double* __restrict__ w,
                                       const double* __restrict__ cwf,
                                       const double* __restrict__ lwt, int32_t* __restrict__ lo,
                                       int32_t* __restrict__ hi, double* __restrict__ coefw,
                                       double* __restrict__ lwtr) {
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mpad) return;
  if (m >= M) {
    lo[m] = 0;
    hi[m] = 0;
    coefw[m] = 0.0;
    lwtr[m] = 0.0;
    return;
  }
  const uint64_t i = row0 + m, base = N / K, big = N % K, cut = big * (base + 1);
  uint64_t f, a;
  if (i < cut) {
    f = i / (base + 1);
    a = f * (base + 1);
    lo[m] = (int32_t)a;
    hi[m] = (int32_t)(a + base + 1);
  } else {
    f = big + (i - cut) / base;
    a = cut + (f - big) * base;
    lo[m] = (int32_t)a;
    hi[m] = (int32_t)(a + base);
  }
  coefw[m] = (w ? w[i] : 1.0) * cwf[f];
  lwtr[m] = lwt[f];
}

// out[g] = sum over the rows of a[na][gt_a] and b[nb][gt_b], g < G, in a fixed order:
// RED_CTAS CTAs each sum a contiguous share of the rows into tmp[cta][32], then one CTA
// adds the shares in order
constexpr int RED_CTAS = 128;
__global__ void __launch_bounds__(256)
    mlcv_reduce_stage1_kernel(const double* __restrict__ a, uint64_t na, int gt_a,
                              const double* __restrict__ b, uint64_t nb, int gt_b, int G,
                              double* __restrict__ tmp) {
  __shared__ double sm[256];
  const uint64_t pa = (na + RED_CTAS - 1) / RED_CTAS, pb = (nb + RED_CTAS - 1) / RED_CTAS;
  const uint64_t a0 = min(na, blockIdx.x * pa), a1 = min(na, a0 + pa);
  const uint64_t b0 = min(nb, blockIdx.x * pb), b1 = min(nb, b0 + pb);
  for (int g = 0; g < G; g++) {
    double s = 0.0;
    for (uint64_t r = a0 + threadIdx.x; r < a1; r += 256) s += a[r * gt_a + g];
    if (b)
      for (uint64_t r = b0 + threadIdx.x; r < b1; r += 256) s += b[r * gt_b + g];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
      if ((int)threadIdx.x < off) sm[threadIdx.x] += sm[threadIdx.x + off];
      __syncthreads();
    }
    if (threadIdx.x == 0) tmp[blockIdx.x * 32 + g] = sm[0];
    __syncthreads();
  }
}

__global__ void mlcv_reduce_stage2_kernel(const double* __restrict__ tmp, int G,
                                          double* __restrict__ out) {
  const int g = threadIdx.x;
  if (g >= G) return;
  double s = 0.0;
  for (int c = 0; c < RED_CTAS; c++) s += tmp[c * 32 + g];
  out[g] = s;
}

template <int D, int GT>
static int launch_rows_f64(const double* data, const double* coef, const double* ih2, uint64_t N,
                           uint64_t row0, uint64_t M, const int32_t* lo, const int32_t* hi,
                           const double* coefw, const double* lwtr, const double* h_grid, int G,
                           double log_shift, const uint32_t* rowmask, void* fix_scratch,
                           double* partial, cudaStream_t st) {
  GridParams<GT> P;
  for (int g = 0; g < GT; g++) P.inv_g2[g] = g < G ? 1.0 / (h_grid[g] * h_grid[g]) : 0.0;
  const unsigned blocks = (unsigned)((M + R64_WARPS - 1) / R64_WARPS);
  if (!rowmask) {  // fp64 mode: every row, one warp each
    mlcv_rows_f64_kernel<D, GT, 0><<<blocks, R64_THREADS, 0, st>>>(
        data, coef, ih2, N, row0, M, lo, hi, coefw, lwtr, P, G, log_shift, nullptr, nullptr,
        nullptr, nullptr, nullptr, partial);
    KDSB_CUDA(cudaGetLastError());
    return 0;
  }
  // recomputation of flagged rows: slots, sliced launch for the first R64_CAP row blocks,
  // whole-list launch for the rest (and the zero fill), combine
  unsigned char* p = (unsigned char*)fix_scratch;
  int32_t* count = (int32_t*)p;
  int32_t* slot_block = (int32_t*)(p + 64);
  int32_t* slots = slot_block + R64_CAP;
  double* sfix = (double*)(p + 64 + ((((uint64_t)R64_CAP + blocks) * 4 + 63) / 64) * 64);
  mlcv_assign_slots_kernel<<<1, 256, 0, st>>>(rowmask, M, G, blocks, slots, slot_block, count);
  KDSB_CUDA(cudaGetLastError());
  mlcv_rows_f64_kernel<D, GT, 0><<<blocks, R64_THREADS, 0, st>>>(
      data, coef, ih2, N, row0, M, lo, hi, coefw, lwtr, P, G, log_shift, rowmask, slots,
      slot_block, count, sfix, partial);
  KDSB_CUDA(cudaGetLastError());
  mlcv_rows_f64_kernel<D, GT, 1><<<dim3(R64_CAP, R64_NSPLIT), R64_THREADS, 0, st>>>(
      data, coef, ih2, N, row0, M, lo, hi, coefw, lwtr, P, G, log_shift, rowmask, slots,
      slot_block, count, sfix, partial);
  KDSB_CUDA(cudaGetLastError());
  mlcv_combine_slices_kernel<GT><<<R64_CAP, R64_THREADS, 0, st>>>(
      sfix, slot_block, count, rowmask, M, G, coefw, lwtr, log_shift, partial);
  KDSB_CUDA(cudaGetLastError());
  return 0;
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
                       const float* h_nk, int G, double* sbuf, double* partial, uint32_t* flagmask,
                       cudaStream_t st) {
  NkParams<GT> P;
  for (int g = 0; g < GT; g++) P.nk[g] = g < G ? h_nk[g] : -1e30f;  // unused slots -> 2^-inf = 0
  const size_t smem = (size_t)CV_STAGES * (D + 2) * TS * sizeof(float) +
                      (size_t)GT * CV_THREADS * sizeof(double);
  KDSB_CUDA(cudaFuncSetAttribute(mlcv_pairs_kernel<D, GT, PM>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(Mpad / CV_THREADS), (unsigned)nsplit);
  const int fused = nsplit == 1;
  mlcv_pairs_kernel<D, GT, PM><<<grid, CV_THREADS, smem, st>>>(
      packed, ntl, tps, qT, M, Mpad, lo, hi, coefw, lwtr, P, G, fused, sbuf, partial, flagmask);
  KDSB_CUDA(cudaGetLastError());
  if (!fused) {
    const size_t fsm = (size_t)GT * CV_THREADS * sizeof(double);
    KDSB_CUDA(cudaFuncSetAttribute(mlcv_finalize_kernel<GT>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsm));
    mlcv_finalize_kernel<GT><<<(unsigned)(Mpad / CV_THREADS), CV_THREADS, fsm, st>>>(
        sbuf, nsplit, M, Mpad, coefw, lwtr, G, partial, flagmask);
    KDSB_CUDA(cudaGetLastError());
  }
  return 0;
}

#define KDSB_MLCV_ARGS packed, ntl, tps, nsplit, qT, M, Mpad, lo, hi, coefw, lwtr, h_nk, G, sbuf, partial, flagmask, st

template <int D>
static int dispatch_gt(int GT, const float* packed, uint32_t ntl, uint32_t tps, int nsplit,
                       const float* qT, uint64_t M, uint64_t Mpad, const int32_t* lo,
                       const int32_t* hi, const double* coefw, const double* lwtr,
                       const float* h_nk, int G, double* sbuf, double* partial, uint32_t* flagmask,
                       cudaStream_t st) {
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
  // CTAs differ in length (tiles inside a row block's own fold are skipped, and a rank's
  // share of the rows sits in one or two folds), so the last wave is only as ragged as one
  // CTA is long: aim for >= 32 CTAs per slot while a CTA keeps >= 16 tiles (a tile is
  // ~0.2 ms of work against microseconds of pipeline fill)
  uint64_t lo = (32 * slots + qt - 1) / qt;
  if (lo > ntl / 16) lo = ntl / 16;
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
                            double* d_sbuf, double* d_partial, uint32_t* d_flagmask,
                            void* stream) {
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
                          d_coefw, d_lwtr, h_nk, G, d_sbuf, d_partial, d_flagmask, st);
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

uint64_t kdsb200_mlcv_f64_blocks(uint64_t M) { return (M + R64_WARPS - 1) / R64_WARPS; }

int kdsb200_mlcv_coef_f64(const double* d_w, const double* d_bw, double bw_scalar, uint64_t N,
                          int D, double w_scale, double* d_coef2n, void* stream) {
  if (N == 0) return 0;
  mlcv_coef_f64_kernel<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_w, d_bw, bw_scalar, N, D, w_scale, d_coef2n, d_coef2n + N);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

uint64_t kdsb200_mlcv_fix_scratch_bytes(uint64_t M) {
  const uint64_t blocks = (M + R64_WARPS - 1) / R64_WARPS;
  return 64 + (((uint64_t)R64_CAP + blocks) * 4 + 63) / 64 * 64 +
         (uint64_t)R64_CAP * R64_NSPLIT * R64_WARPS * 32 * sizeof(double);
}

int kdsb200_mlcv_rows_f64(const double* d_data, const double* d_coef2n, uint64_t N, int D,
                          uint64_t row0, uint64_t M, const int32_t* d_excl_lo,
                          const int32_t* d_excl_hi, const double* d_coefw, const double* d_lwtr,
                          const double* h_grid, int G, double log_shift,
                          const uint32_t* d_rowmask, void* d_fix_scratch, double* d_partial,
                          void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0 || N == 0) return 0;
  if (G < 1 || G > 32) {
    set_error("mlcv_rows_f64: G=%d out of range 1..32 per call", G);
    return 1;
  }
  if (row0 + M > N) {
    set_error("mlcv_rows_f64: rows [%llu, %llu) outside the list", (unsigned long long)row0,
              (unsigned long long)(row0 + M));
    return 1;
  }
  if (d_rowmask && !d_fix_scratch) {
    set_error("mlcv_rows_f64: the recomputation of flagged rows needs its scratch buffer");
    return 1;
  }
  const double* coef = d_coef2n;
  const double* ih2 = d_coef2n + N;
  const int GT = G <= 8 ? 8 : G <= 16 ? 16 : 32;
#define KDSB_R64(d, gt)                                                                        \
  return launch_rows_f64<d, gt>(d_data, coef, ih2, N, row0, M, d_excl_lo, d_excl_hi, d_coefw, \
                                d_lwtr, h_grid, G, log_shift, d_rowmask, d_fix_scratch,       \
                                d_partial, st)
#define KDSB_CASE(d)                 \
  case d:                            \
    if (GT == 8) KDSB_R64(d, 8);     \
    if (GT == 16) KDSB_R64(d, 16);   \
    KDSB_R64(d, 32);
  switch (D) {
    KDSB_CASE(1)
    KDSB_CASE(2)
    KDSB_CASE(3)
    KDSB_CASE(4)
    KDSB_CASE(5)
    KDSB_CASE(6)
    KDSB_CASE(7)
    KDSB_CASE(8)
  }
#undef KDSB_CASE
#undef KDSB_R64
  set_error("mlcv_rows_f64: D=%d out of range 1..8", D);
  return 1;
}

// fold bookkeeping of the test rows [row0, row0+M) on the device: contiguous KFold
// without shuffle (kde.py:134-136), the first N % K folds one longer
int kdsb200_mlcv_rows_setup(uint64_t N, uint64_t K, uint64_t row0, uint64_t M, uint64_t Mpad,
                            const double* d_w, const double* d_cwf, const double* d_lwt,
                            int32_t* d_lo, int32_t* d_hi, double* d_coefw, double* d_lwtr,
                            void* stream) {
  if (Mpad == 0) return 0;
  if (K < 1 || K > N || N >= (1ull << 31)) {
    set_error("mlcv_rows_setup: need 1 <= K <= N < 2^31");
    return 1;
  }
  mlcv_rows_setup_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      N, K, row0, M, Mpad, d_w, d_cwf, d_lwt, d_lo, d_hi, d_coefw, d_lwtr);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

int kdsb200_mlcv_f64_gtile(int G) { return G <= 8 ? 8 : G <= 16 ? 16 : 32; }

uint64_t kdsb200_mlcv_reduce_scratch(void) { return (uint64_t)RED_CTAS * 32; }

int kdsb200_mlcv_reduce(const double* d_a, uint64_t na, int gt_a, const double* d_b, uint64_t nb,
                        int gt_b, int G, double* d_tmp, double* d_out, void* stream) {
  if (G < 1 || G > 32 || G > gt_a || (d_b && G > gt_b)) {
    set_error("mlcv_reduce: G=%d does not fit the partial rows (%d, %d)", G, gt_a, gt_b);
    return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  mlcv_reduce_stage1_kernel<<<RED_CTAS, 256, 0, st>>>(d_a, na, gt_a, d_b, nb, gt_b, G, d_tmp);
  KDSB_CUDA(cudaGetLastError());
  mlcv_reduce_stage2_kernel<<<1, 32, 0, st>>>(d_tmp, G, d_out);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
