// kde_eval: dense Gaussian kernel sum  out[m] = sum_i c_i exp(-|q_m-x_i|^2/(2 h_i^2))
// Replaces KDEpy TreeKDE.evaluate as called from python/kdsource/kdsource.py:239
// (and the per-fold evaluate of kde.py:146-151).  FP32-pipe bound: (2D+1) packed
// FP32 lane-ops + 1 MUFU.EX2 per pair.
//
// Layout: sources are pre-packed (pack_sources_kernel) into tiles of TS=512
// sources x (D+2) rows so that one tile is one contiguous 1-D bulk-TMA copy.
// Each thread keeps Q query points in registers and walks the tile four sources
// at a time with broadcast LDS.128; pairs of sources ride in one f32x2 lane pair
// (FFMA2/FADD2), the query coordinate is a broadcast scalar operand.
#include <math.h>

#include "common.cuh"
#include "kdsb200.h"

namespace kdsb {

struct Center {
  double c[8];
};

// ---------------------------------------------------------------------------
// packing kernels (fp64 in, fp32 tiles out)
// ---------------------------------------------------------------------------
// mode 0 (kde):  a_i = sqrt(log2(e)/2)/h_i ; last row = lc_ref - log2(w_i*w_scale*h_i^-D)
// mode 1 (mlcv): a_i = 1/h_i               ; last row = w_i*w_scale*h_i^-D  (linear)
__global__ void pack_sources_kernel(const double* __restrict__ data, const double* __restrict__ w,
                                    const double* __restrict__ bw, double bw_scalar, uint64_t N,
                                    int D, Center cen, double w_scale, double lc_ref, int mode,
                                    float* __restrict__ packed) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t ntl = n_tiles_for(N);
  if (i >= ntl * TS) return;
  const int rows = D + 2;
  float* tile = packed + (i / TS) * (uint64_t)rows * TS + (i % TS);
  if (i >= N) {
    for (int k = 0; k <= D; k++) tile[(uint64_t)k * TS] = 0.f;
    tile[(uint64_t)(D + 1) * TS] = mode == 0 ? INFINITY : 0.f;
    return;
  }
  const double h = bw ? bw[i] : bw_scalar;
  const double wi = (w ? w[i] : 1.0) * w_scale;
  const double a = (mode == 0 ? 0.84932180028801904272 /* sqrt(log2(e)/2) */ : 1.0) / h;
  for (int k = 0; k < D; k++)
    tile[(uint64_t)k * TS] = (float)(-(data[i * D + k] - cen.c[k]) * a);
  tile[(uint64_t)D * TS] = (float)a;
  if (mode == 0) {
    const double lc = log2(wi) - D * log2(h);  // -inf for w == 0 -> nlc = +inf
    tile[(uint64_t)(D + 1) * TS] = (float)(lc_ref - lc);
  } else {
    tile[(uint64_t)(D + 1) * TS] = (float)(wi * exp2(-D * log2(h) - lc_ref));
  }
}

__global__ void pack_queries_kernel(const double* __restrict__ pts, uint64_t M, uint64_t Mpad,
                                    int D, Center cen, float* __restrict__ qT) {
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mpad) return;
  for (int k = 0; k < D; k++)
    qT[(uint64_t)k * Mpad + m] = m < M ? (float)(pts[m * D + k] - cen.c[k]) : 0.f;
}

// ---------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------
constexpr int EV_THREADS = 256;
constexpr int EV_Q = 4;
constexpr int EV_STAGES = 3;

// KERN: 0 Gaussian (rows as above); 1 Epanechnikov, 2 box -- KDEpy's radially
// symmetric finite-support kernels (kernel names of python/kdsource/kdsource.py:60-65):
// sources packed with mode 1 and h replaced by the support radius R_i, so u = |q-x|^2/R_i^2
// and the last row is the linear coefficient c_i; epa adds c_i*max(0, 1-u), box adds c_i
// where u < 1.
template <int D, bool CUT, int KERN>
__global__ void __launch_bounds__(EV_THREADS, 2)
    kde_eval_kernel(const float* __restrict__ packed, uint32_t n_tiles, uint32_t tiles_per_split,
                    const float* __restrict__ qT, uint64_t Mpad, float cut2, double out_scale,
                    double* __restrict__ out) {
  constexpr int ROWS = D + 2;
  constexpr int TILE_FLOATS = ROWS * TS;
  constexpr uint32_t TILE_BYTES = TILE_FLOATS * sizeof(float);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t full[EV_STAGES];

  const int tid = threadIdx.x;
  const uint32_t t0 = blockIdx.y * tiles_per_split;
  const uint32_t t1 = min(n_tiles, t0 + tiles_per_split);
  const uint32_t nt = t1 > t0 ? t1 - t0 : 0;

  if (tid == 0) {
    for (int s = 0; s < EV_STAGES; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (uint32_t p = 0; p < (uint32_t)(EV_STAGES - 1) && p < nt; p++) {
      mbar_expect_tx(&full[p], TILE_BYTES);
      tma_load_1d(tiles + p * TILE_FLOATS, packed + (uint64_t)(t0 + p) * TILE_FLOATS, TILE_BYTES,
                  &full[p]);
    }
  }

  // query points of this thread (coordinate-major, coalesced)
  const uint64_t mbase = (uint64_t)blockIdx.x * (EV_THREADS * EV_Q) + tid;
  float q[EV_Q][D];
#pragma unroll
  for (int qi = 0; qi < EV_Q; qi++)
#pragma unroll
    for (int k = 0; k < D; k++) q[qi][k] = qT[(uint64_t)k * Mpad + mbase + qi * EV_THREADS];

  double accd[EV_Q];
#pragma unroll
  for (int qi = 0; qi < EV_Q; qi++) accd[qi] = 0.0;

  for (uint32_t it = 0; it < nt; it++) {
    const uint32_t s = it % EV_STAGES;
    const uint32_t parity = (it / EV_STAGES) & 1u;
    if (tid == 0 && it + EV_STAGES - 1 < nt) {
      const uint32_t ps = (it + EV_STAGES - 1) % EV_STAGES;
      mbar_expect_tx(&full[ps], TILE_BYTES);
      tma_load_1d(tiles + ps * TILE_FLOATS,
                  packed + (uint64_t)(t0 + it + EV_STAGES - 1) * TILE_FLOATS, TILE_BYTES,
                  &full[ps]);
    }
    mbar_wait(&full[s], parity);
    const float* tile = tiles + s * TILE_FLOATS;

    f2_t acc[EV_Q];
#pragma unroll
    for (int qi = 0; qi < EV_Q; qi++) acc[qi] = 0ull;

#pragma unroll 2
    for (int j = 0; j < TS; j += 4) {
      float4 r[ROWS];
#pragma unroll
      for (int k = 0; k < ROWS; k++) r[k] = *reinterpret_cast<const float4*>(tile + k * TS + j);
      const f2_t a01 = f2_pack(r[D].x, r[D].y), a23 = f2_pack(r[D].z, r[D].w);
      const f2_t l01 = f2_pack(r[D + 1].x, r[D + 1].y), l23 = f2_pack(r[D + 1].z, r[D + 1].w);
      f2_t c01 = 0ull, c23 = 0ull;
      if (CUT) {
        const f2_t cc = f2_pack(cut2, cut2);
        c01 = f2_mul(f2_mul(a01, a01), cc);
        c23 = f2_mul(f2_mul(a23, a23), cc);
      }
#pragma unroll
      for (int qi = 0; qi < EV_Q; qi++) {
        f2_t tA = (CUT || KERN) ? 0ull : l01, tB = (CUT || KERN) ? 0ull : l23;
#pragma unroll
        for (int k = 0; k < D; k++) {
          const f2_t qq = f2_pack(q[qi][k], q[qi][k]);
          const f2_t dA = f2_fma(qq, a01, f2_pack(r[k].x, r[k].y));
          const f2_t dB = f2_fma(qq, a23, f2_pack(r[k].z, r[k].w));
          tA = f2_fma(dA, dA, tA);
          tB = f2_fma(dB, dB, tB);
        }
        float eA0, eA1, eB0, eB1;
        if (KERN) {
          float u0, u1, u2, u3, g0, g1, g2, g3;
          f2_unpack(tA, u0, u1);
          f2_unpack(tB, u2, u3);
          f2_unpack(l01, g0, g1);
          f2_unpack(l23, g2, g3);
          if (KERN == 1) {
            eA0 = g0 * fmaxf(1.f - u0, 0.f);
            eA1 = g1 * fmaxf(1.f - u1, 0.f);
            eB0 = g2 * fmaxf(1.f - u2, 0.f);
            eB1 = g3 * fmaxf(1.f - u3, 0.f);
          } else {
            eA0 = u0 < 1.f ? g0 : 0.f;
            eA1 = u1 < 1.f ? g1 : 0.f;
            eB0 = u2 < 1.f ? g2 : 0.f;
            eB1 = u3 < 1.f ? g3 : 0.f;
          }
        } else if (CUT) {
          float u0, u1, u2, u3, m0, m1, m2, m3, g0, g1, g2, g3;
          f2_unpack(tA, u0, u1);
          f2_unpack(tB, u2, u3);
          f2_unpack(c01, m0, m1);
          f2_unpack(c23, m2, m3);
          f2_unpack(l01, g0, g1);
          f2_unpack(l23, g2, g3);
          eA0 = u0 <= m0 ? ex2_approx(-(u0 + g0)) : 0.f;
          eA1 = u1 <= m1 ? ex2_approx(-(u1 + g1)) : 0.f;
          eB0 = u2 <= m2 ? ex2_approx(-(u2 + g2)) : 0.f;
          eB1 = u3 <= m3 ? ex2_approx(-(u3 + g3)) : 0.f;
        } else {
          float u0, u1, u2, u3;
          f2_unpack(tA, u0, u1);
          f2_unpack(tB, u2, u3);
          eA0 = ex2_approx(-u0);
          eA1 = ex2_approx(-u1);
          eB0 = ex2_approx(-u2);
          eB1 = ex2_approx(-u3);
        }
        acc[qi] = f2_add(acc[qi], f2_add(f2_pack(eA0, eA1), f2_pack(eB0, eB1)));
      }
    }
#pragma unroll
    for (int qi = 0; qi < EV_Q; qi++) {
      float lo, hi;
      f2_unpack(acc[qi], lo, hi);
      accd[qi] += (double)lo + (double)hi;
    }
    __syncthreads();  // every warp is done with stage s before it is refilled
  }
#pragma unroll
  for (int qi = 0; qi < EV_Q; qi++)
    out[(uint64_t)blockIdx.y * Mpad + mbase + qi * EV_THREADS] = accd[qi] * out_scale;
}

// ---------------------------------------------------------------------------
// split-coordinate variant ("float32x2"): fp32 speed class, accuracy independent of h
// ---------------------------------------------------------------------------
// In the plain fp32 kernel a coordinate carries the rounding error 2^-24 |x - c|, which the
// bandwidth magnifies by (|x - c| / h)(r / h).  Here every recentred
// coordinate v is stored as v_hi + v_lo with v_hi on the grid 2^-S (exact in fp32, and so
// is any difference q_hi - x_hi) and v_lo the remainder, so
//     delta = (q_hi - x_hi) + (q_lo - x_lo)
// is the fp64 difference rounded once.  (4D+1) packed lane-ops + 1 MUFU per pair instead of
// (2D+1).  Tile rows: -x_hi[D], -x_lo[D], a2 = log2(e)/(2 h^2), nlc.
__global__ void pack_sources_split_kernel(const double* __restrict__ data,
                                          const double* __restrict__ w,
                                          const double* __restrict__ bw, double bw_scalar,
                                          uint64_t N, int D, Center cen, double w_scale,
                                          double lc_ref, double grid, float* __restrict__ packed) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t ntl = n_tiles_for(N);
  if (i >= ntl * TS) return;
  const int rows = 2 * D + 2;
  float* tile = packed + (i / TS) * (uint64_t)rows * TS + (i % TS);
  if (i >= N) {
    for (int k = 0; k <= 2 * D; k++) tile[(uint64_t)k * TS] = 0.f;
    tile[(uint64_t)(2 * D + 1) * TS] = INFINITY;
    return;
  }
  const double h = bw ? bw[i] : bw_scalar;
  const double wi = (w ? w[i] : 1.0) * w_scale;
  for (int k = 0; k < D; k++) {
    const double v = data[i * D + k] - cen.c[k];
    const double hi = rint(v * grid) / grid;
    tile[(uint64_t)k * TS] = (float)(-hi);
    tile[(uint64_t)(D + k) * TS] = (float)(-(v - hi));
  }
  tile[(uint64_t)(2 * D) * TS] = (float)(0.72134752044448170368 / (h * h));  // log2(e)/2
  const double lc = log2(wi) - D * log2(h);
  tile[(uint64_t)(2 * D + 1) * TS] = (float)(lc_ref - lc);
}

__global__ void pack_queries_split_kernel(const double* __restrict__ pts, uint64_t M, uint64_t Mpad,
                                          int D, Center cen, double grid, float* __restrict__ qT) {
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mpad) return;
  for (int k = 0; k < D; k++) {
    const double v = m < M ? pts[m * D + k] - cen.c[k] : 0.0;
    const double hi = rint(v * grid) / grid;
    qT[(uint64_t)k * Mpad + m] = (float)hi;
    qT[(uint64_t)(D + k) * Mpad + m] = (float)(v - hi);
  }
}

constexpr int EVS_STAGES = 2;

template <int D>
__global__ void __launch_bounds__(EV_THREADS, 2)
    kde_eval_split_kernel(const float* __restrict__ packed, uint32_t n_tiles,
                          uint32_t tiles_per_split, const float* __restrict__ qT, uint64_t Mpad,
                          double out_scale, double* __restrict__ out) {
  constexpr int ROWS = 2 * D + 2;
  constexpr int TILE_FLOATS = ROWS * TS;
  constexpr uint32_t TILE_BYTES = TILE_FLOATS * sizeof(float);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t full[EVS_STAGES];
  const int tid = threadIdx.x;
  const uint32_t t0 = blockIdx.y * tiles_per_split;
  const uint32_t t1 = min(n_tiles, t0 + tiles_per_split);
  const uint32_t nt = t1 > t0 ? t1 - t0 : 0;
  if (tid == 0) {
    for (int s = 0; s < EVS_STAGES; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    for (uint32_t p = 0; p < (uint32_t)(EVS_STAGES - 1) && p < nt; p++) {
      mbar_expect_tx(&full[p], TILE_BYTES);
      tma_load_1d(tiles + p * TILE_FLOATS, packed + (uint64_t)(t0 + p) * TILE_FLOATS, TILE_BYTES,
                  &full[p]);
    }
  }
  const uint64_t mbase = (uint64_t)blockIdx.x * (EV_THREADS * EV_Q) + tid;
  float qh[EV_Q][D], ql[EV_Q][D];
#pragma unroll
  for (int qi = 0; qi < EV_Q; qi++)
#pragma unroll
    for (int k = 0; k < D; k++) {
      qh[qi][k] = qT[(uint64_t)k * Mpad + mbase + qi * EV_THREADS];
      ql[qi][k] = qT[(uint64_t)(D + k) * Mpad + mbase + qi * EV_THREADS];
    }
  double accd[EV_Q];
#pragma unroll
  for (int qi = 0; qi < EV_Q; qi++) accd[qi] = 0.0;

  for (uint32_t it = 0; it < nt; it++) {
    const uint32_t s = it % EVS_STAGES;
    const uint32_t parity = (it / EVS_STAGES) & 1u;
    if (tid == 0 && it + EVS_STAGES - 1 < nt) {
      const uint32_t ps = (it + EVS_STAGES - 1) % EVS_STAGES;
      mbar_expect_tx(&full[ps], TILE_BYTES);
      tma_load_1d(tiles + ps * TILE_FLOATS,
                  packed + (uint64_t)(t0 + it + EVS_STAGES - 1) * TILE_FLOATS, TILE_BYTES,
                  &full[ps]);
    }
    mbar_wait(&full[s], parity);
    const float* tile = tiles + s * TILE_FLOATS;
    f2_t acc[EV_Q];
#pragma unroll
    for (int qi = 0; qi < EV_Q; qi++) acc[qi] = 0ull;
#pragma unroll 1
    for (int j = 0; j < TS; j += 2) {
      float2 r[ROWS];
#pragma unroll
      for (int k = 0; k < ROWS; k++) r[k] = *reinterpret_cast<const float2*>(tile + k * TS + j);
      const f2_t a2 = f2_pack(r[2 * D].x, r[2 * D].y);
      const f2_t nl = f2_pack(r[2 * D + 1].x, r[2 * D + 1].y);
#pragma unroll
      for (int qi = 0; qi < EV_Q; qi++) {
        f2_t t = 0ull;
#pragma unroll
        for (int k = 0; k < D; k++) {
          const f2_t dh = f2_add(f2_pack(qh[qi][k], qh[qi][k]), f2_pack(r[k].x, r[k].y));
          const f2_t dl = f2_add(f2_pack(ql[qi][k], ql[qi][k]), f2_pack(r[D + k].x, r[D + k].y));
          const f2_t dd = f2_add(dh, dl);
          t = f2_fma(dd, dd, t);
        }
        float e0, e1;
        f2_unpack(f2_fma(t, a2, nl), e0, e1);
        acc[qi] = f2_add(acc[qi], f2_pack(ex2_approx(-e0), ex2_approx(-e1)));
      }
    }
#pragma unroll
    for (int qi = 0; qi < EV_Q; qi++) {
      float lo, hi;
      f2_unpack(acc[qi], lo, hi);
      accd[qi] += (double)lo + (double)hi;
    }
    __syncthreads();
  }
#pragma unroll
  for (int qi = 0; qi < EV_Q; qi++)
    out[(uint64_t)blockIdx.y * Mpad + mbase + qi * EV_THREADS] = accd[qi] * out_scale;
}

__global__ void reduce_splits_kernel(const double* __restrict__ partial, int nsplit, uint64_t M,
                                     uint64_t Mpad, double* __restrict__ out) {
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double s = 0.0;
  for (int p = 0; p < nsplit; p++) s += partial[(uint64_t)p * Mpad + m];
  out[m] = s;
}

template <int D>
static int launch_eval(const float* packed, uint32_t n_tiles, uint32_t tps, int nsplit,
                       const float* qT, uint64_t Mpad, float cut2, int kern, double out_scale,
                       double* out, cudaStream_t st) {
  const size_t smem = (size_t)EV_STAGES * (D + 2) * TS * sizeof(float);
  dim3 grid((unsigned)(Mpad / (EV_THREADS * EV_Q)), (unsigned)nsplit);
#define KDSB_EV(CUTV, KV)                                                                   \
  do {                                                                                      \
    KDSB_CUDA(cudaFuncSetAttribute(kde_eval_kernel<D, CUTV, KV>,                            \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kde_eval_kernel<D, CUTV, KV><<<grid, EV_THREADS, smem, st>>>(packed, n_tiles, tps, qT,   \
                                                                 Mpad, cut2, out_scale, out); \
  } while (0)
  if (kern == 1) KDSB_EV(false, 1);
  else if (kern == 2) KDSB_EV(false, 2);
  else if (cut2 > 0.f) KDSB_EV(true, 0);
  else KDSB_EV(false, 0);
#undef KDSB_EV
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// fp64 variant (1e-12 mode): plain smem-tiled kernel, libdevice exp
// ---------------------------------------------------------------------------
constexpr int E64_THREADS = 128;
constexpr int E64_TILE = 256;

// kern 0: coef*exp(-r2*ih2), ih2 = 1/(2h^2); kern 1/2: ih2 = 1/R^2 (support radius),
// coef*max(0, 1 - r2*ih2) / coef where r2*ih2 < 1
template <int D>
__global__ void __launch_bounds__(E64_THREADS)
    kde_eval_f64_kernel(const double* __restrict__ data, const double* __restrict__ coef,
                        const double* __restrict__ ih2, uint64_t N, uint64_t n_per_split,
                        const double* __restrict__ pts, uint64_t M, double cut2, int kern,
                        double* __restrict__ out, uint64_t Mstride) {
  __shared__ double sx[E64_TILE * D];
  __shared__ double sc[E64_TILE];
  __shared__ double sh[E64_TILE];
  const uint64_t m = (uint64_t)blockIdx.x * E64_THREADS + threadIdx.x;
  double q[D];
#pragma unroll
  for (int k = 0; k < D; k++) q[k] = m < M ? pts[m * D + k] : 0.0;
  const uint64_t i0 = (uint64_t)blockIdx.y * n_per_split;
  const uint64_t i1 = min(N, i0 + n_per_split);
  double acc = 0.0;
  for (uint64_t base = i0; base < i1; base += E64_TILE) {
    const int cnt = (int)min((uint64_t)E64_TILE, i1 - base);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * D; e += E64_THREADS) sx[e] = data[base * D + e];
    for (int e = threadIdx.x; e < cnt; e += E64_THREADS) {
      sc[e] = coef[base + e];
      sh[e] = ih2[base + e];
    }
    __syncthreads();
    for (int j = 0; j < cnt; j++) {
      double r2 = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) {
        const double t = q[k] - sx[j * D + k];
        r2 += t * t;
      }
      if (kern == 0) {
        if (r2 <= cut2) acc += sc[j] * exp(-r2 * sh[j]);
      } else {
        const double u = r2 * sh[j];
        if (u < 1.0) acc += kern == 1 ? sc[j] * (1.0 - u) : sc[j];
      }
    }
  }
  if (m < M) out[(uint64_t)blockIdx.y * Mstride + m] = acc;
}

// log of the volume of the D-dimensional unit ball
__host__ __device__ inline double log_unit_ball(int D) {
  return 0.5 * D * 1.1447298858494001741434273513531 - lgamma(0.5 * D + 1.0);
}

__global__ void coef_f64_kernel(const double* __restrict__ w, const double* __restrict__ bw,
                                double bw_scalar, uint64_t N, int D, double w_scale, int kern,
                                double* __restrict__ coef, double* __restrict__ ih2) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const double h = bw ? bw[i] : bw_scalar;
  const double wi = (w ? w[i] : 1.0) * w_scale;
  if (kern == 0) {
    const double lognorm = -0.5 * D * 1.8378770664093454835606594728112;  // log(2 pi)
    coef[i] = wi * exp(lognorm - D * log(h));
    ih2[i] = 0.5 / (h * h);
  } else {
    // KDEpy scales finite-support kernels to unit variance: R = h / sqrt(var),
    // var = 1/5 (epa), 1/3 (box)
    const double R = h * (kern == 1 ? 2.2360679774997896964 : 1.7320508075688772935);
    const double norm = (kern == 1 ? 0.5 * (D + 2) : 1.0) * exp(-log_unit_ball(D));
    coef[i] = wi * norm * exp(-D * log(R));
    ih2[i] = 1.0 / (R * R);
  }
}

template <int D>
static int launch_eval_split(const float* packed, uint32_t n_tiles, uint32_t tps, int nsplit,
                             const float* qT, uint64_t Mpad, double out_scale, double* out,
                             cudaStream_t st) {
  const size_t smem = (size_t)EVS_STAGES * (2 * D + 2) * TS * sizeof(float);
  dim3 grid((unsigned)(Mpad / (EV_THREADS * EV_Q)), (unsigned)nsplit);
  KDSB_CUDA(cudaFuncSetAttribute(kde_eval_split_kernel<D>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kde_eval_split_kernel<D><<<grid, EV_THREADS, smem, st>>>(packed, n_tiles, tps, qT, Mpad,
                                                           out_scale, out);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace kdsb

using namespace kdsb;

extern "C" {

uint64_t kdsb200_packed_floats(uint64_t N, int D) { return n_tiles_for(N) * (uint64_t)(D + 2) * TS; }

uint64_t kdsb200_eval_mpad(uint64_t M) {
  const uint64_t g = EV_THREADS * EV_Q;
  return (M + g - 1) / g * g;
}

int kdsb200_pack_sources_f32(const double* d_data, const double* d_w, const double* d_bw,
                             double bw_scalar, uint64_t N, int D, const double* h_center,
                             double w_scale, double lc_ref, int mode, float* d_packed,
                             void* stream) {
  if (D < 1 || D > 8) {
    set_error("pack_sources: D=%d out of range 1..8", D);
    return 1;
  }
  Center cen;
  for (int k = 0; k < 8; k++) cen.c[k] = k < D ? h_center[k] : 0.0;
  const uint64_t tot = n_tiles_for(N) * TS;
  if (tot == 0) return 0;
  pack_sources_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_data, d_w, d_bw, bw_scalar, N, D, cen, w_scale, lc_ref, mode, d_packed);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

int kdsb200_pack_queries_f32(const double* d_pts, uint64_t M, uint64_t Mpad, int D,
                             const double* h_center, float* d_qT, void* stream) {
  if (D < 1 || D > 8) {
    set_error("pack_queries: D=%d out of range 1..8", D);
    return 1;
  }
  Center cen;
  for (int k = 0; k < 8; k++) cen.c[k] = k < D ? h_center[k] : 0.0;
  if (Mpad == 0) return 0;
  pack_queries_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_pts, M, Mpad, D, cen, d_qT);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

uint64_t kdsb200_packed_split_floats(uint64_t N, int D) {
  return n_tiles_for(N) * (uint64_t)(2 * D + 2) * TS;
}

int kdsb200_pack_sources_split_f32(const double* d_data, const double* d_w, const double* d_bw,
                                   double bw_scalar, uint64_t N, int D, const double* h_center,
                                   double w_scale, double lc_ref, double grid, float* d_packed,
                                   void* stream) {
  if (D < 1 || D > 8 || !(grid > 0)) {
    set_error("pack_sources_split: D=%d out of range 1..8 or grid <= 0", D);
    return 1;
  }
  Center cen;
  for (int k = 0; k < 8; k++) cen.c[k] = k < D ? h_center[k] : 0.0;
  const uint64_t tot = n_tiles_for(N) * TS;
  if (tot == 0) return 0;
  pack_sources_split_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_data, d_w, d_bw, bw_scalar, N, D, cen, w_scale, lc_ref, grid, d_packed);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

int kdsb200_pack_queries_split_f32(const double* d_pts, uint64_t M, uint64_t Mpad, int D,
                                   const double* h_center, double grid, float* d_qT,
                                   void* stream) {
  if (D < 1 || D > 8 || !(grid > 0)) {
    set_error("pack_queries_split: D=%d out of range 1..8 or grid <= 0", D);
    return 1;
  }
  Center cen;
  for (int k = 0; k < 8; k++) cen.c[k] = k < D ? h_center[k] : 0.0;
  if (Mpad == 0) return 0;
  pack_queries_split_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_pts, M, Mpad, D, cen, grid, d_qT);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

int kdsb200_kde_eval_split_f32(const float* d_packed, uint64_t N, int D, const float* d_qT,
                               uint64_t M, uint64_t Mpad, double out_scale, int nsplit,
                               double* d_partial, double* d_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) return 0;
  if (Mpad % (EV_THREADS * EV_Q) != 0 || Mpad < M) {
    set_error("kde_eval_split_f32: Mpad=%llu must be a multiple of %d and >= M",
              (unsigned long long)Mpad, EV_THREADS * EV_Q);
    return 1;
  }
  const uint32_t ntl = (uint32_t)n_tiles_for(N);
  if (nsplit < 1) nsplit = 1;
  if ((uint32_t)nsplit > ntl && ntl > 0) nsplit = (int)ntl;
  const uint32_t tps = ntl ? (ntl + nsplit - 1) / nsplit : 0;
  if (ntl) nsplit = (int)((ntl + tps - 1) / tps);
  double* target;
  if (nsplit == 1 && Mpad == M) {
    target = d_out;
  } else {
    if (!d_partial) {
      set_error("kde_eval_split_f32: partial buffer required (nsplit=%d)", nsplit);
      return 1;
    }
    target = d_partial;
  }
  int rc = 1;
  switch (D) {
#define KDSB_CASE(d)                                                                              \
  case d:                                                                                         \
    rc = launch_eval_split<d>(d_packed, ntl, tps, nsplit, d_qT, Mpad, out_scale, target, st);     \
    break;
    KDSB_CASE(1)
    KDSB_CASE(2)
    KDSB_CASE(3)
    KDSB_CASE(4)
    KDSB_CASE(5)
    KDSB_CASE(6)
    KDSB_CASE(7)
    KDSB_CASE(8)
#undef KDSB_CASE
    default:
      set_error("kde_eval_split_f32: D=%d out of range 1..8", D);
      return 1;
  }
  if (rc) return rc;
  if (target != d_out) {
    reduce_splits_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(target, nsplit, M, Mpad,
                                                                      d_out);
    KDSB_CUDA(cudaGetLastError());
  }
  return 0;
}

int kdsb200_eval_nsplit(uint64_t N, uint64_t M) {
  // Pick the source split whose CTA count fills whole waves of 2 CTAs/SM best
  // (at least ~4 waves when the problem is large enough).
  const uint64_t qtiles = kdsb200_eval_mpad(M) / (EV_THREADS * EV_Q);
  const uint64_t ntl = n_tiles_for(N);
  if (qtiles == 0 || ntl == 0) return 1;
  const uint64_t slots = (uint64_t)sm_count() * 2;
  uint64_t lo = (4 * slots + qtiles - 1) / qtiles;
  if (lo < 1) lo = 1;
  if (lo > ntl) lo = ntl;
  uint64_t hi = lo * 2 + 4;
  if (hi > ntl) hi = ntl;
  uint64_t best = lo;
  double best_eff = -1.0;
  for (uint64_t ns = lo; ns <= hi; ns++) {
    const uint64_t tps = (ntl + ns - 1) / ns;
    const uint64_t real = (ntl + tps - 1) / tps;  // splits actually launched
    const uint64_t ctas = real * qtiles;
    const uint64_t waves = (ctas + slots - 1) / slots;
    // work of the longest CTA chain relative to a perfect distribution
    const double eff = (double)(qtiles * ntl) / ((double)waves * slots * tps);
    if (eff > best_eff + 1e-9) {
      best_eff = eff;
      best = ns;
    }
  }
  return (int)best;
}

int kdsb200_kde_eval_f32(const float* d_packed, uint64_t N, int D, const float* d_qT, uint64_t M,
                         uint64_t Mpad, double cutoff, double out_scale, int nsplit,
                         double* d_partial, double* d_out, void* stream) {
  return kdsb200_kde_eval_kernel_f32(d_packed, N, D, d_qT, M, Mpad, 0, cutoff, out_scale, nsplit,
                                     d_partial, d_out, stream);
}

int kdsb200_kde_eval_kernel_f32(const float* d_packed, uint64_t N, int D, const float* d_qT,
                                uint64_t M, uint64_t Mpad, int kernel, double cutoff,
                                double out_scale, int nsplit, double* d_partial, double* d_out,
                                void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) return 0;
  if (kernel < 0 || kernel > 2) {
    set_error("kde_eval: kernel=%d must be 0 (gaussian), 1 (epa) or 2 (box)", kernel);
    return 1;
  }
  if (Mpad % (EV_THREADS * EV_Q) != 0 || Mpad < M) {
    set_error("kde_eval_f32: Mpad=%llu must be a multiple of %d and >= M",
              (unsigned long long)Mpad, EV_THREADS * EV_Q);
    return 1;
  }
  const uint32_t ntl = (uint32_t)n_tiles_for(N);
  if (nsplit < 1) nsplit = 1;
  if ((uint32_t)nsplit > ntl && ntl > 0) nsplit = (int)ntl;
  const uint32_t tps = ntl ? (ntl + nsplit - 1) / nsplit : 0;
  if (ntl) nsplit = (int)((ntl + tps - 1) / tps);
  const float cut2 = cutoff > 0 ? (float)(cutoff * cutoff) : -1.f;
  double* target;
  if (nsplit == 1 && Mpad == M) {
    target = d_out;
  } else {
    if (!d_partial) {
      set_error("kde_eval_f32: partial buffer required (nsplit=%d)", nsplit);
      return 1;
    }
    target = d_partial;
  }
  int rc = 1;
  switch (D) {
#define KDSB_CASE(d)                                                                         \
  case d:                                                                                    \
    rc = launch_eval<d>(d_packed, ntl, tps, nsplit, d_qT, Mpad, cut2, kernel, out_scale, target, \
                        st);                                                                 \
    break;
    KDSB_CASE(1)
    KDSB_CASE(2)
    KDSB_CASE(3)
    KDSB_CASE(4)
    KDSB_CASE(5)
    KDSB_CASE(6)
    KDSB_CASE(7)
    KDSB_CASE(8)
#undef KDSB_CASE
    default:
      set_error("kde_eval_f32: D=%d out of range 1..8", D);
      return 1;
  }
  if (rc) return rc;
  if (target != d_out) {
    reduce_splits_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(target, nsplit, M, Mpad,
                                                                      d_out);
    KDSB_CUDA(cudaGetLastError());
  }
  return 0;
}

int kdsb200_kde_eval_f64(const double* d_data, const double* d_w, const double* d_bw,
                         double bw_scalar, uint64_t N, int D, double w_scale,
                         const double* d_pts, uint64_t M, double cutoff, int nsplit,
                         double* d_scratch /* 2N + nsplit*M doubles */, double* d_out,
                         void* stream) {
  return kdsb200_kde_eval_kernel_f64(d_data, d_w, d_bw, bw_scalar, N, D, w_scale, d_pts, M, 0,
                                     cutoff, nsplit, d_scratch, d_out, stream);
}

int kdsb200_kde_eval_kernel_f64(const double* d_data, const double* d_w, const double* d_bw,
                                double bw_scalar, uint64_t N, int D, double w_scale,
                                const double* d_pts, uint64_t M, int kernel, double cutoff,
                                int nsplit, double* d_scratch, double* d_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0 || N == 0) return 0;
  if (kernel < 0 || kernel > 2) {
    set_error("kde_eval: kernel=%d must be 0 (gaussian), 1 (epa) or 2 (box)", kernel);
    return 1;
  }
  double* coef = d_scratch;
  double* ih2 = d_scratch + N;
  double* partial = d_scratch + 2 * N;
  coef_f64_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(d_w, d_bw, bw_scalar, N, D,
                                                               w_scale, kernel, coef, ih2);
  KDSB_CUDA(cudaGetLastError());
  if (nsplit < 1) nsplit = 1;
  const uint64_t nps = (N + nsplit - 1) / nsplit;
  nsplit = (int)((N + nps - 1) / nps);
  dim3 grid((unsigned)((M + E64_THREADS - 1) / E64_THREADS), (unsigned)nsplit);
  const double cut2 = cutoff > 0 ? cutoff * cutoff : INFINITY;
  double* target = nsplit == 1 ? d_out : partial;
  switch (D) {
#define KDSB_CASE(d)                                                                       \
  case d:                                                                                  \
    kde_eval_f64_kernel<d><<<grid, E64_THREADS, 0, st>>>(d_data, coef, ih2, N, nps, d_pts, \
                                                         M, cut2, kernel, target, M);      \
    break;
    KDSB_CASE(1)
    KDSB_CASE(2)
    KDSB_CASE(3)
    KDSB_CASE(4)
    KDSB_CASE(5)
    KDSB_CASE(6)
    KDSB_CASE(7)
    KDSB_CASE(8)
#undef KDSB_CASE
    default:
      set_error("kde_eval_f64: D=%d out of range 1..8", D);
      return 1;
  }
  KDSB_CUDA(cudaGetLastError());
  if (nsplit > 1) {
    reduce_splits_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(partial, nsplit, M, M,
                                                                      d_out);
    KDSB_CUDA(cudaGetLastError());
  }
  return 0;
}

}  // extern "C"
