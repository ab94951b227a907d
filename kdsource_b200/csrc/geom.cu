// geom_transform: particle variables -> scaled KDE variables and jacobian, on the
// device.  Restates Geometry.transform / Geometry.jac and the 14 metric classes of
// python/kdsource/geom.py:158-205, 300-1039 for one particle per thread in fp64,
// fused with the division by `scaling` that KDSource.evaluate applies
// (python/kdsource/kdsource.py:236-239), so query points can go from the raw
// (M,8) particle array to the evaluation kernel without a host round trip.
#include <math.h>

#include "common.cuh"
#include "kdsb200.h"

namespace kdsb {

#define KDS_PI 3.1415926535897932384626433832795028841971694

enum {
  G_ENERGY = 0, G_LETHARGY, G_WAVELENGTH, G_VOL, G_SURFXY, G_SURFR, G_SURFR2, G_SURFCIRCLE,
  G_GUIDE, G_ISOTROP, G_POLAR, G_POLARMU, G_TIME, G_DECADE
};  // order of _metric_names, clib/src/geom.c:13-16

struct XfParams {
  int has_rot;
  double rotinv[9];      // inverse rotation matrix, row-major
  int D;
  double inv_scaling[8];
};

__global__ void __launch_bounds__(256)
    geom_transform_kernel(const double* __restrict__ parts, uint64_t M, kdsb200_geom g,
                          XfParams x, double* __restrict__ vecs, double* __restrict__ jac) {
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const double* p = parts + m * 8;
  const double E = p[0], t = p[7];
  double pos[3] = {p[1], p[2], p[3]};
  double dir[3] = {p[4], p[5], p[6]};
  if (g.has_trasl)  // Geometry.transform geom.py:160-168
    for (int k = 0; k < 3; k++) pos[k] -= g.trasl[k];
  if (x.has_rot) {
    double a[3], b[3];
    for (int r = 0; r < 3; r++) {
      a[r] = x.rotinv[r * 3] * pos[0] + x.rotinv[r * 3 + 1] * pos[1] + x.rotinv[r * 3 + 2] * pos[2];
      b[r] = x.rotinv[r * 3] * dir[0] + x.rotinv[r * 3 + 1] * dir[1] + x.rotinv[r * 3 + 2] * dir[2];
    }
    for (int r = 0; r < 3; r++) {
      pos[r] = a[r];
      dir[r] = b[r];
    }
  }
  double v[8];
  double J = 1.0;
  int col = 0;
  const double DEG = 180.0 / KDS_PI;
  for (int i = 0; i < g.order; i++) {
    const kdsb200_metric& mt = g.ms[i];
    switch (mt.type) {
      case G_ENERGY:  // geom.py:300-307
        v[col++] = E;
        break;
      case G_LETHARGY:  // geom.py:310-353
        v[col++] = log(mt.params[0] / E);
        J *= 1.0 / E;
        break;
      case G_WAVELENGTH:  // geom.py:356-380
        v[col++] = 9.045 / sqrt(E * 1e9);
        J *= 9.045 / (2 * sqrt(E * E * E));
        break;
      case G_TIME:  // geom.py:383-390
        v[col++] = t;
        break;
      case G_DECADE:  // geom.py:393-418
        v[col++] = log10(t);
        J *= 0.43429448190325182765 / t;
        break;
      case G_VOL:  // geom.py:421-468
        v[col++] = pos[0];
        v[col++] = pos[1];
        v[col++] = pos[2];
        break;
      case G_SURFXY:      // geom.py:471-520
      case G_SURFCIRCLE:  // geom.py:663-715
        v[col++] = pos[0];
        v[col++] = pos[1];
        break;
      case G_SURFR:  // geom.py:523-590
        v[col++] = sqrt(pos[0] * pos[0] + pos[1] * pos[1]);
        v[col++] = atan2(pos[1], pos[0]) * 180 / KDS_PI;
        J *= 1.0 / sqrt(pos[0] * pos[0] + pos[1] * pos[1]);
        break;
      case G_SURFR2:  // geom.py:593-660
        v[col++] = pos[0] * pos[0] + pos[1] * pos[1];
        v[col++] = atan2(pos[1], pos[0]) * 180 / KDS_PI;
        J *= 2.0;
        break;
      case G_ISOTROP:  // geom.py:868-966
        v[col++] = dir[0];
        v[col++] = dir[1];
        v[col++] = dir[2];
        break;
      case G_POLAR: {  // geom.py:969-1004
        const double th = acos(dir[2]);
        v[col++] = th * 180 / KDS_PI;
        v[col++] = atan2(dir[1], dir[0]) * 180 / KDS_PI;
        J *= DEG * DEG / sin(th);
      } break;
      case G_POLARMU:  // geom.py:1007-1039
        v[col++] = dir[2];
        v[col++] = atan2(dir[1], dir[0]) * 180 / KDS_PI;
        break;
      case G_GUIDE: {  // geom.py:739-793
        double xs = pos[0], ys = pos[1], zs = pos[2], dxs = dir[0], dys = dir[1], dzs = dir[2];
        const double xw = mt.params[0], yh = mt.params[1], rc = mt.params[3];
        if (rc != 0) {  // rcurv=None is encoded as 0 (see Geometry C descriptors)
          const double rs = sqrt((rc + xs) * (rc + xs) + zs * zs);
          xs = copysign(1.0, rc) * rs - rc;
          zs = fabs(rc) * asin(zs / rs);
          const double ang = zs / rc;
          const double d1 = dxs * cos(ang) + dzs * sin(ang);
          const double d2 = -dxs * sin(ang) + dzs * cos(ang);
          dxs = d1;
          dzs = d2;
        }
        const double a = ys / yh, b = xs / xw;
        double tt = NAN, mu = NAN, ph = NAN;  // np.empty_like leaves unmatched rows undefined
        if (a > -b && a < b) {
          tt = 0.5 * yh + ys; mu = dxs; ph = atan2(-dys, dzs);
        } else if (a > b && a > -b) {
          tt = 1.0 * yh + 0.5 * xw - xs; mu = dys; ph = atan2(dxs, dzs);
        } else if (a < -b && a > b) {
          tt = 1.5 * yh + 1.0 * xw - ys; mu = -dxs; ph = atan2(dys, dzs);
        } else if (a < b && a < -b) {
          tt = 2.0 * yh + 1.5 * xw + xs; mu = -dys; ph = atan2(-dxs, dzs);
        }
        v[col++] = zs;
        v[col++] = tt;
        v[col++] = mu;
        v[col++] = ph * DEG;
      } break;
      default:
        break;
    }
  }
  for (int k = 0; k < x.D; k++) vecs[m * x.D + k] = v[k] * x.inv_scaling[k];
  if (jac) jac[m] = J;
}

// ---------------------------------------------------------------------------
// weighted column moments (Metric.mean / Metric.std, python/kdsource/geom.py:61-70,
// and N_eff = (sum w)^2 / sum w^2, plist.py:349-395) on the device
// ---------------------------------------------------------------------------
constexpr int MOM_BLOCKS = 592;  // 4 per SM on 148 SMs; fixed, so the sums are reproducible
constexpr int MOM_THREADS = 256;

struct MomCenter {
  double c[8];
};

__global__ void __launch_bounds__(MOM_THREADS)
    moments_partial_kernel(const double* __restrict__ vecs, const double* __restrict__ w,
                           uint64_t N, int D, MomCenter cen, int power,
                           double* __restrict__ scratch) {
  double acc[10];
#pragma unroll
  for (int k = 0; k < 10; k++) acc[k] = 0.0;
  for (uint64_t i = (uint64_t)blockIdx.x * MOM_THREADS + threadIdx.x; i < N;
       i += (uint64_t)MOM_BLOCKS * MOM_THREADS) {
    const double wi = w ? w[i] : 1.0;
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k < D) {
        const double d = vecs[i * D + k] - cen.c[k];
        acc[k] += wi * (power == 2 ? d * d : d);
      }
    acc[8] += wi;
    acc[9] += wi * wi;
  }
  __shared__ double sm[10][MOM_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 10; k++) {
    double v = acc[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) sm[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 10) {
    double t = 0.0;
#pragma unroll
    for (int e = 0; e < MOM_THREADS / 32; e++) t += sm[threadIdx.x][e];
    scratch[blockIdx.x * 10 + threadIdx.x] = t;
  }
}

__global__ void moments_final_kernel(const double* __restrict__ scratch, int D,
                                     double* __restrict__ out) {
  // thread k sums column k of the per-block partials in block order
  const int k = threadIdx.x;
  if (k >= 10) return;
  double t = 0.0;
  for (int b = 0; b < MOM_BLOCKS; b++) t += scratch[b * 10 + k];
  if (k < D) out[k] = t;
  if (k == 8) out[D] = t;
  if (k == 9) out[D + 1] = t;
}

}  // namespace kdsb

using namespace kdsb;

extern "C" uint64_t kdsb200_moments_scratch(void) { return (uint64_t)MOM_BLOCKS * 10; }

extern "C" int kdsb200_weighted_moments(const double* d_vecs, const double* d_w, uint64_t N, int D,
                                        const double* h_center, int power, double* d_scratch,
                                        double* d_out, void* stream) {
  if (D < 1 || D > 8 || (power != 1 && power != 2)) {
    set_error("weighted_moments: D=%d (1..8) power=%d (1|2)", D, power);
    return 1;
  }
  MomCenter cen;
  for (int k = 0; k < 8; k++) cen.c[k] = (h_center && k < D) ? h_center[k] : 0.0;
  cudaStream_t st = (cudaStream_t)stream;
  moments_partial_kernel<<<MOM_BLOCKS, MOM_THREADS, 0, st>>>(d_vecs, d_w, N, D, cen, power,
                                                             d_scratch);
  KDSB_CUDA(cudaGetLastError());
  moments_final_kernel<<<1, 32, 0, st>>>(d_scratch, D, d_out);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}


extern "C" int kdsb200_geom_transform(const double* d_parts, uint64_t M,
                                      const struct kdsb200_geom_s* h_geom,
                                      const double* h_rot_inv, const double* h_inv_scaling,
                                      double* d_vecs, double* d_jac, void* stream) {
  if (M == 0) return 0;
  if (h_geom->order < 1 || h_geom->order > KDSB200_MAX_METRICS) {
    set_error("geom_transform: geometry order %d out of range", h_geom->order);
    return 1;
  }
  XfParams x;
  x.D = 0;
  for (int i = 0; i < h_geom->order; i++) x.D += h_geom->ms[i].dim;
  if (x.D < 1 || x.D > 8) {
    set_error("geom_transform: %d KDE variables (1..8 supported)", x.D);
    return 1;
  }
  if (h_geom->has_rot && !h_rot_inv) {
    set_error("geom_transform: geometry has a rotation but no inverse matrix was given");
    return 1;
  }
  x.has_rot = h_geom->has_rot != 0;
  for (int k = 0; k < 9; k++) x.rotinv[k] = x.has_rot ? h_rot_inv[k] : (k % 4 == 0 ? 1.0 : 0.0);
  for (int k = 0; k < 8; k++) x.inv_scaling[k] = (h_inv_scaling && k < x.D) ? h_inv_scaling[k] : 1.0;
  geom_transform_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_parts, M, *h_geom, x, d_vecs, d_jac);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}
