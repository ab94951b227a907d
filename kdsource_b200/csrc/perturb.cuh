// Deviate maps (clib/src/utils.c:9-67) and per-metric perturbations
// (clib/src/geom.c:192-650) of the resampling contract, written once for both sides:
// the device sampler (resample.cu, one particle per thread, Philox or external
// uniforms) and the single-particle host entry points of the C interface
// (X_perturb, Geom_perturb, kds_rand_* in compat.cu, driven by the caller's rng).
//
// S is a uniform-stream type: S::EXT (uniforms arrive as doubles through ext_next())
// or the Philox integer stream (kint()).
#pragma once
#include <math.h>

#include "common.cuh"
#include "kdsb200.h"

namespace kdsb {

#define KDS_PI 3.1415926535897932384626433832795028841971694

enum {
  M_ENERGY = 0, M_LETHARGY, M_WAVELENGTH, M_VOL, M_SURFXY, M_SURFR, M_SURFR2, M_SURFCIRCLE,
  M_GUIDE, M_ISOTROP, M_POLAR, M_POLARMU, M_TIME, M_DECADE
};  // order of _metric_names, clib/src/geom.c:13-16


template <typename R>
struct Mth;
template <>
struct Mth<float> {
  static __host__ __device__ __forceinline__ float exp_(float x) { return expf(x); }
  static __host__ __device__ __forceinline__ float log_(float x) { return logf(x); }
  static __host__ __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
  static __host__ __device__ __forceinline__ float sin_(float x) { return sinf(x); }
  static __host__ __device__ __forceinline__ float cos_(float x) { return cosf(x); }
  static __host__ __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
  static __host__ __device__ __forceinline__ float acos_(float x) { return acosf(x); }
  static __host__ __device__ __forceinline__ float asin_(float x) { return asinf(x); }
  static __host__ __device__ __forceinline__ float pow10_(float x) {
#ifdef __CUDA_ARCH__
    return exp10f(x);
#else
    return powf(10.0f, x);
#endif
  }
};
template <>
struct Mth<double> {
  static __host__ __device__ __forceinline__ double exp_(double x) { return exp(x); }
  static __host__ __device__ __forceinline__ double log_(double x) { return log(x); }
  static __host__ __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
  static __host__ __device__ __forceinline__ double sin_(double x) { return sin(x); }
  static __host__ __device__ __forceinline__ double cos_(double x) { return cos(x); }
  static __host__ __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
  static __host__ __device__ __forceinline__ double acos_(double x) { return acos(x); }
  static __host__ __device__ __forceinline__ double asin_(double x) { return asin(x); }
  static __host__ __device__ __forceinline__ double pow10_(double x) { return pow(10.0, x); }
};

// FAST (fp32 only): MUFU-based approximations (lg2/ex2/rcp/rsq/sin/cos, abs. error
// ~2^-21) for the transcendental steps of the deviate and direction maps.  The
// accept/reject decisions stay exact, so the uniform consumption is unchanged and
// the perturbed coordinates stay within ~1e-6 of the variable's scale.
template <typename R, bool FAST>
struct Mf {
  static __host__ __device__ __forceinline__ R exp_(R x) { return Mth<R>::exp_(x); }
  static __host__ __device__ __forceinline__ R log_(R x) { return Mth<R>::log_(x); }
  static __host__ __device__ __forceinline__ R sqrt_(R x) { return Mth<R>::sqrt_(x); }
  static __host__ __device__ __forceinline__ R div_(R a, R b) { return a / b; }
};
template <>
struct Mf<float, true> {
#ifdef __CUDA_ARCH__
  static __device__ __forceinline__ float exp_(float x) { return __expf(x); }
  static __device__ __forceinline__ float log_(float x) { return __logf(x); }
  static __device__ __forceinline__ float sqrt_(float x) { return __fsqrt_rn(x); }
  static __device__ __forceinline__ float div_(float a, float b) { return __fdividef(a, b); }
#else  // the fast maps exist on the device only; the host never instantiates a kernel with them
  static float exp_(float x) { return expf(x); }
  static float log_(float x) { return logf(x); }
  static float sqrt_(float x) { return sqrtf(x); }
  static float div_(float a, float b) { return a / b; }
#endif
};

// -ln(t) for t in (0, 1], given m = 1 - t to full relative accuracy (FAST fp32 maps):
// MUFU.LG2 has an ABSOLUTE error of ~2^-22, useless where ln t -> 0, so below m = 2^-6
// the series m + m^2/2 + m^3/3 + m^4/4 takes over (rel. error < 2e-8 there).
__host__ __device__ __forceinline__ float fast_neglog(float t, float m) {
  const float ser = m * fmaf(m, fmaf(m, fmaf(m, 0.25f, 0.33333334f), 0.5f), 1.0f);
#ifdef __CUDA_ARCH__
  const float lg = -0.69314718056f * __log2f(t);
#else
  const float lg = -logf(t);
#endif
  return m < 0.015625f ? ser : lg;
}
__host__ __device__ __forceinline__ float fast_rsqrt(float x) {
#ifdef __CUDA_ARCH__
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
__host__ __device__ __forceinline__ void fast_sincos(float a, float* s, float* c) {
#ifdef __CUDA_ARCH__
  *s = __sinf(a);
  *c = __cosf(a);
#else
  *s = sinf(a);
  *c = cosf(a);
#endif
}
__host__ __device__ __forceinline__ void sincos2pi(float u, float* s, float* c) {
#ifdef __CUDA_ARCH__
  sincospif(2.0f * u, s, c);
#else
  *s = sinf(6.2831853071795864769f * u);
  *c = cosf(6.2831853071795864769f * u);
#endif
}

template <typename R, typename S>
__host__ __device__ __forceinline__ R uniform(S& s) {
  if (S::EXT) return (R)s.ext_next();
  return ((R)s.kint() + (R)0.5) * (R)(1.0 / 8388608.0);
}

// kds_rand_norm: Marsaglia polar method, second value discarded (utils.c:9-29).
// Accept/reject is decided exactly (integers for Philox, fp64 for external
// uniforms) so fp32 and fp64 modes consume the same number of uniforms.
//
// Contract 2 (S::V2, counter-based streams only): Box-Muller, z0 = r cos(2 pi u2), z1 =
// r sin(2 pi u2) with r = sqrt(-2 ln u1); z1 is kept and returned by the next call, so a
// pair of deviates costs exactly two uniforms and nothing is rejected.  The usual
// geometries then need 8 Philox words per particle (pick 2, two pairs 4, direction 2).
template <typename R, typename S, bool FAST>
__host__ __device__ R rand_norm(S& s) {
  if constexpr (S::V2) {
    if (s.have_spare) {
      s.have_spare = false;
      return s.spare;
    }
    const int k1 = s.kint(), k2 = s.kint();
    if constexpr (sizeof(R) == 8) {
      const double u1 = ((double)k1 + 0.5) * (1.0 / 8388608.0);
      const double u2 = ((double)k2 + 0.5) * (1.0 / 8388608.0);
      const double r = sqrt(-2.0 * log(u1));
      const double a = (2.0 * KDS_PI) * u2;
      s.spare = (R)(r * sin(a));
      s.have_spare = true;
      return (R)(r * cos(a));
    } else {
      const float u1 = ((float)k1 + 0.5f) * (1.0f / 8388608.0f);
      const float u2 = ((float)k2 + 0.5f) * (1.0f / 8388608.0f);
      // 1 - u1 = (2^24 - (2 k1 + 1)) 2^-24, exact: -ln u1 keeps its relative accuracy
      // where u1 -> 1 and the deviate is small
      const float m = (float)((1 << 24) - (2 * k1 + 1)) * (1.0f / 16777216.0f);
      float L, sn, cs;
      if (FAST) {
        L = fast_neglog(u1, m);
        fast_sincos(6.2831853072f * (u2 - (u2 >= 0.5f ? 1.0f : 0.0f)), &sn, &cs);
      } else {
        L = m < 0.5f ? -log1pf(-m) : -logf(u1);
        sincos2pi(u2, &sn, &cs);
      }
      const float r = Mf<float, FAST>::sqrt_(2.0f * L);
      s.spare = (R)(r * sn);
      s.have_spare = true;
      return (R)(r * cs);
    }
  }
  if (S::EXT) {
    double t, g1, g2;
    do {
      g1 = 2.0 * s.ext_next() - 1.0;
      g2 = 2.0 * s.ext_next() - 1.0;
      t = g1 * g1 + g2 * g2;
    } while (t >= 1.0 || t == 0.0);
    if (sizeof(R) == 8) return (R)(g1 * sqrt((-2.0 * log(t)) / t));
    const float tf = (float)t;
    const float lt = t > 0.5 ? log1pf((float)(t - 1.0)) : logf(tf);
    return (R)((float)g1 * sqrtf((-2.0f * lt) / tf));
  } else {
    int G1, G2;
    long long ti;
    do {
      G1 = 2 * s.kint() + 1 - (1 << 23);
      G2 = 2 * s.kint() + 1 - (1 << 23);
      ti = (long long)G1 * G1 + (long long)G2 * G2;
    } while (ti >= (1ll << 46));
    if (sizeof(R) == 8) {
      const double t = (double)ti * (1.0 / 70368744177664.0);
      const double g1 = (double)G1 * (1.0 / 8388608.0);
      return (R)(g1 * sqrt((-2.0 * log(t)) / t));
    }
    const float tf = (float)ti * (1.0f / 70368744177664.0f);
    const float g1 = (float)G1 * (1.0f / 8388608.0f);
    if (FAST) {
      // g1 * sqrt(-2 ln t / t) = g1 * sqrt(2 L) * rsqrt(t), L = -ln t.  MUFU.LG2 has an
      // ABSOLUTE error of ~2^-22, so close to t = 1 (where L -> 0 and the deviate is
      // g1 sqrt(2L)) it is replaced by the series in m = 1 - t, formed exactly from the
      // integers: -ln(1 - m) = m + m^2/2 + m^3/3 + m^4/4 (m < 2^-6: rel. error < 2e-8).
      const float m = (float)((1ll << 46) - ti) * (1.0f / 70368744177664.0f);
      return (R)(g1 * Mf<float, true>::sqrt_(2.0f * fast_neglog(tf, m)) * fast_rsqrt(tf));
    }
    const float lt = ti > (1ll << 45)
                         ? log1pf(-(float)((1ll << 46) - ti) * (1.0f / 70368744177664.0f))
                         : logf(tf);
    return (R)(g1 * sqrtf((-2.0f * lt) / tf));
  }
}

// kds_rand_epan (utils.c:32-46)
template <typename R, typename S>
__host__ __device__ R rand_epan(S& s) {
  const R x1 = uniform<R, S>(s) * (R)2 - (R)1;
  const R x2 = uniform<R, S>(s) * (R)2 - (R)1;
  const R x3 = uniform<R, S>(s) * (R)2 - (R)1;
  if (fabs(x3) >= fabs(x2) && fabs(x3) >= fabs(x1)) return x2;
  return x3;
}

// kds_rand_type (utils.c:55-67)
template <typename R, typename S, bool FAST>
__host__ __device__ R rand_type(S& s, char kernel) {
  if (kernel == 'g') return rand_norm<R, S, FAST>(s);
  if (kernel == 'e') return rand_epan<R, S>(s);
  if (kernel == 'b') return uniform<R, S>(s) * (R)2 - (R)1;
  return (R)0;
}

template <typename R>
struct Part {
  R ekin, pos[3], dir[3], time;
};

template <typename R>
__host__ __device__ __forceinline__ void reflect(R& x, R lo, R hi) {
  while ((x < lo) || (x > hi)) {
    if (x < lo)
      x += 2 * (lo - x);
    else
      x -= 2 * (x - hi);
  }
}

// rotv (Rodrigues, utils.c:82-103)
template <typename R>
__host__ __device__ void rotv(R* v, const double* rv, int inverse) {
  using M = Mth<R>;
  R r0 = (R)rv[0], r1 = (R)rv[1], r2 = (R)rv[2];
  R theta = M::sqrt_(r0 * r0 + r1 * r1 + r2 * r2);
  if (theta == 0) return;
  if (inverse) theta = -theta;
  const R kdotv = r0 * v[0] + r1 * v[1] + r2 * v[2];
  const R kx[3] = {r1 * v[2] - r2 * v[1], r2 * v[0] - r0 * v[2], r0 * v[1] - r1 * v[0]};
  const R rr[3] = {r0, r1, r2};
  const R ct = M::cos_(theta), st = M::sin_(theta);
  R o[3];
#pragma unroll
  for (int i = 0; i < 3; i++)
    o[i] = v[i] * ct + (kx[i] / fabs(theta)) * st + rr[i] * kdotv / (theta * theta) * (1 - ct);
#pragma unroll
  for (int i = 0; i < 3; i++) v[i] = o[i];
}

// _vMF_perturb (geom.c:557-573)
template <typename R, typename S, bool FAST>
__host__ __device__ void vmf_perturb(S& s, R bw, R& dx, R& dy, R& dz) {
  using M = Mth<R>;
  using F = Mf<R, FAST>;
  const R xi = uniform<R, S>(s);
  R w, uv;
  if (sizeof(R) == 8) {
    w = 1;
    w += bw * bw * M::log_(xi + (1 - xi) * M::exp_(-2 / (bw * bw)));
    uv = M::sqrt_(1 - w * w);
  } else {
    // same map, arranged so 1 - w*w does not cancel in fp32: the argument of the
    // logarithm is 1 - (1 - xi)(1 - e), and 1 - xi is exact for xi = (k + 0.5) 2^-23
    const R e2 = F::exp_(F::div_((R)-2, bw * bw));
    const R om = (1 - xi) * (1 - e2);
    R nl;
    if (FAST)
      nl = (R)fast_neglog((float)(1 - om), (float)om);
    else
      nl = om < (R)0.5 ? -(R)log1pf(-(float)om) : -F::log_(1 - om);
    const R m = bw * bw * nl;
    w = 1 - m;
    uv = F::sqrt_(fmax(m * (2 - m), (R)0));
  }
  const R u01 = uniform<R, S>(s);
  R sphi, cphi;
  if (sizeof(R) == 8) {
    const R phi = (R)(2. * KDS_PI) * u01;
    sphi = M::sin_(phi);
    cphi = M::cos_(phi);
  } else if (FAST) {
    // angle folded to [-pi, pi) before the MUFU sine/cosine
    const float ang = 6.2831853072f * ((float)u01 - ((float)u01 >= 0.5f ? 1.0f : 0.0f));
    float sf, cf;
    fast_sincos(ang, &sf, &cf);
    sphi = sf;
    cphi = cf;
  } else {
    float sf, cf;
    sincos2pi((float)u01, &sf, &cf);
    sphi = sf;
    cphi = cf;
  }
  const R u = uv * cphi, v = uv * sphi;
  const R dx0 = dx, dy0 = dy, dz0 = dz;
  const R cross = v * dx0 - u * dy0;
  if (dz0 > 0) {
    const R den = 1 + dz0;
    dx = u * dz0 + w * dx0 - F::div_(cross * dy0, den);
    dy = v * dz0 + w * dy0 + F::div_(cross * dx0, den);
  } else {
    const R den = 1 - dz0;
    dx = u * dz0 + w * dx0 + F::div_(cross * dy0, den);
    dy = v * dz0 + w * dy0 - F::div_(cross * dx0, den);
  }
  dz = w * dz0 - u * dx0 - v * dy0;
}

// Guide_perturb (geom.c:386-555)
template <typename R, typename S, bool FAST>
__host__ __device__ void guide_perturb(S& s, const kdsb200_metric& m, Part<R>& p, R bw,
                              char kernel) {
  using M = Mth<R>;
  R x = p.pos[0], y = p.pos[1], z = p.pos[2];
  R dx = p.dir[0], dy = p.dir[1], dz = p.dir[2], dx2, dz2;
  const R xw = (R)m.params[0], yh = (R)m.params[1], zmax = (R)m.params[2], rc = (R)m.params[3];
  R t = 0, mu = 0, mu2 = 0, phi = 0;
  int mirror;
  if (rc != 0) {
    const R r = M::sqrt_((rc + x) * (rc + x) + z * z);
    x = copysign((R)1, rc) * r - rc;
    z = fabs(rc) * M::asin_(z / r);
    dx2 = dx;
    dz2 = dz;
    dx = dx2 * M::cos_(z / rc) + dz2 * M::sin_(z / rc);
    dz = -dx2 * M::sin_(z / rc) + dz2 * M::cos_(z / rc);
  }
  if ((y / yh > -x / xw) && (y / yh < x / xw))
    mirror = 0;
  else if ((y / yh > x / xw) && (y / yh > -x / xw))
    mirror = 1;
  else if ((y / yh < -x / xw) && (y / yh > x / xw))
    mirror = 2;
  else
    mirror = 3;
  R lo = 0, hi = 0;
  switch (mirror) {
    case 0: t = (R)0.5 * yh + y; mu = dx; phi = M::atan2_(-dy, dz); lo = 0; hi = yh; break;
    case 1: t = yh + (R)0.5 * xw - x; mu = dy; phi = M::atan2_(dx, dz); lo = yh; hi = yh + xw; break;
    case 2: t = (R)1.5 * yh + xw - y; mu = -dx; phi = M::atan2_(dy, dz); lo = yh + xw; hi = 2 * yh + xw; break;
    default: t = 2 * yh + (R)1.5 * xw + x; mu = -dy; phi = M::atan2_(-dx, dz); lo = 2 * yh + xw; hi = 2 * yh + 2 * xw; break;
  }
  z += bw * (R)m.scaling[0] * rand_type<R, S, FAST>(s, kernel);
  reflect<R>(z, (R)0, zmax);
  t += bw * (R)m.scaling[1] * rand_type<R, S, FAST>(s, kernel);
  reflect<R>(t, lo, hi);
  if (isinf(m.scaling[2]))
    mu = -1 + 2 * uniform<R, S>(s);
  else {
    mu2 = mu + bw * (R)m.scaling[2] * rand_type<R, S, FAST>(s, kernel);
    if (mu >= 0)
      reflect<R>(mu2, (R)0, (R)1);
    else
      reflect<R>(mu2, (R)-1, (R)0);
    mu = mu2;
  }
  if (isinf(m.scaling[3]))
    phi = (R)(2. * KDS_PI) * uniform<R, S>(s);
  else
    phi += bw * (R)m.scaling[3] * (R)(KDS_PI / 180) * rand_type<R, S, FAST>(s, kernel);
  const R sq = M::sqrt_(1 - mu * mu), cp = M::cos_(phi), sp = M::sin_(phi);
  switch (mirror) {
    case 0: x = xw / 2; y = t - (R)0.5 * yh; dx = mu; dz = sq * cp; dy = -sq * sp; break;
    case 1: y = yh / 2; x = -t + yh + (R)0.5 * xw; dy = mu; dz = sq * cp; dx = sq * sp; break;
    case 2: x = -xw / 2; y = -t + (R)1.5 * yh + xw; dx = -mu; dz = sq * cp; dy = sq * sp; break;
    default: y = -yh / 2; x = t - 2 * yh - (R)1.5 * xw; dy = -mu; dz = sq * cp; dx = -sq * sp; break;
  }
  if (rc != 0) {
    const R r = (rc + x) * copysign((R)1, rc), ang = z / rc;
    x = copysign((R)1, rc) * r * M::cos_(ang) - rc;
    z = r * M::sin_(fabs(ang));
    dx2 = dx;
    dz2 = dz;
    dx = dx2 * M::cos_(ang) - dz2 * M::sin_(ang);
    dz = dx2 * M::sin_(ang) + dz2 * M::cos_(ang);
  }
  p.pos[0] = x; p.pos[1] = y; p.pos[2] = z;
  p.dir[0] = dx; p.dir[1] = dy; p.dir[2] = dz;
}

// one metric's perturbation: geom.c:192-650.  TYPE >= 0 fixes the metric at compile
// time (specialised kernels for the usual geometries carry only the code they
// need); TYPE < 0 dispatches on m.type at run time.
template <typename R, typename S, int TYPE, bool FAST>
__host__ __device__ __forceinline__ void metric_perturb(S& s, const kdsb200_metric& m,
                                               Part<R>& p, R bw, char kernel) {
  using M = Mth<R>;
  const R PI180 = (R)(KDS_PI / 180);
  switch (TYPE >= 0 ? TYPE : m.type) {
    case M_ENERGY:  // E_perturb geom.c:192-198
      p.ekin += bw * (R)m.scaling[0] * rand_type<R, S, FAST>(s, kernel);
      if (p.ekin < 0) p.ekin *= -1;
      break;
    case M_LETHARGY: {  // Let_perturb geom.c:200-210
      R E = p.ekin * Mf<R, FAST>::exp_(bw * (R)m.scaling[0] * rand_type<R, S, FAST>(s, kernel));
      while (E > (R)m.params[0])
        E = p.ekin * Mf<R, FAST>::exp_(bw * (R)m.scaling[0] * rand_type<R, S, FAST>(s, kernel));
      p.ekin = E;
    } break;
    case M_WAVELENGTH: {  // wl_perturb geom.c:212-221
      const R wl = (R)9.045 / M::sqrt_(p.ekin * (R)1e9);
      const R wl2 = wl + bw * (R)m.scaling[0] * rand_type<R, S, FAST>(s, kernel);
      p.ekin = (R)81.82e-9 / (wl2 * wl2);
      if (p.ekin < 0) p.ekin *= -1;
    } break;
    case M_TIME:  // t_perturb geom.c:223-227
      p.time += bw * (R)m.scaling[0] * rand_type<R, S, FAST>(s, kernel);
      break;
    case M_DECADE:  // Dec_perturb geom.c:228-233
      p.time *= M::pow10_(bw * (R)m.scaling[0] * rand_type<R, S, FAST>(s, kernel));
      break;
    case M_VOL:  // Vol_perturb geom.c:235-265
#pragma unroll
      for (int a = 0; a < 3; a++) {
        p.pos[a] += bw * (R)m.scaling[a] * rand_type<R, S, FAST>(s, kernel);
        reflect<R>(p.pos[a], (R)m.params[2 * a], (R)m.params[2 * a + 1]);
      }
      break;
    case M_SURFXY:  // SurfXY_perturb geom.c:266-287
#pragma unroll
      for (int a = 0; a < 2; a++) {
        p.pos[a] += bw * (R)m.scaling[a] * rand_type<R, S, FAST>(s, kernel);
        reflect<R>(p.pos[a], (R)m.params[2 * a], (R)m.params[2 * a + 1]);
      }
      break;
    case M_SURFR:    // SurfR_perturb geom.c:288-319
    case M_SURFR2: { // SurfR2_perturb geom.c:320-354
      const bool sq = m.type == M_SURFR2;
      R rmin = (R)m.params[0], rmax = (R)m.params[1];
      if (sq) { rmin *= rmin; rmax *= rmax; }
      R rho = p.pos[0] * p.pos[0] + p.pos[1] * p.pos[1];
      if (!sq) rho = M::sqrt_(rho);
      const R psi = M::atan2_(p.pos[1], p.pos[0]);
      R rho2 = rho + bw * (R)m.scaling[0] * rand_type<R, S, FAST>(s, kernel);
      R psi2 = psi + bw * (R)m.scaling[1] * PI180 * rand_type<R, S, FAST>(s, kernel);
      reflect<R>(rho2, rmin, rmax);
      reflect<R>(psi2, (R)m.params[2], (R)m.params[3]);
      if (sq) rho2 = M::sqrt_(rho2);
      p.pos[0] = rho2 * M::cos_(psi2);
      p.pos[1] = rho2 * M::sin_(psi2);
    } break;
    case M_SURFCIRCLE: {  // SurfCircle_perturb geom.c:355-385
      const R x = p.pos[0] + bw * (R)m.scaling[0] * rand_type<R, S, FAST>(s, kernel);
      const R y = p.pos[1] + bw * (R)m.scaling[1] * rand_type<R, S, FAST>(s, kernel);
      R rho2 = M::sqrt_(x * x + y * y);
      R psi2 = M::atan2_(y, x);
      reflect<R>(rho2, (R)m.params[0], (R)m.params[1]);
      reflect<R>(psi2, (R)m.params[2], (R)m.params[3]);
      p.pos[0] = rho2 * M::cos_(psi2);
      p.pos[1] = rho2 * M::sin_(psi2);
    } break;
    case M_GUIDE:
      guide_perturb<R, S, FAST>(s, m, p, bw, kernel);
      break;
    case M_ISOTROP: {  // Isotrop_perturb geom.c:575-600 (kernel ignored)
      const R sc = bw * (R)m.scaling[0];
      if (isinf(sc)) {
        p.dir[2] = -1 + 2 * uniform<R, S>(s);
        const R dxy = M::sqrt_(1 - p.dir[2] * p.dir[2]);
        const R phi = (R)(2. * KDS_PI) * uniform<R, S>(s);
        p.dir[0] = dxy * M::cos_(phi);
        p.dir[1] = dxy * M::sin_(phi);
      } else if (sc > 0) {
        const R dx = p.dir[0], dy = p.dir[1], dz = p.dir[2];
        vmf_perturb<R, S, FAST>(s, sc, p.dir[0], p.dir[1], p.dir[2]);
        if ((int)m.params[0] && (dx * p.dir[0] < 0)) p.dir[0] *= -1;
        if ((int)m.params[1] && (dy * p.dir[1] < 0)) p.dir[1] *= -1;
        if ((int)m.params[2] && (dz * p.dir[2] < 0)) p.dir[2] *= -1;
      }
    } break;
    case M_POLAR: {  // Polar_perturb geom.c:602-620 (reflected theta is overwritten there too)
      R theta = M::acos_(p.dir[2]);
      R phi = M::atan2_(p.dir[1], p.dir[0]);
      theta = theta + bw * (R)m.scaling[0] * PI180 * rand_type<R, S, FAST>(s, kernel);
      phi += bw * (R)m.scaling[1] * PI180 * rand_type<R, S, FAST>(s, kernel);
      p.dir[0] = M::sin_(theta) * M::cos_(phi);
      p.dir[1] = M::sin_(theta) * M::sin_(phi);
      p.dir[2] = M::cos_(theta);
    } break;
    case M_POLARMU: {  // PolarMu_perturb geom.c:622-650
      R mu = p.dir[2];
      R phi = M::atan2_(p.dir[1], p.dir[0]);
      R mu2 = mu + bw * (R)m.scaling[0] * rand_type<R, S, FAST>(s, kernel);
      if (mu >= 0) {
        while ((mu2 < 0) || (mu2 > 1)) {
          if (mu2 < 0) mu2 *= -1;
          if (mu2 > 1) mu2 -= 2 * (mu2 - 1);
        }
      } else {
        while ((mu2 < -1) || (mu2 > 0)) {
          if (mu2 < -1) mu2 += 2 * (-1 - mu2);
          if (mu2 > 0) mu2 *= -1;
        }
      }
      mu = mu2;
      phi += bw * (R)m.scaling[1] * PI180 * rand_type<R, S, FAST>(s, kernel);
      const R sq = M::sqrt_(1 - mu * mu);
      p.dir[0] = sq * M::cos_(phi);
      p.dir[1] = sq * M::sin_(phi);
      p.dir[2] = mu;
    } break;
    default:
      break;
  }
}

}  // namespace kdsb
