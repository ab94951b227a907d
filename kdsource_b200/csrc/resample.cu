// resample: weighted particle pick (fixed-point prefix sum + binary search) followed
// by the per-metric kernel perturbation of clib/src/geom.c:139-650, one output
// particle per thread, driven by counter-based Philox4x32-10 (or by caller-supplied
// uniforms in the test mode).  Replaces the serial loop of
// clib/src/resamplemcpl.c:40-43 over KDS_rand_sample2 (clib/src/kdsource.c:282-330)
// and packs the result straight into 36-byte single-precision MCPL-3 records, the
// format mcpl_add_particle writes for resamplemcpl.c:31-42.
//
// The pick is this library's own contract (the reference walks the list
// sequentially); the perturbation maps, deviate maps (clib/src/utils.c:9-67) and
// uniform-consumption order follow the reference.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "kdsb200.h"

namespace kdsb {

#define KDS_PI 3.1415926535897932384626433832795028841971694

enum {
  M_ENERGY = 0, M_LETHARGY, M_WAVELENGTH, M_VOL, M_SURFXY, M_SURFR, M_SURFR2, M_SURFCIRCLE,
  M_GUIDE, M_ISOTROP, M_POLAR, M_POLARMU, M_TIME, M_DECADE
};  // order of _metric_names, clib/src/geom.c:13-16

// ---- weights -> fixed-point CDF -------------------------------------------
// c_i = floor(w_i * 2^62 / sum w) (>= 1 for w_i > 0); inclusive prefix sums in
// uint64 are exactly associative, so any scan order gives the same CDF.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint64_t quantise_w(double w, double scale) {
  uint64_t c = (uint64_t)(w * scale);
  if (c == 0 && w > 0) c = 1;
  return c;
}

__device__ uint64_t block_exclusive_scan(uint64_t v, uint64_t* total) {
  __shared__ uint64_t wsum[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint64_t woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SCAN_THREADS / 32; w++) {
    if (w < warp) woff += wsum[w];
    tot += wsum[w];
  }
  __syncthreads();
  *total = tot;
  return woff + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
    cdf_block_sums_kernel(const double* __restrict__ w, uint64_t N, double scale,
                          uint64_t* __restrict__ bsum) {
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint64_t s = 0;
#pragma unroll
  for (int e = 0; e < SCAN_ITEMS; e++)
    if (base + e < N) s += quantise_w(w[base + e], scale);
  uint64_t tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// single-block exclusive scan of the per-block sums (in place)
__global__ void __launch_bounds__(SCAN_THREADS)
    cdf_scan_sums_kernel(uint64_t* __restrict__ bsum, uint64_t nblocks) {
  uint64_t carry = 0;
  for (uint64_t base = 0; base < nblocks; base += SCAN_THREADS) {
    const uint64_t i = base + threadIdx.x;
    const uint64_t v = i < nblocks ? bsum[i] : 0;
    uint64_t tot;
    const uint64_t ex = block_exclusive_scan(v, &tot);
    if (i < nblocks) bsum[i] = carry + ex;
    carry += tot;
  }
}

__global__ void __launch_bounds__(SCAN_THREADS)
    cdf_write_kernel(const double* __restrict__ w, uint64_t N, double scale,
                     const uint64_t* __restrict__ bsum, uint64_t* __restrict__ cdf) {
  const uint64_t base = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint64_t c[SCAN_ITEMS];
  uint64_t s = 0;
#pragma unroll
  for (int e = 0; e < SCAN_ITEMS; e++) {
    c[e] = base + e < N ? quantise_w(w[base + e], scale) : 0;
    s += c[e];
  }
  uint64_t tot;
  uint64_t run = bsum[blockIdx.x] + block_exclusive_scan(s, &tot);
#pragma unroll
  for (int e = 0; e < SCAN_ITEMS; e++) {
    run += c[e];
    if (base + e < N) cdf[base + e] = run;
  }
}

// guide table: guide[b] = first i with cdf[i] > (b << shift), b = 0..K-1.  A pick
// target T falls in bucket b = min(T >> shift, K-1) and its index lies in
// [guide[b], guide[b+1]] (N-1 for the last bucket), so the binary search starts
// from a window of ~N/K entries instead of N.  Same result as the full search.
__global__ void cdf_guide_kernel(const uint64_t* __restrict__ cdf, uint64_t N, int bits,
                                 uint32_t* __restrict__ guide) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t K = 1ull << bits;
  if (b > K) return;
  if (b == K) {
    guide[b] = (uint32_t)(N - 1);
    return;
  }
  const uint64_t T = b << (62 - bits);
  uint64_t lo = 0, hi = N;
  while (lo < hi) {
    const uint64_t mid = lo + ((hi - lo) >> 1);
    if (cdf[mid] > T)
      hi = mid;
    else
      lo = mid + 1;
  }
  guide[b] = (uint32_t)(lo < N ? lo : N - 1);
}

// ---- Philox4x32-10 --------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// per-particle uniform stream: Philox counter (index, block), key = seed; block 0
// words 0,1 form the 64-bit pick, later words give u = ((w >> 9) + 0.5) 2^-23;
// external mode reads U[p*slots + (n % slots)] instead (slot 0 is the pick)
constexpr int RS_THREADS = 256;
constexpr int RS_RING = 16;  // Philox words kept per thread in shared memory (4 blocks)

// rare path of Stream::word(): a lane ran past the words generated up front.  Out
// of line (by value, so the stream state stays in registers) to keep the
// ~60-instruction generator from being replicated at every call site.
__device__ __noinline__ void philox_refill(uint32_t p0, uint32_t p1, uint32_t blk, uint32_t k0,
                                           uint32_t k1, uint32_t* ring) {
  uint32_t o[4];
  philox4x32_10(p0, p1, blk, 0u, k0, k1, o);
  const int s0 = (blk * 4) % RS_RING;
#pragma unroll
  for (int k = 0; k < 4; k++) ring[(s0 + k) * RS_THREADS] = o[k];
}

template <bool EXT>
struct Stream {
  uint32_t k0, k1, p0, p1;
  uint32_t* ring;  // shared memory, word w of this thread at ring[(w % RS_RING) * RS_THREADS]
  const double* ext;
  int slots, next;

  // The first RS_RING words (Philox blocks 0..3) are generated up front by all
  // lanes in lock-step; only a lane that runs past them generates further blocks.
  __device__ __forceinline__ void init(uint64_t seed, uint64_t p, const double* e, int nslots,
                                       uint32_t* ring_base) {
    next = 0;
    if (EXT) {
      ext = e + p * (uint64_t)nslots;
      slots = nslots;
    } else {
      k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
      p0 = (uint32_t)p; p1 = (uint32_t)(p >> 32);
      ring = ring_base + threadIdx.x;
#pragma unroll
      for (uint32_t b = 0; b < RS_RING / 4; b++) fill(b);
    }
  }
  __device__ __forceinline__ void fill(uint32_t blk) {
    uint32_t o[4];
    philox4x32_10(p0, p1, blk, 0u, k0, k1, o);
    const int s0 = (blk * 4) % RS_RING;
#pragma unroll
    for (int k = 0; k < 4; k++) ring[(s0 + k) * RS_THREADS] = o[k];
  }
  // rare: a lane ran past the words generated up front; kept out of line so the
  // ~60-instruction generator is not replicated at every call site of word()
  __device__ __forceinline__ uint32_t word() {
    if (next >= RS_RING && (next & 3) == 0)
      philox_refill(p0, p1, (uint32_t)next >> 2, k0, k1, ring);
    const uint32_t w = ring[(next % RS_RING) * RS_THREADS];
    next++;
    return w;
  }
  __device__ __forceinline__ double ext_next() {
    const double u = ext[next % slots];
    next++;
    return u;
  }
  __device__ __forceinline__ uint64_t pick_target(uint64_t total) {
    if (EXT) {
      uint64_t T = (uint64_t)(ext_next() * (double)total);
      return T >= total ? total - 1 : T;
    }
    const uint64_t hi = word(), lo = word();
    return __umul64hi((hi << 32) | lo, total);
  }
  // 23-bit integer k of the next uniform u = (k + 0.5) 2^-23 (Philox mode only)
  __device__ __forceinline__ int kint() { return (int)(word() >> 9); }
};

template <typename R>
struct Mth;
template <>
struct Mth<float> {
  static __device__ __forceinline__ float exp_(float x) { return expf(x); }
  static __device__ __forceinline__ float log_(float x) { return logf(x); }
  static __device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float sin_(float x) { return sinf(x); }
  static __device__ __forceinline__ float cos_(float x) { return cosf(x); }
  static __device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
  static __device__ __forceinline__ float acos_(float x) { return acosf(x); }
  static __device__ __forceinline__ float asin_(float x) { return asinf(x); }
  static __device__ __forceinline__ float pow10_(float x) { return exp10f(x); }
};
template <>
struct Mth<double> {
  static __device__ __forceinline__ double exp_(double x) { return exp(x); }
  static __device__ __forceinline__ double log_(double x) { return log(x); }
  static __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
  static __device__ __forceinline__ double sin_(double x) { return sin(x); }
  static __device__ __forceinline__ double cos_(double x) { return cos(x); }
  static __device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
  static __device__ __forceinline__ double acos_(double x) { return acos(x); }
  static __device__ __forceinline__ double asin_(double x) { return asin(x); }
  static __device__ __forceinline__ double pow10_(double x) { return pow(10.0, x); }
};

// FAST (fp32 only): MUFU-based approximations (lg2/ex2/rcp/rsq/sin/cos, abs. error
// ~2^-21) for the transcendental steps of the deviate and direction maps.  The
// accept/reject decisions stay exact, so the uniform consumption is unchanged and
// the perturbed coordinates stay within ~1e-6 of the variable's scale.
template <typename R, bool FAST>
struct Mf {
  static __device__ __forceinline__ R exp_(R x) { return Mth<R>::exp_(x); }
  static __device__ __forceinline__ R log_(R x) { return Mth<R>::log_(x); }
  static __device__ __forceinline__ R sqrt_(R x) { return Mth<R>::sqrt_(x); }
  static __device__ __forceinline__ R div_(R a, R b) { return a / b; }
};
template <>
struct Mf<float, true> {
  static __device__ __forceinline__ float exp_(float x) { return __expf(x); }
  static __device__ __forceinline__ float log_(float x) { return __logf(x); }
  static __device__ __forceinline__ float sqrt_(float x) { return __fsqrt_rn(x); }
  static __device__ __forceinline__ float div_(float a, float b) { return __fdividef(a, b); }
};

template <typename R, bool EXT>
__device__ __forceinline__ R uniform(Stream<EXT>& s) {
  if (EXT) return (R)s.ext_next();
  return ((R)s.kint() + (R)0.5) * (R)(1.0 / 8388608.0);
}

// kds_rand_norm: Marsaglia polar method, second value discarded (utils.c:9-29).
// Accept/reject is decided exactly (integers for Philox, fp64 for external
// uniforms) so fp32 and fp64 modes consume the same number of uniforms.
template <typename R, bool EXT, bool FAST>
__device__ R rand_norm(Stream<EXT>& s) {
  if (EXT) {
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
      // g1 * sqrt(-2 ln t / t) = g1 * sqrt(-2 ln2 * lg2 t) * rsqrt(t)
      return (R)(g1 * __fsqrt_rn(-1.3862943611f * __log2f(tf)) * rsqrtf(tf));
    }
    const float lt = ti > (1ll << 45)
                         ? log1pf(-(float)((1ll << 46) - ti) * (1.0f / 70368744177664.0f))
                         : logf(tf);
    return (R)(g1 * sqrtf((-2.0f * lt) / tf));
  }
}

// kds_rand_epan (utils.c:32-46)
template <typename R, bool EXT>
__device__ R rand_epan(Stream<EXT>& s) {
  const R x1 = uniform<R, EXT>(s) * (R)2 - (R)1;
  const R x2 = uniform<R, EXT>(s) * (R)2 - (R)1;
  const R x3 = uniform<R, EXT>(s) * (R)2 - (R)1;
  if (fabs(x3) >= fabs(x2) && fabs(x3) >= fabs(x1)) return x2;
  return x3;
}

// kds_rand_type (utils.c:55-67)
template <typename R, bool EXT, bool FAST>
__device__ R rand_type(Stream<EXT>& s, char kernel) {
  if (kernel == 'g') return rand_norm<R, EXT, FAST>(s);
  if (kernel == 'e') return rand_epan<R, EXT>(s);
  if (kernel == 'b') return uniform<R, EXT>(s) * (R)2 - (R)1;
  return (R)0;
}

template <typename R>
struct Part {
  R ekin, pos[3], dir[3], time;
};

template <typename R>
__device__ __forceinline__ void reflect(R& x, R lo, R hi) {
  while ((x < lo) || (x > hi)) {
    if (x < lo)
      x += 2 * (lo - x);
    else
      x -= 2 * (x - hi);
  }
}

// rotv (Rodrigues, utils.c:82-103)
template <typename R>
__device__ void rotv(R* v, const double* rv, int inverse) {
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
template <typename R, bool EXT, bool FAST>
__device__ void vmf_perturb(Stream<EXT>& s, R bw, R& dx, R& dy, R& dz) {
  using M = Mth<R>;
  using F = Mf<R, FAST>;
  const R xi = uniform<R, EXT>(s);
  R w, uv;
  if (sizeof(R) == 8) {
    w = 1;
    w += bw * bw * M::log_(xi + (1 - xi) * M::exp_(-2 / (bw * bw)));
    uv = M::sqrt_(1 - w * w);
  } else {
    // same map, arranged so 1 - w*w does not cancel in fp32
    const R m = -bw * bw * F::log_(xi + (1 - xi) * F::exp_(F::div_((R)-2, bw * bw)));
    w = 1 - m;
    uv = F::sqrt_(fmax(m * (2 - m), (R)0));
  }
  const R u01 = uniform<R, EXT>(s);
  R sphi, cphi;
  if (sizeof(R) == 8) {
    const R phi = (R)(2. * KDS_PI) * u01;
    sphi = M::sin_(phi);
    cphi = M::cos_(phi);
  } else if (FAST) {
    // angle folded to [-pi, pi) before the MUFU sine/cosine
    const float ang = 6.2831853072f * ((float)u01 - ((float)u01 >= 0.5f ? 1.0f : 0.0f));
    sphi = __sinf(ang);
    cphi = __cosf(ang);
  } else {
    float sf, cf;
    sincospif(2.0f * (float)u01, &sf, &cf);
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
template <typename R, bool EXT, bool FAST>
__device__ void guide_perturb(Stream<EXT>& s, const kdsb200_metric& m, Part<R>& p, R bw,
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
  z += bw * (R)m.scaling[0] * rand_type<R, EXT, FAST>(s, kernel);
  reflect<R>(z, (R)0, zmax);
  t += bw * (R)m.scaling[1] * rand_type<R, EXT, FAST>(s, kernel);
  reflect<R>(t, lo, hi);
  if (isinf(m.scaling[2]))
    mu = -1 + 2 * uniform<R, EXT>(s);
  else {
    mu2 = mu + bw * (R)m.scaling[2] * rand_type<R, EXT, FAST>(s, kernel);
    if (mu >= 0)
      reflect<R>(mu2, (R)0, (R)1);
    else
      reflect<R>(mu2, (R)-1, (R)0);
    mu = mu2;
  }
  if (isinf(m.scaling[3]))
    phi = (R)(2. * KDS_PI) * uniform<R, EXT>(s);
  else
    phi += bw * (R)m.scaling[3] * (R)(KDS_PI / 180) * rand_type<R, EXT, FAST>(s, kernel);
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
template <typename R, bool EXT, int TYPE, bool FAST>
__device__ __forceinline__ void metric_perturb(Stream<EXT>& s, const kdsb200_metric& m,
                                               Part<R>& p, R bw, char kernel) {
  using M = Mth<R>;
  const R PI180 = (R)(KDS_PI / 180);
  switch (TYPE >= 0 ? TYPE : m.type) {
    case M_ENERGY:  // E_perturb geom.c:192-198
      p.ekin += bw * (R)m.scaling[0] * rand_type<R, EXT, FAST>(s, kernel);
      if (p.ekin < 0) p.ekin *= -1;
      break;
    case M_LETHARGY: {  // Let_perturb geom.c:200-210
      R E = p.ekin * Mf<R, FAST>::exp_(bw * (R)m.scaling[0] * rand_type<R, EXT, FAST>(s, kernel));
      while (E > (R)m.params[0])
        E = p.ekin * Mf<R, FAST>::exp_(bw * (R)m.scaling[0] * rand_type<R, EXT, FAST>(s, kernel));
      p.ekin = E;
    } break;
    case M_WAVELENGTH: {  // wl_perturb geom.c:212-221
      const R wl = (R)9.045 / M::sqrt_(p.ekin * (R)1e9);
      const R wl2 = wl + bw * (R)m.scaling[0] * rand_type<R, EXT, FAST>(s, kernel);
      p.ekin = (R)81.82e-9 / (wl2 * wl2);
      if (p.ekin < 0) p.ekin *= -1;
    } break;
    case M_TIME:  // t_perturb geom.c:223-227
      p.time += bw * (R)m.scaling[0] * rand_type<R, EXT, FAST>(s, kernel);
      break;
    case M_DECADE:  // Dec_perturb geom.c:228-233
      p.time *= M::pow10_(bw * (R)m.scaling[0] * rand_type<R, EXT, FAST>(s, kernel));
      break;
    case M_VOL:  // Vol_perturb geom.c:235-265
#pragma unroll
      for (int a = 0; a < 3; a++) {
        p.pos[a] += bw * (R)m.scaling[a] * rand_type<R, EXT, FAST>(s, kernel);
        reflect<R>(p.pos[a], (R)m.params[2 * a], (R)m.params[2 * a + 1]);
      }
      break;
    case M_SURFXY:  // SurfXY_perturb geom.c:266-287
#pragma unroll
      for (int a = 0; a < 2; a++) {
        p.pos[a] += bw * (R)m.scaling[a] * rand_type<R, EXT, FAST>(s, kernel);
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
      R rho2 = rho + bw * (R)m.scaling[0] * rand_type<R, EXT, FAST>(s, kernel);
      R psi2 = psi + bw * (R)m.scaling[1] * PI180 * rand_type<R, EXT, FAST>(s, kernel);
      reflect<R>(rho2, rmin, rmax);
      reflect<R>(psi2, (R)m.params[2], (R)m.params[3]);
      if (sq) rho2 = M::sqrt_(rho2);
      p.pos[0] = rho2 * M::cos_(psi2);
      p.pos[1] = rho2 * M::sin_(psi2);
    } break;
    case M_SURFCIRCLE: {  // SurfCircle_perturb geom.c:355-385
      const R x = p.pos[0] + bw * (R)m.scaling[0] * rand_type<R, EXT, FAST>(s, kernel);
      const R y = p.pos[1] + bw * (R)m.scaling[1] * rand_type<R, EXT, FAST>(s, kernel);
      R rho2 = M::sqrt_(x * x + y * y);
      R psi2 = M::atan2_(y, x);
      reflect<R>(rho2, (R)m.params[0], (R)m.params[1]);
      reflect<R>(psi2, (R)m.params[2], (R)m.params[3]);
      p.pos[0] = rho2 * M::cos_(psi2);
      p.pos[1] = rho2 * M::sin_(psi2);
    } break;
    case M_GUIDE:
      guide_perturb<R, EXT, FAST>(s, m, p, bw, kernel);
      break;
    case M_ISOTROP: {  // Isotrop_perturb geom.c:575-600 (kernel ignored)
      const R sc = bw * (R)m.scaling[0];
      if (isinf(sc)) {
        p.dir[2] = -1 + 2 * uniform<R, EXT>(s);
        const R dxy = M::sqrt_(1 - p.dir[2] * p.dir[2]);
        const R phi = (R)(2. * KDS_PI) * uniform<R, EXT>(s);
        p.dir[0] = dxy * M::cos_(phi);
        p.dir[1] = dxy * M::sin_(phi);
      } else if (sc > 0) {
        const R dx = p.dir[0], dy = p.dir[1], dz = p.dir[2];
        vmf_perturb<R, EXT, FAST>(s, sc, p.dir[0], p.dir[1], p.dir[2]);
        if ((int)m.params[0] && (dx * p.dir[0] < 0)) p.dir[0] *= -1;
        if ((int)m.params[1] && (dy * p.dir[1] < 0)) p.dir[1] *= -1;
        if ((int)m.params[2] && (dz * p.dir[2] < 0)) p.dir[2] *= -1;
      }
    } break;
    case M_POLAR: {  // Polar_perturb geom.c:602-620 (reflected theta is overwritten there too)
      R theta = M::acos_(p.dir[2]);
      R phi = M::atan2_(p.dir[1], p.dir[0]);
      theta = theta + bw * (R)m.scaling[0] * PI180 * rand_type<R, EXT, FAST>(s, kernel);
      phi += bw * (R)m.scaling[1] * PI180 * rand_type<R, EXT, FAST>(s, kernel);
      p.dir[0] = M::sin_(theta) * M::cos_(phi);
      p.dir[1] = M::sin_(theta) * M::sin_(phi);
      p.dir[2] = M::cos_(theta);
    } break;
    case M_POLARMU: {  // PolarMu_perturb geom.c:622-650
      R mu = p.dir[2];
      R phi = M::atan2_(p.dir[1], p.dir[0]);
      R mu2 = mu + bw * (R)m.scaling[0] * rand_type<R, EXT, FAST>(s, kernel);
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
      phi += bw * (R)m.scaling[1] * PI180 * rand_type<R, EXT, FAST>(s, kernel);
      const R sq = M::sqrt_(1 - mu * mu);
      p.dir[0] = sq * M::cos_(phi);
      p.dir[1] = sq * M::sin_(phi);
      p.dir[2] = mu;
    } break;
    default:
      break;
  }
}

enum { SIG_GENERIC = 0, SIG_FLAT = 1, SIG_ACTIV = 2 };

template <typename R, bool EXT, int SIG, bool FAST>
__global__ void __launch_bounds__(RS_THREADS, sizeof(R) == 4 ? 4 : 2)
    resample_kernel(const R* __restrict__ src8, const uint64_t* __restrict__ cdf,
                    const uint32_t* __restrict__ guide, int guide_bits,
                    const float* __restrict__ bw, double bw_scalar, int64_t N, kdsb200_geom g,
                    int32_t pdg, uint64_t seed, uint64_t first, uint64_t count, int64_t seq_start,
                    const double* __restrict__ ext, int slots, int perturb,
                    unsigned char* __restrict__ out36, int64_t* __restrict__ out_idx,
                    double* __restrict__ out_parts) {
  __shared__ __align__(16) uint32_t stage[RS_THREADS * 9];
  __shared__ uint32_t ring[EXT ? 1 : RS_RING * RS_THREADS];
  const uint64_t n0 = (uint64_t)blockIdx.x * RS_THREADS;
  const uint64_t n = n0 + threadIdx.x;
  const bool live = n < count;
  uint32_t rec[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (live) {
  Stream<EXT> s;
  s.init(seed, EXT ? n : first + n, ext, slots, ring);
  const uint64_t total = __ldg(cdf + (N - 1));
  const uint64_t T = s.pick_target(total);
  // upper_bound: first i with cdf[i] > T (window from the guide table if present)
  int64_t lo = 0, hi = N - 1;
  if (guide) {
    const uint64_t K = 1ull << guide_bits;
    uint64_t b = T >> (62 - guide_bits);
    if (b > K - 1) b = K - 1;
    lo = __ldg(guide + b);
    hi = __ldg(guide + b + 1);
  }
  while (lo < hi) {
    const int64_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(cdf + mid) > T)
      hi = mid;
    else
      lo = mid + 1;
  }
  // seq_start >= 0: walk the list in order instead (w_crit <= 0 mode of
  // KDS_rand_sample2, kdsource.c:286-292); the pick number is still consumed
  const int64_t i = seq_start >= 0 ? (int64_t)(((uint64_t)seq_start + first + n) % (uint64_t)N) : lo;

  Part<R> p;
  {
    const R* r = src8 + i * 8;
    if (sizeof(R) == 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(r));
      const float4 b = __ldg(reinterpret_cast<const float4*>(r) + 1);
      p.ekin = a.x; p.pos[0] = a.y; p.pos[1] = a.z; p.pos[2] = a.w;
      p.dir[0] = b.x; p.dir[1] = b.y; p.dir[2] = b.z; p.time = b.w;
    } else {
      p.ekin = r[0]; p.pos[0] = r[1]; p.pos[1] = r[2]; p.pos[2] = r[3];
      p.dir[0] = r[4]; p.dir[1] = r[5]; p.dir[2] = r[6]; p.time = r[7];
    }
  }
  const R h = bw ? (R)__ldg(bw + i) : (R)bw_scalar;
  if (perturb) {  // Geom_perturb geom.c:139-158 (specialised signatures have no trasl/rot)
    if (SIG == SIG_GENERIC && g.has_trasl)
      for (int k = 0; k < 3; k++) p.pos[k] -= (R)g.trasl[k];
    if (SIG == SIG_GENERIC && g.has_rot) {
      rotv<R>(p.pos, g.rot, 1);
      rotv<R>(p.dir, g.rot, 1);
    }
    if (SIG == SIG_FLAT) {          // Lethargy + SurfXY + Isotrop (GeomFlat)
      metric_perturb<R, EXT, M_LETHARGY, FAST>(s, g.ms[0], p, h, 'g');
      metric_perturb<R, EXT, M_SURFXY, FAST>(s, g.ms[1], p, h, 'g');
      metric_perturb<R, EXT, M_ISOTROP, FAST>(s, g.ms[2], p, h, 'g');
    } else if (SIG == SIG_ACTIV) {  // Energy + Vol + Isotrop (GeomActiv)
      metric_perturb<R, EXT, M_ENERGY, FAST>(s, g.ms[0], p, h, 'g');
      metric_perturb<R, EXT, M_VOL, FAST>(s, g.ms[1], p, h, 'g');
      metric_perturb<R, EXT, M_ISOTROP, FAST>(s, g.ms[2], p, h, 'g');
    } else {
      for (int k = 0; k < g.order; k++) metric_perturb<R, EXT, -1, FAST>(s, g.ms[k], p, h, g.kernel);
    }
    if (SIG == SIG_GENERIC && g.has_rot) {
      rotv<R>(p.pos, g.rot, 0);
      rotv<R>(p.dir, g.rot, 0);
    }
    if (SIG == SIG_GENERIC && g.has_trasl)
      for (int k = 0; k < 3; k++) p.pos[k] += (R)g.trasl[k];
  }
  if (out_idx) out_idx[n] = i;
  if (out_parts) {
    double* o = out_parts + n * 8;
    o[0] = p.ekin; o[1] = p.pos[0]; o[2] = p.pos[1]; o[3] = p.pos[2];
    o[4] = p.dir[0]; o[5] = p.dir[1]; o[6] = p.dir[2]; o[7] = p.time;
  }
  if (out36) {
    // MCPL adaptive projection packing (libmcpl mcpl_unitvect_pack_adaptproj)
    const R ax = fabs(p.dir[0]), ay = fabs(p.dir[1]), az = fabs(p.dir[2]);
    R p0, p1, sg;
    if (az < fmax(ax, ay)) {
      const R invz = p.dir[2] != 0 ? Mf<R, FAST>::div_((R)1, p.dir[2]) : (R)INFINITY;
      if (ax >= ay) { p0 = invz; p1 = p.dir[1]; sg = p.dir[0]; }
      else { p0 = p.dir[0]; p1 = invz; sg = p.dir[1]; }
    } else {
      p0 = p.dir[0]; p1 = p.dir[1]; sg = p.dir[2];
    }
    rec[0] = __float_as_uint((float)p.pos[0]);
    rec[1] = __float_as_uint((float)p.pos[1]);
    rec[2] = __float_as_uint((float)p.pos[2]);
    rec[3] = __float_as_uint((float)p0);
    rec[4] = __float_as_uint((float)p1);
    rec[5] = __float_as_uint((float)copysign(fabs(p.ekin), sg));
    rec[6] = __float_as_uint((float)p.time);
    rec[7] = __float_as_uint(1.0f);  // weight = 1/bias, kdsource.c:325
    rec[8] = (uint32_t)pdg;
  }
  }  // live
  if (out36) {
    // stage the CTA's records in shared memory (stride 9 words: conflict-free) and
    // write them out as one contiguous, 16-byte-vectorised block
#pragma unroll
    for (int k = 0; k < 9; k++) stage[threadIdx.x * 9 + k] = rec[k];
    __syncthreads();
    const uint64_t nblk = count - n0 < RS_THREADS ? count - n0 : RS_THREADS;
    const uint32_t words = (uint32_t)nblk * 9;
    unsigned char* base = out36 + n0 * 36;  // n0*36 is a multiple of 16
    if ((reinterpret_cast<uintptr_t>(base) & 15) == 0) {
      uint4* dst = reinterpret_cast<uint4*>(base);
      const uint4* srcv = reinterpret_cast<const uint4*>(stage);
      for (uint32_t v = threadIdx.x; v < words / 4; v += RS_THREADS) dst[v] = srcv[v];
      for (uint32_t wd = (words / 4) * 4 + threadIdx.x; wd < words; wd += RS_THREADS)
        reinterpret_cast<uint32_t*>(base)[wd] = stage[wd];
    } else {
      for (uint32_t wd = threadIdx.x; wd < words; wd += RS_THREADS)
        reinterpret_cast<uint32_t*>(base)[wd] = stage[wd];
    }
  }
}

template <typename R>
__global__ void convert_src_kernel(const double* __restrict__ src, uint64_t n8, R* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n8) out[i] = (R)src[i];
}

}  // namespace kdsb

using namespace kdsb;

extern "C" {

uint64_t kdsb200_cdf_scratch_words(uint64_t N) { return (N + SCAN_BLOCK - 1) / SCAN_BLOCK + 1; }

int kdsb200_weights_cdf(const double* d_w, uint64_t N, double sum_w, uint64_t* d_scratch,
                        uint64_t* d_cdf, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) return 0;
  if (!(sum_w > 0)) {
    set_error("weights_cdf: sum of weights must be positive");
    return 1;
  }
  const double scale = 4611686018427387904.0 / sum_w;
  const uint64_t nb = (N + SCAN_BLOCK - 1) / SCAN_BLOCK;
  cdf_block_sums_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(d_w, N, scale, d_scratch);
  KDSB_CUDA(cudaGetLastError());
  cdf_scan_sums_kernel<<<1, SCAN_THREADS, 0, st>>>(d_scratch, nb);
  KDSB_CUDA(cudaGetLastError());
  cdf_write_kernel<<<(unsigned)nb, SCAN_THREADS, 0, st>>>(d_w, N, scale, d_scratch, d_cdf);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

int kdsb200_cdf_guide(const uint64_t* d_cdf, uint64_t N, int bits, uint32_t* d_guide, void* stream) {
  if (bits < 1 || bits > 26) {
    set_error("cdf_guide: bits=%d out of range 1..26", bits);
    return 1;
  }
  if (N == 0 || N >= (1ull << 32)) {
    set_error("cdf_guide: N must be in 1..2^32-1");
    return 1;
  }
  const uint64_t K1 = (1ull << bits) + 1;
  cdf_guide_kernel<<<(unsigned)((K1 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_cdf, N, bits,
                                                                                  d_guide);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

int kdsb200_convert_particles(const double* d_src8, uint64_t N, int to_f32, void* d_out,
                              void* stream) {
  const uint64_t n8 = N * 8;
  if (!n8) return 0;
  if (to_f32)
    convert_src_kernel<float><<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_src8, n8, (float*)d_out);
  else
    convert_src_kernel<double><<<(unsigned)((n8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_src8, n8, (double*)d_out);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

int kdsb200_resample(const void* d_src8, int src_is_f32, const uint64_t* d_cdf,
                     const uint32_t* d_guide, int guide_bits, const float* d_bw, double bw_scalar, int64_t N, const kdsb200_geom* h_geom, int32_t pdgcode,
                     uint64_t seed, uint64_t first, uint64_t count, int64_t seq_start,
                     const double* d_ext_uniforms, int slots, int perturb, void* d_out36,
                     int64_t* d_out_idx, double* d_out_parts, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (count == 0) return 0;
  if (N <= 0) {
    set_error("resample: empty source list");
    return 1;
  }
  if (h_geom->order < 0 || h_geom->order > KDSB200_MAX_METRICS) {
    set_error("resample: geometry order %d out of range", h_geom->order);
    return 1;
  }
  if (d_guide && (guide_bits < 1 || guide_bits > 26)) {
    set_error("resample: guide_bits=%d out of range 1..26", guide_bits);
    return 1;
  }
  if (d_ext_uniforms && slots < 1) {
    set_error("resample: external uniforms need slots >= 1");
    return 1;
  }
  const unsigned blocks = (unsigned)((count + 255) / 256);
  unsigned char* o36 = (unsigned char*)d_out36;
  // kernels specialised for the usual geometries (Gaussian kernel, no trasl/rot)
  int sig = SIG_GENERIC;
  if (h_geom->kernel == 'g' && !h_geom->has_trasl && !h_geom->has_rot && h_geom->order == 3 &&
      h_geom->ms[2].type == M_ISOTROP) {
    if (h_geom->ms[0].type == M_LETHARGY && h_geom->ms[1].type == M_SURFXY) sig = SIG_FLAT;
    if (h_geom->ms[0].type == M_ENERGY && h_geom->ms[1].type == M_VOL) sig = SIG_ACTIV;
  }
  // fp32 Philox kernels of the specialised signatures use the MUFU fast-math maps
  // unless KDSB200_RESAMPLE_PRECISE is set
  static int precise = -1;
  if (precise < 0) precise = getenv("KDSB200_RESAMPLE_PRECISE") ? 1 : 0;
  const bool fast = !precise && src_is_f32 && !d_ext_uniforms && sig != SIG_GENERIC;
#define KDSB_LAUNCH1(R, EXT, SIG)                                                                \
  if (fast && sizeof(R) == 4 && !EXT && SIG != SIG_GENERIC)                                      \
    resample_kernel<float, false, SIG == SIG_GENERIC ? SIG_FLAT : SIG, true>                     \
        <<<blocks, RS_THREADS, 0, st>>>(                                                         \
            (const float*)d_src8, d_cdf, d_guide, guide_bits, d_bw, bw_scalar, N, *h_geom,       \
            pdgcode, seed, first, count, seq_start, d_ext_uniforms, slots, perturb, o36,         \
            d_out_idx, d_out_parts);                                                             \
  else                                                                                           \
    resample_kernel<R, EXT, SIG, false><<<blocks, RS_THREADS, 0, st>>>(                          \
      (const R*)d_src8, d_cdf, d_guide, guide_bits, d_bw, bw_scalar, N, *h_geom, pdgcode, seed, \
      first, count, seq_start, d_ext_uniforms, slots, perturb, o36, d_out_idx, d_out_parts)
#define KDSB_LAUNCH(R, EXT)                                     \
  do {                                                          \
    if (sig == SIG_FLAT) KDSB_LAUNCH1(R, EXT, SIG_FLAT);        \
    else if (sig == SIG_ACTIV) KDSB_LAUNCH1(R, EXT, SIG_ACTIV); \
    else KDSB_LAUNCH1(R, EXT, SIG_GENERIC);                     \
  } while (0)
  if (src_is_f32) {
    if (d_ext_uniforms) KDSB_LAUNCH(float, true);
    else KDSB_LAUNCH(float, false);
  } else {
    if (d_ext_uniforms) KDSB_LAUNCH(double, true);
    else KDSB_LAUNCH(double, false);
  }
#undef KDSB_LAUNCH1
#undef KDSB_LAUNCH
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
