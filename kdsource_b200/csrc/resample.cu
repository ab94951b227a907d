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
#include "perturb.cuh"

namespace kdsb {

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
// Philox words generated up front per thread and kept in shared memory: 4 blocks under
// contract 1 (the polar method with rejection needs ~12 words per particle of the usual
// geometries), 2 blocks under contract 2 (Box-Muller: exactly 8)
__host__ __device__ constexpr int rs_ring(bool v2) { return v2 ? 8 : 16; }

// rare path of Stream::word(): a lane ran past the words generated up front.  Out
// of line (by value, so the stream state stays in registers) to keep the
// ~60-instruction generator from being replicated at every call site.
__device__ __noinline__ void philox_refill(uint32_t p0, uint32_t p1, uint32_t blk, uint32_t k0,
                                           uint32_t k1, uint32_t* ring, int ring_words) {
  uint32_t o[4];
  philox4x32_10(p0, p1, blk, 0u, k0, k1, o);
  const int s0 = (blk * 4) % ring_words;
#pragma unroll
  for (int k = 0; k < 4; k++) ring[(s0 + k) * RS_THREADS] = o[k];
}

// V2_: uniform-consumption contract 2 (see rand_norm in perturb.cuh): Gaussian deviates by
// Box-Muller, both values of a pair used (`spare`), no rejection.
template <bool EXT_, bool V2_, typename R>
struct Stream {
  static constexpr bool EXT = EXT_;
  static constexpr bool V2 = V2_ && !EXT_;
  static constexpr int RS_RING = rs_ring(V2);
  uint32_t k0, k1, p0, p1;
  uint32_t* ring;  // shared memory, word w of this thread at ring[(w % RS_RING) * RS_THREADS]
  const double* ext;
  int slots, next;
  R spare;            // second value of the last Box-Muller pair (contract 2)
  bool have_spare;

  // The first RS_RING words (Philox blocks 0..3) are generated up front by all
  // lanes in lock-step; only a lane that runs past them generates further blocks.
  __device__ __forceinline__ void init(uint64_t seed, uint64_t p, const double* e, int nslots,
                                       uint32_t* ring_base) {
    next = 0;
    have_spare = false;
    spare = (R)0;
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
      philox_refill(p0, p1, (uint32_t)next >> 2, k0, k1, ring, RS_RING);
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

enum { SIG_GENERIC = 0, SIG_FLAT = 1, SIG_ACTIV = 2 };

template <typename R, bool EXT, int SIG, bool FAST, bool V2>
__global__ void __launch_bounds__(RS_THREADS, sizeof(R) == 4 ? 4 : 2)
    resample_kernel(const R* __restrict__ src8, const uint64_t* __restrict__ cdf,
                    const uint32_t* __restrict__ guide, int guide_bits,
                    const float* __restrict__ bw, double bw_scalar, int64_t N, kdsb200_geom g,
                    int32_t pdg, uint64_t seed, uint64_t first, uint64_t count, int64_t seq_start,
                    const double* __restrict__ ext, int slots, int perturb,
                    unsigned char* __restrict__ out36, int64_t* __restrict__ out_idx,
                    double* __restrict__ out_parts) {
  __shared__ __align__(16) uint32_t stage[RS_THREADS * 9];
  using St = Stream<EXT, V2, R>;
  __shared__ uint32_t ring[EXT ? 1 : St::RS_RING * RS_THREADS];
  const uint64_t n0 = (uint64_t)blockIdx.x * RS_THREADS;
  const uint64_t n = n0 + threadIdx.x;
  const bool live = n < count;
  uint32_t rec[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (live) {
  St s;
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
      metric_perturb<R, St, M_LETHARGY, FAST>(s, g.ms[0], p, h, 'g');
      metric_perturb<R, St, M_SURFXY, FAST>(s, g.ms[1], p, h, 'g');
      metric_perturb<R, St, M_ISOTROP, FAST>(s, g.ms[2], p, h, 'g');
    } else if (SIG == SIG_ACTIV) {  // Energy + Vol + Isotrop (GeomActiv)
      metric_perturb<R, St, M_ENERGY, FAST>(s, g.ms[0], p, h, 'g');
      metric_perturb<R, St, M_VOL, FAST>(s, g.ms[1], p, h, 'g');
      metric_perturb<R, St, M_ISOTROP, FAST>(s, g.ms[2], p, h, 'g');
    } else {
      for (int k = 0; k < g.order; k++) metric_perturb<R, St, -1, FAST>(s, g.ms[k], p, h, g.kernel);
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

static int geom_signature(const kdsb200_geom* g) {
  if (g->kernel == 'g' && !g->has_trasl && !g->has_rot && g->order == 3 &&
      g->ms[2].type == M_ISOTROP) {
    if (g->ms[0].type == M_LETHARGY && g->ms[1].type == M_SURFXY) return SIG_FLAT;
    if (g->ms[0].type == M_ENERGY && g->ms[1].type == M_VOL) return SIG_ACTIV;
  }
  return SIG_GENERIC;
}

extern "C" {

int kdsb200_resample_fp32_ok(const kdsb200_geom* h_geom) {
  return geom_signature(h_geom) != SIG_GENERIC;
}

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
                     const double* d_ext_uniforms, int slots, int perturb, int contract,
                     void* d_out36, int64_t* d_out_idx, double* d_out_parts, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (count == 0) return 0;
  if (contract != 1 && contract != 2) {
    set_error("resample: uniform-consumption contract %d (1 or 2)", contract);
    return 1;
  }
  const bool v2 = contract == 2;
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
  const int sig = geom_signature(h_geom);
  // fp32 Philox kernels of the specialised signatures use the MUFU fast-math maps
  // unless KDSB200_RESAMPLE_PRECISE is set
  static int precise = -1;
  if (precise < 0) precise = getenv("KDSB200_RESAMPLE_PRECISE") ? 1 : 0;
  const bool fast = !precise && src_is_f32 && !d_ext_uniforms && sig != SIG_GENERIC;
#define KDSB_ARGS(R)                                                                            \
  (const R*)d_src8, d_cdf, d_guide, guide_bits, d_bw, bw_scalar, N, *h_geom, pdgcode, seed,     \
      first, count, seq_start, d_ext_uniforms, slots, perturb, o36, d_out_idx, d_out_parts
#define KDSB_LAUNCH2(R, EXT, SIG, V2)                                                           \
  if (fast && sizeof(R) == 4 && !EXT && SIG != SIG_GENERIC)                                     \
    resample_kernel<float, false, SIG == SIG_GENERIC ? SIG_FLAT : SIG, true, V2>                \
        <<<blocks, RS_THREADS, 0, st>>>(KDSB_ARGS(float));                                      \
  else                                                                                          \
    resample_kernel<R, EXT, SIG, false, V2><<<blocks, RS_THREADS, 0, st>>>(KDSB_ARGS(R))
#define KDSB_LAUNCH1(R, EXT, SIG)                   \
  do {                                              \
    if (v2 && !EXT) { KDSB_LAUNCH2(R, EXT, SIG, true); }  \
    else { KDSB_LAUNCH2(R, EXT, SIG, false); }      \
  } while (0)
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
#undef KDSB_ARGS
#undef KDSB_LAUNCH2
#undef KDSB_LAUNCH1
#undef KDSB_LAUNCH
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
