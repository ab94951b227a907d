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

// pick table: the guide table with the answer inlined.  A random access costs a whole
// 128-byte line of HBM on this GPU whatever its width (tools/probes/gather_probe.cu:
// 3.9 DRAM sectors per access for 16..128-byte reads, 4.6e10 lines/s), so the resampler's
// chain guide -> CDF probes -> particle -> bandwidth (5 lines per pick on a list that
// has outgrown the L2) is replaced by ONE line per bucket b of the pick range,
// K = 2^bits >= 2 N buckets:
//   words 0-1  thr0 = cdf[i0]       (~0 if no CDF boundary falls in the bucket)
//   words 2-3  thr1 = cdf[i0 + 1]   (~0 if fewer than two do)
//   word  4    i0   = first i with cdf[i] > (b << shift)
//   word  5    span = i1 - i0, i1 = first i with cdf[i] > ((b+1) << shift), <= N-1
//   words 6-7  bandwidths of particles i0 and i0 + 1
//   words 8-15 fp32 particle i0, words 16-23 fp32 particle i0 + 1, words 24-31 unused
// A pick target T of bucket b is particle i0 if T < thr0, i0 + 1 if T < thr1 -- both in
// hand -- and otherwise (a few percent of the picks) found in cdf[i0+2 .. i0+span]: the
// same index as the plain search in every case.
constexpr int PICK_WORDS = 32;

__device__ __forceinline__ uint64_t cdf_upper_bound(const uint64_t* __restrict__ cdf, uint64_t N,
                                                    uint64_t T) {
  uint64_t lo = 0, hi = N;
  while (lo < hi) {
    const uint64_t mid = lo + ((hi - lo) >> 1);
    if (__ldg(cdf + mid) > T)
      hi = mid;
    else
      lo = mid + 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
    pick_table_kernel(const uint64_t* __restrict__ cdf, uint64_t N, int bits,
                      const float* __restrict__ src8, const float* __restrict__ bw,
                      float bw_scalar, uint32_t* __restrict__ table) {
  const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t K = 1ull << bits;
  if (b >= K) return;
  const int shift = 62 - bits;
  uint64_t i0 = cdf_upper_bound(cdf, N, b << shift);
  if (i0 > N - 1) i0 = N - 1;
  uint64_t i1 = N - 1;
  if (b + 1 < K) {
    i1 = cdf_upper_bound(cdf, N, (b + 1) << shift);
    if (i1 > N - 1) i1 = N - 1;
  }
  const uint64_t ia = i0 + 1 < N ? i0 + 1 : N - 1;
  const uint64_t thr0 = i1 > i0 ? __ldg(cdf + i0) : ~0ull;
  const uint64_t thr1 = i1 > i0 + 1 ? __ldg(cdf + i0 + 1) : ~0ull;
  uint4* e = reinterpret_cast<uint4*>(table + b * PICK_WORDS);
  e[0] = make_uint4((uint32_t)thr0, (uint32_t)(thr0 >> 32), (uint32_t)thr1, (uint32_t)(thr1 >> 32));
  e[1] = make_uint4((uint32_t)i0, (uint32_t)(i1 - i0),
                    __float_as_uint(bw ? __ldg(bw + i0) : bw_scalar),
                    __float_as_uint(bw ? __ldg(bw + ia) : bw_scalar));
  const uint4* s0 = reinterpret_cast<const uint4*>(src8 + i0 * 8);
  const uint4* s1 = reinterpret_cast<const uint4*>(src8 + ia * 8);
  e[2] = __ldg(s0); e[3] = __ldg(s0 + 1);
  e[4] = __ldg(s1); e[5] = __ldg(s1 + 1);
  e[6] = make_uint4(0u, 0u, 0u, 0u);
  e[7] = make_uint4(0u, 0u, 0u, 0u);
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

// MCPL-3 single-precision record (36 bytes) with the adaptive projection packing of the
// direction (libmcpl mcpl_unitvect_pack_adaptproj); weight = 1/bias = 1, kdsource.c:325
template <typename R, bool FAST>
__device__ __forceinline__ void pack_mcpl_record(const Part<R>& p, int32_t pdg, uint32_t (&rec)[9]) {
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
  rec[7] = __float_as_uint(1.0f);
  rec[8] = (uint32_t)pdg;
}

enum { SIG_GENERIC = 0, SIG_FLAT = 1, SIG_ACTIV = 2 };

#ifndef KDSB_RS_MINB
#define KDSB_RS_MINB 5
#endif
template <typename R, bool EXT, int SIG, bool FAST, bool V2, bool TAB = false>
__global__ void __launch_bounds__(RS_THREADS, sizeof(R) == 4 ? (FAST ? KDSB_RS_MINB : 4) : 2)
    resample_kernel(const R* __restrict__ src8, const uint64_t* __restrict__ cdf,
                    const uint32_t* __restrict__ guide, int guide_bits,
                    const float* __restrict__ bw, double bw_scalar, int64_t N, kdsb200_geom g,
                    int32_t pdg, uint64_t seed, uint64_t first, uint64_t count, int64_t seq_start,
                    const double* __restrict__ ext, int slots, int perturb,
                    unsigned char* __restrict__ out36, int64_t* __restrict__ out_idx,
                    double* __restrict__ out_parts) {
  // TAB: one 128-byte table entry per lane, fetched by the warp (below); the records are
  // staged over the entries once these are consumed
  __shared__ __align__(128) uint32_t stage[RS_THREADS * (TAB ? PICK_WORDS : 9)];
  using St = Stream<EXT, V2, R>;
  __shared__ uint32_t ring[EXT ? 1 : St::RS_RING * RS_THREADS];
  const uint64_t n0 = (uint64_t)blockIdx.x * RS_THREADS;
  const uint64_t n = n0 + threadIdx.x;
  const bool live = n < count;
  uint32_t rec[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  // TAB: a lane past the end repeats the last particle (the table fetch is a warp
  // operation) and writes nothing
  const uint64_t ne = TAB && !live ? count - 1 : n;
  if (live || TAB) {
  St s;
  s.init(seed, EXT ? ne : first + ne, ext, slots, ring);
  const uint64_t total = __ldg(cdf + (N - 1));
  const uint64_t T = s.pick_target(total);
  // upper_bound: first i with cdf[i] > T (window from the guide table if present)
  int64_t lo = 0, hi = N - 1;
  bool inl = false;  // TAB: the picked particle came with the table entry
  float4 ea, eb;
  float eh;
  if (TAB) {
    // guide = pick table (pick_table_kernel).  On a list that has outgrown the L2 every
    // request that leaves the L1 for a random line costs an address translation, and the
    // translations bound the rate (tools/probes/gather_probe.cu: 4.6e10 sector requests/s
    // however few bytes each carries).  So eight lanes fetch one entry as ONE 128-byte
    // request (asynchronous copy into the warp's shared-memory area), four entries per
    // instruction, and each lane then reads its header and the chosen particle there.
    const uint64_t K = 1ull << guide_bits;
    uint64_t b = T >> (62 - guide_bits);
    if (b > K - 1) b = K - 1;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t wbase = smem_u32(stage) + (threadIdx.x & ~31u) * (PICK_WORDS * 4);
#pragma unroll
    for (uint32_t m = 0; m < 8; m++) {
      const uint32_t owner = (lane >> 3) + 4u * m, c = lane & 7u;
      const uint64_t bo = __shfl_sync(0xffffffffu, b, owner);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                       wbase + owner * 128u + ((c ^ (owner & 7u)) << 4)),
                   "l"(reinterpret_cast<const uint4*>(guide) + bo * (PICK_WORDS / 4) + c));
    }
    asm volatile("cp.async.commit_group;");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    const unsigned char* ent = reinterpret_cast<const unsigned char*>(stage) +
                               (size_t)threadIdx.x * 128u;
    const uint32_t sw = lane & 7u;
    const uint4 h0 = *reinterpret_cast<const uint4*>(ent + ((0u ^ sw) << 4));
    const uint4 h1 = *reinterpret_cast<const uint4*>(ent + ((1u ^ sw) << 4));
    const uint64_t thr0 = (uint64_t)h0.x | ((uint64_t)h0.y << 32);
    const uint64_t thr1 = (uint64_t)h0.z | ((uint64_t)h0.w << 32);
    lo = h1.x;
    hi = lo + h1.y;
    if (T < thr1 && seq_start < 0) {
      const uint32_t sel = T < thr0 ? 0u : 1u;
      lo += sel;
      hi = lo;
      inl = true;
      ea = *reinterpret_cast<const float4*>(ent + (((2u + 2u * sel) ^ sw) << 4));
      eb = *reinterpret_cast<const float4*>(ent + (((3u + 2u * sel) ^ sw) << 4));
      eh = __uint_as_float(sel ? h1.w : h1.z);
    } else if (T >= thr1) {
      lo += 2;
    }
    __syncwarp();  // entries consumed: the area is reused for the records
  } else if (guide) {
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
  const int64_t i = seq_start >= 0 ? (int64_t)(((uint64_t)seq_start + first + ne) % (uint64_t)N) : lo;

  Part<R> p;
  if (TAB && inl) {
    p.ekin = ea.x; p.pos[0] = ea.y; p.pos[1] = ea.z; p.pos[2] = ea.w;
    p.dir[0] = eb.x; p.dir[1] = eb.y; p.dir[2] = eb.z; p.time = eb.w;
  } else {
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
  const R h = (TAB && inl) ? (R)eh : bw ? (R)__ldg(bw + i) : (R)bw_scalar;
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
  if (out_idx && live) out_idx[n] = i;
  if (out_parts && live) {
    double* o = out_parts + n * 8;
    o[0] = p.ekin; o[1] = p.pos[0]; o[2] = p.pos[1]; o[3] = p.pos[2];
    o[4] = p.dir[0]; o[5] = p.dir[1]; o[6] = p.dir[2]; o[7] = p.time;
  }
  if (out36 && live) {
    pack_mcpl_record<R, FAST>(p, pdg, rec);
  }
  }  // live
  if (out36) {
    // stage each warp's 32 records in shared memory (stride 9 words: conflict-free) and
    // write them out as one contiguous block of 1152 bytes, 16 bytes per lane and store;
    // warps never wait for each other
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t* wst = stage + (threadIdx.x & ~31u) * (TAB ? PICK_WORDS : 9);
#pragma unroll
    for (int k = 0; k < 9; k++) wst[lane * 9 + k] = rec[k];
    __syncwarp();
    const uint64_t nw0 = n0 + (threadIdx.x & ~31u);
    if (nw0 < count) {
      const uint64_t nblk = count - nw0 < 32 ? count - nw0 : 32;
      const uint32_t words = (uint32_t)nblk * 9;
      unsigned char* base = out36 + nw0 * 36;  // nw0*36 is a multiple of 16
      if ((reinterpret_cast<uintptr_t>(base) & 15) == 0) {
        uint4* dst = reinterpret_cast<uint4*>(base);
        const uint4* srcv = reinterpret_cast<const uint4*>(wst);
        for (uint32_t v = lane; v < words / 4; v += 32) dst[v] = srcv[v];
        for (uint32_t wd = (words / 4) * 4 + lane; wd < words; wd += 32)
          reinterpret_cast<uint32_t*>(base)[wd] = wst[wd];
      } else {
        for (uint32_t wd = lane; wd < words; wd += 32)
          reinterpret_cast<uint32_t*>(base)[wd] = wst[wd];
      }
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

uint64_t kdsb200_pick_table_bytes(int bits) { return (1ull << bits) * PICK_WORDS * 4; }

int kdsb200_pick_table(const uint64_t* d_cdf, uint64_t N, int bits, const float* d_src8_f32,
                       const float* d_bw, double bw_scalar, void* d_table, void* stream) {
  if (bits < 1 || bits > 30) {
    set_error("pick_table: bits=%d out of range 1..30", bits);
    return 1;
  }
  if (N == 0 || N >= (1ull << 32)) {
    set_error("pick_table: N must be in 1..2^32-1");
    return 1;
  }
  const uint64_t K = 1ull << bits;
  pick_table_kernel<<<(unsigned)((K + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_cdf, N, bits, d_src8_f32, d_bw, (float)bw_scalar, (uint32_t*)d_table);
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
                     const void* d_guide_v, int guide_bits, int guide_kind, const float* d_bw,
                     double bw_scalar, int64_t N, const kdsb200_geom* h_geom, int32_t pdgcode,
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
  const uint32_t* d_guide = (const uint32_t*)d_guide_v;
  if (guide_kind != 0 && guide_kind != 1) {
    set_error("resample: guide_kind %d (0 = window table, 1 = pick table)", guide_kind);
    return 1;
  }
  if (d_guide && (guide_bits < 1 || guide_bits > (guide_kind ? 30 : 26))) {
    set_error("resample: guide_bits=%d out of range", guide_bits);
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
  // the pick table serves the fp32 Philox kernels of the specialised geometries; any other
  // variant ignores it and searches the whole CDF (same indices)
  const bool tab = d_guide && guide_kind == 1 && fast;
  if (guide_kind == 1 && !tab) d_guide = nullptr;
#define KDSB_ARGS(R)                                                                            \
  (const R*)d_src8, d_cdf, d_guide, guide_bits, d_bw, bw_scalar, N, *h_geom, pdgcode, seed,     \
      first, count, seq_start, d_ext_uniforms, slots, perturb, o36, d_out_idx, d_out_parts
#define KDSB_LAUNCH2(R, EXT, SIG, V2)                                                           \
  if (fast && tab && sizeof(R) == 4 && !EXT && SIG != SIG_GENERIC)                              \
    resample_kernel<float, false, SIG == SIG_GENERIC ? SIG_FLAT : SIG, true, V2, true>          \
        <<<blocks, RS_THREADS, 0, st>>>(KDSB_ARGS(float));                                      \
  else if (fast && sizeof(R) == 4 && !EXT && SIG != SIG_GENERIC)                                \
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
