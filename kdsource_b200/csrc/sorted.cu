// Spatially ordered source tiles: exact tile culling for the hard-cutoff sum and
// tile-local coordinates for the fp32 pair kernel.
//
// The reference's TreeKDE (call sites python/kdsource/kdsource.py:238-240,
// kde.py:146-151) only visits the sources inside the kernel's support radius.  Here the
// sources are ordered along a Hilbert curve when they are packed, so a tile of 512
// consecutive sources is spatially compact; every tile carries its bounding box and a
// centre, and coordinates are stored RELATIVE to that centre.  The evaluation kernel
//   * skips every tile whose box is farther than the cutoff from the box of the CTA's
//     1024 query points (which are ordered the same way) -- exact: a skipped tile holds no
//     pair inside the radius, and pairs inside visited tiles are still masked one by one;
//     a tile stays the unit of the bulk copy, but each of its sixteen runs of 32 sources
//     carries a box of its own, and so does each quarter of the CTA's queries (the qi-th
//     query of every thread: 256 consecutive points of the curve); a (run, quarter) cell
//     is computed only if its two boxes pass the same test -- a CTA-uniform decision, so no
//     lane diverges (in six dimensions a run of 32 is ~0.63 and a quarter ~0.8 of the
//     extent per axis);
//   * forms q - x as (q - c_tile) - (x - c_tile): both terms are small, so the fp32
//     rounding error scales with the tile size instead of the spread of the whole list
//     (the accuracy of the split-coordinate kernel at the (2D+1) lane-op cost).
//
// Pieces: column min/max, Hilbert keys (Skilling's transpose algorithm), an LSD radix
// sort (stable, 8 bits per pass, warp-private counters), the tile packer and the
// persistent evaluation kernel.
#include <math.h>

#include "common.cuh"
#include "kdsb200.h"

namespace kdsb {

// ---------------------------------------------------------------------------
// column min / max
// ---------------------------------------------------------------------------
constexpr int MM_BLOCKS = 592;
constexpr int MM_THREADS = 256;

__global__ void __launch_bounds__(MM_THREADS)
    minmax_partial_kernel(const double* __restrict__ data, uint64_t N, int D,
                          double* __restrict__ scratch) {
  double mn[8], mx[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    mn[k] = INFINITY;
    mx[k] = -INFINITY;
  }
  for (uint64_t i = (uint64_t)blockIdx.x * MM_THREADS + threadIdx.x; i < N;
       i += (uint64_t)MM_BLOCKS * MM_THREADS) {
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (k < D) {
        const double v = data[i * D + k];
        mn[k] = fmin(mn[k], v);
        mx[k] = fmax(mx[k], v);
      }
  }
  __shared__ double smn[8][MM_THREADS / 32], smx[8][MM_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    double a = mn[k], b = mx[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a = fmin(a, __shfl_down_sync(0xffffffffu, a, off));
      b = fmax(b, __shfl_down_sync(0xffffffffu, b, off));
    }
    if (lane == 0) {
      smn[k][warp] = a;
      smx[k][warp] = b;
    }
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double a = INFINITY, b = -INFINITY;
    for (int e = 0; e < MM_THREADS / 32; e++) {
      a = fmin(a, smn[threadIdx.x][e]);
      b = fmax(b, smx[threadIdx.x][e]);
    }
    scratch[blockIdx.x * 16 + threadIdx.x] = a;
    scratch[blockIdx.x * 16 + 8 + threadIdx.x] = b;
  }
}

__global__ void minmax_final_kernel(const double* __restrict__ scratch, int D,
                                    double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= 8) return;
  double a = INFINITY, b = -INFINITY;
  for (int blk = 0; blk < MM_BLOCKS; blk++) {
    a = fmin(a, scratch[blk * 16 + k]);
    b = fmax(b, scratch[blk * 16 + 8 + k]);
  }
  if (k < D) {
    out[k] = a;
    out[D + k] = b;
  }
}

// ---------------------------------------------------------------------------
// Hilbert keys
// ---------------------------------------------------------------------------
struct Quant {
  double lo[8], scale[8];  // cell = clamp(floor((x - lo) * scale), 0, 2^bits - 1)
  int bits;
};

// Skilling, "Programming the Hilbert curve" (2004): axes -> transposed Hilbert index,
// then the bits are interleaved (most significant first) into one integer.
template <int D>
__device__ __forceinline__ uint64_t hilbert_index(uint32_t (&X)[D], int b) {
  const uint32_t M = 1u << (b - 1);
  for (uint32_t Q = M; Q > 1; Q >>= 1) {
    const uint32_t P = Q - 1;
#pragma unroll
    for (int i = 0; i < D; i++) {
      if (X[i] & Q) {
        X[0] ^= P;
      } else {
        const uint32_t t = (X[0] ^ X[i]) & P;
        X[0] ^= t;
        X[i] ^= t;
      }
    }
  }
#pragma unroll
  for (int i = 1; i < D; i++) X[i] ^= X[i - 1];
  uint32_t t = 0;
  for (uint32_t Q = M; Q > 1; Q >>= 1)
    if (X[D - 1] & Q) t ^= Q - 1;
#pragma unroll
  for (int i = 0; i < D; i++) X[i] ^= t;
  uint64_t key = 0;
  for (int bit = b - 1; bit >= 0; bit--)
#pragma unroll
    for (int i = 0; i < D; i++) key = (key << 1) | ((X[i] >> bit) & 1u);
  return key;
}

template <int D>
__global__ void hilbert_keys_kernel(const double* __restrict__ data, uint64_t N, Quant q,
                                    uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  uint32_t X[D];
  const double top = (double)((1u << q.bits) - 1u);
#pragma unroll
  for (int k = 0; k < D; k++) {
    double c = floor((data[i * D + k] - q.lo[k]) * q.scale[k]);
    c = c < 0.0 ? 0.0 : (c > top ? top : c);  // NaN -> top
    X[k] = (uint32_t)c;
  }
  keys[i] = hilbert_index<D>(X, q.bits);
  idx[i] = (uint32_t)i;
}

// ---------------------------------------------------------------------------
// stable LSD radix sort of (uint64 key, uint32 value) pairs, 8 bits per pass.
// A warp owns a segment of 1024 consecutive elements and walks it 32 at a time, so the
// rank of an element among equal digits is: (offset of its digit for this segment, from
// the scan of all (digit, segment) counts) + (count in earlier steps of the segment) +
// (lanes below it with the same digit, from match.any).
// ---------------------------------------------------------------------------
constexpr int RS_SEG = 1024;
constexpr int RS_WARPS = 8;

__global__ void __launch_bounds__(RS_WARPS * 32)
    radix_count_kernel(const uint64_t* __restrict__ keys, uint64_t N, int shift, uint64_t nseg,
                       uint32_t* __restrict__ counts) {
  __shared__ uint32_t cnt[RS_WARPS][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
  __syncwarp();
  const uint64_t seg = (uint64_t)blockIdx.x * RS_WARPS + warp;
  if (seg >= nseg) return;
  const uint64_t base = seg * RS_SEG;
  for (int s = 0; s < RS_SEG / 32; s++) {
    const uint64_t i = base + s * 32 + lane;
    const uint32_t d = i < N ? (uint32_t)((keys[i] >> shift) & 255u) : 256u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (d < 256u && (peers & ((1u << lane) - 1u)) == 0) cnt[warp][d] += __popc(peers);
    __syncwarp();
  }
  for (int d = lane; d < 256; d += 32) counts[(uint64_t)d * nseg + seg] = cnt[warp][d];
}

__global__ void __launch_bounds__(RS_WARPS * 32)
    radix_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                         uint64_t N, int shift, uint64_t nseg, const uint32_t* __restrict__ offs,
                         uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  __shared__ uint32_t cnt[RS_WARPS][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t seg = (uint64_t)blockIdx.x * RS_WARPS + warp;
  if (seg >= nseg) return;
  for (int d = lane; d < 256; d += 32) cnt[warp][d] = offs[(uint64_t)d * nseg + seg];
  __syncwarp();
  const uint64_t base = seg * RS_SEG;
  for (int s = 0; s < RS_SEG / 32; s++) {
    const uint64_t i = base + s * 32 + lane;
    uint64_t k = 0;
    uint32_t d = 256u;
    if (i < N) {
      k = keys[i];
      d = (uint32_t)((k >> shift) & 255u);
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t below = __popc(peers & ((1u << lane) - 1u));
    uint32_t pos = 0;
    if (d < 256u) pos = cnt[warp][d] + below;
    __syncwarp();
    if (d < 256u && below == 0) cnt[warp][d] += __popc(peers);
    __syncwarp();
    if (d < 256u) {
      keys_out[pos] = k;
      vals_out[pos] = vals[i];
    }
  }
}

// exclusive scan of uint32 counters (total < 2^32), three kernels
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 16;
constexpr int SC_BLOCK = SC_THREADS * SC_ITEMS;

__device__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* total) {
  __shared__ uint32_t wsum[SC_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  uint32_t woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SC_THREADS / 32; w++) {
    if (w < warp) woff += wsum[w];
    tot += wsum[w];
  }
  __syncthreads();
  *total = tot;
  return woff + inc - v;
}

__global__ void __launch_bounds__(SC_THREADS)
    scan_block_sums_kernel(const uint32_t* __restrict__ a, uint64_t n,
                           uint32_t* __restrict__ bsum) {
  const uint64_t base = (uint64_t)blockIdx.x * SC_BLOCK + (uint64_t)threadIdx.x * SC_ITEMS;
  uint32_t s = 0;
#pragma unroll
  for (int e = 0; e < SC_ITEMS; e++)
    if (base + e < n) s += a[base + e];
  uint32_t tot;
  block_excl_scan_u32(s, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SC_THREADS)
    scan_sums_kernel(uint32_t* __restrict__ bsum, uint64_t nblocks) {
  uint32_t carry = 0;
  for (uint64_t base = 0; base < nblocks; base += SC_THREADS) {
    const uint64_t i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? bsum[i] : 0;
    uint32_t tot;
    const uint32_t ex = block_excl_scan_u32(v, &tot);
    if (i < nblocks) bsum[i] = carry + ex;
    carry += tot;
  }
}

__global__ void __launch_bounds__(SC_THREADS)
    scan_write_kernel(uint32_t* __restrict__ a, uint64_t n, const uint32_t* __restrict__ bsum) {
  const uint64_t base = (uint64_t)blockIdx.x * SC_BLOCK + (uint64_t)threadIdx.x * SC_ITEMS;
  uint32_t c[SC_ITEMS];
  uint32_t s = 0;
#pragma unroll
  for (int e = 0; e < SC_ITEMS; e++) {
    c[e] = base + e < n ? a[base + e] : 0;
    s += c[e];
  }
  uint32_t tot;
  uint32_t run = bsum[blockIdx.x] + block_excl_scan_u32(s, &tot);
#pragma unroll
  for (int e = 0; e < SC_ITEMS; e++) {
    if (base + e < n) a[base + e] = run;
    run += c[e];
  }
}

static int exclusive_scan_u32(uint32_t* d_a, uint64_t n, uint32_t* d_bsum, cudaStream_t st) {
  const uint64_t nb = (n + SC_BLOCK - 1) / SC_BLOCK;
  scan_block_sums_kernel<<<(unsigned)nb, SC_THREADS, 0, st>>>(d_a, n, d_bsum);
  KDSB_CUDA(cudaGetLastError());
  scan_sums_kernel<<<1, SC_THREADS, 0, st>>>(d_bsum, nb);
  KDSB_CUDA(cudaGetLastError());
  scan_write_kernel<<<(unsigned)nb, SC_THREADS, 0, st>>>(d_a, n, d_bsum);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------
// tile packer: one CTA per tile of TS sorted sources
// ---------------------------------------------------------------------------
struct Center8 {
  double c[8];
};

// rows as in common.cuh, coordinates relative to the tile centre:
//   rows 0..D-1 : -((x_k - c_k) - t_k) * a_i ; row D : a_i ; row D+1 : nlc_i (mode 0) or c_i (mode 1)
// tilebox[t] = {lo[D], hi[D]} of (x - c) over the tile (rounded outwards), followed by
// {lo[D], hi[D], hmax} of each of its SV_NSUB runs of TS / SV_NSUB sources (an empty run
// has lo = +inf); tilecen[t] = t_k (a multiple of 1/grid, exact in fp32).
constexpr int SV_NSUB = 16;
__host__ __device__ constexpr int tilebox_stride(int D) { return 2 * D + SV_NSUB * (2 * D + 1); }

template <int D>
__global__ void __launch_bounds__(TS)
    pack_sorted_kernel(const double* __restrict__ data, const double* __restrict__ w,
                       const double* __restrict__ bw, double bw_scalar, uint64_t N,
                       const uint32_t* __restrict__ perm, Center8 cen, double grid, double w_scale,
                       double lc_ref, int mode, float* __restrict__ packed,
                       float* __restrict__ tilebox, float* __restrict__ tilecen,
                       float* __restrict__ tilehmax) {
  __shared__ double smn[D][TS / 32], smx[D][TS / 32];
  __shared__ double shm[TS / 32];
  __shared__ double tc[D];
  const uint64_t t = blockIdx.x;
  const int j = threadIdx.x, lane = j & 31, warp = j >> 5;
  const uint64_t s = t * TS + j;
  const bool live = s < N;
  const uint64_t i = live ? perm[s] : 0;
  double v[D];
#pragma unroll
  for (int k = 0; k < D; k++) v[k] = live ? data[i * D + k] - cen.c[k] : 0.0;
  {
    double hm = live ? (bw ? bw[i] : bw_scalar) : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) hm = fmax(hm, __shfl_down_sync(0xffffffffu, hm, off));
    if (lane == 0) shm[warp] = hm;
  }
#pragma unroll
  for (int k = 0; k < D; k++) {
    double a = live ? v[k] : INFINITY, b = live ? v[k] : -INFINITY;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a = fmin(a, __shfl_down_sync(0xffffffffu, a, off));
      b = fmax(b, __shfl_down_sync(0xffffffffu, b, off));
    }
    if (lane == 0) {
      smn[k][warp] = a;
      smx[k][warp] = b;
    }
  }
  __syncthreads();
  if (j < D) {
    double a = INFINITY, b = -INFINITY;
    for (int e = 0; e < TS / 32; e++) {
      a = fmin(a, smn[j][e]);
      b = fmax(b, smx[j][e]);
    }
    // outward-rounded fp32 box
    float lo = (float)a, hi = (float)b;
    if ((double)lo > a) lo = nextafterf(lo, -INFINITY);
    if ((double)hi < b) hi = nextafterf(hi, INFINITY);
    tilebox[t * tilebox_stride(D) + j] = lo;
    tilebox[t * tilebox_stride(D) + D + j] = hi;
    const double mid = rint(0.5 * (a + b) * grid) / grid;
    tc[j] = mid;
    tilecen[t * D + j] = (float)mid;
  }
  if (j >= 32 && j < 32 + SV_NSUB * (D + 1)) {  // boxes and largest bandwidths of the runs
    constexpr int WPS = TS / 32 / SV_NSUB;  // warps per run
    const int sub = (j - 32) / (D + 1), k = (j - 32) % (D + 1);
    float* sb = tilebox + t * tilebox_stride(D) + 2 * D + sub * (2 * D + 1);
    if (k < D) {
      double a = INFINITY, b = -INFINITY;
      for (int e = sub * WPS; e < (sub + 1) * WPS; e++) {
        a = fmin(a, smn[k][e]);
        b = fmax(b, smx[k][e]);
      }
      float lo = (float)a, hi = (float)b;
      if ((double)lo > a) lo = nextafterf(lo, -INFINITY);
      if ((double)hi < b) hi = nextafterf(hi, INFINITY);
      sb[k] = lo;
      sb[D + k] = hi;
    } else {
      double hm = 0.0;
      for (int e = sub * WPS; e < (sub + 1) * WPS; e++) hm = fmax(hm, shm[e]);
      float hf = (float)hm;
      if ((double)hf < hm) hf = nextafterf(hf, INFINITY);
      sb[2 * D] = hf;
    }
  }
  if (j == TS - 1) {  // largest bandwidth of the tile, rounded up
    double hm = 0.0;
    for (int e = 0; e < TS / 32; e++) hm = fmax(hm, shm[e]);
    float hf = (float)hm;
    if ((double)hf < hm) hf = nextafterf(hf, INFINITY);
    tilehmax[t] = hf;
  }
  __syncthreads();
  constexpr int ROWS = D + 2;
  float* tile = packed + t * (uint64_t)ROWS * TS + j;
  if (!live) {
#pragma unroll
    for (int k = 0; k <= D; k++) tile[(uint64_t)k * TS] = 0.f;
    tile[(uint64_t)(D + 1) * TS] = mode == 0 ? INFINITY : 0.f;
    return;
  }
  const double h = bw ? bw[i] : bw_scalar;
  const double wi = (w ? w[i] : 1.0) * w_scale;
  const double a = (mode == 0 ? 0.84932180028801904272 /* sqrt(log2(e)/2) */ : 1.0) / h;
#pragma unroll
  for (int k = 0; k < D; k++) tile[(uint64_t)k * TS] = (float)(-(v[k] - tc[k]) * a);
  tile[(uint64_t)D * TS] = (float)a;
  if (mode == 0) {
    const double lc = log2(wi) - D * log2(h);
    tile[(uint64_t)(D + 1) * TS] = (float)(lc_ref - lc);
  } else {
    tile[(uint64_t)(D + 1) * TS] = (float)(wi * exp2(-D * log2(h) - lc_ref));
  }
}

// queries in sorted order, hi part on the grid: qT[k][m] = hi, qT[D+k][m] = lo;
// padding entries repeat the last real query so that they do not widen a CTA's box
__global__ void pack_queries_sorted_kernel(const double* __restrict__ pts, uint64_t M,
                                           uint64_t Mpad, int D, const uint32_t* __restrict__ perm,
                                           Center8 cen, double grid, float* __restrict__ qT) {
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mpad) return;
  const uint64_t src = perm[m < M ? m : M - 1];
  for (int k = 0; k < D; k++) {
    const double v = pts[src * D + k] - cen.c[k];
    const float hi = (float)(rint(v * grid) / grid);
    qT[(uint64_t)k * Mpad + m] = hi;
    qT[(uint64_t)(D + k) * Mpad + m] = (float)(v - (double)hi);
  }
}

// ---------------------------------------------------------------------------
// persistent evaluation kernel
// ---------------------------------------------------------------------------
constexpr int SV_THREADS = 256;
constexpr int SV_Q = 4;
constexpr int SV_QB = SV_THREADS * SV_Q;  // queries per work item
constexpr int SV_STAGES = 3;

struct SortedCtl {
  unsigned int next_item;            // dynamic work counter
  unsigned int pad;
  unsigned long long tiles_visited;  // sum over work items of the tiles they loaded
  unsigned long long cells_visited;  // ... and of the (32-source run x 256-query quarter) cells
                                     //     they computed
};

// KERN 0 gaussian, 1 epa, 2 box (rows as in kde_eval.cu).  CUT 1: one hard radius for all
// sources (cut2 = r^2); CUT 2: a hard radius proportional to each source's bandwidth
// (r_i = c h_i, i.e. u <= cut2 in the kernel's units).  Tiles are culled at cull_r, or at
// tilehmax[t] * rad_scale when rad_scale > 0 (per-tile radius).
template <int D, int CUT, int KERN>
__global__ void __launch_bounds__(SV_THREADS, 2)
    kde_eval_sorted_kernel(const float* __restrict__ packed, const float* __restrict__ tilebox,
                           const float* __restrict__ tilecen,
                           const float* __restrict__ tilehmax, uint32_t n_tiles,
                           uint32_t tiles_per_split, int nsplit, const float* __restrict__ qT,
                           uint64_t Mpad, float cut2, float cull_r, float rad_scale,
                           double out_scale, uint32_t* __restrict__ lists,
                           SortedCtl* __restrict__ ctl, double* __restrict__ out) {
  constexpr int ROWS = D + 2;
  constexpr int TILE_FLOATS = ROWS * TS;
  constexpr uint32_t TILE_BYTES = TILE_FLOATS * sizeof(float);
  constexpr bool CULL = CUT != 0 || KERN != 0;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t full[SV_STAGES];
  __shared__ unsigned int s_item;
  __shared__ float s_box[2 * D];            // box of the item's 1024 queries
  __shared__ float s_qbox[SV_Q][2 * D];     // ... and of each quarter (the qi-th query of every thread)
  __shared__ float s_red[SV_Q][2 * D][SV_THREADS / 32];
  __shared__ uint32_t s_wcnt[SV_THREADS / 32];
  __shared__ uint32_t s_nlist;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < SV_STAGES; s++) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const uint32_t n_qb = (uint32_t)(Mpad / SV_QB);
  const uint32_t n_items = n_qb * (uint32_t)nsplit;
  uint32_t* my_list = lists + (uint64_t)blockIdx.x * tiles_per_split * 4;  // {tile, -, cell mask lo, hi}
  uint32_t itg = 0;  // tiles consumed so far by this CTA (pipeline stage / parity)

  for (;;) {
    if (tid == 0) s_item = atomicAdd(&ctl->next_item, 1u);
    __syncthreads();
    const uint32_t item = s_item;
    if (item >= n_items) break;
    const uint32_t qb = item / (uint32_t)nsplit, sp = item % (uint32_t)nsplit;
    const uint32_t t0 = sp * tiles_per_split;
    const uint32_t t1 = min(n_tiles, t0 + tiles_per_split);
    const uint64_t mbase = (uint64_t)qb * SV_QB + tid;

    uint32_t nlist = t1 > t0 ? t1 - t0 : 0;
    if (CULL) {
      // bounding boxes of the item's queries (hi + lo, recentred coordinates): one per
      // quarter -- the qi-th queries of all threads are 256 consecutive points of the curve
      // -- and their union
#pragma unroll
      for (int qi = 0; qi < SV_Q; qi++)
#pragma unroll
        for (int k = 0; k < D; k++) {
          const float v = __ldg(qT + (uint64_t)k * Mpad + mbase + qi * SV_THREADS) +
                          __ldg(qT + (uint64_t)(D + k) * Mpad + mbase + qi * SV_THREADS);
          float a = v, b = v;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            a = fminf(a, __shfl_down_sync(0xffffffffu, a, off));
            b = fmaxf(b, __shfl_down_sync(0xffffffffu, b, off));
          }
          if (lane == 0) {
            s_red[qi][k][warp] = a;
            s_red[qi][D + k][warp] = b;
          }
        }
      __syncthreads();
      if (tid < SV_Q * D) {
        const int qi = tid / D, k = tid % D;
        float a = INFINITY, b = -INFINITY;
        for (int e = 0; e < SV_THREADS / 32; e++) {
          a = fminf(a, s_red[qi][k][e]);
          b = fmaxf(b, s_red[qi][D + k][e]);
        }
        s_qbox[qi][k] = a;
        s_qbox[qi][D + k] = b;
      }
      if (tid == 0) s_nlist = 0;
      __syncthreads();
      if (tid < D) {
        float a = INFINITY, b = -INFINITY;
        for (int qi = 0; qi < SV_Q; qi++) {
          a = fminf(a, s_qbox[qi][tid]);
          b = fmaxf(b, s_qbox[qi][D + tid]);
        }
        s_box[tid] = a;
        s_box[D + tid] = b;
      }
      __syncthreads();
      float qlo[D], qhi[D];
#pragma unroll
      for (int k = 0; k < D; k++) {
        qlo[k] = s_box[k];
        qhi[k] = s_box[D + k];
      }
      // ordered list of the tiles of [t0, t1) whose box is within the (tile's) radius of
      // the query box
      for (uint32_t c0 = t0; c0 < t1; c0 += SV_THREADS) {
        const uint32_t t = c0 + tid;
        bool keep = false;
        uint64_t mask = 0;  // the tile's (run, quarter) cells that lie within the radius
        if (t < t1) {
          const float* bx = tilebox + (uint64_t)t * tilebox_stride(D);
          float d2 = 0.f;
#pragma unroll
          for (int k = 0; k < D; k++) {
            const float g = fmaxf(fmaxf(__ldg(bx + k) - qhi[k], qlo[k] - __ldg(bx + D + k)), 0.f);
            d2 = fmaf(g, g, d2);
          }
          const float rt = rad_scale > 0.f ? __ldg(tilehmax + t) * rad_scale : cull_r;
          if (d2 <= rt * rt) {
            for (int sub = 0; sub < SV_NSUB; sub++) {
              const float* sb = bx + 2 * D + sub * (2 * D + 1);
              float slo[D], shi[D];
              float e2 = 0.f;
#pragma unroll
              for (int k = 0; k < D; k++) {
                slo[k] = __ldg(sb + k);
                shi[k] = __ldg(sb + D + k);
                const float g = fmaxf(fmaxf(slo[k] - qhi[k], qlo[k] - shi[k]), 0.f);
                e2 = fmaf(g, g, e2);
              }
              const float rs = rad_scale > 0.f ? __ldg(sb + 2 * D) * rad_scale : cull_r;
              if (!(e2 <= rs * rs)) continue;  // an empty run is at infinity
              for (int qi = 0; qi < SV_Q; qi++) {  // the run against each quarter's box
                float f2 = 0.f;
#pragma unroll
                for (int k = 0; k < D; k++) {
                  const float g =
                      fmaxf(fmaxf(slo[k] - s_qbox[qi][D + k], s_qbox[qi][k] - shi[k]), 0.f);
                  f2 = fmaf(g, g, f2);
                }
                if (f2 <= rs * rs) mask |= 1ull << (sub * SV_Q + qi);
              }
            }
            keep = mask != 0;
          }
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_wcnt[warp] = __popc(bal);
        __syncthreads();
        uint32_t off = s_nlist;
        for (int wv = 0; wv < warp; wv++) off += s_wcnt[wv];
        if (keep) {
          const uint32_t pos = off + __popc(bal & ((1u << lane) - 1u));
          my_list[4 * pos] = t;
          my_list[4 * pos + 2] = (uint32_t)mask;
          my_list[4 * pos + 3] = (uint32_t)(mask >> 32);
        }
        __syncthreads();
        if (tid == 0) {
          uint32_t tot = 0;
          for (int wv = 0; wv < SV_THREADS / 32; wv++) tot += s_wcnt[wv];
          s_nlist += tot;
        }
        __syncthreads();
      }
      nlist = s_nlist;
    }
    auto tile_id = [&](uint32_t l) -> uint32_t { return CULL ? my_list[4 * l] : t0 + l; };
    uint32_t ncells = 0;

    if (tid == 0) {
      for (uint32_t p = 0; p < (uint32_t)(SV_STAGES - 1) && p < nlist; p++) {
        const uint32_t st = (itg + p) % SV_STAGES;
        mbar_expect_tx(&full[st], TILE_BYTES);
        tma_load_1d(tiles + st * TILE_FLOATS, packed + (uint64_t)tile_id(p) * TILE_FLOATS,
                    TILE_BYTES, &full[st]);
      }
    }
    double accd[SV_Q];
#pragma unroll
    for (int qi = 0; qi < SV_Q; qi++) accd[qi] = 0.0;

    for (uint32_t it = 0; it < nlist; it++, itg++) {
      const uint32_t s = itg % SV_STAGES;
      const uint32_t parity = (itg / SV_STAGES) & 1u;
      if (tid == 0 && it + SV_STAGES - 1 < nlist) {
        const uint32_t ps = (itg + SV_STAGES - 1) % SV_STAGES;
        mbar_expect_tx(&full[ps], TILE_BYTES);
        tma_load_1d(tiles + ps * TILE_FLOATS,
                    packed + (uint64_t)tile_id(it + SV_STAGES - 1) * TILE_FLOATS, TILE_BYTES,
                    &full[ps]);
      }
      // queries relative to this tile's centre: (hi - c_t) is exact, + lo
      const float* tcp = tilecen + (uint64_t)tile_id(it) * D;
      float q[SV_Q][D];
#pragma unroll
      for (int k = 0; k < D; k++) {
        const float ck = __ldg(tcp + k);
#pragma unroll
        for (int qi = 0; qi < SV_Q; qi++)
          q[qi][k] = (__ldg(qT + (uint64_t)k * Mpad + mbase + qi * SV_THREADS) - ck) +
                     __ldg(qT + (uint64_t)(D + k) * Mpad + mbase + qi * SV_THREADS);
      }
      const uint64_t cmask =
          CULL ? ((uint64_t)my_list[4 * it + 2] | ((uint64_t)my_list[4 * it + 3] << 32)) : ~0ull;
      ncells += __popcll(cmask);
      mbar_wait(&full[s], parity);
      const float* tile = tiles + s * TILE_FLOATS;

      f2_t acc[SV_Q];
#pragma unroll
      for (int qi = 0; qi < SV_Q; qi++) acc[qi] = 0ull;

      constexpr int NRUN = CULL ? SV_NSUB : 1;  // the exact sum walks the tile in one loop
#pragma unroll 1
      for (int sub = 0; sub < NRUN; sub++) {
      const uint32_t m4 = (uint32_t)(cmask >> (sub * SV_Q)) & ((1u << SV_Q) - 1u);
      if (CULL && !m4) continue;
#pragma unroll 2
      for (int j = sub * (TS / NRUN); j < (sub + 1) * (TS / NRUN); j += 4) {
        float4 r[ROWS];
#pragma unroll
        for (int k = 0; k < ROWS; k++) r[k] = *reinterpret_cast<const float4*>(tile + k * TS + j);
        const f2_t a01 = f2_pack(r[D].x, r[D].y), a23 = f2_pack(r[D].z, r[D].w);
        const f2_t l01 = f2_pack(r[D + 1].x, r[D + 1].y), l23 = f2_pack(r[D + 1].z, r[D + 1].w);
        f2_t c01 = 0ull, c23 = 0ull;
        if (CUT == 1) {
          const f2_t cc = f2_pack(cut2, cut2);
          c01 = f2_mul(f2_mul(a01, a01), cc);
          c23 = f2_mul(f2_mul(a23, a23), cc);
        } else if (CUT == 2) {
          c01 = c23 = f2_pack(cut2, cut2);
        }
#pragma unroll
        for (int qi = 0; qi < SV_Q; qi++) {
          if (CULL && !((m4 >> qi) & 1u)) continue;  // the whole CTA skips this quarter
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
          float u0, u1, u2, u3;
          f2_unpack(tA, u0, u1);
          f2_unpack(tB, u2, u3);
          if (KERN) {
            float g0, g1, g2, g3;
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
            float m0, m1, m2, m3, g0, g1, g2, g3;
            f2_unpack(c01, m0, m1);
            f2_unpack(c23, m2, m3);
            f2_unpack(l01, g0, g1);
            f2_unpack(l23, g2, g3);
            eA0 = u0 <= m0 ? ex2_approx(-(u0 + g0)) : 0.f;
            eA1 = u1 <= m1 ? ex2_approx(-(u1 + g1)) : 0.f;
            eB0 = u2 <= m2 ? ex2_approx(-(u2 + g2)) : 0.f;
            eB1 = u3 <= m3 ? ex2_approx(-(u3 + g3)) : 0.f;
          } else {
            eA0 = ex2_approx(-u0);
            eA1 = ex2_approx(-u1);
            eB0 = ex2_approx(-u2);
            eB1 = ex2_approx(-u3);
          }
          acc[qi] = f2_add(acc[qi], f2_add(f2_pack(eA0, eA1), f2_pack(eB0, eB1)));
        }
      }
      }
#pragma unroll
      for (int qi = 0; qi < SV_Q; qi++) {
        float lo, hi;
        f2_unpack(acc[qi], lo, hi);
        accd[qi] += (double)lo + (double)hi;
      }
      __syncthreads();  // every warp is done with stage s before it is refilled
    }
#pragma unroll
    for (int qi = 0; qi < SV_Q; qi++)
      out[(uint64_t)sp * Mpad + mbase + qi * SV_THREADS] = accd[qi] * out_scale;
    if (tid == 0 && nlist) {
      atomicAdd(&ctl->tiles_visited, (unsigned long long)nlist);
      atomicAdd(&ctl->cells_visited, (unsigned long long)ncells);
    }
    __syncthreads();  // s_item / the list are rewritten by the next item
  }
}

// out[perm[m]] = sum over splits of partial[s][m], m < M
__global__ void reduce_scatter_kernel(const double* __restrict__ partial, int nsplit, uint64_t M,
                                      uint64_t Mpad, const uint32_t* __restrict__ perm,
                                      double* __restrict__ out) {
  const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  double s = 0.0;
  for (int p = 0; p < nsplit; p++) s += partial[(uint64_t)p * Mpad + m];
  out[perm[m]] = s;
}

template <int D>
static int launch_sorted(const float* packed, const float* tilebox, const float* tilecen,
                         const float* tilehmax, uint32_t ntl, uint32_t tps, int nsplit,
                         const float* qT, uint64_t Mpad, int kern, int cut, float cut2,
                         float cull_r, float rad_scale, double out_scale, uint32_t* lists,
                         SortedCtl* ctl, double* out, int ctas, cudaStream_t st) {
  const size_t smem = (size_t)SV_STAGES * (D + 2) * TS * sizeof(float);
#define KDSB_SV(CUTV, KV)                                                                      \
  do {                                                                                         \
    KDSB_CUDA(cudaFuncSetAttribute(kde_eval_sorted_kernel<D, CUTV, KV>,                        \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    kde_eval_sorted_kernel<D, CUTV, KV><<<ctas, SV_THREADS, smem, st>>>(                        \
        packed, tilebox, tilecen, tilehmax, ntl, tps, nsplit, qT, Mpad, cut2, cull_r,          \
        rad_scale, out_scale, lists, ctl, out);                                                \
  } while (0)
  if (kern == 1) KDSB_SV(0, 1);
  else if (kern == 2) KDSB_SV(0, 2);
  else if (cut == 1) KDSB_SV(1, 0);
  else if (cut == 2) KDSB_SV(2, 0);
  else KDSB_SV(0, 0);
#undef KDSB_SV
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace kdsb

using namespace kdsb;

extern "C" {

uint64_t kdsb200_minmax_scratch(void) { return (uint64_t)MM_BLOCKS * 16; }

int kdsb200_sorted_tilebox_floats(int D) { return tilebox_stride(D); }

int kdsb200_column_minmax(const double* d_data, uint64_t N, int D, double* d_scratch,
                          double* d_minmax, void* stream) {
  if (D < 1 || D > 8 || N == 0) {
    set_error("column_minmax: D=%d out of range 1..8 or empty input", D);
    return 1;
  }
  cudaStream_t st = (cudaStream_t)stream;
  minmax_partial_kernel<<<MM_BLOCKS, MM_THREADS, 0, st>>>(d_data, N, D, d_scratch);
  KDSB_CUDA(cudaGetLastError());
  minmax_final_kernel<<<1, 32, 0, st>>>(d_scratch, D, d_minmax);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

// scratch of the ordering: two key buffers, one value buffer, the (digit, segment)
// counters and the block sums of their scan
uint64_t kdsb200_order_scratch_bytes(uint64_t N) {
  const uint64_t nseg = (N + RS_SEG - 1) / RS_SEG;
  const uint64_t ncnt = 256 * nseg;
  const uint64_t nb = (ncnt + SC_BLOCK - 1) / SC_BLOCK + 1;
  return 2 * N * 8 + N * 4 + ncnt * 4 + nb * 4 + 64;
}

int kdsb200_hilbert_bits(int D) {
  int b = 60 / D;
  return b > 16 ? 16 : b;
}

int kdsb200_hilbert_order(const double* d_data, uint64_t N, int D, const double* h_lo,
                          const double* h_hi, void* d_scratch, uint32_t* d_perm, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (D < 1 || D > 8) {
    set_error("hilbert_order: D=%d out of range 1..8", D);
    return 1;
  }
  if (N == 0) return 0;
  if (N >= (1ull << 32)) {
    set_error("hilbert_order: N must be below 2^32");
    return 1;
  }
  Quant q;
  q.bits = kdsb200_hilbert_bits(D);
  for (int k = 0; k < 8; k++) {
    q.lo[k] = k < D ? h_lo[k] : 0.0;
    const double span = k < D ? h_hi[k] - h_lo[k] : 0.0;
    q.scale[k] = span > 0 ? (double)(1u << q.bits) / span : 0.0;
  }
  const uint64_t nseg = (N + RS_SEG - 1) / RS_SEG;
  const uint64_t ncnt = 256 * nseg;
  unsigned char* p = (unsigned char*)d_scratch;
  uint64_t* keys_a = (uint64_t*)p;
  uint64_t* keys_b = keys_a + N;
  uint32_t* vals_b = (uint32_t*)(keys_b + N);
  uint32_t* counts = vals_b + N;
  uint32_t* bsum = counts + ncnt;
  uint32_t* vals_a = d_perm;
  const unsigned kb = (unsigned)((N + 255) / 256);
  switch (D) {
#define KDSB_CASE(d)                                                                 \
  case d:                                                                            \
    hilbert_keys_kernel<d><<<kb, 256, 0, st>>>(d_data, N, q, keys_a, vals_a);        \
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
  }
  KDSB_CUDA(cudaGetLastError());
  int passes = (q.bits * D + 7) / 8;
  if (passes & 1) passes++;  // an even number of passes leaves the result in (keys_a, d_perm)
  const unsigned sb = (unsigned)((nseg + RS_WARPS - 1) / RS_WARPS);
  for (int ps = 0; ps < passes; ps++) {
    const uint64_t* kin = (ps & 1) ? keys_b : keys_a;
    uint64_t* kout = (ps & 1) ? keys_a : keys_b;
    const uint32_t* vin = (ps & 1) ? vals_b : vals_a;
    uint32_t* vout = (ps & 1) ? vals_a : vals_b;
    radix_count_kernel<<<sb, RS_WARPS * 32, 0, st>>>(kin, N, ps * 8, nseg, counts);
    KDSB_CUDA(cudaGetLastError());
    if (exclusive_scan_u32(counts, ncnt, bsum, st)) return 1;
    radix_scatter_kernel<<<sb, RS_WARPS * 32, 0, st>>>(kin, vin, N, ps * 8, nseg, counts, kout,
                                                       vout);
    KDSB_CUDA(cudaGetLastError());
  }
  return 0;
}

uint64_t kdsb200_sorted_qpad(uint64_t M) { return (M + SV_QB - 1) / SV_QB * SV_QB; }

int kdsb200_pack_sources_sorted_f32(const double* d_data, const double* d_w, const double* d_bw,
                                    double bw_scalar, uint64_t N, int D, const uint32_t* d_perm,
                                    const double* h_center, double grid, double w_scale,
                                    double lc_ref, int mode, float* d_packed, float* d_tilebox,
                                    float* d_tilecen, float* d_tilehmax, void* stream) {
  if (D < 1 || D > 8 || !(grid > 0)) {
    set_error("pack_sources_sorted: D=%d out of range 1..8 or grid <= 0", D);
    return 1;
  }
  Center8 cen;
  for (int k = 0; k < 8; k++) cen.c[k] = k < D ? h_center[k] : 0.0;
  const uint64_t ntl = n_tiles_for(N);
  if (ntl == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (D) {
#define KDSB_CASE(d)                                                                         \
  case d:                                                                                    \
    pack_sorted_kernel<d><<<(unsigned)ntl, TS, 0, st>>>(d_data, d_w, d_bw, bw_scalar, N,     \
                                                        d_perm, cen, grid, w_scale, lc_ref,  \
                                                        mode, d_packed, d_tilebox, d_tilecen,  \
                                                        d_tilehmax);                         \
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
  }
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

int kdsb200_pack_queries_sorted_f32(const double* d_pts, uint64_t M, uint64_t Mpad, int D,
                                    const uint32_t* d_perm, const double* h_center, double grid,
                                    float* d_qT, void* stream) {
  if (D < 1 || D > 8 || !(grid > 0)) {
    set_error("pack_queries_sorted: D=%d out of range 1..8 or grid <= 0", D);
    return 1;
  }
  if (M == 0 || Mpad == 0) return 0;
  Center8 cen;
  for (int k = 0; k < 8; k++) cen.c[k] = k < D ? h_center[k] : 0.0;
  pack_queries_sorted_kernel<<<(unsigned)((Mpad + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      d_pts, M, Mpad, D, d_perm, cen, grid, d_qT);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

// persistent CTAs of the sorted evaluation kernel on this device
int kdsb200_sorted_ctas(void) { return 2 * sm_count(); }

uint64_t kdsb200_sorted_scratch_bytes(uint64_t N, uint64_t Mpad, int nsplit) {
  const uint64_t ntl = n_tiles_for(N);
  if (nsplit < 1) nsplit = 1;
  const uint64_t tps = (ntl + nsplit - 1) / nsplit;
  return 64 + (uint64_t)kdsb200_sorted_ctas() * tps * 16 + (uint64_t)nsplit * Mpad * 8 + 64;
}

int kdsb200_sorted_nsplit(uint64_t N, uint64_t M) {
  // enough work items (query blocks x source splits) to keep every persistent CTA busy
  // several times over; items are handed out dynamically, so no wave arithmetic is needed
  const uint64_t qb = kdsb200_sorted_qpad(M) / SV_QB;
  const uint64_t ntl = n_tiles_for(N);
  if (!qb || !ntl) return 1;
  const uint64_t want = (uint64_t)kdsb200_sorted_ctas() * 6;
  uint64_t ns = (want + qb - 1) / qb;
  if (ns > ntl) ns = ntl;
  if (ns > 256) ns = 256;
  return (int)(ns < 1 ? 1 : ns);
}

int kdsb200_kde_eval_sorted_f32(const float* d_packed, const float* d_tilebox,
                                const float* d_tilecen, const float* d_tilehmax, uint64_t N,
                                int D, const float* d_qT, const uint32_t* d_perm_q, uint64_t M,
                                uint64_t Mpad, int kernel, double cutoff, double cutoff_rel,
                                double out_scale, int nsplit, void* d_scratch, double* d_out,
                                uint64_t* d_tiles_visited, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (M == 0) return 0;
  if (kernel < 0 || kernel > 2) {
    set_error("kde_eval_sorted: kernel=%d must be 0 (gaussian), 1 (epa) or 2 (box)", kernel);
    return 1;
  }
  if (Mpad % SV_QB != 0 || Mpad < M) {
    set_error("kde_eval_sorted: Mpad=%llu must be a multiple of %d and >= M",
              (unsigned long long)Mpad, SV_QB);
    return 1;
  }
  const uint32_t ntl = (uint32_t)n_tiles_for(N);
  if (ntl == 0) {
    KDSB_CUDA(cudaMemsetAsync(d_out, 0, M * sizeof(double), st));
    return 0;
  }
  if (nsplit < 1) nsplit = 1;
  if ((uint32_t)nsplit > ntl) nsplit = (int)ntl;
  const uint32_t tps = (ntl + nsplit - 1) / nsplit;
  nsplit = (int)((ntl + tps - 1) / tps);
  const int ctas = kdsb200_sorted_ctas();
  unsigned char* p = (unsigned char*)d_scratch;
  SortedCtl* ctl = (SortedCtl*)p;
  uint32_t* lists = (uint32_t*)(p + 64);
  uint64_t off = 64 + (uint64_t)ctas * tps * 16;
  off = (off + 63) / 64 * 64;
  double* partial = (double*)(p + off);
  KDSB_CUDA(cudaMemsetAsync(ctl, 0, sizeof(SortedCtl), st));
  // Gaussian kernel: cutoff > 0 is one hard radius for every source; cutoff_rel > 0 a hard
  // radius c h_i per source (u = |q-x|^2 log2(e) / (2 h_i^2) <= c^2 log2(e) / 2 in the
  // kernel's units), with tiles culled at c times the tile's largest bandwidth.  Finite
  // kernels are packed with their support radius R_i in place of h_i and culled at the
  // tile's largest R_i.  A relative margin keeps fp32 rounding of the boxes from dropping a
  // boundary tile.
  int cut = 0;
  float cut2 = -1.f, cull_r = INFINITY, rad_scale = 0.f;
  if (kernel == 0 && cutoff > 0) {
    cut = 1;
    cut2 = (float)(cutoff * cutoff);
    cull_r = (float)(cutoff * (1.0 + 1e-5));
  } else if (kernel == 0 && cutoff_rel > 0) {
    cut = 2;
    cut2 = (float)(cutoff_rel * cutoff_rel * 0.72134752044448170368);
    rad_scale = (float)(cutoff_rel * (1.0 + 1e-5));
  } else if (kernel != 0) {
    rad_scale = (float)(1.0 + 1e-5);
  }
  if (rad_scale > 0.f && !d_tilehmax) {
    set_error("kde_eval_sorted: per-tile radii requested without d_tilehmax");
    return 1;
  }
  int rc = 1;
  switch (D) {
#define KDSB_CASE(d)                                                                          \
  case d:                                                                                     \
    rc = launch_sorted<d>(d_packed, d_tilebox, d_tilecen, d_tilehmax, ntl, tps, nsplit, d_qT, \
                          Mpad, kernel, cut, cut2, cull_r, rad_scale, out_scale, lists, ctl,  \
                          partial, ctas, st);                                                 \
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
      set_error("kde_eval_sorted: D=%d out of range 1..8", D);
      return 1;
  }
  if (rc) return rc;
  reduce_scatter_kernel<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(partial, nsplit, M, Mpad,
                                                                     d_perm_q, d_out);
  KDSB_CUDA(cudaGetLastError());
  if (d_tiles_visited)  // {tiles loaded, (32-source run x 256-query quarter) cells computed}
    KDSB_CUDA(cudaMemcpyAsync(d_tiles_visited, &ctl->tiles_visited, 2 * sizeof(uint64_t),
                              cudaMemcpyDeviceToDevice, st));
  return 0;
}

}  // extern "C"
