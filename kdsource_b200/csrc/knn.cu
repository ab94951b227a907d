// knn_kth_dist: brute-force k-nearest-neighbour search inside contiguous batches,
// the arithmetic of sklearn NearestNeighbors(n_neighbors=k+1).fit(b).kneighbors(b)
// at python/kdsource/kde.py:93-98 (self point included, distances ascending).
//
// One CTA owns a chunk of query rows of one batch; each warp owns a subset of the
// rows.  Candidates stream through shared memory in tiles; every lane keeps JL
// candidates in registers and the row coordinates arrive as broadcast LDS.  Per
// row the state in shared memory is a sorted (distance^2, index) list of KP
// entries, a 32-entry candidate queue and the current threshold tau (the kk-th
// best so far).  Lanes ballot `d2 <= tau`; survivors are compacted into the
// queue, and a full queue is sorted with a warp-shuffle bitonic network and
// merged into the list with bitonic merges (one per 32-entry chunk).
// Neighbour order is (distance^2, index) lexicographic, which makes indices
// deterministic under ties.
//
// T = double reproduces numpy/sklearn bit for bit (sub, mul, add in coordinate
// order, no FMA contraction); T = float is the fast mode (FMA chain).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "kdsb200.h"

namespace kdsb {

constexpr int KNN_THREADS = 256;
constexpr int KNN_WARPS = KNN_THREADS / 32;

template <typename T>
struct Ent {
  T d;
  int i;
};

template <typename T>
__device__ __forceinline__ bool ent_less(const Ent<T>& a, const Ent<T>& b) {
  return a.d < b.d || (a.d == b.d && a.i < b.i);
}

template <typename T>
__device__ __forceinline__ Ent<T> ent_shfl_xor(const Ent<T>& a, int m) {
  Ent<T> r;
  r.d = __shfl_xor_sync(0xffffffffu, a.d, m);
  r.i = __shfl_xor_sync(0xffffffffu, a.i, m);
  return r;
}
template <typename T>
__device__ __forceinline__ Ent<T> ent_shfl(const Ent<T>& a, int src) {
  Ent<T> r;
  r.d = __shfl_sync(0xffffffffu, a.d, src);
  r.i = __shfl_sync(0xffffffffu, a.i, src);
  return r;
}

// compare-exchange step of a bitonic network across lanes
template <typename T>
__device__ __forceinline__ Ent<T> cex(const Ent<T>& v, int j, bool take_min) {
  const Ent<T> o = ent_shfl_xor(v, j);
  const bool o_less = ent_less(o, v);
  return (o_less == take_min) ? o : v;
}

// full bitonic sort of 32 entries (one per lane), ascending
template <typename T>
__device__ __forceinline__ Ent<T> warp_sort32(Ent<T> v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const bool up = (lane & k) == 0;
      const bool lower = (lane & j) == 0;
      v = cex(v, j, lower == up);
    }
  }
  return v;
}

// sorts a bitonic sequence of 32 entries ascending
template <typename T>
__device__ __forceinline__ Ent<T> warp_bitonic_merge32(Ent<T> v, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) v = cex(v, j, (lane & j) == 0);
  return v;
}

template <typename T>
struct KnnSmem {
  T* rowq;     // [RC][D]
  T* tau;      // [RC]
  int* qcnt;   // [RC]
  T* list_d;   // [RC][KP]
  int* list_i; // [RC][KP]
  T* que_d;    // [RC][32]
  int* que_i;  // [RC][32]
  T* cand;     // [D][CT]
};

// merge the 32 queue entries of row r (padded with +inf beyond `cnt`) into its list
template <typename T>
__device__ __forceinline__ void merge_queue(const KnnSmem<T>& S, int r, int cnt, int KP, int kk,
                                            int lane) {
  Ent<T> b;
  if (lane < cnt) {
    b.d = S.que_d[r * 32 + lane];
    b.i = S.que_i[r * 32 + lane];
  } else {
    b.d = (T)INFINITY;
    b.i = 0x7fffffff;
  }
  b = warp_sort32(b, lane);
  for (int c = 0; c < KP; c += 32) {
    Ent<T> a;
    a.d = S.list_d[r * KP + c + lane];
    a.i = S.list_i[r * KP + c + lane];
    const Ent<T> br = ent_shfl(b, 31 - lane);
    const bool a_less = ent_less(a, br);
    Ent<T> lo = a_less ? a : br;
    Ent<T> hi = a_less ? br : a;
    lo = warp_bitonic_merge32(lo, lane);
    hi = warp_bitonic_merge32(hi, lane);
    S.list_d[r * KP + c + lane] = lo.d;
    S.list_i[r * KP + c + lane] = lo.i;
    b = hi;
  }
  __syncwarp();
  if (lane == 0) S.tau[r] = S.list_d[r * KP + kk - 1];
  __syncwarp();
}

// whole-warp call: compact the lanes with `pred` into the row's queue, merging
// into the sorted list whenever 32 entries have accumulated
template <typename T>
__device__ __noinline__ void push_candidates(KnnSmem<T> S, int r, T d2, int j, bool pred, int KP,
                                             int kk, int lane) {
  const unsigned mask = __ballot_sync(0xffffffffu, pred);
  if (!mask) return;
  const int cnt = S.qcnt[r];
  const int pos = cnt + __popc(mask & ((1u << lane) - 1u));
  if (pred && pos < 32) {
    S.que_d[r * 32 + pos] = d2;
    S.que_i[r * 32 + pos] = j;
  }
  int total = cnt + __popc(mask);
  __syncwarp();
  if (total >= 32) {
    merge_queue<T>(S, r, 32, KP, kk, lane);
    if (pred && pos >= 32) {
      S.que_d[r * 32 + pos - 32] = d2;
      S.que_i[r * 32 + pos - 32] = j;
    }
    total -= 32;
  }
  if (lane == 0) S.qcnt[r] = total;
  __syncwarp();
}

template <typename T>
__device__ __noinline__ void flush_queue(KnnSmem<T> S, int r, int KP, int kk, int lane) {
  const int cnt = S.qcnt[r];
  if (cnt > 0) merge_queue<T>(S, r, cnt, KP, kk, lane);
}

template <typename T, int D>
__device__ __forceinline__ T dist2(const T (&q)[D], const T* x) {
  if constexpr (sizeof(T) == 8) {
    // numpy order: ((d0*d0 + d1*d1) + d2*d2) ... without contraction
    double t = __dsub_rn(q[0], x[0]);
    double acc = __dmul_rn(t, t);
#pragma unroll
    for (int k = 1; k < D; k++) {
      t = __dsub_rn(q[k], x[k]);
      acc = __dadd_rn(acc, __dmul_rn(t, t));
    }
    return acc;
  } else {
    float t = q[0] - x[0];
    float acc = t * t;
#pragma unroll
    for (int k = 1; k < D; k++) {
      t = q[k] - x[k];
      acc = fmaf(t, t, acc);
    }
    return acc;
  }
}

template <typename T, int D, int JL>
__global__ void __launch_bounds__(KNN_THREADS)
    knn_kernel(const T* __restrict__ dataT /* [D][Nstride] */, uint64_t Nstride,
               const int64_t* __restrict__ batch_off, int chunks, int RC, int KP, int kk,
               double* __restrict__ out_kth, double* __restrict__ out_dist,
               int32_t* __restrict__ out_idx) {
  constexpr int CT = 32 * JL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  KnnSmem<T> S;
  {
    unsigned char* p = smem_raw;
    S.rowq = reinterpret_cast<T*>(p);   p += sizeof(T) * RC * D;
    S.tau = reinterpret_cast<T*>(p);    p += sizeof(T) * RC;
    S.list_d = reinterpret_cast<T*>(p); p += sizeof(T) * RC * KP;
    S.que_d = reinterpret_cast<T*>(p);  p += sizeof(T) * RC * 32;
    S.cand = reinterpret_cast<T*>(p);   p += sizeof(T) * D * CT;
    S.qcnt = reinterpret_cast<int*>(p); p += sizeof(int) * RC;
    S.list_i = reinterpret_cast<int*>(p); p += sizeof(int) * RC * KP;
    S.que_i = reinterpret_cast<int*>(p);
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int bidx = blockIdx.x / chunks;  // batch
  const int64_t b0 = batch_off[bidx], b1 = batch_off[bidx + 1];
  const int bn = (int)(b1 - b0);
  const int r0 = (blockIdx.x % chunks) * RC;
  if (r0 >= bn) return;
  const int nrows = min(RC, bn - r0);

  for (int e = tid; e < nrows * D; e += KNN_THREADS) {
    const int r = e / D, k = e % D;
    S.rowq[r * D + k] = dataT[(uint64_t)k * Nstride + b0 + r0 + r];
  }
  for (int e = tid; e < nrows; e += KNN_THREADS) {
    S.tau[e] = (T)INFINITY;
    S.qcnt[e] = 0;
  }
  for (int e = tid; e < nrows * KP; e += KNN_THREADS) {
    S.list_d[e] = (T)INFINITY;
    S.list_i[e] = 0x7fffffff;
  }

  for (int cbase = 0; cbase < bn; cbase += CT) {
    __syncthreads();
    for (int e = tid; e < D * CT; e += KNN_THREADS) {
      const int k = e / CT, c = e % CT;
      S.cand[e] = (cbase + c < bn) ? dataT[(uint64_t)k * Nstride + b0 + cbase + c] : (T)0;
    }
    __syncthreads();
    T x[JL][D];
#pragma unroll
    for (int e = 0; e < JL; e++)
#pragma unroll
      for (int k = 0; k < D; k++) x[e][k] = S.cand[k * CT + e * 32 + lane];

    for (int r = warp; r < nrows; r += KNN_WARPS) {
      T q[D];
#pragma unroll
      for (int k = 0; k < D; k++) q[k] = S.rowq[r * D + k];
      const T tau = S.tau[r];
      T d2[JL];
      bool any = false;
#pragma unroll
      for (int e = 0; e < JL; e++) {
        d2[e] = dist2<T, D>(q, x[e]);
        any |= (cbase + e * 32 + lane < bn) && (d2[e] <= tau);
      }
      // rare path (kept out of line so the hot loop stays in the instruction cache)
      if (__any_sync(0xffffffffu, any)) {
#pragma unroll
        for (int e = 0; e < JL; e++) {
          const int j = cbase + e * 32 + lane;
          push_candidates<T>(S, r, d2[e], j, (j < bn) && (d2[e] <= tau), KP, kk, lane);
        }
      }
    }
  }
  __syncwarp();
  // flush leftovers and write results
  for (int r = warp; r < nrows; r += KNN_WARPS) {
    flush_queue<T>(S, r, KP, kk, lane);
    const uint64_t row = (uint64_t)(b0 + r0 + r);
    for (int e = lane; e < kk; e += 32) {
      const double dd = sqrt((double)S.list_d[r * KP + e]);
      if (out_dist) out_dist[row * kk + e] = dd;
      if (out_idx) out_idx[row * kk + e] = S.list_i[r * KP + e] == 0x7fffffff ? -1 : S.list_i[r * KP + e];
      if (e == kk - 1 && out_kth) out_kth[row] = dd;
    }
  }
}


// ---------------------------------------------------------------------------
// small-k kernel (k+1 <= KK <= 24): one query row per thread, its KK best
// (distance^2, index) pairs sorted in registers.  Candidates stream through a
// double-buffered shared-memory tile (cp.async) and are read as broadcast
// LDS.128; no cross-lane traffic at all.  Candidates arrive in index order and
// the insertion test is strict, so equal distances keep the smaller index first
// ((distance, index) order, as the oracle and sklearn's brute force give).
// ---------------------------------------------------------------------------
constexpr int KS_THREADS = 256;
template <typename T>
struct KsPend {
  // pending survivors per thread between warp-wide inserts
  static constexpr int N = sizeof(T) == 4 ? 16 : 8;
};
template <typename T>
struct KsTile {
  static constexpr int CT = sizeof(T) == 4 ? 512 : 128;  // candidates per tile
};

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
// stores through 32-bit shared-window addresses (one STS, no 64-bit pointer arithmetic)
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void sts_s32(uint32_t addr, int v) {
  asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename T, int D, int KK, int KS_CT>
__global__ void __launch_bounds__(KS_THREADS, sizeof(T) == 4 ? 4 : 2)
    knn_small_kernel(const T* __restrict__ dataT, uint64_t Nstride,
                     const int64_t* __restrict__ batch_off, int chunks, int kk,
                     double* __restrict__ out_kth, double* __restrict__ out_dist,
                     int32_t* __restrict__ out_idx) {
  constexpr int KS_PEND = KsPend<T>::N;
  extern __shared__ __align__(16) unsigned char ks_smem[];
  T(*tile)[D][KS_CT] = reinterpret_cast<T(*)[D][KS_CT]>(ks_smem);
  T(*pend_d)[KS_THREADS] =
      reinterpret_cast<T(*)[KS_THREADS]>(ks_smem + sizeof(T) * 2 * D * KS_CT);
  int(*pend_i)[KS_THREADS] = reinterpret_cast<int(*)[KS_THREADS]>(
      ks_smem + sizeof(T) * 2 * D * KS_CT + sizeof(T) * KS_PEND * KS_THREADS);
  const int tid = threadIdx.x;
  const int bidx = blockIdx.x / chunks;
  const int64_t b0 = batch_off[bidx], b1 = batch_off[bidx + 1];
  const int bn = (int)(b1 - b0);
  const int r0 = (blockIdx.x % chunks) * KS_THREADS;
  if (r0 >= bn) return;
  const int row = r0 + tid;
  const bool live = row < bn;

  T q[D];
#pragma unroll
  for (int k = 0; k < D; k++) q[k] = live ? dataT[(uint64_t)k * Nstride + b0 + row] : (T)0;
  T bd[KK];
  int bi[KK];
#pragma unroll
  for (int s = 0; s < KK; s++) {
    bd[s] = (T)INFINITY;
    bi[s] = -1;
  }

  // per-thread pending candidates; entries are in candidate (index) order, and the
  // insert below is strict, so equal distances keep the smaller index first.  pd / pi
  // point at the next free slot of this thread's column.
  const uint32_t pd0 = smem_u32(&pend_d[0][tid]), pi0 = smem_u32(&pend_i[0][tid]);
  uint32_t pd = pd0, pi = pi0;
  constexpr uint32_t PDSTEP = KS_THREADS * sizeof(T), PISTEP = KS_THREADS * sizeof(int);
  auto flush_pending = [&]() {
    const int npend = (int)((pd - pd0) / PDSTEP);
    for (int e = 0; e < KS_PEND; e++) {
      if (!__any_sync(0xffffffffu, e < npend)) break;
      if (e < npend) {
        T nd = pend_d[e][tid];
        int ni = pend_i[e][tid];
        if (nd < bd[KK - 1]) {
          bool shifting = false;  // once placed, everything behind moves down one slot
#pragma unroll
          for (int s = 0; s < KK; s++) {
            shifting = shifting || (nd < bd[s]);
            if (shifting) {
              const T td = bd[s];
              const int ti = bi[s];
              bd[s] = nd;
              bi[s] = ni;
              nd = td;
              ni = ti;
            }
          }
        }
      }
    }
    pd = pd0;
    pi = pi0;
  };

  // padding candidates sit at infinity: their distance is +inf and never survives
  auto load_tile = [&](int buf, int cbase) {
    for (int e = tid; e < D * KS_CT; e += KS_THREADS) {
      const int k = e / KS_CT, c = e % KS_CT;
      if (cbase + c < bn) {
        const T* src = dataT + (uint64_t)k * Nstride + b0 + cbase + c;
        if (sizeof(T) == 4) cp_async4(&tile[buf][k][c], src);
        else cp_async8(&tile[buf][k][c], src);
      } else {
        tile[buf][k][c] = (T)INFINITY;
      }
    }
    cp_async_commit();
  };

  const int ntile = (bn + KS_CT - 1) / KS_CT;
  load_tile(0, 0);
  for (int t = 0; t < ntile; t++) {
    cp_async_wait_all();
    __syncthreads();  // tile t visible; everyone is done with the other buffer
    if (t + 1 < ntile) load_tile((t + 1) & 1, (t + 1) * KS_CT);
    const int buf = t & 1;
    const int cbase = t * KS_CT;
    const int cnt = min(KS_CT, bn - cbase);
    constexpr int V = 4;  // candidates per step (one LDS.128 per coordinate in fp32, two in fp64)
    for (int c = 0; c < cnt; c += V) {
      T x[D][V];
#pragma unroll
      for (int k = 0; k < D; k++) {
        if constexpr (sizeof(T) == 4) {
          const float4 v = *reinterpret_cast<const float4*>(&tile[buf][k][c]);
          x[k][0] = v.x; x[k][1] = v.y; x[k][2] = v.z; x[k][3] = v.w;
        } else {
          const double2 v0 = *reinterpret_cast<const double2*>(&tile[buf][k][c]);
          const double2 v1 = *reinterpret_cast<const double2*>(&tile[buf][k][c + 2]);
          x[k][0] = v0.x; x[k][1] = v0.y; x[k][2] = v1.x; x[k][3] = v1.y;
        }
      }
      T d2[V];
      bool any = false;
      if constexpr (sizeof(T) == 4) {
        // two candidates per f32x2 lane pair: (q - x)^2 accumulated with FADD2/FFMA2
        f2_t acc01 = 0ull, acc23 = 0ull;
#pragma unroll
        for (int k = 0; k < D; k++) {
          const f2_t qq = f2_pack(-q[k], -q[k]);
          const f2_t e01 = f2_add(f2_pack(x[k][0], x[k][1]), qq);
          const f2_t e23 = f2_add(f2_pack(x[k][2], x[k][3]), qq);
          acc01 = k == 0 ? f2_mul(e01, e01) : f2_fma(e01, e01, acc01);
          acc23 = k == 0 ? f2_mul(e23, e23) : f2_fma(e23, e23, acc23);
        }
        float a, b, cc, dd;
        f2_unpack(acc01, a, b);
        f2_unpack(acc23, cc, dd);
        d2[0] = a; d2[1] = b; d2[2] = cc; d2[3] = dd;
#pragma unroll
        for (int v = 0; v < V; v++) any |= d2[v] < bd[KK - 1];
      } else {
#pragma unroll
        for (int v = 0; v < V; v++) {
          T xv[D];
#pragma unroll
          for (int k = 0; k < D; k++) xv[k] = x[k][v];
          d2[v] = dist2<T, D>(q, xv);
          any |= d2[v] < bd[KK - 1];
        }
      }
      // One warp-uniform branch guards everything that is not distance arithmetic: the
      // survivors are parked in this thread's pending list (shared memory, predicated
      // stores, no divergence), and the expensive sorted insert runs later for the whole
      // warp at once.  Most steps of a long candidate stream skip all of it.
      if (__any_sync(0xffffffffu, any)) {
        const T thr = bd[KK - 1];
#pragma unroll
        for (int v = 0; v < V; v++) {
          if (d2[v] < thr) {  // two predicated stores and two offset bumps
            if constexpr (sizeof(T) == 4) sts_f32(pd, d2[v]);
            else sts_f64(pd, d2[v]);
            sts_s32(pi, cbase + c + v);
            pd += PDSTEP;
            pi += PISTEP;
          }
        }
        if (__any_sync(0xffffffffu, pd > pd0 + (KS_PEND - V) * PDSTEP)) flush_pending();
      }
    }
  }
  flush_pending();
  if (live) {
    const uint64_t grow = (uint64_t)(b0 + row);
#pragma unroll
    for (int s = 0; s < KK; s++) {
      if (s < kk) {
        const double dd = sqrt((double)bd[s]);
        if (out_dist) out_dist[grow * kk + s] = dd;
        if (out_idx) out_idx[grow * kk + s] = bi[s];
        if (s == kk - 1 && out_kth) out_kth[grow] = dd;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// fp32 distances-only variant, TWO query rows per thread.  With one row per thread every
// candidate coordinate is a broadcast LDS.128 that feeds only 4 pairs per lane: 6 LDS.128
// (512 B each) per 128 pairs keep the 128 B/clk shared-memory port busy for 24 cycles,
// twice what the FP32 pipe needs for the arithmetic -- the kernel was LSU-bound at ~50 %
// of the FP32 roofline.  Two rows per thread halve the bytes per pair.  The neighbour
// indices are dropped (bw_knn only needs the k-th distance), which pays for the second
// row's sorted list in registers; when indices are requested the one-row kernel runs.
// ---------------------------------------------------------------------------
template <int D, int KK, int CT, int K2_PEND>
__global__ void __launch_bounds__(KS_THREADS, 3)
    knn_kth2_kernel(const float* __restrict__ dataT, uint64_t Nstride,
                    const int64_t* __restrict__ batch_off, int chunks, int kk,
                    double* __restrict__ out_kth, double* __restrict__ out_dist) {
  extern __shared__ __align__(16) unsigned char ks_smem[];
  float(*tile)[D][CT] = reinterpret_cast<float(*)[D][CT]>(ks_smem);
  float(*pend)[K2_PEND][KS_THREADS] = reinterpret_cast<float(*)[K2_PEND][KS_THREADS]>(
      ks_smem + sizeof(float) * 2 * D * CT);
  const int tid = threadIdx.x;
  const int bidx = blockIdx.x / chunks;
  const int64_t b0 = batch_off[bidx], b1 = batch_off[bidx + 1];
  const int bn = (int)(b1 - b0);
  const int r0 = (blockIdx.x % chunks) * (2 * KS_THREADS);
  if (r0 >= bn) return;
  const int row[2] = {r0 + tid, r0 + KS_THREADS + tid};

  float q[2][D];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int k = 0; k < D; k++)
      q[r][k] = row[r] < bn ? dataT[(uint64_t)k * Nstride + b0 + row[r]] : 0.f;
  float bd[2][KK];
#pragma unroll
  for (int r = 0; r < 2; r++)
#pragma unroll
    for (int s = 0; s < KK; s++) bd[r][s] = INFINITY;
  // pending survivors: pp[r] points at the next free slot of this thread's column
  const uint32_t pbase[2] = {smem_u32(&pend[0][0][tid]), smem_u32(&pend[1][0][tid])};
  uint32_t pp[2] = {pbase[0], pbase[1]};
  constexpr uint32_t PSTEP = KS_THREADS * sizeof(float);

  auto flush_pending = [&](int r) {
    const int np = (int)((pp[r] - pbase[r]) / PSTEP);
    for (int e = 0; e < K2_PEND; e++) {
      if (!__any_sync(0xffffffffu, e < np)) break;
      if (e < np) {
        float nd = pend[r][e][tid];
        if (nd < bd[r][KK - 1]) {
          bool shifting = false;
#pragma unroll
          for (int s = 0; s < KK; s++) {
            shifting = shifting || (nd < bd[r][s]);
            if (shifting) {
              const float td = bd[r][s];
              bd[r][s] = nd;
              nd = td;
            }
          }
        }
      }
    }
    pp[r] = pbase[r];
  };

  // padding candidates sit at infinity: their distance is +inf and never survives, so the
  // hot loop needs no bounds test
  auto load_tile = [&](int buf, int cbase) {
    for (int e = tid; e < D * CT; e += KS_THREADS) {
      const int k = e / CT, c = e % CT;
      if (cbase + c < bn)
        cp_async4(&tile[buf][k][c], dataT + (uint64_t)k * Nstride + b0 + cbase + c);
      else
        tile[buf][k][c] = INFINITY;
    }
    cp_async_commit();
  };

  const int ntile = (bn + CT - 1) / CT;
  load_tile(0, 0);
  for (int t = 0; t < ntile; t++) {
    cp_async_wait_all();
    __syncthreads();
    if (t + 1 < ntile) load_tile((t + 1) & 1, (t + 1) * CT);
    const int buf = t & 1;
    const int cbase = t * CT;
    const int cnt = min(CT, bn - cbase);
    for (int c = 0; c < cnt; c += 4) {
      float4 x[D];
#pragma unroll
      for (int k = 0; k < D; k++) x[k] = *reinterpret_cast<const float4*>(&tile[buf][k][c]);
      float d2[2][4];
      bool any[2];
#pragma unroll
      for (int r = 0; r < 2; r++) {
        f2_t acc01 = 0ull, acc23 = 0ull;
#pragma unroll
        for (int k = 0; k < D; k++) {
          const f2_t nq = f2_pack(-q[r][k], -q[r][k]);  // scalar broadcast operand of FADD2
          const f2_t e01 = f2_add(f2_pack(x[k].x, x[k].y), nq);
          const f2_t e23 = f2_add(f2_pack(x[k].z, x[k].w), nq);
          acc01 = k == 0 ? f2_mul(e01, e01) : f2_fma(e01, e01, acc01);
          acc23 = k == 0 ? f2_mul(e23, e23) : f2_fma(e23, e23, acc23);
        }
        f2_unpack(acc01, d2[r][0], d2[r][1]);
        f2_unpack(acc23, d2[r][2], d2[r][3]);
        any[r] = fminf(fminf(d2[r][0], d2[r][1]), fminf(d2[r][2], d2[r][3])) < bd[r][KK - 1];
      }
      if (__any_sync(0xffffffffu, any[0] || any[1])) {
        // warp-uniform slow path: predicated appends (a store and a pointer bump each)
#pragma unroll
        for (int r = 0; r < 2; r++) {
          const float thr = bd[r][KK - 1];
#pragma unroll
          for (int v = 0; v < 4; v++) {
            if (d2[r][v] < thr) {
              sts_f32(pp[r], d2[r][v]);
              pp[r] += PSTEP;
            }
          }
        }
        const bool full0 = pp[0] > pbase[0] + (K2_PEND - 4) * PSTEP;
        const bool full1 = pp[1] > pbase[1] + (K2_PEND - 4) * PSTEP;
        if (__any_sync(0xffffffffu, full0 || full1)) {
          flush_pending(0);
          flush_pending(1);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; r++) {
    flush_pending(r);
    if (row[r] < bn) {
      const uint64_t grow = (uint64_t)(b0 + row[r]);
#pragma unroll
      for (int s = 0; s < KK; s++) {
        if (s < kk) {
          const double dd = sqrt((double)bd[r][s]);
          if (out_dist) out_dist[grow * kk + s] = dd;
          if (s == kk - 1 && out_kth) out_kth[grow] = dd;
        }
      }
    }
  }
}

template <typename T>
__global__ void knn_transpose_kernel(const double* __restrict__ data, uint64_t N, int D,
                                     uint64_t Nstride, T* __restrict__ dataT) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  for (int k = 0; k < D; k++) dataT[(uint64_t)k * Nstride + i] = (T)data[i * D + k];
}

template <typename T>
static size_t knn_smem_bytes(int RC, int KP, int D, int CT) {
  return sizeof(T) * ((size_t)RC * D + RC + (size_t)RC * KP + (size_t)RC * 32 + (size_t)D * CT) +
         sizeof(int) * ((size_t)RC + (size_t)RC * KP + (size_t)RC * 32);
}

template <typename T, int D>
static int launch_knn(const double* d_data, uint64_t N, const int64_t* d_batch_off, int nb,
                      int64_t max_batch, int kk, void* d_scratch, double* out_kth,
                      double* out_dist, int32_t* out_idx, cudaStream_t st) {
  constexpr int JL = sizeof(T) == 8 ? 4 : 8;
  constexpr int CT = 32 * JL;
  const int KP = (kk + 31) / 32 * 32;
  int RC = 128;
  while (RC > KNN_WARPS && knn_smem_bytes<T>(RC, KP, D, CT) > 200 * 1024) RC /= 2;
  const size_t smem = knn_smem_bytes<T>(RC, KP, D, CT);
  if (smem > 227 * 1024) {
    set_error("knn: k+1=%d needs %zu bytes of shared memory per CTA (limit 227 KB)", kk, smem);
    return 1;
  }
  T* dataT = reinterpret_cast<T*>(d_scratch);
  knn_transpose_kernel<T><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(d_data, N, D, N, dataT);
  KDSB_CUDA(cudaGetLastError());
  if (sizeof(T) == 4 && kk <= 12 && !out_idx && !getenv("KDSB200_KNN_ONE_ROW")) {
    // fp32, distances only: two rows per thread (see knn_kth2_kernel)
    constexpr int CT2 = 256;
    const int64_t ch = (max_batch + 2 * KS_THREADS - 1) / (2 * KS_THREADS);
    if (ch * nb >= (1ll << 31)) {
      set_error("knn: grid too large (%lld CTAs)", (long long)(ch * nb));
      return 1;
    }
    static int pend_sel = -1;  // pending slots per row (KDSB200_KNN_PEND: 12 | 16 | 24, tuning)
    if (pend_sel < 0) {
      const char* e = getenv("KDSB200_KNN_PEND");
      pend_sel = e ? atoi(e) : 16;
    }
    const int pn = pend_sel == 12 ? 12 : pend_sel == 24 ? 24 : 16;
    const size_t ksm = sizeof(float) * 2 * D * CT2 + sizeof(float) * 2 * pn * KS_THREADS;
    const float* dT = reinterpret_cast<const float*>(dataT);
#define KDSB_K2P(KKV, PN)                                                                       \
  do {                                                                                          \
    KDSB_CUDA(cudaFuncSetAttribute(knn_kth2_kernel<D, KKV, CT2, PN>,                            \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksm));     \
    knn_kth2_kernel<D, KKV, CT2, PN><<<(unsigned)(ch * nb), KS_THREADS, ksm, st>>>(             \
        dT, N, d_batch_off, (int)ch, kk, out_kth, out_dist);                                    \
  } while (0)
#define KDSB_K2(KKV)                   \
  do {                                 \
    if (pn == 12) KDSB_K2P(KKV, 12);   \
    else if (pn == 24) KDSB_K2P(KKV, 24); \
    else KDSB_K2P(KKV, 16);            \
  } while (0)
    if (kk <= 4) KDSB_K2(4);
    else KDSB_K2(12);
#undef KDSB_K2P
#undef KDSB_K2
    KDSB_CUDA(cudaGetLastError());
    return 0;
  }
  if (kk <= 24) {
    const int64_t ch = (max_batch + KS_THREADS - 1) / KS_THREADS;
    if (ch * nb >= (1ll << 31)) {
      set_error("knn: grid too large (%lld CTAs)", (long long)(ch * nb));
      return 1;
    }
    const unsigned g = (unsigned)(ch * nb);
    // candidates per shared-memory tile: 256 in fp32 keeps four CTAs per SM resident next
    // to the pending lists (KDSB200_KNN_TILE overrides, for tuning)
    static int tile_sel = -1;
    if (tile_sel < 0) {
      const char* e = getenv("KDSB200_KNN_TILE");
      tile_sel = e ? atoi(e) : 0;
    }
    const int ct = sizeof(T) == 4 ? (tile_sel == 512 ? 512 : 256) : (tile_sel == 256 ? 256 : 128);
    const size_t ksm = sizeof(T) * 2 * D * ct +
                       (sizeof(T) + sizeof(int)) * KsPend<T>::N * KS_THREADS;
#define KDSB_KS2(KKV, CTV)                                                                     \
  do {                                                                                         \
    KDSB_CUDA(cudaFuncSetAttribute(knn_small_kernel<T, D, KKV, CTV>,                           \
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksm));    \
    knn_small_kernel<T, D, KKV, CTV><<<g, KS_THREADS, ksm, st>>>(dataT, N, d_batch_off,        \
                                                                 (int)ch, kk, out_kth,         \
                                                                 out_dist, out_idx);           \
  } while (0)
#define KDSB_KS(KKV)                                             \
  do {                                                           \
    if (sizeof(T) == 4) {                                        \
      if (ct == 512) KDSB_KS2(KKV, 512);                         \
      else KDSB_KS2(KKV, 256);                                   \
    } else {                                                     \
      if (ct == 256) KDSB_KS2(KKV, 256);                         \
      else KDSB_KS2(KKV, 128);                                   \
    }                                                            \
  } while (0)
    if (kk <= 4) KDSB_KS(4);
    else if (kk <= 12) KDSB_KS(12);
    else KDSB_KS(24);
#undef KDSB_KS2
#undef KDSB_KS
    KDSB_CUDA(cudaGetLastError());
    return 0;
  }
  KDSB_CUDA(cudaFuncSetAttribute(knn_kernel<T, D, JL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)smem));
  const int64_t chunks = (max_batch + RC - 1) / RC;
  if (chunks * nb >= (1ll << 31)) {
    set_error("knn: grid too large (%lld CTAs)", (long long)(chunks * nb));
    return 1;
  }
  knn_kernel<T, D, JL><<<(unsigned)(chunks * nb), KNN_THREADS, smem, st>>>(
      dataT, N, d_batch_off, (int)chunks, RC, KP, kk, out_kth, out_dist, out_idx);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace kdsb

using namespace kdsb;

extern "C" {

int kdsb200_knn(const double* d_data, uint64_t N, int D, const int64_t* d_batch_off, int nb,
                int64_t max_batch, int kk, int use_f64, void* d_scratch, double* d_kth,
                double* d_dist, int32_t* d_idx, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0 || nb == 0) return 0;
  if (kk < 1 || kk > 1024) {
    set_error("knn: k+1=%d out of range 1..1024", kk);
    return 1;
  }
  if (max_batch >= (1ll << 31)) {
    set_error("knn: batch too large (%lld rows)", (long long)max_batch);
    return 1;
  }
  switch (D) {
#define KDSB_CASE(d)                                                                            \
  case d:                                                                                       \
    return use_f64 ? launch_knn<double, d>(d_data, N, d_batch_off, nb, max_batch, kk, d_scratch, \
                                           d_kth, d_dist, d_idx, st)                            \
                   : launch_knn<float, d>(d_data, N, d_batch_off, nb, max_batch, kk, d_scratch,  \
                                          d_kth, d_dist, d_idx, st);
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
  set_error("knn: D=%d out of range 1..8", D);
  return 1;
}

}  // extern "C"
