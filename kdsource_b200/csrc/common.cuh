// Shared device/host helpers for the kdsb200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace kdsb {

// ---- error reporting (C ABI: int status + kdsb200_last_error()) ----------
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int sm_count();

#define KDSB_CUDA(call)                                   \
  do {                                                    \
    if (kdsb::check_cuda((call), #call)) return 1;        \
  } while (0)

// Packed fp32 source layout shared by kde_eval / mlcv_score:
//   tile t (TS sources) = ROWS x TS floats, ROWS = D + 2
//   rows 0..D-1 : -(x_k - c_k) * a_i        (negated, pre-scaled coordinate)
//   row  D      : a_i                        (per-source inverse length scale)
//   row  D+1    : kde: nlc_i = lc_ref - log2(c_i)  (>= 0, +inf for padding)
//                 mlcv: c_i / c_ref                 (0 for padding)
constexpr int TS = 512;  // sources per tile

__host__ __device__ inline uint64_t n_tiles_for(uint64_t n) { return (n + TS - 1) / TS; }

// ---- packed f32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2) -------------
typedef unsigned long long f2_t;  // two packed floats {lo, hi}

__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
  f2_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) {
  f2_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
  f2_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// MUFU.EX2 (flush-to-zero): 2^x, max rel. error 2^-22
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= 0, two lanes at once, entirely on the FMA pipe (plus one FMNMX and
// one shift-add per lane on the ALU pipe): Cody-Waite split by the magic-number
// trick (t = x + 1.5*2^23 leaves round(x) in the low mantissa bits), degree-5
// minimax polynomial on [-0.5, 0.5] (max rel. error 2.3e-7 in fp32, the same as
// MUFU.EX2's 2^-22), exponent inserted by adding round(x) << 23 to the bits.
// Used to take part of the exponentials off the SFU when MUFU.EX2 is the limit.
__device__ __forceinline__ f2_t ex2_poly2(f2_t x) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  const f2_t xc = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const f2_t magic = f2_pack(12582912.f, 12582912.f);
  const f2_t nmagic = f2_pack(-12582912.f, -12582912.f);
  const f2_t t = f2_add(xc, magic);
  const f2_t n = f2_add(t, nmagic);
  const f2_t mone = f2_pack(-1.f, -1.f);
  const f2_t f = f2_fma(n, mone, xc);  // x - round(x)
  f2_t p = f2_pack(0.0013276472f, 0.0013276472f);
  p = f2_fma(p, f, f2_pack(0.009675541f, 0.009675541f));
  p = f2_fma(p, f, f2_pack(0.055507131f, 0.055507131f));
  p = f2_fma(p, f, f2_pack(0.2402212f, 0.2402212f));
  p = f2_fma(p, f, f2_pack(0.69314694f, 0.69314694f));
  p = f2_fma(p, f, f2_pack(1.0000001f, 1.0000001f));
  float p0, p1, t0, t1;
  f2_unpack(p, p0, p1);
  f2_unpack(t, t0, t1);
  const float r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  const float r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
  return f2_pack(r0, r1);
}

// ---- mbarrier + 1-D bulk TMA copy ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy (TMA engine, UBLKCP in SASS); bytes % 16 == 0
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace kdsb
