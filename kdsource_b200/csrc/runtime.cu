// Error reporting, device info and pipe-rate calibration for libkdsb200.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "kdsb200.h"

namespace kdsb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return 1;
}

int sm_count() {
  static int cached = 0;
  if (!cached) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

// ---- calibration: dependent-free FFMA2 / FFMA / MUFU.EX2 loops ------------
// the FP32 and SFU pipe rates behind the roofline fractions are
// calibrated on the box the benchmark runs on, not assumed.
template <int MODE>
__global__ void __launch_bounds__(256) calib_kernel(float* out, int iters, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
  const float m = 0.999f, c = 1e-3f;
  if (MODE == 0) {  // scalar FFMA
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
        a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
      }
    }
  } else if (MODE == 1) {  // packed FFMA2
    f2_t p0 = f2_pack(a0, a1), p1 = f2_pack(a2, a3), p2 = f2_pack(a4, a5), p3 = f2_pack(a6, a7);
    f2_t p4 = f2_pack(a1, a0), p5 = f2_pack(a3, a2), p6 = f2_pack(a5, a4), p7 = f2_pack(a7, a6);
    const f2_t mm = f2_pack(m, m), cc = f2_pack(c, c);
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        p0 = f2_fma(p0, mm, cc); p1 = f2_fma(p1, mm, cc); p2 = f2_fma(p2, mm, cc); p3 = f2_fma(p3, mm, cc);
        p4 = f2_fma(p4, mm, cc); p5 = f2_fma(p5, mm, cc); p6 = f2_fma(p6, mm, cc); p7 = f2_fma(p7, mm, cc);
      }
    }
    float x, y;
    f2_unpack(f2_add(f2_add(p0, p1), f2_add(p2, p3)), x, y); a0 = x + y;
    f2_unpack(f2_add(f2_add(p4, p5), f2_add(p6, p7)), x, y); a1 = x + y;
  } else if (MODE == 3) {  // FFMA2 and scalar FFMA streams interleaved (heavy + lite pipes?)
    f2_t p0 = f2_pack(a0, a1), p1 = f2_pack(a2, a3), p2 = f2_pack(a4, a5), p3 = f2_pack(a6, a7);
    const f2_t mm = f2_pack(m, m), cc = f2_pack(c, c);
    float s0 = a0 + 9.f, s1 = a1 + 9.f, s2 = a2 + 9.f, s3 = a3 + 9.f;
    float s4 = a4 + 9.f, s5 = a5 + 9.f, s6 = a6 + 9.f, s7 = a7 + 9.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        p0 = f2_fma(p0, mm, cc); s0 = fmaf(s0, m, c); s1 = fmaf(s1, m, c);
        p1 = f2_fma(p1, mm, cc); s2 = fmaf(s2, m, c); s3 = fmaf(s3, m, c);
        p2 = f2_fma(p2, mm, cc); s4 = fmaf(s4, m, c); s5 = fmaf(s5, m, c);
        p3 = f2_fma(p3, mm, cc); s6 = fmaf(s6, m, c); s7 = fmaf(s7, m, c);
      }
    }
    float x, y;
    f2_unpack(f2_add(f2_add(p0, p1), f2_add(p2, p3)), x, y);
    a0 = x + y + s0 + s1 + s2 + s3 + s4 + s5 + s6 + s7;
  } else if (MODE == 5 || MODE == 6 || MODE == 7) {
    // packed FADD2 (5), FMUL2 (6), and the kNN distance mix FADD2 -> FFMA2 (7)
    f2_t p0 = f2_pack(a0, a1), p1 = f2_pack(a2, a3), p2 = f2_pack(a4, a5), p3 = f2_pack(a6, a7);
    f2_t p4 = f2_pack(a1, a0), p5 = f2_pack(a3, a2), p6 = f2_pack(a5, a4), p7 = f2_pack(a7, a6);
    const f2_t mm = f2_pack(m, m), cc = f2_pack(c, c);
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (MODE == 5) {
          p0 = f2_add(p0, cc); p1 = f2_add(p1, cc); p2 = f2_add(p2, cc); p3 = f2_add(p3, cc);
          p4 = f2_add(p4, cc); p5 = f2_add(p5, cc); p6 = f2_add(p6, cc); p7 = f2_add(p7, cc);
        } else if (MODE == 6) {
          p0 = f2_mul(p0, mm); p1 = f2_mul(p1, mm); p2 = f2_mul(p2, mm); p3 = f2_mul(p3, mm);
          p4 = f2_mul(p4, mm); p5 = f2_mul(p5, mm); p6 = f2_mul(p6, mm); p7 = f2_mul(p7, mm);
        } else {
          p0 = f2_add(p0, cc); p4 = f2_fma(p0, p0, p4); p1 = f2_add(p1, cc); p5 = f2_fma(p1, p1, p5);
          p2 = f2_add(p2, cc); p6 = f2_fma(p2, p2, p6); p3 = f2_add(p3, cc); p7 = f2_fma(p3, p3, p7);
        }
      }
    }
    float x, y;
    f2_unpack(f2_add(f2_add(p0, p1), f2_add(p2, p3)), x, y); a0 = x + y;
    f2_unpack(f2_add(f2_add(p4, p5), f2_add(p6, p7)), x, y); a1 = x + y;
  } else if (MODE == 4) {  // fp64 pipe: DFMA
    double d0 = a0, d1 = a1, d2 = a2, d3 = a3, d4 = a4, d5 = a5, d6 = a6, d7 = a7;
    const double dm = 0.999, dc = 1e-3;
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        d0 = fma(d0, dm, dc); d1 = fma(d1, dm, dc); d2 = fma(d2, dm, dc); d3 = fma(d3, dm, dc);
        d4 = fma(d4, dm, dc); d5 = fma(d5, dm, dc); d6 = fma(d6, dm, dc); d7 = fma(d7, dm, dc);
      }
    }
    a0 = (float)(d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7);
    a1 = a2 = a3 = a4 = a5 = a6 = a7 = 0.f;
  } else {  // MUFU.EX2
    a0 = -1e-3f * a0; a1 = -1e-3f * a1; a2 = -1e-3f * a2; a3 = -1e-3f * a3;
    a4 = -1e-3f * a4; a5 = -1e-3f * a5; a6 = -1e-3f * a6; a7 = -1e-3f * a7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int u = 0; u < 8; u++) {
        a0 = ex2_approx(a0); a1 = ex2_approx(a1); a2 = ex2_approx(a2); a3 = ex2_approx(a3);
        a4 = ex2_approx(a4); a5 = ex2_approx(a5); a6 = ex2_approx(a6); a7 = ex2_approx(a7);
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

}  // namespace kdsb

using namespace kdsb;

extern "C" {

const char* kdsb200_last_error(void) { return g_err; }
int kdsb200_version(void) { return 1000; }

int kdsb200_device_info(int* sm, int* cc_major, int* cc_minor, int* clock_khz) {
  int dev = 0;
  KDSB_CUDA(cudaGetDevice(&dev));
  KDSB_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
  KDSB_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  KDSB_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  KDSB_CUDA(cudaDeviceGetAttribute(clock_khz, cudaDevAttrClockRate, dev));
  return 0;
}

// mode 0: FFMA, 1: FFMA2, 2: MUFU.EX2, 3: FFMA2+FFMA interleaved, 4: DFMA, 5: FADD2, 6: FMUL2,
// 7: FADD2+FFMA2 (the kNN distance mix).  Returns lane-operations per second
// (FFMA2 counts two per lane-instruction) measured with CUDA events.
int kdsb200_calibrate_pipe(int mode, float* d_scratch /* >= sm*8*256 floats */, int iters,
                           double* ops_per_s, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = sm_count() * 8;
  cudaEvent_t e0, e1;
  KDSB_CUDA(cudaEventCreate(&e0));
  KDSB_CUDA(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {
    KDSB_CUDA(cudaEventRecord(e0, st));
    if (mode == 0) calib_kernel<0><<<blocks, 256, 0, st>>>(d_scratch, iters, 1.0f);
    else if (mode == 1) calib_kernel<1><<<blocks, 256, 0, st>>>(d_scratch, iters, 1.0f);
    else if (mode == 3) calib_kernel<3><<<blocks, 256, 0, st>>>(d_scratch, iters, 1.0f);
    else if (mode == 4) calib_kernel<4><<<blocks, 256, 0, st>>>(d_scratch, iters, 1.0f);
    else if (mode == 5) calib_kernel<5><<<blocks, 256, 0, st>>>(d_scratch, iters, 1.0f);
    else if (mode == 6) calib_kernel<6><<<blocks, 256, 0, st>>>(d_scratch, iters, 1.0f);
    else if (mode == 7) calib_kernel<7><<<blocks, 256, 0, st>>>(d_scratch, iters, 1.0f);
    else calib_kernel<2><<<blocks, 256, 0, st>>>(d_scratch, iters, 1.0f);
    KDSB_CUDA(cudaEventRecord(e1, st));
    KDSB_CUDA(cudaEventSynchronize(e1));
  }
  float ms = 0.f;
  KDSB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double per_thread = (double)iters * 64.0;  // 64 lane-ops per iteration in every mode
  *ops_per_s = per_thread * blocks * 256.0 / (ms * 1e-3);
  return 0;
}

}  // extern "C"
