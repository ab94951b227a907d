// mcpl_unpack: raw MCPL-3 particle records -> (N,8) fp64 particles, weights, pdg codes,
// one record per thread.  Device-side counterpart of the decode step of PList_get /
// mcpl_read (clib/src/plist.c:69-89; python/kdsource/plist.py:349-395): the resampler
// consumes the list on the device, so the records are uploaded as they lie in the file
// (36 B per particle in single precision instead of 72 B of fp64 columns) and decoded
// here, including libmcpl's adaptive-projection unit-vector unpacking.
//
// Arithmetic is fp64 without FMA contraction, so the result is bit-identical to the
// host (numpy) decoder in kdsource_b200/mcpl.py.
#include <math.h>

#include "common.cuh"
#include "kdsb200.h"

namespace kdsb {

template <typename F>
__device__ __forceinline__ double ld_fp(const unsigned char* p) {
  F v;
  memcpy(&v, p, sizeof(F));
  return (double)v;
}

template <typename F>
__global__ void __launch_bounds__(256)
    mcpl_unpack_kernel(const unsigned char* __restrict__ raw, uint64_t n, kdsb200_mcpl_layout L,
                       double* __restrict__ parts, double* __restrict__ w,
                       int32_t* __restrict__ pdg) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned char* r = raw + i * L.record_bytes;
  constexpr int fp = sizeof(F);
  const double x = ld_fp<F>(r + L.off_pos), y = ld_fp<F>(r + L.off_pos + fp),
               z = ld_fp<F>(r + L.off_pos + 2 * fp);
  const double p0 = ld_fp<F>(r + L.off_dir), p1 = ld_fp<F>(r + L.off_dir + fp);
  const double sek = ld_fp<F>(r + L.off_dir + 2 * fp);
  const double t = ld_fp<F>(r + L.off_time);
  // mcpl_unitvect_unpack_adaptproj: (1/z, y | sign x), (x, 1/z | sign y), (x, y | sign z)
  const double sgn = signbit(sek) ? -1.0 : 1.0;
  double ux, uy, uz;
  if (fabs(p0) > 1.0) {
    uz = 1.0 / p0;
    uy = p1;
    ux = sgn * sqrt(fmax(0.0, __dsub_rn(1.0, __dadd_rn(__dmul_rn(p1, p1), __dmul_rn(uz, uz)))));
  } else if (fabs(p1) > 1.0) {
    uz = 1.0 / p1;
    ux = p0;
    uy = sgn * sqrt(fmax(0.0, __dsub_rn(1.0, __dadd_rn(__dmul_rn(p0, p0), __dmul_rn(uz, uz)))));
  } else {
    ux = p0;
    uy = p1;
    uz = sgn * sqrt(fmax(0.0, __dsub_rn(1.0, __dadd_rn(__dmul_rn(p0, p0), __dmul_rn(p1, p1)))));
  }
  double* o = parts + i * 8;
  o[0] = fabs(sek);
  o[1] = x; o[2] = y; o[3] = z;
  o[4] = ux; o[5] = uy; o[6] = uz;
  o[7] = t;
  w[i] = L.off_weight >= 0 ? ld_fp<F>(r + L.off_weight) : L.universal_weight;
  if (L.off_pdg >= 0) {
    int32_t c;
    memcpy(&c, r + L.off_pdg, 4);
    pdg[i] = c;
  } else {
    pdg[i] = L.universal_pdg;
  }
}

}  // namespace kdsb

using namespace kdsb;

extern "C" int kdsb200_mcpl_unpack(const void* d_raw, uint64_t n, const kdsb200_mcpl_layout* h_layout,
                                   double* d_parts8, double* d_weights, int32_t* d_pdg,
                                   void* stream) {
  if (n == 0) return 0;
  const kdsb200_mcpl_layout& L = *h_layout;
  const int fp = L.single_precision ? 4 : 8;
  const int need = (L.off_pdg >= 0 ? L.off_pdg + 4 : 0);
  if (L.record_bytes <= 0 || L.off_pos < 0 || L.off_dir < 0 || L.off_time < 0 ||
      L.off_pos + 3 * fp > L.record_bytes || L.off_dir + 3 * fp > L.record_bytes ||
      L.off_time + fp > L.record_bytes || need > L.record_bytes ||
      (L.off_weight >= 0 && L.off_weight + fp > L.record_bytes)) {
    set_error("mcpl_unpack: inconsistent record layout (record_bytes=%d)", L.record_bytes);
    return 1;
  }
  const unsigned blocks = (unsigned)((n + 255) / 256);
  if (L.single_precision)
    mcpl_unpack_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const unsigned char*)d_raw, n, L, d_parts8, d_weights, d_pdg);
  else
    mcpl_unpack_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(
        (const unsigned char*)d_raw, n, L, d_parts8, d_weights, d_pdg);
  KDSB_CUDA(cudaGetLastError());
  return 0;
}
