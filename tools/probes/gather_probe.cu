// Random-gather microbenchmark: how many DRAM bytes does a random access of W bytes at
// stride S cost on this GPU, and what rate does it reach with high memory-level
// parallelism?  (Design input for the resampler's pick table, DESIGN.md.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe.bin gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}

template <int W>  // bytes read per access: 16, 32, 64, 128
__global__ void __launch_bounds__(256) gather(const uint4* __restrict__ buf, uint64_t nslots,
                                              int stride16, int iters, uint4* __restrict__ out) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int it = 0; it < iters; it += 4) {
    uint4 v[4][W / 16];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint64_t slot = mix(tid * 131 + it + u) % nslots;
      const uint4* p = buf + slot * stride16;
#pragma unroll
      for (int c = 0; c < W / 16; c++) v[u][c] = __ldg(p + c);
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int c = 0; c < W / 16; c++) { acc.x ^= v[u][c].x; acc.y ^= v[u][c].y; acc.z ^= v[u][c].z; acc.w ^= v[u][c].w; }
  }
  if (acc.x == 0x12345678u) out[tid & 1023] = acc;
}

// G lanes share one access of G*16 bytes: one coalesced request per line instead of G
template <int G>
__global__ void __launch_bounds__(256) gather_coop(const uint4* __restrict__ buf, uint64_t nslots,
                                                   int stride16, int iters, uint4* __restrict__ out) {
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t grp = tid / G;
  const int sub = (int)(tid % G);
  uint4 acc = make_uint4(0, 0, 0, 0);
  for (int it = 0; it < iters * G; it += 4) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint64_t slot = mix(grp * 131 + it + u) % nslots;
      v[u] = __ldg(buf + slot * stride16 + sub);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
  }
  if (acc.x == 0x12345678u) out[tid & 1023] = acc;
}

int main(int argc, char** argv) {
  const uint64_t bytes = (argc > 1 ? atoll(argv[1]) : 4096ll) << 20;
  uint4* buf; uint4* out;
  cudaMalloc(&buf, bytes); cudaMalloc(&out, 1024 * 16);
  cudaMemset(buf, 1, bytes);
  const int iters = 16;
  const unsigned blocks = 148 * 8 * 16;
  const double accesses = (double)blocks * 256 * iters;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int cfg[][2] = {{16, 32}, {32, 32}, {16, 64}, {32, 64}, {64, 64}, {16, 128}, {32, 128}, {64, 128}, {128, 128}};
  for (auto& c : cfg) {
    const int W = c[0], S = c[1];
    const uint64_t nslots = bytes / S;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      if (W == 16) gather<16><<<blocks, 256>>>(buf, nslots, S / 16, iters, out);
      else if (W == 32) gather<32><<<blocks, 256>>>(buf, nslots, S / 16, iters, out);
      else if (W == 64) gather<64><<<blocks, 256>>>(buf, nslots, S / 16, iters, out);
      else gather<128><<<blocks, 256>>>(buf, nslots, S / 16, iters, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    printf("buf %llu MB  W=%3d B at stride %3d B: %.3f ms  %.3e accesses/s  useful %.1f GB/s\n",
           (unsigned long long)(bytes >> 20), W, S, best, accesses / best * 1e3, accesses * W / best * 1e3 / 1e9);
  }
  for (int G = 2; G <= 8; G *= 2) {
    const uint64_t nslots = bytes / 128;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      if (G == 2) gather_coop<2><<<blocks, 256>>>(buf, nslots, 8, iters, out);
      else if (G == 4) gather_coop<4><<<blocks, 256>>>(buf, nslots, 8, iters, out);
      else gather_coop<8><<<blocks, 256>>>(buf, nslots, 8, iters, out);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep && ms < best) best = ms;
    }
    printf("buf %llu MB  cooperative: %d lanes x 16 B of one 128 B line: %.3f ms  %.3e accesses/s  useful %.1f GB/s\n",
           (unsigned long long)(bytes >> 20), G, best, accesses / best * 1e3, accesses * G * 16 / best * 1e3 / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
