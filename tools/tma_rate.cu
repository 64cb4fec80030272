// tma_rate.cu -- how fast can ONE thread per SM stream global memory into shared memory with cp.async.bulk,
// and where do the producer thread's cycles go (wait / expect_tx / bulk-copy issue)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_rate tools/tma_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../osqp-python_b200/csrc/device_common.cuh"
using namespace b200qp;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

template <int S, int NCP>
__global__ void __launch_bounds__(32, 1) k_rate(const char* src, unsigned per_cta_mask, int bytes, int ncopy, long long* prof) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ alignas(8) uint64_t full[S];
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
    const char* base = src + (size_t)blockIdx.x * ((size_t)per_cta_mask + 1);
    long long tw = 0, te = 0, tb = 0;
    unsigned off = 0;
    int st = 0, ph = 0;
    const uint32_t part = (uint32_t)bytes / NCP;
    for (int t = 0; t < ncopy; ++t) {
      const long long c0 = clock64();
      if (t >= S) mbar_wait(&full[st], ph ^ 1);
      const long long c1 = clock64();
      mbar_expect_tx(&full[st], part * NCP);
      const long long c2 = clock64();
#pragma unroll
      for (int c = 0; c < NCP; ++c) bulk_g2s(smem + (size_t)st * bytes + c * part, base + off + c * part, part, &full[st]);
      const long long c3 = clock64();
      tw += c1 - c0; te += c2 - c1; tb += c3 - c2;
      off = (off + (unsigned)bytes) & per_cta_mask;
      if (++st == S) { st = 0; ph ^= 1; }
    }
    for (int k = 0; k < S; ++k) { mbar_wait(&full[st], ph ^ 1); if (++st == S) { st = 0; ph ^= 1; } }
    if (blockIdx.x == 0) { prof[0] = tw; prof[1] = te; prof[2] = tb; }
  }
}

template <int S, int NCP>
static void run(const char* src, size_t total, int bytes, int ctas_per_sm, long long* d_prof) {
  if ((size_t)S * bytes * ctas_per_sm > 200 * 1024) return;
  auto kern = k_rate<S, NCP>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int G = 148 * ctas_per_sm;
  size_t per_cta = 1; while (per_cta * 2 * G <= total) per_cta *= 2;
  const int ncopy = (int)((8u << 20) / bytes / ctas_per_sm);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e9;
  for (int r = 0; r < 3; ++r) {
    CK(cudaEventRecord(e0));
    kern<<<G, 32, (size_t)S * bytes>>>(src, (unsigned)(per_cta - 1), bytes, ncopy, d_prof);
    CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best;
  }
  long long h[3]; CK(cudaMemcpy(h, d_prof, sizeof h, cudaMemcpyDeviceToHost));
  printf("copies/tile=%d bytes=%6d S=%2d cta/SM=%d : %7.1f us %6.2f TB/s %6.1f ns/tile/producer | clk per tile: wait %5.0f expect %4.0f issue %4.0f\n",
         NCP, bytes, S, ctas_per_sm, best * 1e3, (double)ncopy * bytes * G / (best * 1e-3) * 1e-12, best * 1e6 / ncopy,
         (double)h[0] / ncopy, (double)h[1] / ncopy, (double)h[2] / ncopy);
  CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

int main() {
  const size_t total = (size_t)2 << 30;
  char* src; CK(cudaMalloc(&src, total)); CK(cudaMemset(src, 1, total));
  long long* d_prof; CK(cudaMalloc(&d_prof, 64));
  const int sizes[] = {1536, 3840, 7680, 15360};
  for (int bytes : sizes)
    for (int c = 1; c <= 2; ++c) {
      run<4, 1>(src, total, bytes, c, d_prof);
      run<8, 1>(src, total, bytes, c, d_prof);
      run<16, 1>(src, total, bytes, c, d_prof);
      run<32, 1>(src, total, bytes, c, d_prof);
      run<16, 2>(src, total, bytes, c, d_prof);
      run<16, 3>(src, total, bytes, c, d_prof);
    }
  return 0;
}
