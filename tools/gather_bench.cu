// gather_bench.cu -- microbenchmarks behind the SpMV design decisions:
//   A) random fp64 gather rate per SM from  (0) L2 through L1TEX (ld.global.cg)
//                                            (1) the CTA's own shared memory
//                                            (2) distributed shared memory of a thread-block cluster
//   B) a column-blocked CSR SpMV whose gather operand lives in cluster shared memory, against the
//      plain L2-gather CSR SpMV on the same uniformly random matrix (n = m = 1e6, 10 nnz/row).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/gather_bench tools/gather_bench.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../osqp-python_b200/csrc/tile_spmv.cuh"
#include "../osqp-python_b200/csrc/tiling.h"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, \
              cudaGetErrorString(e_));                                                          \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ double ld_dsmem(uint32_t addr) {
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ double ld_cg(const double* p) { return __ldcg(p); }
__device__ __forceinline__ double ld_stream(const double* p) {
  double v;
  asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
  int v;
  asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// ------------------------------------------------------------------------------------------ A
// MODE 0: L2 gather; 1: own shared memory; 2: cluster shared memory (mapa table); 3: cluster, linear address
template <int MODE>
__global__ void __launch_bounds__(512, 1)
k_gather_rate(const double* __restrict__ xg, uint32_t range, int W, int per_thread, double* out, unsigned* info) {
  extern __shared__ __align__(16) double sx[];
  __shared__ uint32_t s_base[16];
  const int tid = threadIdx.x;
  uint32_t C = 1, rank = 0;
  if (MODE >= 1) {
    for (int i = tid; i < W; i += blockDim.x) sx[i] = 1.0 + 1e-6 * i;
  }
  if (MODE >= 2) {
    C = cluster_nctarank();
    rank = cluster_ctarank();
    if (tid < 16) s_base[tid] = mapa(smem_u32(sx), tid < (int)C ? tid : 0);
    __syncthreads();
    if (info && blockIdx.x == 0 && tid < 16) info[tid] = s_base[tid];
    cluster_sync();
  } else {
    __syncthreads();
  }
  const uint32_t base0 = (MODE >= 2) ? s_base[0] : 0;
  const uint32_t stride = (MODE >= 2 && C > 1) ? s_base[1] - s_base[0] : 0;
  uint32_t s = (blockIdx.x * blockDim.x + tid) * 2654435761u + 12345u;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int it = 0; it < per_thread; it += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s = s * 1664525u + 1013904223u;
      const uint32_t h = s ^ (s >> 15);
      double v;
      if (MODE == 0) v = ld_cg(xg + (uint32_t)(((uint64_t)h * range) >> 32));
      else if (MODE == 1) v = sx[(uint32_t)(((uint64_t)h * (uint32_t)W) >> 32)];
      else {
        const uint32_t r = (h >> 27) & (C - 1), off = (uint32_t)(((uint64_t)(h & 0x7ffffffu) * (uint32_t)W) >> 27);
        v = (MODE == 2) ? ld_dsmem(s_base[r] + off * 8u) : ld_dsmem(base0 + r * stride + off * 8u);
      }
      acc[u] += v;
    }
  }
  double t = 0;
  for (int u = 0; u < 8; ++u) t += acc[u];
  if (t == -1.0) out[0] = t;
  if (MODE >= 2) cluster_sync();
}

// ------------------------------------------------------------------------------------------ B
// plain CSR, LG lanes per row, operand gathered from L2
template <int LG>
__global__ void __launch_bounds__(256) k_spmv_l2(int rows, const int* __restrict__ ptr, const int* __restrict__ idx,
                                                 const double* __restrict__ val, const double* __restrict__ x,
                                                 double* __restrict__ y) {
  const int lane = threadIdx.x & (LG - 1);
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) / LG);
  double acc = 0.0;
  if (row < rows) {
    const int a = __ldg(ptr + row), b = __ldg(ptr + row + 1);
#pragma unroll 4
    for (int k = a + lane; k < b; k += LG) acc = fma(ld_stream(val + k), ld_cg(x + ld_stream(idx + k)), acc);
  }
#pragma unroll
  for (int off = LG / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (row < rows && lane == 0) y[row] = acc;
}

// persistent grid, contiguous row range per CTA, two rows per group in flight (the in-kernel scheme of admm_kernel.cuh)
template <int LG>
__global__ void __launch_bounds__(512, 2) k_spmv_pers(int rows, const int* __restrict__ ptr, const int* __restrict__ idx,
                                                      const double* __restrict__ val, const double* __restrict__ x,
                                                      double* __restrict__ y) {
  const int G = gridDim.x, cta = blockIdx.x, tid = threadIdx.x;
  const int m0 = (int)((long long)rows * cta / G), m1 = (int)((long long)rows * (cta + 1) / G);
  constexpr int NG = 512 / LG;
  const int lane = tid & (LG - 1), grp = tid / LG;
  for (int base = m0; base < m1; base += 2 * NG) {
    const int rowA = base + grp, rowB = rowA + NG;
    const bool okA = rowA < m1, okB = rowB < m1;
    int a1 = 0, b1 = 0, a2 = 0, b2 = 0;
    if (okA) { a1 = __ldg(ptr + rowA); b1 = __ldg(ptr + rowA + 1); }
    if (okB) { a2 = __ldg(ptr + rowB); b2 = __ldg(ptr + rowB + 1); }
    const int k10 = a1 + lane, k11 = k10 + LG, k20 = a2 + lane, k21 = k20 + LG;
    const bool h10 = k10 < b1, h11 = k11 < b1, h20 = k20 < b2, h21 = k21 < b2;
    int j10 = 0, j11 = 0, j20 = 0, j21 = 0;
    double v10 = 0.0, v11 = 0.0, v20 = 0.0, v21 = 0.0;
    if (h10) { j10 = ld_stream(idx + k10); v10 = ld_stream(val + k10); }
    if (h20) { j20 = ld_stream(idx + k20); v20 = ld_stream(val + k20); }
    if (h11) { j11 = ld_stream(idx + k11); v11 = ld_stream(val + k11); }
    if (h21) { j21 = ld_stream(idx + k21); v21 = ld_stream(val + k21); }
    double y10 = 0.0, y11 = 0.0, y20 = 0.0, y21 = 0.0;
    if (h10) y10 = ld_cg(x + j10);
    if (h20) y20 = ld_cg(x + j20);
    if (h11) y11 = ld_cg(x + j11);
    if (h21) y21 = ld_cg(x + j21);
    double accA = fma(v11, y11, v10 * y10), accB = fma(v21, y21, v20 * y20);
#pragma unroll 4
    for (int k = k11 + LG; k < b1; k += LG) accA = fma(ld_stream(val + k), ld_cg(x + ld_stream(idx + k)), accA);
#pragma unroll 4
    for (int k = k21 + LG; k < b2; k += LG) accB = fma(ld_stream(val + k), ld_cg(x + ld_stream(idx + k)), accB);
#pragma unroll
    for (int off = LG / 2; off > 0; off >>= 1) {
      accA += __shfl_xor_sync(0xffffffffu, accA, off);
      accB += __shfl_xor_sync(0xffffffffu, accB, off);
    }
    if (lane == 0) { if (okA) y[rowA] = accA; if (okB) y[rowB] = accB; }
  }
}

// column-blocked CSR; one cluster per row slab; the operand block sits in the cluster's shared memory.
// bptr: B x (rows+1) segment pointers into idx/val (entries sorted by block, then row); idx packs (rank<<16 | offset).
template <int LG>
__global__ void __launch_bounds__(512, 1)
k_spmv_dsmem(int rows, int cols, int B, int W, const int* __restrict__ bptr, const int* __restrict__ idx,
             const double* __restrict__ val, const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ __align__(16) double sx[];
  __shared__ uint32_t s_base[16];
  const int tid = threadIdx.x;
  const uint32_t C = cluster_nctarank(), rank = cluster_ctarank();
  const int ncl = gridDim.x / C, cl = blockIdx.x / C;
  if (tid < 16) s_base[tid] = mapa(smem_u32(sx), tid < (int)C ? tid : 0);
  __syncthreads();
  const uint32_t base0 = s_base[0], stride = C > 1 ? s_base[1] - s_base[0] : 0;
  // rows of this CTA
  const long long slab0 = (long long)rows * cl / ncl, slab1 = (long long)rows * (cl + 1) / ncl;
  const int r0 = (int)(slab0 + (slab1 - slab0) * rank / C), r1 = (int)(slab0 + (slab1 - slab0) * (rank + 1) / C);
  const int lane = tid & (LG - 1), grp = tid / LG, ngrp = 512 / LG;
  for (int b = 0; b < B; ++b) {
    const long long c0 = ((long long)b * C + rank) * W;
    for (int i = tid; i < W; i += 512) sx[i] = (c0 + i < cols) ? ld_cg(x + c0 + i) : 0.0;
    cluster_sync();
    const int* p = bptr + (size_t)b * (rows + 1);
    for (int rb = r0; rb < r1; rb += 2 * ngrp) {
      const int rowA = rb + grp, rowB = rowA + ngrp;
      int a0 = 0, e0 = 0, a1 = 0, e1 = 0;
      if (rowA < r1) { a0 = __ldg(p + rowA); e0 = __ldg(p + rowA + 1); }
      if (rowB < r1) { a1 = __ldg(p + rowB); e1 = __ldg(p + rowB + 1); }
      double accA = 0.0, accB = 0.0;
      int kA = a0 + lane, kB = a1 + lane;
      while (kA < e0 || kB < e1) {
        const bool pa = kA < e0, pb = kB < e1;
        int ia = 0, ib = 0;
        double va = 0.0, vb = 0.0;
        if (pa) { ia = ld_stream(idx + kA); va = ld_stream(val + kA); }
        if (pb) { ib = ld_stream(idx + kB); vb = ld_stream(val + kB); }
        if (pa) accA = fma(va, ld_dsmem(base0 + (uint32_t)(ia >> 16) * stride + (uint32_t)(ia & 0xffff) * 8u), accA);
        if (pb) accB = fma(vb, ld_dsmem(base0 + (uint32_t)(ib >> 16) * stride + (uint32_t)(ib & 0xffff) * 8u), accB);
        kA += LG; kB += LG;
      }
#pragma unroll
      for (int off = LG / 2; off > 0; off >>= 1) {
        accA += __shfl_xor_sync(0xffffffffu, accA, off);
        accB += __shfl_xor_sync(0xffffffffu, accB, off);
      }
      if (lane == 0) {
        if (rowA < r1) y[rowA] = (b == 0) ? accA : y[rowA] + accA;
        if (rowB < r1) y[rowB] = (b == 0) ? accB : y[rowB] + accB;
      }
    }
    cluster_sync();
  }
}

// same layout but the operand block is gathered from this CTA's own shared memory only (C == 1): measures the
// local-shared-memory ceiling with many column blocks (partials accumulate in y)
// (covered by k_spmv_dsmem with cluster size 1)

static double now_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

template <typename K>
static void launch_cluster(K kern, int grid, int C, size_t smem, void** args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(512);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelExC(&cfg, (const void*)kern, args));
}

template <typename K>
static int max_clusters(K kern, int C, size_t smem) {
  CK(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (C > 8) CK(cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C * 64);
  cfg.blockDim = dim3(512);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int nc = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, (const void*)kern, &cfg);
  if (e != cudaSuccess) { printf("  maxActiveClusters(C=%d) failed: %s\n", C, cudaGetErrorString(e)); cudaGetLastError(); return 0; }
  return nc;
}

int main(int argc, char** argv) {
  int n = 1000000, m = 1000000, per_row = 10;
  if (argc > 1) n = m = atoi(argv[1]);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int SMs = prop.multiProcessorCount;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  printf("device %s SMs %d clock %d kHz smem/block optin %zu\n", prop.name, SMs, clk_khz, prop.sharedMemPerBlockOptin);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double *d_x, *d_out; unsigned* d_info;
  CK(cudaMalloc(&d_x, (size_t)n * 8)); CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_info, 64));
  std::vector<double> hx(n);
  std::mt19937_64 rng(7);
  std::normal_distribution<double> nd(0.0, 1.0);
  for (auto& v : hx) v = nd(rng);
  CK(cudaMemcpy(d_x, hx.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  char* d_flush; const size_t FL = 512ull << 20;
  CK(cudaMalloc(&d_flush, FL));

  // ---------------------------------------------------------------- A: gather rates
  const int W = 27136;   // doubles per CTA in shared memory (212 KB)
  const size_t smem = (size_t)W * 8;
  const int per_thread = 4096;
  auto report = [&](const char* name, int ctas, double ms) {
    const double g = (double)ctas * 512 * per_thread;
    printf("A  %-28s ctas %4d  %8.3f ms  %7.2f Ggather/s  %.3f gathers/clk/SM (at %d MHz)  -> 1e7 gathers on these SMs: %.1f us\n",
           name, ctas, ms, g / ms * 1e-6, g / (ms * 1e-3) / ctas / (clk_khz * 1e3), clk_khz / 1000, 1e7 / (g / ms) * 1e3);
  };
  {
    CK(cudaFuncSetAttribute((const void*)k_gather_rate<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute((const void*)k_gather_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_gather_rate<0><<<SMs, 512, 0>>>(d_x, (uint32_t)n, W, per_thread, d_out, nullptr);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      if (rep) report("L2 (ld.global.cg) 1 CTA/SM", SMs, now_ms(e0, e1));
    }
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_gather_rate<0><<<SMs * 2, 512, 0>>>(d_x, (uint32_t)n, W, per_thread / 2, d_out, nullptr);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      if (rep) { const double ms = now_ms(e0, e1); const double g = (double)SMs * 512 * per_thread;
        printf("A  %-28s ctas %4d  %8.3f ms  %7.2f Ggather/s  %.3f gathers/clk/SM\n", "L2 2 CTA/SM (1024 thr)", SMs * 2, ms, g / ms * 1e-6,
               g / (ms * 1e-3) / SMs / (clk_khz * 1e3)); }
    }
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      k_gather_rate<1><<<SMs, 512, smem>>>(d_x, (uint32_t)W, W, per_thread, d_out, nullptr);
      CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      if (rep) report("own shared memory", SMs, now_ms(e0, e1));
    }
    for (int C : {2, 4, 8, 16}) {
      if (!getenv("GB_DSMEM")) break;
      for (int mode = 2; mode <= 3; ++mode) {
        auto kern = mode == 2 ? k_gather_rate<2> : k_gather_rate<3>;
        const int nc = max_clusters(kern, C, smem);
        if (nc <= 0) continue;
        const int ctas = nc * C;
        uint32_t range = (uint32_t)(C * W);
        int Wv = W, pt = per_thread;
        double* outp = d_out; const double* xp = d_x; unsigned* ip = d_info;
        void* args[] = {&xp, &range, &Wv, &pt, &outp, &ip};
        for (int rep = 0; rep < 2; ++rep) {
          CK(cudaEventRecord(e0));
          launch_cluster(kern, ctas, C, smem, args);
          CK(cudaEventRecord(e1));
          cudaError_t e = cudaEventSynchronize(e1);
          if (e != cudaSuccess) { printf("  cluster %d launch failed: %s\n", C, cudaGetErrorString(e)); return 1; }
          if (rep) {
            char nm[64]; snprintf(nm, sizeof nm, "cluster smem C=%d %s", C, mode == 2 ? "table" : "linear");
            report(nm, ctas, now_ms(e0, e1));
          }
        }
        if (mode == 2) {
          unsigned hi[16]; CK(cudaMemcpy(hi, d_info, 64, cudaMemcpyDeviceToHost));
          printf("   max active clusters %d; mapa bases:", nc);
          for (int i = 0; i < C; ++i) printf(" %08x", hi[i]);
          printf("\n");
        }
      }
    }
  }

  // ---------------------------------------------------------------- B: SpMV
  printf("building %d x %d matrix, %d nnz/row ...\n", m, n, per_row); fflush(stdout);
  std::vector<int> ptr(m + 1), idx((size_t)m * per_row);
  std::vector<double> val((size_t)m * per_row);
  {
    std::uniform_int_distribution<int> ud(0, n - 1);
    std::vector<int> cols;
    for (int i = 0; i < m; ++i) {
      cols.clear(); cols.push_back(i % n);
      while ((int)cols.size() < per_row) { int c = ud(rng); if (std::find(cols.begin(), cols.end(), c) == cols.end()) cols.push_back(c); }
      std::sort(cols.begin(), cols.end());
      ptr[i] = i * per_row;
      for (int k = 0; k < per_row; ++k) { idx[(size_t)i * per_row + k] = cols[k]; val[(size_t)i * per_row + k] = nd(rng); }
    }
    ptr[m] = m * per_row;
  }
  const size_t nnz = idx.size();
  std::vector<double> yref(m);
  for (int i = 0; i < m; ++i) { double a = 0; for (int k = ptr[i]; k < ptr[i + 1]; ++k) a += val[k] * hx[idx[k]]; yref[i] = a; }
  int *d_ptr, *d_idx; double *d_val, *d_y;
  CK(cudaMalloc(&d_ptr, (size_t)(m + 1) * 4)); CK(cudaMalloc(&d_idx, (nnz + 8) * 4)); CK(cudaMalloc(&d_val, (nnz + 8) * 8));
  CK(cudaMalloc(&d_y, (size_t)m * 8));
  CK(cudaMemcpy(d_ptr, ptr.data(), (size_t)(m + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx, idx.data(), nnz * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_val, val.data(), nnz * 8, cudaMemcpyHostToDevice));
  std::vector<double> hy(m);
  auto check = [&](const char* name) {
    CK(cudaMemcpy(hy.data(), d_y, (size_t)m * 8, cudaMemcpyDeviceToHost));
    double err = 0, sc = 0;
    for (int i = 0; i < m; ++i) { err = std::max(err, std::fabs(hy[i] - yref[i])); sc = std::max(sc, std::fabs(yref[i])); }
    printf("   %s max abs err %.3e (scale %.3e)\n", name, err, sc);
  };
  const double algo = 12.0 * nnz + 4.0 * (m + 1) + 8.0 * m + 8.0 * n;
  auto time_it = [&](const char* name, auto&& launch) {
    double best = 1e30, sum = 0; const int R = 6;
    for (int rep = 0; rep < R + 1; ++rep) {
      CK(cudaMemsetAsync(d_flush, rep, FL));
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      cudaError_t e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) { printf("  %s failed: %s\n", name, cudaGetErrorString(e)); exit(1); }
      const double ms = now_ms(e0, e1);
      if (rep) { best = std::min(best, ms); sum += ms; }
    }
    printf("B  %-34s best %7.1f us  mean %7.1f us  -> %.2f TB/s algorithmic (best)\n", name, best * 1e3, sum / R * 1e3, algo / (best * 1e-3) * 1e-12);
  };
  time_it("csr L2-gather LG=8", [&] { k_spmv_l2<8><<<(int)(((long long)m * 8 + 255) / 256), 256>>>(m, d_ptr, d_idx, d_val, d_x, d_y); });
  check("csr L2-gather");
  time_it("csr L2-gather LG=4", [&] { k_spmv_l2<4><<<(int)(((long long)m * 4 + 255) / 256), 256>>>(m, d_ptr, d_idx, d_val, d_x, d_y); });
  time_it("csr L2-gather LG=2", [&] { k_spmv_l2<2><<<(int)(((long long)m * 2 + 255) / 256), 256>>>(m, d_ptr, d_idx, d_val, d_x, d_y); });
  check("csr L2-gather LG=2");
  time_it("csr L2-gather LG=2 128thr", [&] { k_spmv_l2<2><<<(int)(((long long)m * 2 + 127) / 128), 128>>>(m, d_ptr, d_idx, d_val, d_x, d_y); });
  time_it("csr L2-gather LG=4 128thr", [&] { k_spmv_l2<4><<<(int)(((long long)m * 4 + 127) / 128), 128>>>(m, d_ptr, d_idx, d_val, d_x, d_y); });
  for (int per_sm = 2; per_sm <= 4; ++per_sm) {
    char nm[64];
    snprintf(nm, sizeof nm, "csr persistent LG=4 pair %d CTA/SM", per_sm);
    time_it(nm, [&] { k_spmv_pers<4><<<SMs * per_sm, 512>>>(m, d_ptr, d_idx, d_val, d_x, d_y); });
    check(nm);
    snprintf(nm, sizeof nm, "csr persistent LG=2 pair %d CTA/SM", per_sm);
    time_it(nm, [&] { k_spmv_pers<2><<<SMs * per_sm, 512>>>(m, d_ptr, d_idx, d_val, d_x, d_y); });
    snprintf(nm, sizeof nm, "csr persistent LG=8 pair %d CTA/SM", per_sm);
    time_it(nm, [&] { k_spmv_pers<8><<<SMs * per_sm, 512>>>(m, d_ptr, d_idx, d_val, d_x, d_y); });
  }

  {
    using namespace b200qp;
    std::vector<int4> tiles; std::vector<int> cta_tile, part;
    make_tiles(m, ptr.data(), TS_TILE, TS_MAXROWS, tiles);
    CK(cudaFuncSetAttribute((const void*)k_spmv_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_tile, NT, TS_SMEM));
    int4* d_tiles; int* d_ct;
    CK(cudaMalloc(&d_tiles, tiles.size() * sizeof(int4)));
    CK(cudaMemcpy(d_tiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice));
    TileMat M{m, n, (int)nnz, (int)tiles.size(), d_ptr, d_idx, d_val, d_tiles, nullptr};
    for (int per_sm = 1; per_sm <= occ; ++per_sm) {
      const int G = SMs * per_sm;
      partition_tiles(tiles, m, G, cta_tile, part);
      CK(cudaMalloc(&d_ct, (G + 1) * 4));
      CK(cudaMemcpy(d_ct, cta_tile.data(), (G + 1) * 4, cudaMemcpyHostToDevice));
      CK(cudaMemset(d_y, 0, (size_t)m * 8));
      char nm[96]; snprintf(nm, sizeof nm, "tile-stream TILE=%d S=%d grid=%dx%d (occ %d)", TS_TILE, TS_STAGES, G, NT, occ);
      time_it(nm, [&] { k_spmv_tile<<<G, NT, TS_SMEM>>>(M, d_ct, d_x, d_y); });
      check(nm);
      CK(cudaFree(d_ct));
    }
    CK(cudaFree(d_tiles));
  }

  for (int C : {16}) {
    if (!getenv("GB_DSMEM")) break;
    auto kern = k_spmv_dsmem<4>;
    const int nc = max_clusters(kern, C, smem);
    if (nc <= 0) continue;
    const long long bc = (long long)C * W;
    const int B = (int)((n + bc - 1) / bc);
    if (B > 64) { printf("B  cluster %d: %d column blocks, skipped\n", C, B); continue; }
    // blocked layout
    std::vector<int> bptr((size_t)B * (m + 1)), bidx(nnz);
    std::vector<double> bval(nnz);
    {
      std::vector<size_t> cnt((size_t)B * m + 1, 0);
      for (int i = 0; i < m; ++i) for (int k = ptr[i]; k < ptr[i + 1]; ++k) cnt[(size_t)(idx[k] / bc) * m + i + 1]++;
      for (size_t t = 0; t < (size_t)B * m; ++t) cnt[t + 1] += cnt[t];
      for (int b = 0; b < B; ++b) { for (int i = 0; i < m; ++i) bptr[(size_t)b * (m + 1) + i] = (int)cnt[(size_t)b * m + i]; bptr[(size_t)b * (m + 1) + m] = (int)cnt[(size_t)b * m + m]; }
      std::vector<size_t> pos(cnt.begin(), cnt.end() - 1);
      for (int i = 0; i < m; ++i) for (int k = ptr[i]; k < ptr[i + 1]; ++k) {
        const int b = (int)(idx[k] / bc); const long long lc = idx[k] - b * bc;
        const size_t p = pos[(size_t)b * m + i]++;
        bidx[p] = (int)(((lc / W) << 16) | (lc % W)); bval[p] = val[k];
      }
    }
    int *d_bptr, *d_bidx; double* d_bval;
    CK(cudaMalloc(&d_bptr, bptr.size() * 4)); CK(cudaMalloc(&d_bidx, (nnz + 8) * 4)); CK(cudaMalloc(&d_bval, (nnz + 8) * 8));
    CK(cudaMemcpy(d_bptr, bptr.data(), bptr.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bidx, bidx.data(), nnz * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_bval, bval.data(), nnz * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(d_y, 0, (size_t)m * 8));
    int rows = m, cols = n, Bv = B, Wv = W;
    const int* a1 = d_bptr; const int* a2 = d_bidx; const double* a3 = d_bval; const double* a4 = d_x; double* a5 = d_y;
    void* args[] = {&rows, &cols, &Bv, &Wv, &a1, &a2, &a3, &a4, &a5};
    char nm[96]; snprintf(nm, sizeof nm, "col-blocked, cluster smem C=%d B=%d ncl=%d", C, B, nc);
    time_it(nm, [&] { launch_cluster(kern, nc * C, C, smem, args); });
    check(nm);
    CK(cudaFree(d_bptr)); CK(cudaFree(d_bidx)); CK(cudaFree(d_bval));
  }
  printf("done\n");
  return 0;
}
