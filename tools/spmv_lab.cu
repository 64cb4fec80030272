// spmv_lab.cu -- round-2 microbenchmark behind the "compacted slice" SpMV of the persistent ADMM kernel.
//
// Layout under test (CSL, compacted slice layout): rows are cut into slices of 32 (one warp); inside a slice the
// entries are stored column-major and compacted -- first entry 0 of every row that has one (in row order), then
// entry 1 of every row that has one, ... -- so lane t of the warp owns row t, the warp's loads of one "column"
// are contiguous (coalesced / conflict-free), no padding is stored and no row pointer is chased.
// Variants:
//   direct : thread per row, index/value read with coalesced global loads
//   ring   : persistent CTAs; one producer thread streams slice tiles into a shared-memory ring with
//            cp.async.bulk + mbarrier complete_tx; every consumer warp owns a slice at a time and hands the
//            stage back through an "empty" mbarrier (no CTA-wide barrier inside a phase)
// Matrices: n = m = 1e6, 10 nnz/row uniform random columns (+diag) = BASELINE config 4's A; its transpose (row
// lengths vary); the banded counterpart (columns within +-64 of the diagonal).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/spmv_lab tools/spmv_lab.cu
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../osqp-python_b200/csrc/device_common.cuh"

using namespace b200qp;

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, __LINE__, \
              cudaGetErrorString(e_));                                                          \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

struct Csl {   // device view
  int rows, nslices;
  const int* soff;            // nslices+1 entry offsets (multiples of 4)
  const unsigned char* len;   // rows (padded to 32*nslices)
  const int* idx;
  const double* val;
};

// volatile: ptxas otherwise sinks every load next to its use (load -> gather -> fma serialised per entry)
__device__ __forceinline__ double ld_gather(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
// predicated gather: inactive lanes issue no memory request at all (a shared dummy address would make one L2
// slice a hot spot: measured 53 -> 72 us when 6 of 16 gather instructions per row were dummies)
__device__ __forceinline__ double ld_gather_if(const double* p, bool act) {
  double v;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\tmov.f64 %0, 0d0000000000000000;\n\t@p ld.global.cg.f64 %0, [%1];\n\t}"
               : "=d"(v) : "l"(p), "r"((int)act));
  return v;
}
__device__ __forceinline__ int ld_stream_v(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double ld_stream_v(const double* p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// one slice, entries read through `LD` accessors (global or shared)
template <int U, class FI, class FV>
__device__ __forceinline__ double slice_dot(int len, int maxlen, const double* __restrict__ vec, const double* __restrict__ zero,
                                            FI&& ld_i, FV&& ld_v) {
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  int pos = 0;
  double acc = 0.0;
  for (int j0 = 0; j0 < maxlen; j0 += U) {
    int k[U], c[U];
    double g[U], v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool act = len > j0 + u;
      const unsigned mask = __ballot_sync(0xffffffffu, act);
      k[u] = act ? pos + __popc(mask & lt) : 0;     // slot 0 of the slice: always a valid entry
      pos += __popc(mask);
    }
    // unconditional loads, batched: all indices, then all gathers, then all values
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = ld_i(k[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) g[u] = ld_gather_if(vec + c[u], len > j0 + u);
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = ld_v(k[u]);
    // descending: the first FMA needs the LAST gather, so ptxas cannot sink the loads next to their uses
#pragma unroll
    for (int u = U - 1; u >= 0; --u) acc = fma(v[u], g[u], acc);
  }
  return acc;
}

// shared-memory flavour: the whole chunk (indices AND values) is copied to registers first; after the last chunk
// `release()` hands the stage back to the producer BEFORE the gathers are waited for (stage hold time = a few
// LDS instead of an L2 round trip)
template <int U, class REL>
__device__ __forceinline__ double slice_dot_smem(int len, int maxlen, const double* __restrict__ vec, const int* si, const double* sv,
                                                 REL&& release) {
  const unsigned lt = (1u << (threadIdx.x & 31)) - 1u;
  int pos = 0;
  double acc = 0.0;
  for (int j0 = 0; j0 < maxlen; j0 += U) {
    int k[U], c[U];
    double g[U], v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool act = len > j0 + u;
      const unsigned mask = __ballot_sync(0xffffffffu, act);
      k[u] = act ? pos + __popc(mask & lt) : 0;
      pos += __popc(mask);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = si[k[u]];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = sv[k[u]];
    // the stage may only be handed back once the LDS results have ARRIVED (they queue behind the gathers in the
    // LSU): make the release depend on the last-issued index and value
    if (j0 + U >= maxlen) release((c[U - 1] >> 31) | (__popc(__double2hiint(v[U - 1])) >> 6));
#pragma unroll
    for (int u = 0; u < U; ++u) g[u] = ld_gather_if(vec + c[u], len > j0 + u);
#pragma unroll
    for (int u = U - 1; u >= 0; --u) acc = fma(v[u], g[u], acc);
  }
  if (maxlen == 0) release(0);
  return acc;
}

// ---- ring variant 2: packed tiles of TS slices ([values | indices] contiguous -> ONE bulk copy per tile), early release
struct CslP {   // device view of the packed layout
  int rows, nslices;
  const int* soff;            // nslices+1 entry offsets (multiples of 4), as in Csl
  const unsigned char* len;
  const unsigned char* blob;  // tile t (slices t*TS..): bytes [12*soff[t*TS], 12*soff[(t+1)*TS]) = values then indices
};
template <int NWC, int U, int TS, int CAPB, int S>
__global__ void __launch_bounds__(NWC * 32 + 32, 1)
k_csl_ring2(const CslP M, const int* __restrict__ cta_tile, const double* __restrict__ x, double* __restrict__ y) {
  extern __shared__ __align__(128) unsigned char smem[];   // [S][CAPB]
  __shared__ alignas(8) uint64_t full[S];
  __shared__ alignas(8) uint64_t empty[S];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], TS); }
    mbar_fence_init();
  }
  __syncthreads();
  const int t0 = cta_tile[blockIdx.x], t1 = cta_tile[blockIdx.x + 1];   // tiles of this CTA
  const int T = t1 - t0;
  if (warp == NWC) {
    const uint64_t pol = policy_evict_first();
    for (int tb = 0; tb < T; tb += 32) {
      const int mine = M.soff[min((t0 + tb + lane) * TS, M.nslices)];
      const int last = M.soff[min((t0 + tb + 32) * TS, M.nslices)];
      const int nt = min(32, T - tb);
      for (int i = 0; i < nt; ++i) {
        const int a = __shfl_sync(0xffffffffu, mine, i);
        const int b = (i == 31) ? last : __shfl_sync(0xffffffffu, mine, (i + 1) & 31);
        if (lane == 0) {
          const int t = tb + i, st = t % S, cnt = b - a;
          if (t >= S) mbar_wait(&empty[st], ((t / S) - 1) & 1);
          mbar_expect_tx(&full[st], (uint32_t)cnt * 12u);
          if (cnt) bulk_g2s_hint(smem + (size_t)st * CAPB, M.blob + (size_t)a * 12, (uint32_t)cnt * 12u, &full[st], pol);
        }
        __syncwarp();
      }
    }
    return;
  }
  // consumer warp w: slices (local) w, w + NWC, ...; slice ls lives in tile ls / TS at sub-slot ls % TS
  const int nsl = T * TS;
  for (int ls = warp; ls < nsl; ls += NWC) {
    const int t = ls / TS, st = t % S;
    const int slice = t0 * TS + ls;
    const bool exists = slice < M.nslices;
    const int row = slice * 32 + lane;
    const int len = exists ? M.len[row] : 0;
    int maxlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    const int tile_a = M.soff[min((t0 + t) * TS, M.nslices)];
    const int tile_b = M.soff[min((t0 + t + 1) * TS, M.nslices)];
    const int sl_a = exists ? M.soff[slice] : tile_b;
    mbar_wait(&full[st], (t / S) & 1);
    const unsigned char* stage = smem + (size_t)st * CAPB;
    const double* sv = reinterpret_cast<const double*>(stage) + (sl_a - tile_a);
    const int* si = reinterpret_cast<const int*>(stage + (size_t)(tile_b - tile_a) * 8) + (sl_a - tile_a);
    const double acc = slice_dot_smem<U>(len, maxlen, x, si, sv, [&](int dep) { __syncwarp(); if (lane == 0) mbar_arrive(&empty[st] + dep); });
    if (exists && row < M.rows) y[row] = acc;
  }
}

template <int U>
__global__ void __launch_bounds__(256) k_csl_direct(const Csl M, const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ zero) {
  const int slice = (int)((blockIdx.x * 256u + threadIdx.x) >> 5);
  if (slice >= M.nslices) return;
  const int lane = threadIdx.x & 31, row = slice * 32 + lane;
  const int len = M.len[row];
  int maxlen = len;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
  const int* gi = M.idx + M.soff[slice];
  const double* gv = M.val + M.soff[slice];
  const double acc = slice_dot<U>(len, maxlen, x, zero, [&](int k) { return ld_stream_v(gi + k); }, [&](int k) { return ld_stream_v(gv + k); });
  if (row < M.rows) y[row] = acc;
}

// ---- ring variant --------------------------------------------------------------------------------------
// NWC consumer warps + 1 producer warp; stage = one slice (<= CAP entries)
template <int NWC, int U, int CAP, int S>
__global__ void __launch_bounds__(NWC * 32 + 32, 1)
k_csl_ring(const Csl M, const int* __restrict__ cta_slice, const double* __restrict__ x, double* __restrict__ y, int hint, const double* __restrict__ zero) {
  extern __shared__ __align__(128) unsigned char smem[];
  double* sval = reinterpret_cast<double*>(smem);                      // [S][CAP]
  int* sidx = reinterpret_cast<int*>(smem + (size_t)S * CAP * 8);      // [S][CAP]
  __shared__ alignas(8) uint64_t full[S];
  __shared__ alignas(8) uint64_t empty[S];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_fence_init();
  }
  __syncthreads();
  const int s0 = cta_slice[blockIdx.x], s1 = cta_slice[blockIdx.x + 1];
  const int T = s1 - s0;
  if (warp == NWC) {
    // producer warp: the slice offsets of 32 tiles are fetched by the 32 lanes at once (one dependent global
    // load per tile on a single thread made the producer the bottleneck: 60 us even on the banded matrix)
    const uint64_t pol = policy_evict_first();
    for (int tb = 0; tb < T; tb += 32) {
      const int mine = M.soff[min(s0 + tb + lane, s1)];
      const int last = M.soff[min(s0 + tb + 32, s1)];
      const int nt = min(32, T - tb);
      for (int i = 0; i < nt; ++i) {
        const int a = __shfl_sync(0xffffffffu, mine, i);
        const int b = (i == 31) ? last : __shfl_sync(0xffffffffu, mine, (i + 1) & 31);
        if (lane == 0) {
          const int t = tb + i, st = t % S, cnt = b - a;
          if (t >= S) mbar_wait(&empty[st], ((t / S) - 1) & 1);
          mbar_expect_tx(&full[st], (uint32_t)cnt * 12u);
          if (cnt) {
            if (hint) {
              bulk_g2s_hint(sval + (size_t)st * CAP, M.val + a, (uint32_t)cnt * 8u, &full[st], pol);
              bulk_g2s_hint(sidx + (size_t)st * CAP, M.idx + a, (uint32_t)cnt * 4u, &full[st], pol);
            } else {
              bulk_g2s(sval + (size_t)st * CAP, M.val + a, (uint32_t)cnt * 8u, &full[st]);
              bulk_g2s(sidx + (size_t)st * CAP, M.idx + a, (uint32_t)cnt * 4u, &full[st]);
            }
          }
        }
        __syncwarp();
      }
    }
    return;
  }
  for (int t = warp; t < T; t += NWC) {
    const int st = t % S;
    const int row = (s0 + t) * 32 + lane;
    const int len = M.len[row];
    int maxlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, o));
    mbar_wait(&full[st], (t / S) & 1);
    const int* si = sidx + (size_t)st * CAP;
    const double* sv = sval + (size_t)st * CAP;
    const double acc = slice_dot<U>(len, maxlen, x, zero, [&](int k) { return si[k]; }, [&](int k) { return sv[k]; });
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    if (row < M.rows) y[row] = acc;
  }
}

// rewrite the vector in place (what the persistent kernel does between phases): lines end up DIRTY in L2
__global__ void k_rewrite(double* x, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = x[i] * 1.0;
}
__global__ void k_touch(const double* x, int n, double* out) {
  double a = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) a += __ldcg(x + i);
  if (a == -1.0) out[0] = a;
}

// plain CSR LG=4 reference (round-1 "group" flavour, simplified)
template <int LG>
__global__ void __launch_bounds__(256) k_csr_group(int rows, const int* __restrict__ ptr, const int* __restrict__ idx,
                                                   const double* __restrict__ val, const double* __restrict__ x,
                                                   double* __restrict__ y) {
  const int lane = threadIdx.x & (LG - 1);
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) / LG);
  double acc = 0.0;
  if (row < rows) {
    const int a = ptr[row], b = ptr[row + 1];
    for (int k = a + lane; k < b; k += LG) acc = fma(ld_stream(val + k), __ldcg(x + ld_stream(idx + k)), acc);
  }
#pragma unroll
  for (int o = LG / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (row < rows && lane == 0) y[row] = acc;
}

// ---- host ----------------------------------------------------------------------------------------------
struct HostCsr { int rows, cols; std::vector<int> ptr, idx; std::vector<double> val; };

static HostCsr make_matrix(int n, int per_row, bool banded, unsigned seed) {
  HostCsr M; M.rows = n; M.cols = n; M.ptr.resize(n + 1); M.idx.resize((size_t)n * per_row); M.val.resize((size_t)n * per_row);
  std::mt19937_64 rng(seed);
  std::normal_distribution<double> nd;
  for (int r = 0; r < n; ++r) {
    M.ptr[r] = r * per_row;
    int* c = &M.idx[(size_t)r * per_row];
    c[0] = r;
    for (int k = 1; k < per_row;) {
      int j = banded ? r + (int)(rng() % 129) - 64 : (int)(rng() % (unsigned)n);
      if (j < 0 || j >= n) continue;
      bool dup = false;
      for (int t = 0; t < k; ++t) dup |= c[t] == j;
      if (dup) continue;
      c[k++] = j;
    }
    std::sort(c, c + per_row);
    for (int k = 0; k < per_row; ++k) M.val[(size_t)r * per_row + k] = nd(rng);
  }
  M.ptr[n] = n * per_row;
  return M;
}
static HostCsr transpose(const HostCsr& A) {
  HostCsr T; T.rows = A.cols; T.cols = A.rows; T.ptr.assign(T.rows + 1, 0);
  const size_t nnz = A.idx.size();
  T.idx.resize(nnz); T.val.resize(nnz);
  for (size_t k = 0; k < nnz; ++k) T.ptr[A.idx[k] + 1]++;
  for (int i = 0; i < T.rows; ++i) T.ptr[i + 1] += T.ptr[i];
  std::vector<int> fill(T.ptr.begin(), T.ptr.end() - 1);
  for (int r = 0; r < A.rows; ++r)
    for (int k = A.ptr[r]; k < A.ptr[r + 1]; ++k) { const int p = fill[A.idx[k]]++; T.idx[p] = r; T.val[p] = A.val[k]; }
  return T;
}

struct DevCsl {
  Csl v; int maxslice = 0; long long nnz = 0;
  std::vector<int> soff, h_idx;
  std::vector<double> h_val;
  int *d_soff, *d_idx; unsigned char* d_len; double* d_val;
};
static DevCsl to_csl(const HostCsr& M) {
  DevCsl D;
  const int ns = (M.rows + 31) / 32;
  D.soff.assign(ns + 1, 0);
  std::vector<unsigned char> len((size_t)ns * 32, 0);
  std::vector<int> idx; std::vector<double> val;
  idx.reserve(M.idx.size() + 4 * ns); val.reserve(M.idx.size() + 4 * ns);
  for (int s = 0; s < ns; ++s) {
    D.soff[s] = (int)idx.size();
    int maxlen = 0;
    for (int t = 0; t < 32; ++t) {
      const int r = s * 32 + t;
      const int l = r < M.rows ? M.ptr[r + 1] - M.ptr[r] : 0;
      if (l > 255) { fprintf(stderr, "row too long\n"); exit(1); }
      len[(size_t)s * 32 + t] = (unsigned char)l;
      maxlen = std::max(maxlen, l);
    }
    for (int j = 0; j < maxlen; ++j)
      for (int t = 0; t < 32; ++t) {
        const int r = s * 32 + t;
        if (r < M.rows && M.ptr[r + 1] - M.ptr[r] > j) { idx.push_back(M.idx[M.ptr[r] + j]); val.push_back(M.val[M.ptr[r] + j]); }
      }
    while (idx.size() & 3) { idx.push_back(0); val.push_back(0.0); }
    D.maxslice = std::max(D.maxslice, (int)idx.size() - D.soff[s]);
  }
  D.soff[ns] = (int)idx.size();
  D.nnz = (long long)M.idx.size();
  CK(cudaMalloc(&D.d_soff, (ns + 1) * 4)); CK(cudaMalloc(&D.d_len, (size_t)ns * 32));
  CK(cudaMalloc(&D.d_idx, idx.size() * 4 + 64)); CK(cudaMalloc(&D.d_val, val.size() * 8 + 64));
  CK(cudaMemcpy(D.d_soff, D.soff.data(), (ns + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(D.d_len, len.data(), (size_t)ns * 32, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(D.d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(D.d_val, val.data(), val.size() * 8, cudaMemcpyHostToDevice));
  D.v = Csl{M.rows, ns, D.d_soff, D.d_len, D.d_idx, D.d_val};
  D.h_idx = idx; D.h_val = val;
  return D;
}

static int g_warm_mode = 1;
static char* g_flush; static const size_t FL = (size_t)512 << 20;
static double *g_x, *g_y, *g_scratch, *g_zero;
static std::vector<double> h_x, h_ref;
static cudaEvent_t e0, e1;

template <class F>
static void bench(const char* name, const HostCsr& M, long long nnz, int warm_x, F&& launch) {
  const int reps = 12;
  double best = 1e9, sum = 0;
  for (int r = 0; r < reps + 2; ++r) {
    // flush by READING a buffer larger than L2: a memset flush leaves 126 MB of dirty lines whose write-back then
    // competes with the timed kernel's reads (banded SpMV: 35 us instead of ~22 us)
    k_touch<<<148 * 4, 512>>>(reinterpret_cast<const double*>(g_flush), (int)(FL / 8), g_scratch);
    if (warm_x == 1) k_touch<<<148, 512>>>(g_x, M.cols, g_scratch);
    if (warm_x == 2) k_rewrite<<<296, 288>>>(g_x, M.cols);
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 2) { best = std::min(best, (double)ms); sum += ms; }
  }
  std::vector<double> y(M.rows);
  CK(cudaMemcpy(y.data(), g_y, (size_t)M.rows * 8, cudaMemcpyDeviceToHost));
  double err = 0, sc = 0;
  for (int r = 0; r < M.rows; r += 97) {
    double a = 0;
    for (int k = M.ptr[r]; k < M.ptr[r + 1]; ++k) a += M.val[k] * h_x[M.idx[k]];
    err = std::max(err, std::fabs(a - y[r])); sc = std::max(sc, std::fabs(a));
  }
  const double bytes = 12.0 * nnz + 4.0 * (M.rows + 1) + 8.0 * M.rows + 8.0 * M.cols;
  printf("%-46s best %7.1f us  mean %7.1f us  -> %5.2f TB/s algorithmic (best)  err %.1e/%.1e%s\n", name, best * 1e3,
         sum / reps * 1e3, bytes / (best * 1e-3) * 1e-12, err, sc, warm_x == 1 ? "  [x warm in L2]" : warm_x == 2 ? "  [x rewritten: dirty in L2]" : "");
  fflush(stdout);
  CK(cudaMemset(g_y, 0, (size_t)M.rows * 8));
}

template <int NWC, int U, int CAP, int S>
static void run_ring(const char* tag, const HostCsr& M, const DevCsl& D, int ctas_per_sm_unused, int hint) {
  (void)ctas_per_sm_unused;
  if (D.maxslice > CAP) { printf("%s: skipped (max slice %d > CAP %d)\n", tag, D.maxslice, CAP); return; }
  const size_t smem = (size_t)S * CAP * 12;
  auto kern = k_csl_ring<NWC, U, CAP, S>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NWC * 32 + 32, smem));
  if (occ < 1) { printf("%s: does not fit\n", tag); return; }
  const int G = 148 * occ;
  // slices per CTA balanced by entries
  std::vector<int> cs(G + 1, D.v.nslices);
  cs[0] = 0;
  { int s = 0; const long long tot = D.soff[D.v.nslices];
    for (int g = 1; g < G; ++g) { const long long target = tot * g / G; while (s < D.v.nslices && D.soff[s] < target) ++s; cs[g] = s; } }
  int* d_cs; CK(cudaMalloc(&d_cs, (G + 1) * 4)); CK(cudaMemcpy(d_cs, cs.data(), (G + 1) * 4, cudaMemcpyHostToDevice));
  char name[160];
  snprintf(name, sizeof name, "%s ring NWC=%d U=%d CAP=%d S=%d occ=%d hint=%d", tag, NWC, U, CAP, S, occ, hint);
  bench(name, M, D.nnz, g_warm_mode, [&] { kern<<<G, NWC * 32 + 32, smem>>>(D.v, d_cs, g_x, g_y, hint, g_zero); });
  CK(cudaFree(d_cs));
}

template <int NWC, int U, int TS, int CAPB, int S>
static void run_ring2(const char* tag, const HostCsr& M, const DevCsl& D, const std::vector<int>& h_idx, const std::vector<double>& h_val) {
  const int ns = D.v.nslices, NT_ = (ns + TS - 1) / TS;
  // packed blob
  std::vector<unsigned char> blob((size_t)D.soff[ns] * 12 + 64);
  int maxtile = 0;
  for (int t = 0; t < NT_; ++t) {
    const int a = D.soff[t * TS], b = D.soff[std::min((t + 1) * TS, ns)];
    maxtile = std::max(maxtile, b - a);
    memcpy(&blob[(size_t)a * 12], &h_val[a], (size_t)(b - a) * 8);
    memcpy(&blob[(size_t)a * 12 + (size_t)(b - a) * 8], &h_idx[a], (size_t)(b - a) * 4);
  }
  if (maxtile * 12 > CAPB) { printf("%s ring2 TS=%d: skipped (tile %d B > %d)\n", tag, TS, maxtile * 12, CAPB); return; }
  const size_t smem = (size_t)S * CAPB;
  auto kern = k_csl_ring2<NWC, U, TS, CAPB, S>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NWC * 32 + 32, smem));
  if (occ < 1) { printf("%s ring2: does not fit\n", tag); return; }
  const int G = 148 * occ;
  std::vector<int> ct(G + 1, NT_);
  ct[0] = 0;
  { int t = 0; const long long tot = D.soff[ns];
    for (int g = 1; g < G; ++g) { const long long target = tot * g / G; while (t < NT_ && D.soff[std::min(t * TS, ns)] < target) ++t; ct[g] = t; } }
  unsigned char* d_blob; int* d_ct;
  CK(cudaMalloc(&d_blob, blob.size())); CK(cudaMemcpy(d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_ct, (G + 1) * 4)); CK(cudaMemcpy(d_ct, ct.data(), (G + 1) * 4, cudaMemcpyHostToDevice));
  CslP P{M.rows, ns, D.d_soff, D.d_len, d_blob};
  char name[160];
  snprintf(name, sizeof name, "%s ring2 NWC=%d U=%d TS=%d CAPB=%d S=%d occ=%d", tag, NWC, U, TS, CAPB, S, occ);
  bench(name, M, D.nnz, g_warm_mode, [&] { kern<<<G, NWC * 32 + 32, smem>>>(P, d_ct, g_x, g_y); });
  CK(cudaFree(d_blob)); CK(cudaFree(d_ct));
}

template <int U>
static void run_direct(const char* tag, const HostCsr& M, const DevCsl& D) {
  char name[160];
  snprintf(name, sizeof name, "%s direct U=%d", tag, U);
  const int grid = (D.v.nslices * 32 + 255) / 256;
  bench(name, M, D.nnz, g_warm_mode, [&] { k_csl_direct<U><<<grid, 256>>>(D.v, g_x, g_y, g_zero); });
}

static void run_all(const char* tag, const HostCsr& M) {
  DevCsl D = to_csl(M);
  printf("== %s: rows %d nnz %lld max slice entries %d\n", tag, M.rows, D.nnz, D.maxslice);
  int *d_ptr, *d_idx; double* d_val;
  CK(cudaMalloc(&d_ptr, (M.rows + 1) * 4)); CK(cudaMalloc(&d_idx, M.idx.size() * 4)); CK(cudaMalloc(&d_val, M.val.size() * 8));
  CK(cudaMemcpy(d_ptr, M.ptr.data(), (M.rows + 1) * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_idx, M.idx.data(), M.idx.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_val, M.val.data(), M.val.size() * 8, cudaMemcpyHostToDevice));
  char name[160];
  snprintf(name, sizeof name, "%s csr LG=4 (round-1 scheme)", tag);
  bench(name, M, D.nnz, g_warm_mode, [&] { k_csr_group<4><<<(int)(((long long)M.rows * 4 + 255) / 256), 256>>>(M.rows, d_ptr, d_idx, d_val, g_x, g_y); });
  snprintf(name, sizeof name, "%s csr LG=4 (round-1 scheme) x cold", tag);
  bench(name, M, D.nnz, 0, [&] { k_csr_group<4><<<(int)(((long long)M.rows * 4 + 255) / 256), 256>>>(M.rows, d_ptr, d_idx, d_val, g_x, g_y); });
  for (g_warm_mode = 1; g_warm_mode <= 2; ++g_warm_mode) {
    run_direct<5>(tag, M, D);
    run_direct<10>(tag, M, D);
    run_ring<8, 5, 512, 16>(tag, M, D, 2, 1);
    run_ring<8, 10, 512, 16>(tag, M, D, 2, 1);
    run_ring<7, 10, 512, 16>(tag, M, D, 2, 1);
    run_ring<16, 5, 512, 24>(tag, M, D, 1, 1);
    run_ring<31, 5, 512, 36>(tag, M, D, 1, 1);
  }
  g_warm_mode = 1;
  cudaFree(d_ptr); cudaFree(d_idx); cudaFree(d_val);
  cudaFree(D.d_soff); cudaFree(D.d_len); cudaFree(D.d_idx); cudaFree(D.d_val);
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1000000;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s SMs %d clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  CK(cudaMalloc(&g_flush, FL)); CK(cudaMemset(g_flush, 0, FL)); CK(cudaMalloc(&g_x, (size_t)n * 8)); CK(cudaMalloc(&g_y, (size_t)n * 8)); CK(cudaMalloc(&g_scratch, 64)); CK(cudaMalloc(&g_zero, 64)); CK(cudaMemset(g_zero, 0, 64));
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  h_x.resize(n);
  { std::mt19937_64 rng(7); std::normal_distribution<double> nd; for (auto& v : h_x) v = nd(rng); }
  CK(cudaMemcpy(g_x, h_x.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  {
    HostCsr A = make_matrix(n, 10, false, 3);
    run_all("A random", A);
    HostCsr At = transpose(A);
    run_all("At random", At);
  }
  {
    HostCsr B = make_matrix(n, 10, true, 5);
    run_all("A banded", B);
  }
  printf("done\n");
  return 0;
}
