// agather_lab.cu -- round-2 microbenchmark: SpMV gathers issued as cp.async (SASS LDGSTS) into shared memory.
//
// Question.  The persistent kernel's SpMV on uniformly random columns is bound by the L1TEX gather rate (about one
// 128-byte line per clock per SM: profiles/r1b_gather_bench_rates.log), yet it only keeps that pipe ~50 % busy
// (profiles/r2a_*): a register-destination gather pins two registers until it returns, so a lane holds 5 gathers in
// flight, stalls for the whole round trip, and starts over.  cp.async has no register destination and no scoreboard:
// a warp can put the gathers of a whole 32-row tile (10 x 32) in flight at once and read them back from shared memory.
//
// Variants (same tile blobs as the production kernel, built by csrc/csl_host.h; matrix stream by cp.async.bulk ring):
//   reg      production scheme: register gathers, 5 per lane per round (part_dot of admm_kernel_csl.cuh)
//   async8   cp.async.ca 8 B per gather  (L1-allocating; the phases of the persistent kernel are separated by
//            acquire fences, which is what makes weak .ca loads of vectors written by other SMs legal)
//   async16  cp.async.cg 16 B per gather (the aligned pair that contains the element; L2 only)
// plus the pure gather rate of LDGSTS with no matrix at all.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/agather_lab tools/agather_lab.cu
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../osqp-python_b200/csrc/admm_kernel_csl.cuh"
#include "../osqp-python_b200/csrc/csl_host.h"

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

__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- pure LDGSTS gather rate: every thread keeps 2 groups of 8 random 8-byte gathers in flight
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_rate(const double* __restrict__ x, unsigned n, int per_thread, double* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  constexpr int ES = MODE ? 16 : 8;
  unsigned char* mine = sm + (size_t)threadIdx.x * 16 * ES;   // 2 groups x 8 slots
  unsigned s = (blockIdx.x * 512u + threadIdx.x) * 2654435761u + 12345u;
  double acc = 0.0;
  auto issue = [&](int g) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s = s * 1664525u + 1013904223u;
      const unsigned j = (unsigned)(((unsigned long long)s * n) >> 32);
      if (MODE) cp_async16(mine + (g * 8 + u) * ES, x + (j & ~1u));
      else cp_async8(mine + (g * 8 + u) * ES, x + j);
    }
    cp_commit();
  };
  issue(0);
  int g = 1;
  for (int it = 8; it < per_thread; it += 8) {
    issue(g);
    cp_wait<1>();
    g ^= 1;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += *reinterpret_cast<const double*>(mine + (g * 8 + u) * ES);
  }
  cp_wait<0>();
  g ^= 1;
#pragma unroll
  for (int u = 0; u < 8; ++u) acc += *reinterpret_cast<const double*>(mine + (g * 8 + u) * ES);
  if (acc == 123.456) out[0] = acc;
}

// ---- producer warp of the tile ring (same as k_spmv_csl)
__device__ __forceinline__ void produce(const CslStream& M, int t0, int T, int S, int capb, unsigned char* ring, uint64_t* bar_full,
                                        uint64_t* bar_empty) {
  const int lane = threadIdx.x & 31;
  const uint64_t pol = policy_evict_first();
  int q = 0;
  for (int ib = 0; ib < T; ib += 31) {
    const unsigned mine = __ldg(M.toff + t0 + min(ib + lane, T));
    const int nt = min(31, T - ib);
    for (int i = 0; i < nt; ++i) {
      const unsigned o0 = __shfl_sync(0xffffffffu, mine, i), o1 = __shfl_sync(0xffffffffu, mine, i + 1);
      if (lane == 0) {
        const int stg = q >= S ? q - S : q;
        mbar_wait(&bar_empty[stg], q >= S ? 1u : 0u);
        const uint32_t bytes = (o1 - o0) * 16u;
        mbar_expect_tx(&bar_full[stg], bytes);
        bulk_g2s_hint(ring + (size_t)stg * capb, M.blob + (size_t)o0 * 16u, bytes, &bar_full[stg], pol);
      }
      q = q + 1 == 2 * S ? 0 : q + 1;
      __syncwarp();
    }
  }
}

// MODE 0: register gathers (production part_dot).  1: cp.async.ca 8 B.  2: cp.async.cg 16 B.
template <int NWC, int CH, int MODE, int MINB>
__global__ void __launch_bounds__(NWC * 32 + 32, MINB)
k_spmv(const CslStream M, const int* __restrict__ part, int S, int capb, const double* __restrict__ x, double* __restrict__ y,
       const double* __restrict__ x2) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ alignas(8) uint64_t bar_full[64];
  __shared__ alignas(8) uint64_t bar_empty[64];
  constexpr int ES = MODE == 2 ? 16 : 8;
  unsigned char* ring = smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
    mbar_fence_init();
    for (int s = 0; s < S; ++s) csl_mbar_arrive(&bar_empty[s]);
  }
  __syncthreads();
  const int t0 = part[blockIdx.x], T = part[blockIdx.x + 1] - t0;
  if (warp == NWC) { produce(M, t0, T, S, capb, ring, bar_full, bar_empty); return; }
  unsigned char* gb = smem + (size_t)S * capb + (size_t)warp * CH * 32 * ES + (size_t)lane * ES;   // this lane's slot of column 0
  const unsigned lt = (1u << lane) - 1u;
  int q = warp;
  for (int i = warp; i < T; i += NWC) {
    const int stg = q >= S ? q - S : q;
    mbar_wait(&bar_empty[stg], q >= S ? 1u : 0u);      // see admm_kernel_csl.cuh: pins full[stg] to the phase waited for
    mbar_wait(&bar_full[stg], q >= S ? 1u : 0u);
    const TileView tv = tile_view<true>(ring + (size_t)stg * capb, lane);
    double acc = 0.0;
    if (MODE == 0) {
      double dummy;
      part_dot<1, true, CH>(tv.i1, tv.v1, tv.len1, tv.max1, x, nullptr, acc, dummy);
    } else if (MODE == 3) {   // n-side tile: K u = P u + A' t (phase C of the persistent kernel)
      double accA, dummy;
      part_dot<1, true, CH>(tv.i1, tv.v1, tv.len1, tv.max1, x, nullptr, acc, dummy);
      part_dot<1, true, CH>(tv.i2, tv.v2, tv.len2, tv.max2, x2, nullptr, accA, dummy);
      acc += accA;
    } else {
      const bool uniform = __all_sync(0xffffffffu, tv.len1 == tv.max1);
      int pos_i = 0, pos_f = 0;
      for (int j0 = 0; j0 < tv.max1; j0 += CH) {
        // issue: every gather of this chunk of columns, no register destination
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const bool act = tv.len1 > j0 + j;
          int k;
          if (uniform) k = (j0 + j) * 32 + lane;
          else { const unsigned mask = __ballot_sync(0xffffffffu, act); k = pos_i + __popc(mask & lt); pos_i += __popc(mask); }
          if (act) {
            const int c = tv.i1[k];
            if (MODE == 2) cp_async16(gb + j * 32 * ES, x + (c & ~1));
            else cp_async8(gb + j * 32 * ES, x + c);
          }
        }
        cp_commit();
        cp_wait<0>();
        // consume: values from the tile, gathered elements from this lane's own slots, fixed order
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const bool act = tv.len1 > j0 + j;
          int k;
          if (uniform) k = (j0 + j) * 32 + lane;
          else { const unsigned mask = __ballot_sync(0xffffffffu, act); k = pos_f + __popc(mask & lt); pos_f += __popc(mask); }
          if (act) {
            const double g = MODE == 2 ? *reinterpret_cast<const double*>(gb + j * 32 * ES + ((tv.i1[k] & 1) << 3))
                                       : *reinterpret_cast<const double*>(gb + j * 32 * ES);
            acc = fma(tv.v1[k], g, acc);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) csl_mbar_arrive(&bar_empty[stg]);
    const int row = (t0 + i) * 32 + lane;
    if (row < M.rows) y[row] = acc;
    q += NWC;
    while (q >= 2 * S) q -= 2 * S;
  }
}

__global__ void k_rewrite(double* x, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = x[i] * 1.0000001;
}
__global__ void k_touch(const double* x, size_t n, double* out) {
  double a = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) a += x[i];
  if (a == 1.2345) out[0] = a;
}

// ---- host
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

struct DevTiles {
  CslStream st; CslHost H; unsigned* d_toff; unsigned char* d_blob;
};
// one-part (M2 == nullptr) or two-part (part 1 = M, part 2 = M2; same row count) tiles
static DevTiles upload(const HostCsr& M, const HostCsr* M2 = nullptr) {
  DevTiles D;
  csl_build(M.rows, M.ptr.data(), M.idx.data(), M2 ? M2->ptr.data() : nullptr, M2 ? M2->idx.data() : nullptr, D.H);
  double* blobd = reinterpret_cast<double*>(D.H.blob.data());
  for (size_t e = 0; e < D.H.src1.size(); ++e) blobd[D.H.dst1[e]] = M.val[D.H.src1[e]];
  if (M2) for (size_t e = 0; e < D.H.src2.size(); ++e) blobd[D.H.dst2[e]] = M2->val[D.H.src2[e]];
  CK(cudaMalloc(&D.d_toff, D.H.toff.size() * 4));
  CK(cudaMalloc(&D.d_blob, D.H.blob.size() + 256));
  CK(cudaMemcpy(D.d_toff, D.H.toff.data(), D.H.toff.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(D.d_blob, D.H.blob.data(), D.H.blob.size(), cudaMemcpyHostToDevice));
  D.st = CslStream{M.rows, D.H.ntiles, D.d_toff, D.d_blob};
  return D;
}

static char* g_flush; static const size_t FL = (size_t)512 << 20;
static double *g_x, *g_x2, *g_y, *g_scratch;
static std::vector<double> h_x;
static cudaEvent_t e0, e1;

// OCC: CTAs per SM wanted (dynamic shared memory is padded so that exactly OCC fit)
template <int NWC, int CH, int MODE, int OCC>
static void run(const char* tag, const HostCsr& M, const HostCsr* M2, DevTiles& D, int S) {
  constexpr int ES = MODE == 2 ? 16 : 8;
  const int capb = (int)((D.H.max_tile_bytes + 127) & ~127u);
  size_t smem = (size_t)S * capb + ((MODE == 1 || MODE == 2) ? (size_t)NWC * CH * 32 * ES : 0);
  auto fn = k_spmv<NWC, CH, MODE, OCC>;
  const size_t per_cta_max = (size_t)(233472 / OCC - 2048) & ~(size_t)127;
  if (smem > per_cta_max) { printf("%-50s skipped (smem %zu > %zu)\n", tag, smem, per_cta_max); return; }
  const size_t need = smem;
  smem = std::max(smem, OCC < 4 ? (size_t)(233472 / (OCC + 1)) : smem);   // pad: no more than OCC CTAs per SM
  smem = std::min(smem, per_cta_max);
  CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, NWC * 32 + 32, smem));
  if (occ < 1) { printf("%-50s skipped (occupancy 0, smem %zu)\n", tag, smem); return; }
  const int G = occ * 148;
  std::vector<int> part;
  csl_partition(D.H, G, part);
  int* d_part; CK(cudaMalloc(&d_part, part.size() * 4));
  CK(cudaMemcpy(d_part, part.data(), part.size() * 4, cudaMemcpyHostToDevice));
  const int reps = 10;
  double best = 1e9, sum = 0;
  for (int r = 0; r < reps + 2; ++r) {
    k_touch<<<148 * 4, 512>>>(reinterpret_cast<const double*>(g_flush), FL / 8, g_scratch);   // flush by reading
    k_rewrite<<<296, 288>>>(g_x, M.cols);                                                    // x dirty in L2, as between phases
    if (M2) k_rewrite<<<296, 288>>>(g_x2, M2->cols);
    CK(cudaEventRecord(e0));
    fn<<<G, NWC * 32 + 32, smem>>>(D.st, d_part, S, capb, g_x, g_y, g_x2);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r >= 2) { best = std::min(best, (double)ms); sum += ms; }
  }
  std::vector<double> y(M.rows), xh(M.cols), xh2(M2 ? M2->cols : 0);
  CK(cudaMemcpy(y.data(), g_y, (size_t)M.rows * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(xh.data(), g_x, (size_t)M.cols * 8, cudaMemcpyDeviceToHost));
  if (M2) CK(cudaMemcpy(xh2.data(), g_x2, (size_t)M2->cols * 8, cudaMemcpyDeviceToHost));
  double err = 0, sc = 0;
  for (int r = 0; r < M.rows; r += 61) {
    double a = 0;
    for (int k = M.ptr[r]; k < M.ptr[r + 1]; ++k) a += M.val[k] * xh[M.idx[k]];
    if (M2) for (int k = M2->ptr[r]; k < M2->ptr[r + 1]; ++k) a += M2->val[k] * xh2[M2->idx[k]];
    err = std::max(err, std::fabs(a - y[r])); sc = std::max(sc, std::fabs(a));
  }
  const double nnz = (double)M.idx.size() + (M2 ? (double)M2->idx.size() : 0.0);
  const double bytes = 12.0 * nnz + (M2 ? 8.0 : 4.0) * (M.rows + 1) + 8.0 * M.rows + 8.0 * M.cols + (M2 ? 8.0 * M2->cols : 0.0);
  printf("%-50s warps/SM %2d  occ %d S %2d smem %6zu (pad %6zu)  best %6.1f us  mean %6.1f us  %5.2f TB/s  %.3f gathers/clk/SM  err %.1e\n",
         tag, occ * NWC, occ, S, need, smem, best * 1e3, sum / reps * 1e3, bytes / (best * 1e-3) / 1e12,
         nnz / (best * 1e-3) / 148 / 1.965e9, err / std::max(sc, 1e-300));
  fflush(stdout);
  CK(cudaFree(d_part));
}

static void run_all(const char* tag, const HostCsr& M, const HostCsr* M2, bool with_async) {
  DevTiles D = upload(M, M2);
  printf("== %s: %d rows, nnz %zu%s, %d tiles, max tile %u B\n", tag, M.rows, M.idx.size() + (M2 ? M2->idx.size() : 0), M2 ? " (two parts)" : "",
         D.H.ntiles, D.H.max_tile_bytes);
  char nm[96];
  const int capb = (int)((D.H.max_tile_bytes + 127) & ~127u);
#define RUN(NWC, CH, MODE, OCC, SLACK) do { \
    const int cap_s = (int)(((size_t)(233472 / OCC - 2048) & ~(size_t)127) / capb); \
    const int S_ = std::min(std::min(NWC + SLACK, cap_s), 60); \
    snprintf(nm, sizeof nm, "%s %s NWC=%d U=%d occ=%d", tag, (MODE == 0 || MODE == 3) ? "reg" : (MODE == 1 ? "async8" : "async16"), NWC, CH, OCC); \
    if (S_ > NWC) { if (M2) run<NWC, CH, 3, OCC>(nm, M, M2, D, S_); else run<NWC, CH, MODE, OCC>(nm, M, nullptr, D, S_); } \
    else printf("%-50s skipped (no ring slack: S %d)\n", nm, S_); } while (0)
  // occupancy / warps at U = 5
  RUN(11, 5, 0, 2, 5);
  RUN(11, 5, 0, 2, 14);
  RUN(11, 5, 0, 3, 5);
  RUN(8, 5, 0, 3, 5);
  RUN(8, 5, 0, 4, 4);
  RUN(15, 5, 0, 2, 7);
  RUN(17, 5, 0, 2, 8);
  RUN(23, 5, 0, 1, 12);
  RUN(31, 5, 0, 1, 16);
  // gathers in flight per lane
  RUN(11, 3, 0, 2, 5);
  RUN(11, 7, 0, 2, 5);
  RUN(11, 8, 0, 2, 5);
  RUN(11, 10, 0, 2, 5);
  RUN(11, 10, 0, 3, 5);
  RUN(15, 10, 0, 2, 7);
  RUN(23, 10, 0, 1, 12);
  RUN(31, 10, 0, 1, 16);
  RUN(15, 8, 0, 2, 7);
  RUN(23, 8, 0, 1, 12);
  if (with_async) {
    RUN(16, 10, 1, 1, 8);
    RUN(11, 10, 1, 2, 3);
    RUN(16, 16, 2, 1, 6);
  }
#undef RUN
  CK(cudaFree(D.d_toff)); CK(cudaFree(D.d_blob));
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 1000000;
  CK(cudaMalloc(&g_flush, FL)); CK(cudaMemset(g_flush, 1, FL));
  CK(cudaMalloc(&g_x, (size_t)n * 8 + 64)); CK(cudaMalloc(&g_x2, (size_t)n * 8 + 64)); CK(cudaMalloc(&g_y, (size_t)n * 8 + 64));
  CK(cudaMalloc(&g_scratch, 64));
  h_x.resize(n);
  std::mt19937_64 rng(5); std::normal_distribution<double> nd;
  for (auto& v : h_x) v = nd(rng);
  CK(cudaMemcpy(g_x, h_x.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  for (auto& v : h_x) v = nd(rng);
  CK(cudaMemcpy(g_x2, h_x.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  // pure LDGSTS rate
  for (int mode = 0; mode < 2; ++mode) {
    const int per_thread = 4096;
    const size_t smem = 512 * 16 * (mode ? 16 : 8);
    auto fn = mode ? k_rate<1> : k_rate<0>;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < 2; ++rep) {
      CK(cudaEventRecord(e0));
      fn<<<148, 512, smem>>>(g_x, (unsigned)n, per_thread, g_scratch);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double g = 148.0 * 512 * per_thread;
      if (rep) printf("LDGSTS rate %s: %.3f ms  %.1f Ggather/s  %.3f gathers/clk/SM -> 1e7 gathers: %.1f us\n", mode ? "cg 16B" : "ca  8B", ms,
                      g / (ms * 1e-3) / 1e9, g / (ms * 1e-3) / 148 / 1.965e9, 1e7 / (g / (ms * 1e-3)) * 1e6);
    }
  }
  HostCsr A = make_matrix(n, 10, false, 1);
  run_all("A random", A, nullptr, true);
  HostCsr At = transpose(A);
  {   // P: symmetric pattern, diagonal + one random symmetric pair per row (config 4's P_full has ~3 per row)
    HostCsr Ph = make_matrix(n, 2, false, 9);
    HostCsr Pt = transpose(Ph);
    HostCsr P; P.rows = n; P.cols = n; P.ptr.assign(n + 1, 0);
    for (int r = 0; r < n; ++r) {
      std::vector<int> c(Ph.idx.begin() + Ph.ptr[r], Ph.idx.begin() + Ph.ptr[r + 1]);
      for (int k = Pt.ptr[r]; k < Pt.ptr[r + 1]; ++k) if (Pt.idx[k] != r) c.push_back(Pt.idx[k]);
      std::sort(c.begin(), c.end()); c.erase(std::unique(c.begin(), c.end()), c.end());
      for (int j : c) { P.idx.push_back(j); P.val.push_back(0.01 * (1 + (j % 7))); }
      P.ptr[r + 1] = (int)P.idx.size();
    }
    run_all("K = P | A' random", P, &At, false);
  }
  {
    HostCsr B = make_matrix(n, 10, true, 2);
    run_all("A banded", B, nullptr, false);
  }
  return 0;
}
