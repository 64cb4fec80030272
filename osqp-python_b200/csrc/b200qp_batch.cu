// b200qp_batch.cu -- host driver + C ABI of the batched solver (include/b200qp.h, "batched independent QPs").
//
// A batch owns one prototype solver (b200qp_setup on element 0: format conversion, Ruiz equilibration, the
// shared scaled matrices) and the per-element arrays: scaled q, l, u; the ADMM state x, z, y, rho; results.
// solve() is ONE launch of admm_batch_kernel (admm_batch.cuh): one CTA per QP, all vectors in shared memory.
// No CPU compute path: every numerical step runs in the kernels of this translation unit / b200qp.cu.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "admm_batch.cuh"
#include "internal_view.h"

using namespace b200qp;

namespace {

#define BCK(call)                                                                                   \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) {                                                                        \
      std::fprintf(stderr, "[b200qp] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e_), __FILE__, \
                   __LINE__, cudaGetErrorString(e_));                                               \
      throw (int)B200QP_ALGEBRA_LOAD_ERROR;                                                         \
    }                                                                                               \
  } while (0)

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// out[b][i] = scalar * in[b][i] * s[i]   (s may be null)
__global__ void kb_scale_rows(long long total, int len, double* __restrict__ out, const double* __restrict__ in,
                              const double* __restrict__ s, double scalar) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) out[t] = scalar * in[t] * (s ? s[t % len] : 1.0);
}
__global__ void kb_fill(long long total, double* __restrict__ out, double v) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) out[t] = v;
}
// zs[b] = A xs[b]   (8 lanes per row, one grid row-block per element)
__global__ void kb_spmv(CsrMat A, int batch, const double* __restrict__ xs, double* __restrict__ zs) {
  const int lane = threadIdx.x & 7;
  const long long g = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const long long rows = (long long)batch * A.rows;
  double acc = 0.0;
  int row = 0;
  long long b = 0;
  if (g < rows) {
    b = g / A.rows; row = (int)(g % A.rows);
    const double* x = xs + b * A.cols;
    for (int k = A.ptr[row] + lane; k < A.ptr[row + 1]; k += 8) acc = fma(A.val[k], x[A.idx[k]], acc);
  }
#pragma unroll
  for (int off = 4; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (g < rows && lane == 0) zs[b * A.rows + row] = acc;
}

// dst[b][i] = src[i]  (replicate the shared scaled values before the first per-element update)
__global__ void kb_tile_rows(long long total, int len, double* __restrict__ dst, const double* __restrict__ src) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < total) dst[t] = src[t % len];
}
// per-element A values, UNSCALED, CSC order [batch][nnzA] -> scaled CSR(A') (same order) and CSR(A) (mapA)
// (single-solver counterpart: k_update_A in aux_kernels.cuh; reference _osqp.py:1455-1470)
__global__ void kb_update_A(long long total, int nnz, const double* __restrict__ vals, const int* __restrict__ colptr,
                            const int* __restrict__ rowidx, int ncols, const int* __restrict__ mapA,
                            const double* __restrict__ D, const double* __restrict__ E, double* __restrict__ Atv,
                            double* __restrict__ Av) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long b = t / nnz;
  const int k = (int)(t % nnz);
  int lo = 0, hi = ncols;               // column j with colptr[j] <= k < colptr[j+1]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (colptr[mid] <= k) lo = mid; else hi = mid; }
  const double v = E[rowidx[k]] * vals[t] * D[lo];
  Atv[b * nnz + k] = v;
  Av[b * nnz + mapA[k]] = v;
}
// per-element P values, UNSCALED, upper triangle in CSC order [batch][nnzPt] -> both mirrored positions of the
// full CSR values [batch][nnzP] and the diagonal [batch][n]  (k_update_P; reference _osqp.py:1437-1452)
__global__ void kb_update_P(long long total, int nnzPt, int nnzP, int n, const double* __restrict__ vals,
                            const int* __restrict__ colptr, const int* __restrict__ rowidx,
                            const int* __restrict__ posU, const int* __restrict__ posL, const double* __restrict__ D,
                            double c, double* __restrict__ Pv, double* __restrict__ pdv) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const long long b = t / nnzPt;
  const int k = (int)(t % nnzPt);
  int lo = 0, hi = n;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (colptr[mid] <= k) lo = mid; else hi = mid; }
  const int i = rowidx[k];
  const double v = c * D[i] * vals[t] * D[lo];
  Pv[b * nnzP + posU[k]] = v;
  if (posL[k] >= 0) Pv[b * nnzP + posL[k]] = v; else pdv[b * n + i] = v;
}

// New bounds of a whole batch: clip to +-INFTY (interface.py:237-238), scale by E (_osqp.py:1353), check l <= u
// against the other bound -- the new one if given, else the stored (already scaled) one -- and raise *bad
// (reference: _osqp.py:1348, 1388, 1419).  The scaled values stay in the staging buffers until the check passed.
__global__ void kb_bounds(long long total, int m, double* __restrict__ l_new, double* __restrict__ u_new,
                          const double* __restrict__ l_cur, const double* __restrict__ u_cur,
                          const double* __restrict__ E, int* __restrict__ bad) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const double e = E[t % m];
  double lo = l_cur[t], hi = u_cur[t];
  if (l_new) { lo = e * fmin(fmax(l_new[t], -QP_INFTY), QP_INFTY); l_new[t] = lo; }
  if (u_new) { hi = e * fmin(fmax(u_new[t], -QP_INFTY), QP_INFTY); u_new[t] = hi; }
  if (!(lo <= hi + 1e-12 * fabs(hi))) *bad = 1;
}

typedef void (*batch_fn)(const BArgs);
int pick_blg(double avg) { return avg <= 12.0 ? 4 : 16; }
batch_fn pick_batch_kernel(int lgn, int lgm, bool dense, int nt) {
  if (dense && nt == 512) {
    if (lgn == 4 && lgm == 4) return admm_batch_kernel<4, 4, 512, true>;
    if (lgn == 4 && lgm == 16) return admm_batch_kernel<4, 16, 512, true>;
    if (lgn == 16 && lgm == 4) return admm_batch_kernel<16, 4, 512, true>;
    return admm_batch_kernel<16, 16, 512, true>;
  }
  if (dense) {
    if (lgn == 4 && lgm == 4) return admm_batch_kernel<4, 4, 1024, true>;
    if (lgn == 4 && lgm == 16) return admm_batch_kernel<4, 16, 1024, true>;
    if (lgn == 16 && lgm == 4) return admm_batch_kernel<16, 4, 1024, true>;
    return admm_batch_kernel<16, 16, 1024, true>;
  }
  if (lgn == 4 && lgm == 4) return admm_batch_kernel<4, 4, NTB_PCG, false>;
  if (lgn == 4 && lgm == 16) return admm_batch_kernel<4, 16, NTB_PCG, false>;
  if (lgn == 16 && lgm == 4) return admm_batch_kernel<16, 4, NTB_PCG, false>;
  return admm_batch_kernel<16, 16, NTB_PCG, false>;
}

}  // namespace

struct b200qp_batch {
  b200qp_solver* proto = nullptr;
  InternalView v{};
  int batch = 0, n = 0, m = 0, device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  double *q = nullptr, *l = nullptr, *u = nullptr;             // scaled, [batch][n|m]
  double *xs = nullptr, *zs = nullptr, *ys = nullptr, *rho_b = nullptr;
  double *sol_x = nullptr, *sol_y = nullptr, *cert_n = nullptr, *cert_m = nullptr;
  double* stage = nullptr;                                      // batch * max(n, m) doubles
  double *kcache = nullptr, *rcache = nullptr;                  // dense mode: per-CTA inverse cache
  double *Pv = nullptr, *Av = nullptr, *Atv = nullptr, *pdv = nullptr, *stage_mat = nullptr;   // per-element matrix values
  bool per_mat = false;
  BInfo* info = nullptr;
  int* next = nullptr;
  // pinned host mirrors of the last solve
  double *h_x = nullptr, *h_y = nullptr, *h_cn = nullptr, *h_cm = nullptr;
  BInfo* h_info = nullptr;
  int lgn = 4, lgm = 4, grid = 1, nt = NTB_PCG;
  bool dense = false;
  size_t smem = 0;
  bool solved = false, any_cert = false;
  b200qp_batch_stats stats{};

  ~b200qp_batch() {
    cudaSetDevice(device);
    for (double* p : {q, l, u, xs, zs, ys, rho_b, sol_x, sol_y, cert_n, cert_m, stage, kcache, rcache, Pv, Av, Atv, pdv, stage_mat})
      if (p) cudaFree(p);
    if (info) cudaFree(info);
    if (next) cudaFree(next);
    for (double* p : {h_x, h_y, h_cn, h_cm})
      if (p) cudaFreeHost(p);
    if (h_info) cudaFreeHost(h_info);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
    if (proto) b200qp_cleanup(proto);
  }
};

namespace {

double* dalloc(size_t cnt) {
  double* p = nullptr;
  BCK(cudaMalloc(&p, std::max<size_t>(cnt, 1) * sizeof(double)));
  return p;
}
double* halloc(size_t cnt) {
  double* p = nullptr;
  BCK(cudaMallocHost(&p, std::max<size_t>(cnt, 1) * sizeof(double)));
  return p;
}

// host array [batch][len] -> device, scaled: out = scalar * in .* s
void upload_scaled(b200qp_batch* b, const double* h, int len, double* out, const double* s, double scalar) {
  const long long total = (long long)b->batch * len;
  if (!total) return;
  BCK(cudaMemcpyAsync(b->stage, h, total * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  kb_scale_rows<<<cdiv(total, 256), 256, 0, b->stream>>>(total, len, out, b->stage, s, scalar);
}

// validate + install new bounds (either may be null); returns 0 or B200QP_DATA_VALIDATION_ERROR, in which case
// the stored bounds are untouched.  Device-side: no host pass over the batch x m entries.
int set_bounds(b200qp_batch* b, const double* l_new, const double* u_new) {
  const long long bm = (long long)b->batch * b->m;
  if (!bm || (!l_new && !u_new)) return 0;
  double* sl = l_new ? b->stage : nullptr;
  double* su = u_new ? b->stage + bm : nullptr;
  if (l_new) BCK(cudaMemcpyAsync(sl, l_new, bm * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  if (u_new) BCK(cudaMemcpyAsync(su, u_new, bm * sizeof(double), cudaMemcpyHostToDevice, b->stream));
  BCK(cudaMemsetAsync(b->next, 0, sizeof(int), b->stream));        // the work counter doubles as the flag
  kb_bounds<<<cdiv(bm, 256), 256, 0, b->stream>>>(bm, b->m, sl, su, b->l, b->u, b->v.E, b->next);
  int bad = 0;
  BCK(cudaMemcpyAsync(&bad, b->next, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  BCK(cudaStreamSynchronize(b->stream));
  if (bad) return B200QP_DATA_VALIDATION_ERROR;
  if (l_new) BCK(cudaMemcpyAsync(b->l, sl, bm * sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
  if (u_new) BCK(cudaMemcpyAsync(b->u, su, bm * sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
  BCK(cudaStreamSynchronize(b->stream));
  return 0;
}

}  // namespace

extern "C" {

b200qp_int b200qp_batch_cleanup(b200qp_batch* b) {
  delete b;
  return 0;
}

b200qp_int b200qp_batch_setup(b200qp_batch** batchp, b200qp_int batch, const b200qp_csc* P, const b200qp_float* q,
                              const b200qp_csc* A, const b200qp_float* l, const b200qp_float* u, b200qp_int m,
                              b200qp_int n, const b200qp_settings* settings) {
  if (!batchp) return B200QP_DATA_VALIDATION_ERROR;
  *batchp = nullptr;
  const double t0 = now_s();
  if (batch <= 0 || !settings || !q || n <= 0 || m < 0 || (m > 0 && (!l || !u))) return B200QP_DATA_VALIDATION_ERROR;
  b200qp_batch* b = new b200qp_batch();
  b200qp_settings st = *settings;
  st.verbose = 0;
  st.polishing = 0;
  // element 0 fixes the scaling (the reference's setup(); later elements are update()s of it)
  // element 0, bounds clipped to +-INFTY as the single-solver facade does (interface.py:237-238)
  std::vector<double> l0(m), u0(m);
  for (int i = 0; i < m; ++i) { l0[i] = std::min(std::max(l[i], -B200QP_INFTY), B200QP_INFTY); u0[i] = std::min(std::max(u[i], -B200QP_INFTY), B200QP_INFTY); }
  int rc = b200qp_setup(&b->proto, P, q, A, l0.data(), u0.data(), m, n, &st);
  if (rc) { delete b; return rc; }
  b->proto->settings->verbose = settings->verbose;
  try {
    rc = b200qp_internal_view(b->proto, &b->v);
    if (rc) throw rc;
    b->batch = batch; b->n = n; b->m = m; b->device = b->v.device;
    BCK(cudaSetDevice(b->device));
    BCK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking));
    BCK(cudaEventCreate(&b->ev0));
    BCK(cudaEventCreate(&b->ev1));
    const size_t bn = (size_t)batch * n, bm = (size_t)batch * m;
    b->q = dalloc(bn); b->xs = dalloc(bn); b->sol_x = dalloc(bn); b->cert_n = dalloc(bn);
    b->l = dalloc(bm); b->u = dalloc(bm); b->zs = dalloc(bm); b->ys = dalloc(bm); b->sol_y = dalloc(bm); b->cert_m = dalloc(bm);
    b->rho_b = dalloc(batch);
    b->stage = dalloc(std::max(bn, 2 * bm));
    BCK(cudaMalloc(&b->info, (size_t)batch * sizeof(BInfo)));
    BCK(cudaMalloc(&b->next, sizeof(int)));
    b->h_x = halloc(bn); b->h_y = halloc(bm); b->h_cn = halloc(bn); b->h_cm = halloc(bm);
    BCK(cudaMallocHost(&b->h_info, (size_t)batch * sizeof(BInfo)));
    std::memset(b->h_info, 0, (size_t)batch * sizeof(BInfo));
    for (int i = 0; i < batch; ++i) b->h_info[i].status = B200QP_UNSOLVED;
    std::fill(b->h_x, b->h_x + bn, 0.0); std::fill(b->h_cn, b->h_cn + bn, 0.0);
    std::fill(b->h_y, b->h_y + bm, 0.0); std::fill(b->h_cm, b->h_cm + bm, 0.0);
    BCK(cudaMemsetAsync(b->xs, 0, std::max<size_t>(bn, 1) * sizeof(double), b->stream));
    BCK(cudaMemsetAsync(b->zs, 0, std::max<size_t>(bm, 1) * sizeof(double), b->stream));
    BCK(cudaMemsetAsync(b->ys, 0, std::max<size_t>(bm, 1) * sizeof(double), b->stream));
    kb_fill<<<cdiv(batch, 256), 256, 0, b->stream>>>(batch, b->rho_b, b->proto->settings->rho);
    upload_scaled(b, q, n, b->q, b->v.D, b->v.c);                 // q <- c D q   (_osqp.py:1328)
    BCK(cudaStreamSynchronize(b->stream));                        // stage is re-used
    if (m) {
      BCK(cudaMemsetAsync(b->l, 0, bm * sizeof(double), b->stream));
      BCK(cudaMemsetAsync(b->u, 0, bm * sizeof(double), b->stream));
      rc = set_bounds(b, l, u);                                   // l <- E l, u <- E u, every element l <= u
      if (rc) throw rc;
    }
    // launch geometry: every vector of one QP in shared memory
    b->lgn = pick_blg(n ? (double)(b->v.P.nnz + b->v.At.nnz) / n / 2 : 1.0);
    b->lgm = pick_blg(m ? (double)b->v.A.nnz / m : 1.0);
    cudaDeviceProp prop;
    BCK(cudaGetDeviceProperties(&prop, b->device));
    // inner solver: K^-1 on chip when it fits next to the vectors (small QPs), else the engine's PCG.
    // B200QP_BATCH_DENSE=0 / 1 forces the choice (1 still requires that it fits).
    int nt_dense = NTB_DENSE;
    if (const char* e = std::getenv("B200QP_BATCH_NT")) { const int v = std::atoi(e); if (v == 512 || v == 1024) nt_dense = v; }
    b->dense = n <= DENSE_NMAX && batch_smem_bytes(n, m, true, nt_dense) <= (size_t)prop.sharedMemPerBlockOptin;
    if (const char* e = std::getenv("B200QP_BATCH_DENSE")) b->dense = b->dense && std::atoi(e) != 0;
    b->nt = b->dense ? nt_dense : NTB_PCG;
    b->smem = batch_smem_bytes(n, m, b->dense, b->nt);
    if (b->smem > (size_t)prop.sharedMemPerBlockOptin) {
      std::fprintf(stderr, "[b200qp] batch: n = %d, m = %d needs %zu bytes of shared memory per QP (limit %zu); "
                           "use one b200qp_solver per QP for problems of this size\n",
                   n, m, b->smem, (size_t)prop.sharedMemPerBlockOptin);
      throw (int)B200QP_MEM_ALLOC_ERROR;
    }
    batch_fn fn = pick_batch_kernel(b->lgn, b->lgm, b->dense, b->nt);
    BCK(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smem));
    int occ = 0;
    BCK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)fn, b->nt, b->smem));
    if (occ < 1) throw (int)B200QP_ALGEBRA_LOAD_ERROR;
    b->grid = std::min(batch, occ * prop.multiProcessorCount);
    if (b->dense) {
      b->kcache = dalloc((size_t)b->grid * n * batch_ldk(n));
      b->rcache = dalloc((size_t)b->grid * std::max(m, 1));
    }
    b->stats.grid = b->grid; b->stats.block = b->nt; b->stats.smem_bytes = (int)b->smem; b->stats.dense = b->dense ? 1 : 0;
    b->stats.batch = batch; b->stats.n = n; b->stats.m = m;
  } catch (int code) {
    delete b;
    return code;
  }
  b->stats.setup_time = now_s() - t0;
  if (settings->verbose) {
    std::printf("b200qp batch: %d QPs, n = %d, m = %d, nnz(P) + nnz(A) = %d; %s, grid %d x %d, %zu B shared memory per QP\n",
                batch, n, m, P->p[n] + A->p[n], b->dense ? "on-chip K^-1" : "Jacobi-PCG", b->grid, b->nt, b->smem);
    std::fflush(stdout);
  }
  *batchp = b;
  return 0;
}

b200qp_int b200qp_batch_solve(b200qp_batch* b) {
  if (!b || !b->proto) return B200QP_WORKSPACE_NOT_INIT_ERROR;
  const double t0 = now_s();
  const b200qp_settings* st = b->proto->settings;
  try {
    BCK(cudaSetDevice(b->device));
    BArgs a{};
    a.n = b->n; a.m = b->m; a.batch = b->batch;
    a.P = b->v.P; a.A = b->v.A; a.At = b->v.At;
    a.pdiag = b->v.pdiag; a.D = b->v.D; a.Dinv = b->v.Dinv; a.E = b->v.E; a.Einv = b->v.Einv;
    a.q = b->q; a.l = b->l; a.u = b->u;
    a.xs = b->xs; a.zs = b->zs; a.ys = b->ys; a.rho_b = b->rho_b;
    a.sol_x = b->sol_x; a.sol_y = b->sol_y; a.cert_n = b->cert_n; a.cert_m = b->cert_m;
    a.info = b->info; a.next = b->next;
    a.kcache = b->kcache; a.rcache = b->rcache;
    a.Pv = b->Pv; a.Av = b->Av; a.Atv = b->Atv; a.pdv = b->pdv; a.per_mat = b->per_mat ? 1 : 0;
    a.sigma = st->sigma; a.alpha = st->alpha; a.eps_abs = st->eps_abs; a.eps_rel = st->eps_rel;
    a.eps_pinf = st->eps_prim_inf; a.eps_dinf = st->eps_dual_inf; a.c = b->v.c; a.adapt_tol = st->adaptive_rho_tolerance;
    a.cg_tol_frac = st->cg_tol_fraction; a.rho_eq_factor = b->v.rho_eq_factor; a.time_limit_ns = st->time_limit * 1e9;
    a.max_iter = st->max_iter; a.check_term = st->check_termination; a.adaptive_rho = st->adaptive_rho;
    a.adapt_interval = st->adaptive_rho_interval > 0 ? st->adaptive_rho_interval : 50;
    a.cg_max_iter = st->cg_max_iter; a.cg_precond = st->cg_precond == B200QP_DIAGONAL_PRECONDITIONER;
    a.unscale = (st->scaling && !st->scaled_termination) ? 1 : 0;
    a.rho_is_vec = st->rho_is_vec; a.check_dualgap = st->check_dualgap;
    a.warm = st->warm_starting ? 1 : 0;
    BCK(cudaMemsetAsync(b->next, 0, sizeof(int), b->stream));
    batch_fn fn = pick_batch_kernel(b->lgn, b->lgm, b->dense, b->nt);
    BCK(cudaEventRecord(b->ev0, b->stream));
    fn<<<b->grid, b->nt, b->smem, b->stream>>>(a);
    BCK(cudaEventRecord(b->ev1, b->stream));
    BCK(cudaGetLastError());
    const size_t bn = (size_t)b->batch * b->n, bm = (size_t)b->batch * b->m;
    BCK(cudaMemcpyAsync(b->h_info, b->info, (size_t)b->batch * sizeof(BInfo), cudaMemcpyDeviceToHost, b->stream));
    BCK(cudaMemcpyAsync(b->h_x, b->sol_x, bn * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    if (bm) BCK(cudaMemcpyAsync(b->h_y, b->sol_y, bm * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
    BCK(cudaStreamSynchronize(b->stream));
    float ms = 0.f;
    BCK(cudaEventElapsedTime(&ms, b->ev0, b->ev1));
    long long it = 0, cg = 0;
    bool any_cert = false;
    for (int i = 0; i < b->batch; ++i) {
      it += b->h_info[i].iter; cg += b->h_info[i].cg_iters;
      const int s = b->h_info[i].status;
      any_cert = any_cert || (s >= B200QP_PRIMAL_INFEASIBLE && s <= B200QP_DUAL_INFEASIBLE_INACCURATE);
    }
    b->any_cert = any_cert;
    if (any_cert) {   // certificates are rare: fetch them only when some element is infeasible
      BCK(cudaMemcpy(b->h_cn, b->cert_n, bn * sizeof(double), cudaMemcpyDeviceToHost));
      if (bm) BCK(cudaMemcpy(b->h_cm, b->cert_m, bm * sizeof(double), cudaMemcpyDeviceToHost));
    }
    b->stats.admm_iters = it; b->stats.cg_iters = cg; b->stats.kernel_ms = ms;
    b->solved = true;
  } catch (int code) {
    return code;
  }
  b->stats.solve_time = now_s() - t0;
  if (st->verbose) {
    int ns = 0;
    for (int i = 0; i < b->batch; ++i) ns += b->h_info[i].status == B200QP_SOLVED;
    std::printf("b200qp batch: %d / %d solved, %lld ADMM iterations, %lld PCG iterations, kernel %.3f ms\n", ns, b->batch,
                b->stats.admm_iters, b->stats.cg_iters, b->stats.kernel_ms);
    std::fflush(stdout);
  }
  return 0;
}

b200qp_int b200qp_batch_update_data_vec(b200qp_batch* b, const b200qp_float* q_new, const b200qp_float* l_new,
                                        const b200qp_float* u_new) {
  if (!b || !b->proto) return B200QP_WORKSPACE_NOT_INIT_ERROR;
  const double t0 = now_s();
  try {
    BCK(cudaSetDevice(b->device));
    const int rcb = set_bounds(b, l_new, u_new);
    if (rcb) return rcb;
    if (q_new) { upload_scaled(b, q_new, b->n, b->q, b->v.D, b->v.c); BCK(cudaStreamSynchronize(b->stream)); }
    for (int i = 0; i < b->batch; ++i) b->h_info[i].status = B200QP_UNSOLVED;
  } catch (int code) {
    return code;
  }
  b->stats.update_time = now_s() - t0;
  return 0;
}

b200qp_int b200qp_batch_update_data_mat(b200qp_batch* b, const b200qp_float* Px_new, const b200qp_float* Ax_new) {
  if (!b || !b->proto) return B200QP_WORKSPACE_NOT_INIT_ERROR;
  if (!Px_new && !Ax_new) return 0;
  const double t0 = now_s();
  try {
    BCK(cudaSetDevice(b->device));
    const InternalView& v = b->v;
    const long long B = b->batch;
    const int n = b->n, nnzP = v.P.nnz, nnzA = v.A.nnz, nnzPt = v.nnzPt;
    if (!b->per_mat) {
      // first per-element update: every element starts from the shared (scaled) values of the setup
      b->Pv = dalloc((size_t)B * std::max(nnzP, 1)); b->Av = dalloc((size_t)B * std::max(nnzA, 1));
      b->Atv = dalloc((size_t)B * std::max(nnzA, 1)); b->pdv = dalloc((size_t)B * n);
      b->stage_mat = dalloc((size_t)B * std::max(std::max(nnzPt, nnzA), 1));
      if (nnzP) kb_tile_rows<<<cdiv(B * nnzP, 256), 256, 0, b->stream>>>(B * nnzP, nnzP, b->Pv, v.P.val);
      if (nnzA) {
        kb_tile_rows<<<cdiv(B * nnzA, 256), 256, 0, b->stream>>>(B * nnzA, nnzA, b->Av, v.A.val);
        kb_tile_rows<<<cdiv(B * nnzA, 256), 256, 0, b->stream>>>(B * nnzA, nnzA, b->Atv, v.At.val);
      }
      kb_tile_rows<<<cdiv(B * n, 256), 256, 0, b->stream>>>(B * n, n, b->pdv, v.pdiag);
      b->per_mat = true;
    }
    if (Px_new && nnzPt) {
      BCK(cudaMemcpyAsync(b->stage_mat, Px_new, (size_t)B * nnzPt * sizeof(double), cudaMemcpyHostToDevice, b->stream));
      kb_update_P<<<cdiv(B * nnzPt, 256), 256, 0, b->stream>>>(B * nnzPt, nnzPt, nnzP, n, b->stage_mat, v.Pt_colptr, v.Pt_rowidx,
                                                               v.posU, v.posL, v.D, v.c, b->Pv, b->pdv);
      BCK(cudaStreamSynchronize(b->stream));
    }
    if (Ax_new && nnzA) {
      BCK(cudaMemcpyAsync(b->stage_mat, Ax_new, (size_t)B * nnzA * sizeof(double), cudaMemcpyHostToDevice, b->stream));
      kb_update_A<<<cdiv(B * nnzA, 256), 256, 0, b->stream>>>(B * nnzA, nnzA, b->stage_mat, v.At.ptr, v.At.idx, n, v.mapA, v.D, v.E,
                                                              b->Atv, b->Av);
      BCK(cudaStreamSynchronize(b->stream));
    }
    BCK(cudaStreamSynchronize(b->stream));
    for (int i = 0; i < b->batch; ++i) b->h_info[i].status = B200QP_UNSOLVED;
  } catch (int code) {
    return code;
  }
  b->stats.update_time = now_s() - t0;
  return 0;
}

b200qp_int b200qp_batch_warm_start(b200qp_batch* b, const b200qp_float* x, const b200qp_float* y) {
  if (!b || !b->proto) return B200QP_WORKSPACE_NOT_INIT_ERROR;
  try {
    BCK(cudaSetDevice(b->device));
    b->proto->settings->warm_starting = 1;                                  // reference: _osqp.py:1498
    const long long bn = (long long)b->batch * b->n, bm = (long long)b->batch * b->m;
    if (x) upload_scaled(b, x, b->n, b->xs, b->v.Dinv, 1.0);                // x <- Dinv x
    else BCK(cudaMemsetAsync(b->xs, 0, bn * sizeof(double), b->stream));
    BCK(cudaStreamSynchronize(b->stream));
    if (bm) {
      if (y) upload_scaled(b, y, b->m, b->ys, b->v.Einv, b->v.c);           // y <- c Einv y
      else BCK(cudaMemsetAsync(b->ys, 0, bm * sizeof(double), b->stream));
      kb_spmv<<<cdiv(bm * 8, 256), 256, 0, b->stream>>>(b->v.A, b->batch, b->xs, b->zs);   // z <- A x
      BCK(cudaStreamSynchronize(b->stream));
    }
  } catch (int code) {
    return code;
  }
  return 0;
}

b200qp_int b200qp_batch_update_settings(b200qp_batch* b, const b200qp_settings* ns) {
  if (!b || !b->proto || !ns) return B200QP_WORKSPACE_NOT_INIT_ERROR;
  b200qp_settings t = *ns;
  t.polishing = 0;
  return b200qp_update_settings(b->proto, &t);
}

b200qp_int b200qp_batch_update_rho(b200qp_batch* b, b200qp_float rho_new) {
  if (!b || !b->proto) return B200QP_WORKSPACE_NOT_INIT_ERROR;
  const int rc = b200qp_update_rho(b->proto, rho_new);
  if (rc) return rc;
  try {
    BCK(cudaSetDevice(b->device));
    kb_fill<<<cdiv(b->batch, 256), 256, 0, b->stream>>>(b->batch, b->rho_b, b->proto->settings->rho);
    BCK(cudaStreamSynchronize(b->stream));
  } catch (int code) {
    return code;
  }
  return 0;
}

b200qp_int b200qp_batch_get_solution(const b200qp_batch* b, b200qp_float* x, b200qp_float* y,
                                     b200qp_float* prim_inf_cert, b200qp_float* dual_inf_cert) {
  if (!b || !b->proto) return B200QP_WORKSPACE_NOT_INIT_ERROR;
  const size_t bn = (size_t)b->batch * b->n, bm = (size_t)b->batch * b->m;
  if (x) std::memcpy(x, b->h_x, bn * sizeof(double));
  if (y) std::memcpy(y, b->h_y, bm * sizeof(double));
  const double nanv = std::nan("");   // no element was infeasible: all-NaN certificates, as the reference returns
  if (prim_inf_cert) { if (b->any_cert) std::memcpy(prim_inf_cert, b->h_cm, bm * sizeof(double)); else std::fill(prim_inf_cert, prim_inf_cert + bm, nanv); }
  if (dual_inf_cert) { if (b->any_cert) std::memcpy(dual_inf_cert, b->h_cn, bn * sizeof(double)); else std::fill(dual_inf_cert, dual_inf_cert + bn, nanv); }
  return 0;
}

b200qp_int b200qp_batch_get_info(const b200qp_batch* b, b200qp_batch_info* info) {
  if (!b || !b->proto || !info) return B200QP_WORKSPACE_NOT_INIT_ERROR;
  static_assert(sizeof(b200qp_batch_info) == sizeof(BInfo), "info layouts must match");
  std::memcpy(info, b->h_info, (size_t)b->batch * sizeof(BInfo));
  return 0;
}

b200qp_int b200qp_batch_get_stats(const b200qp_batch* b, b200qp_batch_stats* stats) {
  if (!b || !b->proto || !stats) return B200QP_WORKSPACE_NOT_INIT_ERROR;
  *stats = b->stats;
  return 0;
}

}  // extern "C"
