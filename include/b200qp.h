/*
 * b200qp.h -- C ABI of the Blackwell-native OSQP ADMM engine (libb200qp.so).
 *
 * Drop-in boundary: these entry points are what the reference's pybind11 layer
 * binds for the solve() path.  Each declaration cites the reference call site it
 * replaces (paths relative to the reference repo, osqp/osqp-python):
 *
 *   reference (upstream C API, un-vendored osqp v1.0.0)     ->  this library
 *   osqp_setup            src/bindings.cpp.in:153           ->  b200qp_setup
 *   osqp_solve            src/bindings.cpp.in:198           ->  b200qp_solve
 *   osqp_warm_start       src/bindings.cpp.in:193           ->  b200qp_warm_start
 *   osqp_update_data_vec  src/bindings.cpp.in:237           ->  b200qp_update_data_vec
 *   osqp_update_data_mat  src/bindings.cpp.in:280           ->  b200qp_update_data_mat
 *   osqp_update_settings  src/bindings.cpp.in:204           ->  b200qp_update_settings
 *   osqp_update_rho       src/bindings.cpp.in:213           ->  b200qp_update_rho
 *   osqp_cleanup          src/bindings.cpp.in:160           ->  b200qp_cleanup
 *   osqp_set_default_settings  src/bindings.cpp.in:449      ->  b200qp_set_default_settings
 *   osqp_capabilities     src/bindings.cpp.in:402           ->  b200qp_capabilities
 *   osqp_adjoint_derivative_compute  src/bindings.cpp.in:301  ->  b200qp_adjoint_derivative_compute
 *   osqp_adjoint_derivative_get_mat  src/bindings.cpp.in:309  ->  b200qp_adjoint_derivative_get_mat
 *   osqp_adjoint_derivative_get_vec  src/bindings.cpp.in:318  ->  b200qp_adjoint_derivative_get_vec
 *   OSQPCscMatrix / OSQPSettings / OSQPInfo / OSQPSolution / OSQPSolver
 *                         src/bindings.cpp.in:41-48, 405-447, 473-492, 79-105, 163-176
 *
 * Plain pointers and sizes only; every array argument is a HOST buffer that is
 * consumed before the call returns (the reference binding passes temporaries,
 * src/bindings.cpp.in:185-190,224-234,251-277).  Integers are 32-bit, reals are
 * fp64 (reference build: OSQP_USE_LONG=OFF, no OSQP_USE_FLOAT; CMakeLists.txt:9).
 *
 * There is no CPU fallback: without a CUDA device b200qp_setup returns
 * B200QP_ALGEBRA_LOAD_ERROR and the compute entry points return
 * B200QP_WORKSPACE_NOT_INIT_ERROR.
 */
#ifndef B200QP_H
#define B200QP_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int b200qp_int;      /* OSQPInt   */
typedef double b200qp_float; /* OSQPFloat */

#define B200QP_INFTY 1e30 /* OSQP_INFTY, src/bindings.cpp.in:340 */

/* osqp_status_type, src/bindings.cpp.in:349-361 (same order, values 1..11) */
enum b200qp_status_type {
  B200QP_SOLVED = 1,
  B200QP_SOLVED_INACCURATE,
  B200QP_PRIMAL_INFEASIBLE,
  B200QP_PRIMAL_INFEASIBLE_INACCURATE,
  B200QP_DUAL_INFEASIBLE,
  B200QP_DUAL_INFEASIBLE_INACCURATE,
  B200QP_MAX_ITER_REACHED,
  B200QP_TIME_LIMIT_REACHED,
  B200QP_NON_CVX,
  B200QP_SIGINT,
  B200QP_UNSOLVED
};

/* osqp_error_type, src/bindings.cpp.in:364-375 */
enum b200qp_error_type {
  B200QP_NO_ERROR = 0,
  B200QP_DATA_VALIDATION_ERROR,
  B200QP_SETTINGS_VALIDATION_ERROR,
  B200QP_LINSYS_SOLVER_INIT_ERROR,
  B200QP_NONCVX_ERROR,
  B200QP_MEM_ALLOC_ERROR,
  B200QP_WORKSPACE_NOT_INIT_ERROR,
  B200QP_ALGEBRA_LOAD_ERROR,
  B200QP_CODEGEN_DEFINES_ERROR,
  B200QP_DATA_NOT_INITIALIZED,
  B200QP_FUNC_NOT_IMPLEMENTED
};

/* osqp_linsys_solver_type, src/bindings.cpp.in:343-346 */
enum b200qp_linsys_solver_type { B200QP_UNKNOWN_SOLVER = 0, B200QP_DIRECT_SOLVER = 1, B200QP_INDIRECT_SOLVER = 2 };
/* osqp_precond_type, src/bindings.cpp.in:378-381 */
enum b200qp_precond_type { B200QP_NO_PRECONDITIONER = 0, B200QP_DIAGONAL_PRECONDITIONER = 1 };
/* osqp_capabilities_type, src/bindings.cpp.in:395-400 */
enum b200qp_capabilities_type {
  B200QP_CAPABILITY_DIRECT_SOLVER = 0x01,
  B200QP_CAPABILITY_INDIRECT_SOLVER = 0x02,
  B200QP_CAPABILITY_CODEGEN = 0x04,
  B200QP_CAPABILITY_UPDATE_MATRICES = 0x08,
  B200QP_CAPABILITY_DERIVATIVES = 0x10
};

/* OSQPCscMatrix, src/bindings.cpp.in:41-48.  P is UPPER-TRIANGULAR CSC (interface.py:221-222),
 * A is CSC with sorted row indices (interface.py:225-235). */
typedef struct {
  b200qp_int m, n;
  b200qp_int *p; /* column pointers, n+1 */
  b200qp_int *i; /* row indices, nzmax   */
  b200qp_float *x;
  b200qp_int nzmax;
  b200qp_int nz; /* -1: compressed column */
} b200qp_csc;

/* OSQPSettings, src/bindings.cpp.in:405-447 (all 29 fields, same names) */
typedef struct {
  b200qp_int device;
  b200qp_int linsys_solver; /* only B200QP_INDIRECT_SOLVER is implemented */
  b200qp_int verbose;
  b200qp_int warm_starting;
  b200qp_int scaling;
  b200qp_int polishing;
  b200qp_float rho;
  b200qp_int rho_is_vec;
  b200qp_float sigma;
  b200qp_float alpha;
  b200qp_int cg_max_iter;
  b200qp_int cg_tol_reduction; /* accepted, unused: the PCG stops at |r|inf <= cg_tol_fraction * |r0|inf */
  b200qp_float cg_tol_fraction;
  b200qp_int cg_precond;
  b200qp_int adaptive_rho;
  b200qp_int adaptive_rho_interval; /* 0 = engine default (50) */
  b200qp_float adaptive_rho_fraction; /* accepted, unused: no setup-time heuristic */
  b200qp_float adaptive_rho_tolerance;
  b200qp_int max_iter;
  b200qp_float eps_abs;
  b200qp_float eps_rel;
  b200qp_float eps_prim_inf;
  b200qp_float eps_dual_inf;
  b200qp_int scaled_termination;
  b200qp_int check_termination; /* every k iterations; 0 = only at max_iter */
  b200qp_int check_dualgap;
  b200qp_float time_limit; /* seconds of device time; checked at termination checks */
  b200qp_float delta;
  b200qp_int polish_refine_iter;
} b200qp_settings;

/* OSQPInfo, src/bindings.cpp.in:473-492 */
typedef struct {
  char status[32];
  b200qp_int status_val;
  b200qp_int status_polish;
  b200qp_float obj_val;
  b200qp_float dual_obj_val;
  b200qp_float prim_res;
  b200qp_float dual_res;
  b200qp_float duality_gap;
  b200qp_int iter;
  b200qp_int rho_updates;
  b200qp_float rho_estimate;
  b200qp_float setup_time;
  b200qp_float solve_time;
  b200qp_float update_time;
  b200qp_float polish_time;
  b200qp_float run_time;
  b200qp_float primdual_int;
  b200qp_float rel_kkt_error;
} b200qp_info;

/* OSQPSolution, src/bindings.cpp.in:79-105: host arrays owned by the solver, refreshed by solve() */
typedef struct {
  b200qp_float *x;             /* n */
  b200qp_float *y;             /* m */
  b200qp_float *prim_inf_cert; /* m */
  b200qp_float *dual_inf_cert; /* n */
} b200qp_solution;

/* OSQPSolver, src/bindings.cpp.in:163-176 */
typedef struct {
  b200qp_settings *settings;
  b200qp_solution *solution;
  b200qp_info *info;
  void *work; /* opaque device workspace */
} b200qp_solver;

/* Engine counters of the last solve (not in the reference; used by bench.py's roofline) */
typedef struct {
  long long admm_iters;
  long long cg_iters;     /* total PCG iterations                                      */
  long long grid_barriers;
  double kernel_ms;       /* device time of the ADMM kernel, CUDA events on its stream */
  double algo_bytes;      /* algorithmic HBM bytes of that launch (DESIGN.md byte model) */
  b200qp_int grid, block; /* launch shape                                               */
  b200qp_int n, m, nnzP_full, nnzA;
  double phase_ms[8];     /* device time per phase kind, measured on CTA 0 (barrier waits included):
                             CSR kernel R, C1 (A p), C2 (K p), C3, C4, Z; tile kernel R, B (A u), C (K u), V, -, Z */
  double algo_bytes_engine; /* same launch, counted with the engine's own storage (tile blobs, its vector passes) */
  double spmv_bytes[2];   /* algorithmic bytes of ONE A-product phase / ONE K-product phase (SURVEY 8d model)  */
  b200qp_int kernel_kind; /* 0 = CSR lanes-per-row kernel, 1 = compacted-slice tiles + TMA ring, 2 = multi-GPU (CSR),
                             3 = multi-GPU (tiles + TMA ring, flat NVLink barrier)                                 */
  b200qp_int ring_stages, ring_stage_bytes;
  double autotune_ms[2];  /* setup-time measurement behind the ring depth: 3 ADMM iterations with the 94 KB ring
                             (196 KB shared-memory carve-out) / the 108 KB ring (228 KB); 0 = not measured          */
} b200qp_stats;

/* ---- the reference path ------------------------------------------------------------------ */
void b200qp_set_default_settings(b200qp_settings *settings);
b200qp_int b200qp_capabilities(void);
const char *b200qp_version(void);
const char *b200qp_error_message(b200qp_int code);

b200qp_int b200qp_setup(b200qp_solver **solverp, const b200qp_csc *P, const b200qp_float *q,
                        const b200qp_csc *A, const b200qp_float *l, const b200qp_float *u,
                        b200qp_int m, b200qp_int n, const b200qp_settings *settings);
b200qp_int b200qp_solve(b200qp_solver *solver);
b200qp_int b200qp_cleanup(b200qp_solver *solver);
b200qp_int b200qp_warm_start(b200qp_solver *solver, const b200qp_float *x, const b200qp_float *y);
void b200qp_cold_start(b200qp_solver *solver);
b200qp_int b200qp_update_data_vec(b200qp_solver *solver, const b200qp_float *q_new,
                                  const b200qp_float *l_new, const b200qp_float *u_new);
b200qp_int b200qp_update_data_mat(b200qp_solver *solver, const b200qp_float *Px_new,
                                  const b200qp_int *Px_new_idx, b200qp_int P_new_n,
                                  const b200qp_float *Ax_new, const b200qp_int *Ax_new_idx,
                                  b200qp_int A_new_n);
b200qp_int b200qp_update_settings(b200qp_solver *solver, const b200qp_settings *new_settings);
b200qp_int b200qp_update_rho(b200qp_solver *solver, b200qp_float rho_new);
b200qp_int b200qp_get_stats(const b200qp_solver *solver, b200qp_stats *stats);

/* ---- adjoint derivatives of a loss L(x*, y*) (reference: src/osqp/interface.py:519-598,
 * src/bindings.cpp.in:283-319; semantics pinned by src/osqp/tests/derivative_test.py) -----------------------
 * After a solve() that ended SOLVED: compute(dx = dL/dx, dy = dL/dy or NULL) solves the adjoint KKT system of the
 * active set ON DEVICE (the same persistent ADMM kernel, run on the equality-constrained auxiliary QP
 * min 1/2 r'P r - dx'r  s.t. a_i'r = dy_i on the active rows); get_vec / get_mat then return
 *   dq = -rx,  dl_i / du_i = ry_i on lower- / upper-active rows,
 *   dP = -(rx x' + x rx')/2 on the upper-triangular pattern passed in dP,  dA = -(y rx' + ry x') on the pattern in dA
 * (dP->x and dA->x are overwritten; p and i describe the patterns of the matrices given at setup). */
b200qp_int b200qp_adjoint_derivative_compute(b200qp_solver *solver, const b200qp_float *dx, const b200qp_float *dy);
b200qp_int b200qp_adjoint_derivative_get_mat(b200qp_solver *solver, b200qp_csc *dP, b200qp_csc *dA);
b200qp_int b200qp_adjoint_derivative_get_vec(b200qp_solver *solver, b200qp_float *dq, b200qp_float *dl, b200qp_float *du);

/* ---- multi-GPU: ONE QP row-partitioned over `world` GPUs (BASELINE.json configs[3]) ----
 * Not in the reference (its CUDA algebra is single-GPU: OSQPSettings.device is one int,
 * src/bindings.cpp.in:409).  Every rank passes the full problem and settings->device = its GPU; rank g
 * keeps a slab of constraint rows.  Connect the ranks once (cudaIpc handles between processes, or
 * b200qp_dist_connect_local inside one process), then call b200qp_solve on every rank concurrently
 * (or b200qp_dist_solve_group).  After a solve solution->x is complete on every rank; solution->y and
 * prim_inf_cert are valid on the rank's own rows [mlo, mhi) only. */
b200qp_int b200qp_dist_setup(b200qp_solver **solverp, b200qp_int rank, b200qp_int world,
                             b200qp_int ranks_on_device, const b200qp_csc *P, const b200qp_float *q,
                             const b200qp_csc *A, const b200qp_float *l, const b200qp_float *u, b200qp_int m,
                             b200qp_int n, const b200qp_settings *settings);
b200qp_int b200qp_dist_row_range(const b200qp_solver *solver, b200qp_int *mlo, b200qp_int *mhi);
b200qp_int b200qp_dist_get_handle(b200qp_solver *solver, void *handle64);
b200qp_int b200qp_dist_connect_ipc(b200qp_solver *solver, const void *handles_world_x_64);
b200qp_int b200qp_dist_connect_local(b200qp_solver **group, b200qp_int world);
b200qp_int b200qp_dist_solve_group(b200qp_solver **group, b200qp_int world);

/* ---- batched independent QPs (BASELINE.json configs[4]; SURVEY.md 8e "replicas only") ------------------
 * `batch` QPs that share one (P, A) and differ in q, l, u; one CTA per QP, every vector of a QP in shared
 * memory (and, for n up to ~150, the inverse of the reduced KKT matrix too, so the inner solve is one dense
 * mat-vec; larger QPs use the engine's Jacobi-PCG).  Replaces the reference's batched caller, which loops solver objects over a thread pool and
 * re-uses one setup through update() (src/osqp/nn/torch.py:128-164, 200-217): element b is the QP a
 * reference user gets from setup(P, q[0], A, l[0], u[0]) followed by update(q=q[b], l=l[b], u=u[b]) and
 * solve() -- same scaling (D, E, c of element 0; update() never re-equilibrates, _osqp.py:1312-1367), own
 * ADMM state and own rho.  q is batch x n, l and u are batch x m, row-major, HOST buffers consumed before
 * the call returns; bounds beyond +-B200QP_INFTY are clipped on device.  Multi-GPU: one b200qp_batch per GPU (settings->device) over a shard of the batch; no
 * collective.  Polish is not run on batch elements (status_polish stays 0).  Matrix VALUES may also differ per
 * element (b200qp_batch_update_data_mat); the sparsity pattern is shared. */
typedef struct b200qp_batch b200qp_batch;
typedef struct {            /* per element; mirrors the OSQPInfo fields that are per solve */
  b200qp_int status_val, iter, rho_updates, cg_iter;
  b200qp_float obj_val, dual_obj_val, prim_res, dual_res, rho, rho_estimate;
} b200qp_batch_info;
typedef struct {
  long long admm_iters, cg_iters;   /* summed over the batch, last solve */
  double kernel_ms;                 /* device time of the batch kernel, CUDA events on its stream */
  double setup_time, update_time, solve_time; /* host wall clock, seconds */
  b200qp_int grid, block, smem_bytes, batch, n, m;
  b200qp_int dense;                 /* 1: K^-1 held in shared memory (small QPs); 0: Jacobi-PCG */
} b200qp_batch_stats;
b200qp_int b200qp_batch_setup(b200qp_batch **batchp, b200qp_int batch, const b200qp_csc *P, const b200qp_float *q,
                              const b200qp_csc *A, const b200qp_float *l, const b200qp_float *u, b200qp_int m,
                              b200qp_int n, const b200qp_settings *settings);
b200qp_int b200qp_batch_solve(b200qp_batch *b);
b200qp_int b200qp_batch_update_data_vec(b200qp_batch *b, const b200qp_float *q_new, const b200qp_float *l_new,
                                        const b200qp_float *u_new);
/* per-element matrix values on the shared pattern (the reference's batched caller passes P_val / A_val per
 * element, src/osqp/nn/torch.py:136-140): Px_new is batch x nnz(triu P), Ax_new is batch x nnz(A), CSC order,
 * unscaled, either may be NULL.  Same semantics as osqp_update_data_mat without index arrays
 * (src/bindings.cpp.in:240-281): the scaling of the setup is kept. */
b200qp_int b200qp_batch_update_data_mat(b200qp_batch *b, const b200qp_float *Px_new, const b200qp_float *Ax_new);
b200qp_int b200qp_batch_warm_start(b200qp_batch *b, const b200qp_float *x, const b200qp_float *y);
b200qp_int b200qp_batch_update_settings(b200qp_batch *b, const b200qp_settings *new_settings);
b200qp_int b200qp_batch_update_rho(b200qp_batch *b, b200qp_float rho_new);
b200qp_int b200qp_batch_get_solution(const b200qp_batch *b, b200qp_float *x, b200qp_float *y,
                                     b200qp_float *prim_inf_cert, b200qp_float *dual_inf_cert);
b200qp_int b200qp_batch_get_info(const b200qp_batch *b, b200qp_batch_info *info);
b200qp_int b200qp_batch_get_stats(const b200qp_batch *b, b200qp_batch_stats *stats);
b200qp_int b200qp_batch_cleanup(b200qp_batch *b);

/* ---- hot operator, standalone (parity tests and the SpMV roofline line of bench.py) --------
 * y = M x for a CSR matrix (fp64 values, int32 indices).  *_host: host buffers in and out.
 * The device-handle form keeps the matrix and vectors resident and times `reps` launches with
 * CUDA events on the launching stream. */
b200qp_int b200qp_spmv_csr_host(b200qp_int rows, b200qp_int cols, const b200qp_int *ptr,
                                const b200qp_int *idx, const b200qp_float *val, const b200qp_float *x,
                                b200qp_float *y, b200qp_int variant);
typedef struct b200qp_spmv_plan b200qp_spmv_plan;
b200qp_int b200qp_spmv_plan_create(b200qp_spmv_plan **plan, b200qp_int rows, b200qp_int cols,
                                   const b200qp_int *ptr, const b200qp_int *idx, const b200qp_float *val,
                                   b200qp_int device);
b200qp_int b200qp_spmv_plan_run(b200qp_spmv_plan *plan, const b200qp_float *x_host, b200qp_float *y_host,
                                b200qp_int variant, b200qp_int reps, b200qp_int flush_l2, double *avg_ms);
void b200qp_spmv_plan_destroy(b200qp_spmv_plan *plan);
b200qp_int b200qp_spmv_num_variants(void);
const char *b200qp_spmv_variant_name(b200qp_int variant);

b200qp_int b200qp_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200QP_H */
