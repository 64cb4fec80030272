/*
 * osqp_api_types.h -- "Option A" shim: lets the reference's UNMODIFIED pybind11 binding,
 * src/bindings.cpp.in, compile against libb200qp.so.  The binding includes "osqp_api_functions.h" and
 * "osqp_api_types.h" (bindings.cpp.in:9-10); put this directory first on the include path and it gets these
 * two files instead of the upstream osqp v1.0.0 headers.  Every name below is one the binding uses; the line
 * numbers are the places in bindings.cpp.in that use it.
 */
#ifndef B200QP_OSQP_API_TYPES_SHIM_H
#define B200QP_OSQP_API_TYPES_SHIM_H

#include "../b200qp.h"

typedef b200qp_int OSQPInt;     /* bindings.cpp.in:16-21 */
typedef b200qp_float OSQPFloat;

typedef b200qp_csc OSQPCscMatrix;    /* m, n, p, i, x, nzmax, nz          :41-48   */
typedef b200qp_settings OSQPSettings; /* 29 fields, same names              :405-447 */
typedef b200qp_info OSQPInfo;         /* 18 fields, same names              :473-492 */
typedef b200qp_solution OSQPSolution; /* x, y, prim_inf_cert, dual_inf_cert :79-105  */
typedef b200qp_solver OSQPSolver;     /* ->settings, ->solution, ->info     :163-176 */

#define OSQP_INFTY B200QP_INFTY /* :340 */

/* enum types the binding registers with py::enum_ (:343-400); same members, same values */
typedef enum b200qp_linsys_solver_type osqp_linsys_solver_type;
#define OSQP_DIRECT_SOLVER B200QP_DIRECT_SOLVER
#define OSQP_INDIRECT_SOLVER B200QP_INDIRECT_SOLVER

typedef enum b200qp_status_type osqp_status_type;
#define OSQP_SOLVED B200QP_SOLVED
#define OSQP_SOLVED_INACCURATE B200QP_SOLVED_INACCURATE
#define OSQP_PRIMAL_INFEASIBLE B200QP_PRIMAL_INFEASIBLE
#define OSQP_PRIMAL_INFEASIBLE_INACCURATE B200QP_PRIMAL_INFEASIBLE_INACCURATE
#define OSQP_DUAL_INFEASIBLE B200QP_DUAL_INFEASIBLE
#define OSQP_DUAL_INFEASIBLE_INACCURATE B200QP_DUAL_INFEASIBLE_INACCURATE
#define OSQP_MAX_ITER_REACHED B200QP_MAX_ITER_REACHED
#define OSQP_TIME_LIMIT_REACHED B200QP_TIME_LIMIT_REACHED
#define OSQP_NON_CVX B200QP_NON_CVX
#define OSQP_SIGINT B200QP_SIGINT
#define OSQP_UNSOLVED B200QP_UNSOLVED

typedef enum b200qp_error_type osqp_error_type;
#define OSQP_NO_ERROR B200QP_NO_ERROR
#define OSQP_DATA_VALIDATION_ERROR B200QP_DATA_VALIDATION_ERROR
#define OSQP_SETTINGS_VALIDATION_ERROR B200QP_SETTINGS_VALIDATION_ERROR
#define OSQP_LINSYS_SOLVER_INIT_ERROR B200QP_LINSYS_SOLVER_INIT_ERROR
#define OSQP_NONCVX_ERROR B200QP_NONCVX_ERROR
#define OSQP_MEM_ALLOC_ERROR B200QP_MEM_ALLOC_ERROR
#define OSQP_WORKSPACE_NOT_INIT_ERROR B200QP_WORKSPACE_NOT_INIT_ERROR
#define OSQP_ALGEBRA_LOAD_ERROR B200QP_ALGEBRA_LOAD_ERROR
#define OSQP_CODEGEN_DEFINES_ERROR B200QP_CODEGEN_DEFINES_ERROR
#define OSQP_DATA_NOT_INITIALIZED B200QP_DATA_NOT_INITIALIZED
#define OSQP_FUNC_NOT_IMPLEMENTED B200QP_FUNC_NOT_IMPLEMENTED

typedef enum b200qp_precond_type osqp_precond_type;
#define OSQP_NO_PRECONDITIONER B200QP_NO_PRECONDITIONER
#define OSQP_DIAGONAL_PRECONDITIONER B200QP_DIAGONAL_PRECONDITIONER

typedef enum b200qp_capabilities_type osqp_capabilities_type;
#define OSQP_CAPABILITY_DIRECT_SOLVER B200QP_CAPABILITY_DIRECT_SOLVER
#define OSQP_CAPABILITY_INDIRECT_SOLVER B200QP_CAPABILITY_INDIRECT_SOLVER
#define OSQP_CAPABILITY_CODEGEN B200QP_CAPABILITY_CODEGEN
#define OSQP_CAPABILITY_UPDATE_MATRICES B200QP_CAPABILITY_UPDATE_MATRICES
#define OSQP_CAPABILITY_DERIVATIVES B200QP_CAPABILITY_DERIVATIVES

/* code generation is outside the engine (capability bit not set; the facade checks it before calling,
 * interface.py:463); the binding still needs the type to exist (:452-463) */
typedef struct {
  OSQPInt embedded_mode;
  OSQPInt float_type;
  OSQPInt printing_enable;
  OSQPInt profiling_enable;
  OSQPInt interrupt_enable;
  OSQPInt derivatives_enable;
} OSQPCodegenDefines;

#endif
