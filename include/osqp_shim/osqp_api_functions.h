/*
 * osqp_api_functions.h -- "Option A" shim, function half (see osqp_api_types.h in this directory): the upstream
 * entry points the reference binding calls, mapped 1:1 onto the C ABI of libb200qp.so.
 */
#ifndef B200QP_OSQP_API_FUNCTIONS_SHIM_H
#define B200QP_OSQP_API_FUNCTIONS_SHIM_H

#include "osqp_api_types.h"

#define osqp_setup b200qp_setup                               /* bindings.cpp.in:153 */
#define osqp_solve b200qp_solve                               /* :198 */
#define osqp_warm_start b200qp_warm_start                     /* :193 */
#define osqp_update_data_vec b200qp_update_data_vec           /* :237 */
#define osqp_update_data_mat b200qp_update_data_mat           /* :280 */
#define osqp_update_settings b200qp_update_settings           /* :204 */
#define osqp_update_rho b200qp_update_rho                     /* :213 */
#define osqp_cleanup b200qp_cleanup                           /* :160 */
#define osqp_set_default_settings b200qp_set_default_settings /* :449 */
#define osqp_capabilities b200qp_capabilities                 /* :402 */
#define osqp_adjoint_derivative_compute b200qp_adjoint_derivative_compute /* :301 */
#define osqp_adjoint_derivative_get_mat b200qp_adjoint_derivative_get_mat /* :309 */
#define osqp_adjoint_derivative_get_vec b200qp_adjoint_derivative_get_vec /* :318 */

/* code generation: not part of the engine */
static inline void osqp_set_default_codegen_defines(OSQPCodegenDefines* d) { /* :463 */
  if (!d) return;
  d->embedded_mode = 1; d->float_type = 0; d->printing_enable = 0; d->profiling_enable = 0;
  d->interrupt_enable = 0; d->derivatives_enable = 0;
}
static inline OSQPInt osqp_codegen(OSQPSolver* s, const char* dir, const char* prefix, OSQPCodegenDefines* d) { /* :322 */
  (void)s; (void)dir; (void)prefix; (void)d;
  return OSQP_FUNC_NOT_IMPLEMENTED;
}

#endif
