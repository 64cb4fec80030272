// b200qp_csl.cu -- translation unit of the compacted-slice / TMA-ring ADMM kernel (compiled side by side with b200qp.cu).
#include "admm_kernel_csl.cuh"

const void* b200qp_csl_kernel() { return (const void*)b200qp::admm_kernel_csl; }
