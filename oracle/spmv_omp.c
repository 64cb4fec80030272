/*
 * spmv_omp.c -- TEST INFRASTRUCTURE (CPU oracle), NOT PRODUCT.
 *
 * The hot operator of the oracle's engine-spec (indirect) mode, y = M x for a CSR matrix, parallelised over the
 * host cores with OpenMP so that the CPU baseline / `bench.py --impl reference` arm uses every host thread it
 * can (scipy's csr_matvec holds the GIL and runs on one core).  Same arithmetic as scipy: rows summed in index
 * order.  Reference counterpart: the SpMVs of the KKT solve, src/osqppurepy/_osqp.py:644-658 (P.dot / A.dot).
 * Built by osqp-python_b200/build.py into oracle/_build/liboracle_spmv.so; only oracle/osqp_oracle.py loads it.
 */
#include <omp.h>

void oracle_csr_matvec(int rows, const int* ptr, const int* idx, const double* val, const double* x, double* y) {
#pragma omp parallel for schedule(static)
  for (int r = 0; r < rows; ++r) {
    double acc = 0.0;
    for (int k = ptr[r]; k < ptr[r + 1]; ++k) acc += val[k] * x[idx[k]];
    y[r] = acc;
  }
}

int oracle_omp_threads(void) { return omp_get_max_threads(); }
void oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
