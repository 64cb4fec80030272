// internal_view.h -- library-internal window into a solver's device workspace, shared by the translation
// units of libb200qp.so (b200qp.cu owns the workspace, b200qp_batch.cu borrows the scaled matrices and the
// equilibration vectors of its prototype solver).  Not part of include/b200qp.h.
#pragma once
#include "../../include/b200qp.h"
#include "device_common.cuh"

namespace b200qp {
struct InternalView {
  int n, m, device;
  CsrMat P, A, At;                          // scaled: c D P D (full symmetric CSR), E A D, D A' E
  const double *pdiag, *D, *Dinv, *E, *Einv;
  double c, rho_eq_factor;
  // value-update maps (b200qp_update_data_mat): triu-CSC(P) position -> both CSR(P) positions, CSC(A) -> CSR(A)
  const int *mapA, *Pt_colptr, *Pt_rowidx, *posU, *posL;
  int nnzPt;
};
}  // namespace b200qp

extern "C" int b200qp_internal_view(const b200qp_solver* solver, b200qp::InternalView* view);
