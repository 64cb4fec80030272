// bindings.cpp -- pybind11 module `osqp_cuda`: the extension module the reference's Python facade
// loads for algebra='cuda' (reference: src/osqp/interface.py:20-24 maps 'cuda' -> 'osqp_cuda';
// the surface it consumes is listed in SURVEY.md 8b and comes from src/bindings.cpp.in:325-514).
// Everything here is marshalling; all numerics are behind the C ABI of include/b200qp.h.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include <cstring>

#include "../../include/b200qp.h"

namespace py = pybind11;
using namespace pybind11::literals;
using darray = py::array_t<b200qp_float, py::array::c_style | py::array::forcecast>;
using iarray = py::array_t<b200qp_int, py::array::c_style | py::array::forcecast>;

namespace {

// scipy CSC -> b200qp_csc view; keeps the (possibly converted) index/value arrays alive
struct CscHandle {
  iarray p, i;
  darray x;
  b200qp_int m = 0, n = 0, nzmax = 0, nz = -1;
  b200qp_csc view{};

  explicit CscHandle(py::object mat) {
    py::module_ sp = py::module_::import("scipy.sparse");
    if (!sp.attr("isspmatrix_csc")(mat).cast<bool>()) mat = sp.attr("csc_matrix")(mat);
    py::tuple shape = mat.attr("shape");
    m = shape[0].cast<b200qp_int>();
    n = shape[1].cast<b200qp_int>();
    p = iarray::ensure(mat.attr("indptr"));
    i = iarray::ensure(mat.attr("indices"));
    x = darray::ensure(mat.attr("data"));
    if (!p || !i || !x) throw py::type_error("CSC: indptr/indices/data must be convertible to int32/int32/float64");
    nzmax = mat.attr("nnz").cast<b200qp_int>();
    view.m = m; view.n = n; view.nzmax = nzmax; view.nz = -1;
    view.p = const_cast<b200qp_int*>(p.data());
    view.i = const_cast<b200qp_int*>(i.data());
    view.x = const_cast<b200qp_float*>(x.data());
  }
};

struct SolutionView {
  b200qp_solver* s;
  b200qp_int n, m;
  // fresh copies on every access, like the reference getters (src/bindings.cpp.in:79-105)
  py::array_t<b200qp_float> copy(const b200qp_float* src, b200qp_int len) const {
    py::array_t<b200qp_float> out(len);
    if (len) std::memcpy(out.mutable_data(), src, sizeof(b200qp_float) * (size_t)len);
    return out;
  }
  py::array_t<b200qp_float> x() const { return copy(s->solution->x, n); }
  py::array_t<b200qp_float> y() const { return copy(s->solution->y, m); }
  py::array_t<b200qp_float> prim_inf_cert() const { return copy(s->solution->prim_inf_cert, m); }
  py::array_t<b200qp_float> dual_inf_cert() const { return copy(s->solution->dual_inf_cert, n); }
};

const b200qp_float* opt_doubles(const py::object& o, darray& keep, py::ssize_t* len = nullptr) {
  if (o.is_none()) return nullptr;
  keep = darray::ensure(o);
  if (!keep) throw py::type_error("expected an array of float64");
  if (len) *len = keep.size();
  return keep.data();
}
const b200qp_int* opt_ints(const py::object& o, iarray& keep, py::ssize_t* len = nullptr) {
  if (o.is_none()) return nullptr;
  keep = iarray::ensure(o);
  if (!keep) throw py::type_error("expected an array of int32");
  if (len) *len = keep.size();
  return keep.data();
}

class Solver {
 public:
  Solver(const CscHandle& P, const py::array_t<b200qp_float>& q, const CscHandle& A,
         const py::array_t<b200qp_float>& l, const py::array_t<b200qp_float>& u, b200qp_int m, b200qp_int n,
         const b200qp_settings* settings, b200qp_int rank = 0, b200qp_int world = 0, b200qp_int ranks_on_device = 1)
      : m_(m), n_(n) {
    darray qc = darray::ensure(q), lc = darray::ensure(l), uc = darray::ensure(u);
    if (qc.size() != n || lc.size() != m || uc.size() != m) throw py::value_error(std::to_string(B200QP_DATA_VALIDATION_ERROR));
    b200qp_int status;
    {
      py::gil_scoped_release nogil;
      if (world >= 1)   // multi-GPU rank of a row-partitioned QP (world == 0: the reference's single-GPU solver)
        status = b200qp_dist_setup(&s_, rank, world, ranks_on_device, &P.view, qc.data(), &A.view, lc.data(), uc.data(), m, n, settings);
      else
        status = b200qp_setup(&s_, &P.view, qc.data(), &A.view, lc.data(), uc.data(), m, n, settings);
    }
    // nonzero status -> ValueError(str(code)); the facade turns it into OSQPException (interface.py:128-141)
    if (status) throw py::value_error(std::to_string(status));
    sol_ = SolutionView{s_, n, m};
  }
  ~Solver() {
    if (s_) b200qp_cleanup(s_);
  }
  Solver(const Solver&) = delete;
  Solver& operator=(const Solver&) = delete;

  b200qp_int solve() {
    py::gil_scoped_release nogil;   // reference: src/bindings.cpp.in:196-201
    return b200qp_solve(s_);
  }
  b200qp_int warm_start(py::object x, py::object y) {
    darray kx, ky;
    py::ssize_t lx = n_, ly = m_;
    const b200qp_float* px = opt_doubles(x, kx, &lx);
    const b200qp_float* pyv = opt_doubles(y, ky, &ly);
    // the C ABI copies n / m doubles from these pointers: a wrong length must not reach it
    if (lx != n_ || ly != m_) throw py::value_error(std::to_string(B200QP_DATA_VALIDATION_ERROR));
    return b200qp_warm_start(s_, px, pyv);
  }
  b200qp_int update_data_vec(py::object q, py::object l, py::object u) {
    darray kq, kl, ku;
    py::ssize_t lq = n_, ll = m_, lu = m_;
    const b200qp_float* pq = opt_doubles(q, kq, &lq);
    const b200qp_float* pl = opt_doubles(l, kl, &ll);
    const b200qp_float* pu = opt_doubles(u, ku, &lu);
    if (lq != n_ || ll != m_ || lu != m_) throw py::value_error(std::to_string(B200QP_DATA_VALIDATION_ERROR));
    return b200qp_update_data_vec(s_, pq, pl, pu);
  }
  b200qp_int update_data_mat(py::object Px, py::object Pi, py::object Ax, py::object Ai) {
    darray kpx, kax;
    iarray kpi, kai;
    py::ssize_t np_ = 0, na_ = 0;
    const b200qp_float* ppx = opt_doubles(Px, kpx, &np_);
    const b200qp_int* ppi = opt_ints(Pi, kpi, &np_);
    const b200qp_float* pax = opt_doubles(Ax, kax, &na_);
    const b200qp_int* pai = opt_ints(Ai, kai, &na_);
    return b200qp_update_data_mat(s_, ppx, ppi, (b200qp_int)np_, pax, pai, (b200qp_int)na_);
  }
  b200qp_int update_settings(const b200qp_settings& st) {
    const b200qp_int status = b200qp_update_settings(s_, &st);
    if (status) throw py::value_error(std::to_string(status));
    return status;
  }
  b200qp_int update_rho(b200qp_float rho) { return b200qp_update_rho(s_, rho); }
  // adjoint derivatives (reference: src/bindings.cpp.in:283-319)
  b200qp_int adjoint_derivative_compute(py::object dx, py::object dy) {
    darray kx, ky;
    const b200qp_float* px = opt_doubles(dx, kx);
    const b200qp_float* pyv = opt_doubles(dy, ky);
    if ((px && kx.size() != n_) || (pyv && ky.size() != m_)) throw py::value_error(std::to_string(B200QP_DATA_VALIDATION_ERROR));
    py::gil_scoped_release nogil;
    return b200qp_adjoint_derivative_compute(s_, px, pyv);
  }
  b200qp_int adjoint_derivative_get_mat(CscHandle& dP, CscHandle& dA) {
    // the handles own writable copies of the patterns; their x arrays are overwritten in place
    dP.view.x = dP.x.mutable_data();
    dA.view.x = dA.x.mutable_data();
    return b200qp_adjoint_derivative_get_mat(s_, &dP.view, &dA.view);
  }
  b200qp_int adjoint_derivative_get_vec(py::array_t<b200qp_float, py::array::c_style> dq,
                                        py::array_t<b200qp_float, py::array::c_style> dl,
                                        py::array_t<b200qp_float, py::array::c_style> du) {
    if (dq.size() != n_ || dl.size() != m_ || du.size() != m_) throw py::value_error(std::to_string(B200QP_DATA_VALIDATION_ERROR));
    return b200qp_adjoint_derivative_get_vec(s_, dq.mutable_data(), dl.mutable_data(), du.mutable_data());
  }
  b200qp_settings* settings() { return s_->settings; }
  b200qp_info* info() { return s_->info; }
  SolutionView& solution() { return sol_; }
  b200qp_solver* raw() const { return s_; }
  py::bytes dist_handle() {
    char h[64];
    const b200qp_int rc = b200qp_dist_get_handle(s_, h);
    if (rc) throw py::value_error(std::to_string(rc));
    return py::bytes(h, 64);
  }
  void dist_connect_ipc(const std::string& handles) {
    const b200qp_int rc = b200qp_dist_connect_ipc(s_, handles.data());
    if (rc) throw py::value_error(std::to_string(rc));
  }
  py::tuple dist_row_range() const {
    b200qp_int lo = 0, hi = 0;
    b200qp_dist_row_range(s_, &lo, &hi);
    return py::make_tuple(lo, hi);
  }
  py::dict stats() const {
    b200qp_stats st{};
    b200qp_get_stats(s_, &st);
    py::list ph;
    for (int i = 0; i < 8; ++i) ph.append(st.phase_ms[i]);
    return py::dict("phase_ms"_a = ph, "admm_iters"_a = st.admm_iters, "cg_iters"_a = st.cg_iters, "grid_barriers"_a = st.grid_barriers,
                    "kernel_ms"_a = st.kernel_ms, "algo_bytes"_a = st.algo_bytes, "grid"_a = st.grid,
                    "block"_a = st.block, "n"_a = st.n, "m"_a = st.m, "nnzP_full"_a = st.nnzP_full, "nnzA"_a = st.nnzA,
                    "algo_bytes_engine"_a = st.algo_bytes_engine, "spmv_bytes_A"_a = st.spmv_bytes[0],
                    "spmv_bytes_K"_a = st.spmv_bytes[1], "kernel_kind"_a = st.kernel_kind,
                    "ring_stages"_a = st.ring_stages, "ring_stage_bytes"_a = st.ring_stage_bytes,
                    "autotune_ms"_a = py::make_tuple(st.autotune_ms[0], st.autotune_ms[1]));
  }

 private:
  b200qp_int m_, n_;
  b200qp_solver* s_ = nullptr;
  SolutionView sol_{};
};

// batched solver: `batch` QPs sharing (P, A) (include/b200qp.h: b200qp_batch_*); replaces the reference's
// thread-pool loop over solver objects (src/osqp/nn/torch.py:128-164, 200-217)
class BatchSolver {
 public:
  BatchSolver(const CscHandle& P, py::object q, const CscHandle& A, py::object l, py::object u, b200qp_int batch,
              b200qp_int m, b200qp_int n, const b200qp_settings* settings)
      : batch_(batch), m_(m), n_(n), nnzP_(P.nzmax), nnzA_(A.nzmax) {
    darray qc = darray::ensure(q), lc = darray::ensure(l), uc = darray::ensure(u);
    if (!qc || !lc || !uc || batch <= 0 || qc.size() != (py::ssize_t)batch * n || lc.size() != (py::ssize_t)batch * m ||
        uc.size() != (py::ssize_t)batch * m)
      throw py::value_error(std::to_string(B200QP_DATA_VALIDATION_ERROR));
    b200qp_int status;
    {
      py::gil_scoped_release nogil;
      status = b200qp_batch_setup(&b_, batch, &P.view, qc.data(), &A.view, lc.data(), uc.data(), m, n, settings);
    }
    if (status) throw py::value_error(std::to_string(status));
  }
  ~BatchSolver() {
    if (b_) b200qp_batch_cleanup(b_);
  }
  BatchSolver(const BatchSolver&) = delete;
  BatchSolver& operator=(const BatchSolver&) = delete;

  b200qp_int solve() {
    py::gil_scoped_release nogil;
    return b200qp_batch_solve(b_);
  }
  const b200qp_float* opt_batch(const py::object& o, darray& keep, py::ssize_t len) const {
    const b200qp_float* p = opt_doubles(o, keep);
    if (p && keep.size() != (py::ssize_t)batch_ * len) throw py::value_error(std::to_string(B200QP_DATA_VALIDATION_ERROR));
    return p;
  }
  b200qp_int update_data_vec(py::object q, py::object l, py::object u) {
    darray kq, kl, ku;
    const b200qp_float* pq = opt_batch(q, kq, n_);
    const b200qp_float* pl = opt_batch(l, kl, m_);
    const b200qp_float* pu = opt_batch(u, ku, m_);
    py::gil_scoped_release nogil;
    return b200qp_batch_update_data_vec(b_, pq, pl, pu);
  }
  b200qp_int update_data_mat(py::object Px, py::object Ax) {
    darray kp, ka;
    const b200qp_float* pp = opt_doubles(Px, kp);
    const b200qp_float* pa = opt_doubles(Ax, ka);
    if ((pp && (kp.ndim() != 2 || kp.shape(0) != batch_ || kp.shape(1) != nnzP_)) ||
        (pa && (ka.ndim() != 2 || ka.shape(0) != batch_ || ka.shape(1) != nnzA_)))
      throw py::value_error(std::to_string(B200QP_DATA_VALIDATION_ERROR));
    py::gil_scoped_release nogil;
    return b200qp_batch_update_data_mat(b_, pp, pa);
  }
  b200qp_int warm_start(py::object x, py::object y) {
    darray kx, ky;
    const b200qp_float* px = opt_batch(x, kx, n_);
    const b200qp_float* pyv = opt_batch(y, ky, m_);
    return b200qp_batch_warm_start(b_, px, pyv);
  }
  b200qp_int update_settings(const b200qp_settings& st) {
    const b200qp_int status = b200qp_batch_update_settings(b_, &st);
    if (status) throw py::value_error(std::to_string(status));
    return status;
  }
  b200qp_int update_rho(b200qp_float rho) { return b200qp_batch_update_rho(b_, rho); }
  py::dict solution(bool certificates) const {
    py::array_t<b200qp_float> x({(py::ssize_t)batch_, (py::ssize_t)n_}), y({(py::ssize_t)batch_, (py::ssize_t)m_});
    if (!certificates) {
      b200qp_batch_get_solution(b_, x.mutable_data(), y.mutable_data(), nullptr, nullptr);
      return py::dict("x"_a = x, "y"_a = y);
    }
    py::array_t<b200qp_float> pc({(py::ssize_t)batch_, (py::ssize_t)m_}), dc({(py::ssize_t)batch_, (py::ssize_t)n_});
    b200qp_batch_get_solution(b_, x.mutable_data(), y.mutable_data(), pc.mutable_data(), dc.mutable_data());
    return py::dict("x"_a = x, "y"_a = y, "prim_inf_cert"_a = pc, "dual_inf_cert"_a = dc);
  }
  py::dict info() const {
    std::vector<b200qp_batch_info> v((size_t)batch_);
    b200qp_batch_get_info(b_, v.data());
    py::array_t<b200qp_int> status(batch_), iter(batch_), rho_updates(batch_), cg_iter(batch_);
    py::array_t<b200qp_float> obj(batch_), dobj(batch_), pres(batch_), dres(batch_), rho(batch_), rho_est(batch_);
    for (b200qp_int i = 0; i < batch_; ++i) {
      status.mutable_at(i) = v[i].status_val; iter.mutable_at(i) = v[i].iter;
      rho_updates.mutable_at(i) = v[i].rho_updates; cg_iter.mutable_at(i) = v[i].cg_iter;
      obj.mutable_at(i) = v[i].obj_val; dobj.mutable_at(i) = v[i].dual_obj_val; pres.mutable_at(i) = v[i].prim_res;
      dres.mutable_at(i) = v[i].dual_res; rho.mutable_at(i) = v[i].rho; rho_est.mutable_at(i) = v[i].rho_estimate;
    }
    return py::dict("status_val"_a = status, "iter"_a = iter, "rho_updates"_a = rho_updates, "cg_iter"_a = cg_iter,
                    "obj_val"_a = obj, "dual_obj_val"_a = dobj, "prim_res"_a = pres, "dual_res"_a = dres, "rho"_a = rho,
                    "rho_estimate"_a = rho_est);
  }
  py::dict stats() const {
    b200qp_batch_stats st{};
    b200qp_batch_get_stats(b_, &st);
    return py::dict("admm_iters"_a = st.admm_iters, "cg_iters"_a = st.cg_iters, "kernel_ms"_a = st.kernel_ms,
                    "setup_time"_a = st.setup_time, "update_time"_a = st.update_time, "solve_time"_a = st.solve_time,
                    "grid"_a = st.grid, "block"_a = st.block, "smem_bytes"_a = st.smem_bytes, "batch"_a = st.batch,
                    "n"_a = st.n, "m"_a = st.m, "dense"_a = st.dense);
  }

 private:
  b200qp_int batch_, m_, n_, nnzP_ = 0, nnzA_ = 0;   // nnz of the triu-CSC P and of the CSC A given at setup
  b200qp_batch* b_ = nullptr;
};

}  // namespace

PYBIND11_MODULE(osqp_cuda, mod) {
  mod.doc() = "B200-native OSQP ADMM engine (extension module behind osqp.OSQP(algebra='cuda'))";
  mod.attr("OSQP_USE_FLOAT") = 0;
  mod.attr("OSQP_USE_LONG") = 0;
  mod.attr("OSQP_INFTY") = B200QP_INFTY;
  mod.attr("__version__") = b200qp_version();

  py::enum_<b200qp_linsys_solver_type>(mod, "osqp_linsys_solver_type", py::module_local())
      .value("OSQP_DIRECT_SOLVER", B200QP_DIRECT_SOLVER)
      .value("OSQP_INDIRECT_SOLVER", B200QP_INDIRECT_SOLVER)
      .export_values();
  py::enum_<b200qp_status_type>(mod, "osqp_status_type", py::module_local())
      .value("OSQP_SOLVED", B200QP_SOLVED)
      .value("OSQP_SOLVED_INACCURATE", B200QP_SOLVED_INACCURATE)
      .value("OSQP_PRIMAL_INFEASIBLE", B200QP_PRIMAL_INFEASIBLE)
      .value("OSQP_PRIMAL_INFEASIBLE_INACCURATE", B200QP_PRIMAL_INFEASIBLE_INACCURATE)
      .value("OSQP_DUAL_INFEASIBLE", B200QP_DUAL_INFEASIBLE)
      .value("OSQP_DUAL_INFEASIBLE_INACCURATE", B200QP_DUAL_INFEASIBLE_INACCURATE)
      .value("OSQP_MAX_ITER_REACHED", B200QP_MAX_ITER_REACHED)
      .value("OSQP_TIME_LIMIT_REACHED", B200QP_TIME_LIMIT_REACHED)
      .value("OSQP_NON_CVX", B200QP_NON_CVX)
      .value("OSQP_SIGINT", B200QP_SIGINT)
      .value("OSQP_UNSOLVED", B200QP_UNSOLVED)
      .export_values();
  py::enum_<b200qp_error_type>(mod, "osqp_error_type", py::module_local())
      .value("OSQP_NO_ERROR", B200QP_NO_ERROR)
      .value("OSQP_DATA_VALIDATION_ERROR", B200QP_DATA_VALIDATION_ERROR)
      .value("OSQP_SETTINGS_VALIDATION_ERROR", B200QP_SETTINGS_VALIDATION_ERROR)
      .value("OSQP_LINSYS_SOLVER_INIT_ERROR", B200QP_LINSYS_SOLVER_INIT_ERROR)
      .value("OSQP_NONCVX_ERROR", B200QP_NONCVX_ERROR)
      .value("OSQP_MEM_ALLOC_ERROR", B200QP_MEM_ALLOC_ERROR)
      .value("OSQP_WORKSPACE_NOT_INIT_ERROR", B200QP_WORKSPACE_NOT_INIT_ERROR)
      .value("OSQP_ALGEBRA_LOAD_ERROR", B200QP_ALGEBRA_LOAD_ERROR)
      .value("OSQP_CODEGEN_DEFINES_ERROR", B200QP_CODEGEN_DEFINES_ERROR)
      .value("OSQP_DATA_NOT_INITIALIZED", B200QP_DATA_NOT_INITIALIZED)
      .value("OSQP_FUNC_NOT_IMPLEMENTED", B200QP_FUNC_NOT_IMPLEMENTED);
  py::enum_<b200qp_precond_type>(mod, "osqp_precond_type", py::module_local())
      .value("OSQP_NO_PRECONDITIONER", B200QP_NO_PRECONDITIONER)
      .value("OSQP_DIAGONAL_PRECONDITIONER", B200QP_DIAGONAL_PRECONDITIONER)
      .export_values();
  py::enum_<b200qp_capabilities_type>(mod, "osqp_capabilities_type", py::module_local())
      .value("OSQP_CAPABILITY_DIRECT_SOLVER", B200QP_CAPABILITY_DIRECT_SOLVER)
      .value("OSQP_CAPABILITY_INDIRECT_SOLVER", B200QP_CAPABILITY_INDIRECT_SOLVER)
      .value("OSQP_CAPABILITY_CODEGEN", B200QP_CAPABILITY_CODEGEN)
      .value("OSQP_CAPABILITY_UPDATE_MATRICES", B200QP_CAPABILITY_UPDATE_MATRICES)
      .value("OSQP_CAPABILITY_DERIVATIVES", B200QP_CAPABILITY_DERIVATIVES);
  mod.def("osqp_capabilities", &b200qp_capabilities);
  mod.def("device_count", &b200qp_device_count);

  py::class_<CscHandle>(mod, "CSC", py::module_local())
      .def(py::init<py::object>())
      .def_readonly("m", &CscHandle::m)
      .def_readonly("n", &CscHandle::n)
      .def_readonly("p", &CscHandle::p)
      .def_readonly("i", &CscHandle::i)
      .def_readonly("x", &CscHandle::x)
      .def_readonly("nzmax", &CscHandle::nzmax)
      .def_readonly("nz", &CscHandle::nz);

  // field discovery by the facade iterates the class __dict__ (interface.py:318-322): real attributes
  py::class_<b200qp_settings>(mod, "OSQPSettings", py::module_local())
      .def(py::init([]() {
        auto* s = new b200qp_settings();
        b200qp_set_default_settings(s);
        return s;
      }))
      .def_readwrite("device", &b200qp_settings::device)
      .def_property(
          "linsys_solver", [](const b200qp_settings& s) { return (b200qp_linsys_solver_type)s.linsys_solver; },
          [](b200qp_settings& s, b200qp_linsys_solver_type v) { s.linsys_solver = (b200qp_int)v; })
      .def_readwrite("verbose", &b200qp_settings::verbose)
      .def_readwrite("warm_starting", &b200qp_settings::warm_starting)
      .def_readwrite("scaling", &b200qp_settings::scaling)
      .def_readwrite("polishing", &b200qp_settings::polishing)
      .def_readwrite("rho", &b200qp_settings::rho)
      .def_readwrite("rho_is_vec", &b200qp_settings::rho_is_vec)
      .def_readwrite("sigma", &b200qp_settings::sigma)
      .def_readwrite("alpha", &b200qp_settings::alpha)
      .def_readwrite("cg_max_iter", &b200qp_settings::cg_max_iter)
      .def_readwrite("cg_tol_reduction", &b200qp_settings::cg_tol_reduction)
      .def_readwrite("cg_tol_fraction", &b200qp_settings::cg_tol_fraction)
      .def_property(
          "cg_precond", [](const b200qp_settings& s) { return (b200qp_precond_type)s.cg_precond; },
          [](b200qp_settings& s, b200qp_precond_type v) { s.cg_precond = (b200qp_int)v; })
      .def_readwrite("adaptive_rho", &b200qp_settings::adaptive_rho)
      .def_readwrite("adaptive_rho_interval", &b200qp_settings::adaptive_rho_interval)
      .def_readwrite("adaptive_rho_fraction", &b200qp_settings::adaptive_rho_fraction)
      .def_readwrite("adaptive_rho_tolerance", &b200qp_settings::adaptive_rho_tolerance)
      .def_readwrite("max_iter", &b200qp_settings::max_iter)
      .def_readwrite("eps_abs", &b200qp_settings::eps_abs)
      .def_readwrite("eps_rel", &b200qp_settings::eps_rel)
      .def_readwrite("eps_prim_inf", &b200qp_settings::eps_prim_inf)
      .def_readwrite("eps_dual_inf", &b200qp_settings::eps_dual_inf)
      .def_readwrite("scaled_termination", &b200qp_settings::scaled_termination)
      .def_readwrite("check_termination", &b200qp_settings::check_termination)
      .def_readwrite("check_dualgap", &b200qp_settings::check_dualgap)
      .def_readwrite("time_limit", &b200qp_settings::time_limit)
      .def_readwrite("delta", &b200qp_settings::delta)
      .def_readwrite("polish_refine_iter", &b200qp_settings::polish_refine_iter);
  mod.def("osqp_set_default_settings", &b200qp_set_default_settings);

  py::class_<SolutionView>(mod, "OSQPSolution", py::module_local())
      .def_property_readonly("x", &SolutionView::x)
      .def_property_readonly("y", &SolutionView::y)
      .def_property_readonly("prim_inf_cert", &SolutionView::prim_inf_cert)
      .def_property_readonly("dual_inf_cert", &SolutionView::dual_inf_cert);

  py::class_<b200qp_info>(mod, "OSQPInfo", py::module_local())
      .def_property_readonly("status", [](const b200qp_info& i) { return std::string(i.status); })
      .def_readonly("status_val", &b200qp_info::status_val)
      .def_readonly("status_polish", &b200qp_info::status_polish)
      .def_readwrite("obj_val", &b200qp_info::obj_val)   // the facade overwrites it for NON_CVX (interface.py:415)
      .def_readwrite("dual_obj_val", &b200qp_info::dual_obj_val)
      .def_readonly("prim_res", &b200qp_info::prim_res)
      .def_readonly("dual_res", &b200qp_info::dual_res)
      .def_readonly("duality_gap", &b200qp_info::duality_gap)
      .def_readonly("iter", &b200qp_info::iter)
      .def_readonly("rho_updates", &b200qp_info::rho_updates)
      .def_readonly("rho_estimate", &b200qp_info::rho_estimate)
      .def_readonly("setup_time", &b200qp_info::setup_time)
      .def_readonly("solve_time", &b200qp_info::solve_time)
      .def_readonly("update_time", &b200qp_info::update_time)
      .def_readonly("polish_time", &b200qp_info::polish_time)
      .def_readonly("run_time", &b200qp_info::run_time)
      .def_readonly("primdual_int", &b200qp_info::primdual_int)
      .def_readonly("rel_kkt_error", &b200qp_info::rel_kkt_error);

  py::class_<Solver>(mod, "OSQPSolver", py::module_local())
      .def(py::init<const CscHandle&, const py::array_t<b200qp_float>&, const CscHandle&,
                    const py::array_t<b200qp_float>&, const py::array_t<b200qp_float>&, b200qp_int, b200qp_int,
                    const b200qp_settings*, b200qp_int, b200qp_int, b200qp_int>(),
           "P"_a, "q"_a.noconvert(), "A"_a, "l"_a.noconvert(), "u"_a.noconvert(), "m"_a, "n"_a, "settings"_a,
           "rank"_a = 0, "world"_a = 0, "ranks_on_device"_a = 1)
      .def("dist_handle", &Solver::dist_handle)
      .def("dist_connect_ipc", &Solver::dist_connect_ipc)
      .def("dist_row_range", &Solver::dist_row_range)
      .def_property_readonly("solution", &Solver::solution, py::return_value_policy::reference_internal)
      .def_property_readonly("info", &Solver::info, py::return_value_policy::reference_internal)
      .def("warm_start", &Solver::warm_start, "x"_a.none(true), "y"_a.none(true))
      .def("solve", &Solver::solve)
      .def("update_data_vec", &Solver::update_data_vec, "q"_a.none(true), "l"_a.none(true), "u"_a.none(true))
      .def("update_data_mat", &Solver::update_data_mat, "P_x"_a.none(true), "P_i"_a.none(true), "A_x"_a.none(true),
           "A_i"_a.none(true))
      .def("update_settings", &Solver::update_settings)
      .def("update_rho", &Solver::update_rho)
      .def("adjoint_derivative_compute", &Solver::adjoint_derivative_compute, "dx"_a.none(true), "dy"_a.none(true))
      .def("adjoint_derivative_get_mat", &Solver::adjoint_derivative_get_mat, "dP"_a, "dA"_a)
      .def("adjoint_derivative_get_vec", &Solver::adjoint_derivative_get_vec, "dq"_a.noconvert(), "dl"_a.noconvert(), "du"_a.noconvert())
      .def("get_settings", &Solver::settings, py::return_value_policy::reference_internal)
      .def("get_stats", &Solver::stats);

  py::class_<BatchSolver>(mod, "OSQPBatchSolver", py::module_local())
      .def(py::init<const CscHandle&, py::object, const CscHandle&, py::object, py::object, b200qp_int, b200qp_int,
                    b200qp_int, const b200qp_settings*>(),
           "P"_a, "q"_a, "A"_a, "l"_a, "u"_a, "batch"_a, "m"_a, "n"_a, "settings"_a)
      .def("solve", &BatchSolver::solve)
      .def("update_data_vec", &BatchSolver::update_data_vec, "q"_a.none(true) = py::none(), "l"_a.none(true) = py::none(),
           "u"_a.none(true) = py::none())
      .def("update_data_mat", &BatchSolver::update_data_mat, "P_x"_a.none(true) = py::none(), "A_x"_a.none(true) = py::none())
      .def("warm_start", &BatchSolver::warm_start, "x"_a.none(true) = py::none(), "y"_a.none(true) = py::none())
      .def("update_settings", &BatchSolver::update_settings)
      .def("update_rho", &BatchSolver::update_rho)
      .def("get_solution", &BatchSolver::solution, "certificates"_a = true)
      .def("get_info", &BatchSolver::info)
      .def("get_stats", &BatchSolver::stats);

  // one process driving every rank of a row-partitioned QP (include/b200qp.h: b200qp_dist_*)
  mod.def("dist_connect_local", [](py::list solvers) {
    std::vector<b200qp_solver*> grp;
    for (auto h : solvers) grp.push_back(h.cast<Solver&>().raw());
    const b200qp_int rc = b200qp_dist_connect_local(grp.data(), (b200qp_int)grp.size());
    if (rc) throw py::value_error(std::to_string(rc));
  });
  mod.def("dist_solve_group", [](py::list solvers) {
    std::vector<b200qp_solver*> grp;
    for (auto h : solvers) grp.push_back(h.cast<Solver&>().raw());
    b200qp_int rc;
    {
      py::gil_scoped_release nogil;
      rc = b200qp_dist_solve_group(grp.data(), (b200qp_int)grp.size());
    }
    return rc;
  });
}
