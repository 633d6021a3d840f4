// c_rigid.cpp -- pybind11 host class `CManyBodies`, the drop-in for the reference's
// nanobind class of the same name (/root/reference/src/c_rigid_obj.cpp:997-1027; nanobind
// is not available here).  Same module-level names, same 14 methods +
// `precision`, same argument meaning and error behaviour; the arithmetic happens on the
// GPU behind the C ABI of include/rbl.h.  Built twice: -DDOUBLE_PRECISION selects `real`
// exactly like the reference's eigen_defines.h:5-37.
//
// Extensions beyond the reference's bound surface are grouped at the end of the class and
// marked EXT (fused apply_saddle, Kinv products, GMRES / Lanczos drivers, ctx handle).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <cstdint>
#include <stdexcept>
#include <string>

#include "../../include/rbl.h"

namespace py = pybind11;

#ifdef DOUBLE_PRECISION
using real = double;
static constexpr int kPrecision = RBL_F64;
static constexpr const char* kPrecisionName = "double";
#else
using real = float;
static constexpr int kPrecision = RBL_F32;
static constexpr const char* kPrecisionName = "single";
#endif

#ifndef RBL_MODULE_NAME
#error "compile with -DRBL_MODULE_NAME=<python module name>"
#endif

// numpy input of either float width, converted to `real` (tests/test_precision.py feeds both)
using Arr = py::array_t<real, py::array::c_style | py::array::forcecast>;

class CManyBodies {
  rbl_ctx* ctx_ = nullptr;

  void check(int status) const {
    if (status != RBL_OK) throw std::runtime_error(rbl_last_error(ctx_));
  }
  py::ssize_t n3() const { return 3 * (py::ssize_t)rbl_n_bodies(ctx_) * rbl_blobs_per_body(ctx_); }
  py::ssize_t n6() const { return 6 * (py::ssize_t)rbl_n_bodies(ctx_); }
  static void want(const Arr& a, py::ssize_t n, const char* what) {
    if (a.size() != n)
      throw std::runtime_error(std::string(what) + ": expected " + std::to_string(n) + " values, got " +
                               std::to_string(a.size()));
  }

 public:
  CManyBodies() {
    int st = rbl_create(kPrecision, -1, &ctx_);
    if (st != RBL_OK) throw std::runtime_error(rbl_last_error(nullptr));
  }
  ~CManyBodies() { rbl_destroy(ctx_); }
  CManyBodies(const CManyBodies&) = delete;
  CManyBodies& operator=(const CManyBodies&) = delete;

  // ---- the reference's bound surface (c_rigid_obj.cpp:1001-1026) -------------------------
  void setParameters(real a, real dt, real kBT, real eta, Arr cfg) {
    if (cfg.size() == 0 || cfg.size() % 3 != 0) throw std::runtime_error("Rigid config must have length 3N");
    check(rbl_set_parameters(ctx_, a, dt, kBT, eta, cfg.data(), (int)(cfg.size() / 3)));
  }
  void setBlkPC(bool v) { check(rbl_set_flags(ctx_, v ? 1 : 0, -1)); }
  void setWallPC(bool v) { check(rbl_set_flags(ctx_, -1, v ? 1 : 0)); }
  void setConfig(Arr X, Arr Q) {
    if (X.size() % 3 != 0 || Q.size() != 4 * (X.size() / 3))
      throw std::runtime_error("setConfig: X must have 3*N_bodies and Q 4*N_bodies entries");
    check(rbl_set_config(ctx_, X.data(), Q.data(), (int)(X.size() / 3)));
  }
  py::tuple getConfig() {
    py::array_t<real> X(n6() / 2), Q(4 * (n6() / 6));
    check(rbl_get_config(ctx_, X.mutable_data(), Q.mutable_data()));
    return py::make_tuple(X, Q);
  }
  void set_K_mats() { check(rbl_set_K_mats(ctx_)); }
  py::array_t<real> K_x_U(Arr U) {
    want(U, n6(), "K_x_U");
    py::array_t<real> out(n3());
    check(rbl_K_dot(ctx_, U.data(), out.mutable_data()));
    return out;
  }
  py::array_t<real> KT_x_Lam(Arr lam) {
    want(lam, n3(), "KT_x_Lam");
    py::array_t<real> out(n6());
    check(rbl_KT_dot(ctx_, lam.data(), out.mutable_data()));
    return out;
  }
  // the reference returns a Python list of 3N floats; an ndarray is accepted by every caller
  // (Rigid.py wraps it in np.array) and avoids boxing 3N objects
  py::array_t<real> multi_body_pos() {
    py::array_t<real> out(n3());
    check(rbl_blob_positions(ctx_, out.mutable_data()));
    return out;
  }
  py::array_t<real> apply_PC(Arr in) {
    want(in, n3() + n6(), "apply_PC");
    py::array_t<real> out(n3() + n6());
    check(rbl_apply_PC(ctx_, in.data(), out.mutable_data()));
    return out;
  }
  py::object sparse(bool inverse) {
    const py::ssize_t rows = inverse ? n6() : n3(), cols = inverse ? n3() : n6();
    const py::ssize_t nnz = inverse ? 4 * n3() : 3 * n3();
    py::array_t<int64_t> indptr(cols + 1);
    py::array_t<int32_t> indices(nnz);
    py::array_t<real> data(nnz);
    check(inverse ? rbl_export_Kinv_csc(ctx_, indptr.mutable_data(), indices.mutable_data(), data.mutable_data())
                  : rbl_export_K_csc(ctx_, indptr.mutable_data(), indices.mutable_data(), data.mutable_data()));
    py::object csc = py::module_::import("scipy.sparse").attr("csc_matrix");
    return csc(py::make_tuple(data, indices, indptr), py::arg("shape") = py::make_tuple(rows, cols));
  }
  py::object get_K() { return sparse(false); }
  py::object get_Kinv() { return sparse(true); }
  void evolve_X_Q(Arr U) {
    want(U, n6(), "evolve_X_Q");
    check(rbl_evolve(ctx_, U.data()));
  }
  py::array_t<real> apply_M(Arr F, Arr r_vecs) {
    if (F.size() != r_vecs.size() || F.size() % 3 != 0)
      throw std::runtime_error("apply_M: F and r_vecs must both have 3*N_blobs entries");
    py::array_t<real> out(F.size());
    check(rbl_apply_M(ctx_, F.data(), r_vecs.data(), (int)(F.size() / 3), out.mutable_data()));
    return out;
  }

  // ---- EXT: not bound by the reference -------------------------------------------------------
  py::array_t<real> apply_saddle(Arr x) {  // Rigid.py:73-80 fused on the device
    want(x, n3() + n6(), "apply_saddle");
    py::array_t<real> out(n3() + n6());
    check(rbl_apply_saddle(ctx_, x.data(), out.mutable_data()));
    return out;
  }
  py::array_t<real> Kinv_x_V(Arr V) {  // :406
    want(V, n3(), "Kinv_x_V");
    py::array_t<real> out(n6());
    check(rbl_Kinv_dot(ctx_, V.data(), out.mutable_data()));
    return out;
  }
  py::array_t<real> KTinv_x_F(Arr F) {  // :408
    want(F, n6(), "KTinv_x_F");
    py::array_t<real> out(n3());
    check(rbl_KTinv_dot(ctx_, F.data(), out.mutable_data()));
    return out;
  }
  py::tuple gmres(Arr rhs, double tol, int restart, int max_iter) {
    want(rhs, n3() + n6(), "gmres");
    py::array_t<real> x(n3() + n6());
    int iters = 0;
    double relres = 0;
    {
      py::gil_scoped_release nogil;
      int st = rbl_gmres(ctx_, rhs.data(), x.mutable_data(), tol, restart, max_iter, &iters, &relres);
      py::gil_scoped_acquire gil;
      check(st);
    }
    return py::make_tuple(x, iters, relres);
  }
  py::tuple lanczos_sqrt(Arr W, double tol, int max_iter) {
    want(W, n3(), "lanczos_sqrt");
    py::array_t<real> out(n3());
    int iters = 0;
    check(rbl_lanczos_sqrt(ctx_, W.data(), out.mutable_data(), tol, max_iter, &iters));
    return py::make_tuple(out, iters);
  }
  py::tuple apply_M2(Arr F1, Arr F2, Arr r_vecs) {  // two right-hand sides in one pass over the pairs
    if (F1.size() != r_vecs.size() || F2.size() != r_vecs.size() || F1.size() % 3 != 0)
      throw std::runtime_error("apply_M2: F1, F2 and r_vecs must all have 3*N_blobs entries");
    py::array_t<real> o1(F1.size()), o2(F1.size());
    check(rbl_apply_M2(ctx_, F1.data(), F2.data(), r_vecs.data(), (int)(F1.size() / 3), o1.mutable_data(), o2.mutable_data()));
    return py::make_tuple(o1, o2);
  }
  py::tuple lanczos_sqrt2(Arr W1, Arr W2, double tol, int max_iter) {
    want(W1, n3(), "lanczos_sqrt2 W1");
    want(W2, n3(), "lanczos_sqrt2 W2");
    py::array_t<real> o1(n3()), o2(n3());
    int iters[2] = {0, 0};
    check(rbl_lanczos_sqrt2(ctx_, W1.data(), W2.data(), o1.mutable_data(), o2.mutable_data(), tol, max_iter, iters));
    return py::make_tuple(o1, o2, iters[0], iters[1]);
  }
  // one Brownian-dynamics step (noise supplied by the caller; None -> deterministic)
  py::tuple bd_step(Arr F_ext, py::object slip, py::object W1, py::object W2, py::object Wr, double kBT, double tol,
                    int restart, int max_iter, double ltol, int lmax) {
    want(F_ext, n6(), "bd_step F_ext");
    auto opt = [&](py::object o, const char* what, Arr& keep) -> const real* {
      if (o.is_none()) return nullptr;
      keep = Arr::ensure(o);
      if (!keep) throw std::runtime_error(std::string(what) + ": not convertible to a float array");
      want(keep, n3(), what);
      return keep.data();
    };
    Arr k0, k1, k2, k3;
    const real* ps = opt(slip, "bd_step slip", k0);
    const real* p1 = opt(W1, "bd_step W1", k1);
    const real* p2 = opt(W2, "bd_step W2", k2);
    const real* pr = opt(Wr, "bd_step Wr", k3);
    py::array_t<real> U(n6());
    int iters = 0;
    double relres = 0;
    check(rbl_bd_step(ctx_, F_ext.data(), ps, p1, p2, pr, kBT, tol, restart, max_iter, ltol, lmax, U.mutable_data(),
                      &iters, &relres));
    return py::make_tuple(U, iters, relres);
  }
  py::tuple bd_step_seeded(Arr F_ext, py::object slip, std::uint64_t seed, std::uint64_t step, double kBT, double tol,
                           int restart, int max_iter, double ltol, int lmax) {
    want(F_ext, n6(), "bd_step F_ext");
    Arr keep;
    const real* ps = nullptr;
    if (!slip.is_none()) {
      keep = Arr::ensure(slip);
      if (!keep) throw std::runtime_error("bd_step slip: not convertible to a float array");
      want(keep, n3(), "bd_step slip");
      ps = keep.data();
    }
    py::array_t<real> U(n6());
    int iters = 0;
    double relres = 0;
    check(rbl_bd_step_seeded(ctx_, F_ext.data(), ps, seed, step, kBT, tol, restart, max_iter, ltol, lmax, U.mutable_data(),
                             &iters, &relres));
    return py::make_tuple(U, iters, relres);
  }
  py::tuple normals(std::uint64_t seed, std::uint64_t step, std::uint64_t first, py::ssize_t n) {
    py::array_t<real> a(n), b(n), c(n);
    check(rbl_normals(ctx_, seed, step, first, (size_t)n, a.mutable_data(), b.mutable_data(), c.mutable_data()));
    return py::make_tuple(a, b, c);
  }
  void set_noise_preconditioner(int mode) { check(rbl_set_noise_preconditioner(ctx_, mode)); }
  void set_lanczos_pairing(bool on) { check(rbl_set_lanczos_pairing(ctx_, on ? 1 : 0)); }
  std::uintptr_t handle() const { return reinterpret_cast<std::uintptr_t>(ctx_); }
};

PYBIND11_MODULE(RBL_MODULE_NAME, m) {
  m.doc() = "Rigid code (B200-native drop-in for Rigid_Body_Light's c_rigid)";
  py::class_<CManyBodies>(m, "CManyBodies", py::module_local())  // both precisions load in one process
      .def(py::init<>())
      .def("getConfig", &CManyBodies::getConfig, "get the X and Q vectors for the current position")
      .def("setParameters", &CManyBodies::setParameters, "Set parameters for the module")
      .def("setBlkPC", &CManyBodies::setBlkPC, "set PC type")
      .def("setWallPC", &CManyBodies::setWallPC, "use wall corrections")
      .def("set_K_mats", &CManyBodies::set_K_mats, "Set the K,K^T,K^-1 matrices for the module")
      .def("K_x_U", &CManyBodies::K_x_U, "Multiply K by U", py::arg("U"))
      .def("KT_x_Lam", &CManyBodies::KT_x_Lam, "Multiply K^T by lambda", py::arg("lambda"))
      .def("multi_body_pos", &CManyBodies::multi_body_pos, "Get the blob positions")
      .def("apply_PC", &CManyBodies::apply_PC, "apply for PC")
      .def("setConfig", &CManyBodies::setConfig, "Set the X and Q vectors for the current position",
           py::arg("X"), py::arg("Q"))
      .def("get_K", &CManyBodies::get_K, "get K")
      .def("get_Kinv", &CManyBodies::get_Kinv, "get Kinv")
      .def("evolve_X_Q", &CManyBodies::evolve_X_Q, "evolve rigid bodies", py::arg("U"))
      .def("apply_M", &CManyBodies::apply_M, "mobility matrix mult", py::arg("F"), py::arg("r_vecs"))
      .def_property_readonly_static(
          "precision", [](py::object) { return std::string(kPrecisionName); },
          "Compilation precision, a string holding either single or double.")
      // EXT
      .def("apply_saddle", &CManyBodies::apply_saddle, "fused [M lam - K U ; K^T lam]", py::arg("x"))
      .def("Kinv_x_V", &CManyBodies::Kinv_x_V, py::arg("V"))
      .def("KTinv_x_F", &CManyBodies::KTinv_x_F, py::arg("F"))
      .def("gmres", &CManyBodies::gmres, py::arg("rhs"), py::arg("tol") = 1e-8, py::arg("restart") = 60,
           py::arg("max_iter") = 300)
      .def("lanczos_sqrt", &CManyBodies::lanczos_sqrt, py::arg("W"), py::arg("tol") = 1e-6,
           py::arg("max_iter") = 100)
      .def("apply_M2", &CManyBodies::apply_M2, py::arg("F1"), py::arg("F2"), py::arg("r_vecs"))
      .def("lanczos_sqrt2", &CManyBodies::lanczos_sqrt2, py::arg("W1"), py::arg("W2"), py::arg("tol") = 1e-6,
           py::arg("max_iter") = 100)
      .def("bd_step", &CManyBodies::bd_step, py::arg("F_ext"), py::arg("slip") = py::none(), py::arg("W1") = py::none(),
           py::arg("W2") = py::none(), py::arg("Wr") = py::none(), py::arg("kBT") = 0.0, py::arg("tol") = 1e-8,
           py::arg("restart") = 60, py::arg("max_iter") = 300, py::arg("lanczos_tol") = 1e-6,
           py::arg("lanczos_max_iter") = 100)
      .def("bd_step_seeded", &CManyBodies::bd_step_seeded, py::arg("F_ext"), py::arg("slip") = py::none(), py::arg("seed") = 0,
           py::arg("step") = 0, py::arg("kBT") = 0.0, py::arg("tol") = 1e-8, py::arg("restart") = 60, py::arg("max_iter") = 300,
           py::arg("lanczos_tol") = 1e-6, py::arg("lanczos_max_iter") = 100)
      .def("normals", &CManyBodies::normals, py::arg("seed"), py::arg("step"), py::arg("first"), py::arg("n"))
      .def("set_noise_preconditioner", &CManyBodies::set_noise_preconditioner, py::arg("mode"),
           "0: symmetric square root; 1: block-Cholesky preconditioned noise in bd_step (default); 2: also in lanczos_sqrt")
      .def("set_lanczos_pairing", &CManyBodies::set_lanczos_pairing, py::arg("on"))
      .def("handle", &CManyBodies::handle, "address of the rbl_ctx (for ctypes users of include/rbl.h)");
}
