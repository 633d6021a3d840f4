// pair_host.cpp -- compiles the kernel's per-pair arithmetic (rbl_pair.cuh) for the
// HOST so the division-free / branch-free reformulation can be checked against the
// oracle on a machine without a GPU (tests/test_pair_math_host.py).  Test code only:
// this is not a CPU fallback and is never loaded by the package.
#include <vector>

#include "../../rigid_body_light_b200/csrc/rbl_pair.cuh"

template <typename real>
static void matvec(const real* F, const real* r, int n, double a, double eta, int wall,
                   int near, real* U) {
  rbl::PairConsts<real> C = rbl::make_pair_consts<real>(a, eta);
  for (int i = 0; i < n; ++i) {
    real ux = 0, uy = 0, uz = 0;
    for (int j = 0; j < n; ++j) {
      real zj = r[3 * j + 2];
      real bj = wall ? (zj >= (real)a ? (real)1 : zj * C.inv_a) : (real)1;
      real fx = bj * F[3 * j], fy = bj * F[3 * j + 1], fz = bj * F[3 * j + 2];
      real z2 = 2 * zj, zz4 = 4 * zj * zj;
      if (wall) {
        if (near) rbl::pair<real, true, true>(C, r[3*i], r[3*i+1], r[3*i+2], r[3*j], r[3*j+1], zj, fx, fy, fz, z2, zz4, ux, uy, uz);
        else      rbl::pair<real, true, false>(C, r[3*i], r[3*i+1], r[3*i+2], r[3*j], r[3*j+1], zj, fx, fy, fz, z2, zz4, ux, uy, uz);
      } else {
        if (near) rbl::pair<real, false, true>(C, r[3*i], r[3*i+1], r[3*i+2], r[3*j], r[3*j+1], zj, fx, fy, fz, z2, zz4, ux, uy, uz);
        else      rbl::pair<real, false, false>(C, r[3*i], r[3*i+1], r[3*i+2], r[3*j], r[3*j+1], zj, fx, fy, fz, z2, zz4, ux, uy, uz);
      }
    }
    real zi = r[3 * i + 2];
    real bi = wall ? (zi >= (real)a ? (real)1 : zi * C.inv_a) : (real)1;
    U[3 * i] = ux * C.out_scale * bi;
    U[3 * i + 1] = uy * C.out_scale * bi;
    U[3 * i + 2] = uz * C.out_scale * bi;
  }
}

extern "C" void pair_matvec_host_f64(const double* F, const double* r, int n, double a, double eta, int wall, int near, double* U) { matvec<double>(F, r, n, a, eta, wall, near, U); }
extern "C" void pair_matvec_host_f32(const float* F, const float* r, int n, double a, double eta, int wall, int near, float* U) { matvec<float>(F, r, n, a, eta, wall, near, U); }

// targets and sources in separate arrays (no self pairs): lets the NEAR=false fast path,
// which is only legal when every pair has r >= 2a, be compared with the general path
template <typename real>
static void cross(const real* rt, int nt, const real* F, const real* r, int n, double a, double eta,
                  int wall, int near, real* U) {
  rbl::PairConsts<real> C = rbl::make_pair_consts<real>(a, eta);
  for (int i = 0; i < nt; ++i) {
    real ux = 0, uy = 0, uz = 0;
    for (int j = 0; j < n; ++j) {
      real zj = r[3 * j + 2];
      real fx = F[3 * j], fy = F[3 * j + 1], fz = F[3 * j + 2];
      real z2 = 2 * zj, zz4 = 4 * zj * zj;
      if (wall) {
        if (near) rbl::pair<real, true, true>(C, rt[3*i], rt[3*i+1], rt[3*i+2], r[3*j], r[3*j+1], zj, fx, fy, fz, z2, zz4, ux, uy, uz);
        else      rbl::pair<real, true, false>(C, rt[3*i], rt[3*i+1], rt[3*i+2], r[3*j], r[3*j+1], zj, fx, fy, fz, z2, zz4, ux, uy, uz);
      } else {
        if (near) rbl::pair<real, false, true>(C, rt[3*i], rt[3*i+1], rt[3*i+2], r[3*j], r[3*j+1], zj, fx, fy, fz, z2, zz4, ux, uy, uz);
        else      rbl::pair<real, false, false>(C, rt[3*i], rt[3*i+1], rt[3*i+2], r[3*j], r[3*j+1], zj, fx, fy, fz, z2, zz4, ux, uy, uz);
      }
    }
    U[3 * i] = ux; U[3 * i + 1] = uy; U[3 * i + 2] = uz;
  }
}
extern "C" void pair_cross_host_f64(const double* rt, int nt, const double* F, const double* r, int n, double a, double eta, int wall, int near, double* U) { cross<double>(rt, nt, F, r, n, a, eta, wall, near, U); }

// symmetric (unordered-pair) evaluation: every pair i<j once through pair_sym, the diagonal
// through the ordered general path -- the arithmetic of the symmetric CUDA kernel
template <typename real>
static void matvec_sym(const real* F, const real* r, int n, double a, double eta, int wall, int near, real* U) {
  rbl::PairConsts<real> C = rbl::make_pair_consts<real>(a, eta);
  std::vector<real> f(3 * (size_t)n), acc(3 * (size_t)n, 0);
  for (int j = 0; j < n; ++j) {
    real zj = r[3 * j + 2];
    real bj = wall ? (zj >= (real)a ? (real)1 : zj * C.inv_a) : (real)1;
    for (int c = 0; c < 3; ++c) f[3 * j + c] = bj * F[3 * j + c];
  }
  for (int i = 0; i < n; ++i) {
    real zi = r[3 * i + 2];
    // self term: ordered general path
    if (wall) rbl::pair<real, true, true>(C, r[3*i], r[3*i+1], zi, r[3*i], r[3*i+1], zi, f[3*i], f[3*i+1], f[3*i+2], 2*zi, 4*zi*zi, acc[3*i], acc[3*i+1], acc[3*i+2]);
    else      rbl::pair<real, false, true>(C, r[3*i], r[3*i+1], zi, r[3*i], r[3*i+1], zi, f[3*i], f[3*i+1], f[3*i+2], 2*zi, 4*zi*zi, acc[3*i], acc[3*i+1], acc[3*i+2]);
    for (int j = i + 1; j < n; ++j) {
      real zj = r[3 * j + 2];
#define ARGS r[3*i], r[3*i+1], zi, f[3*i], f[3*i+1], f[3*i+2], -4*zi*zi, r[3*j], r[3*j+1], zj, f[3*j], f[3*j+1], f[3*j+2], -4*zj*zj, acc[3*i], acc[3*i+1], acc[3*i+2], acc[3*j], acc[3*j+1], acc[3*j+2]
      if (wall) { if (near) rbl::pair_sym<real, true, true>(C, ARGS); else rbl::pair_sym<real, true, false>(C, ARGS); }
      else      { if (near) rbl::pair_sym<real, false, true>(C, ARGS); else rbl::pair_sym<real, false, false>(C, ARGS); }
#undef ARGS
    }
  }
  for (int i = 0; i < n; ++i) {
    real zi = r[3 * i + 2];
    real bi = wall ? (zi >= (real)a ? (real)1 : zi * C.inv_a) : (real)1;
    for (int c = 0; c < 3; ++c) U[3 * i + c] = acc[3 * i + c] * C.out_scale * bi;
  }
}
extern "C" void pair_matvec_sym_host_f64(const double* F, const double* r, int n, double a, double eta, int wall, int near, double* U) { matvec_sym<double>(F, r, n, a, eta, wall, near, U); }
extern "C" void pair_matvec_sym_host_f32(const float* F, const float* r, int n, double a, double eta, int wall, int near, float* U) { matvec_sym<float>(F, r, n, a, eta, wall, near, U); }

// two right-hand sides per unordered pair through pair_symR<.., 2> (the arithmetic of the
// two-right-hand-side symmetric kernel): U1 = M F1, U2 = M F2
template <typename real>
static void matvec_sym2(const real* F1, const real* F2, const real* r, int n, double a, double eta, int wall, int near,
                        real* U1, real* U2) {
  rbl::PairConsts<real> C = rbl::make_pair_consts<real>(a, eta);
  std::vector<real> f(6 * (size_t)n), acc(6 * (size_t)n, 0);
  for (int j = 0; j < n; ++j) {
    real zj = r[3 * j + 2];
    real bj = wall ? (zj >= (real)a ? (real)1 : zj * C.inv_a) : (real)1;
    for (int c = 0; c < 3; ++c) { f[6 * j + c] = bj * F1[3 * j + c]; f[6 * j + 3 + c] = bj * F2[3 * j + c]; }
  }
  for (int i = 0; i < n; ++i) {
    real zi = r[3 * i + 2];
    for (int k = 0; k < 2; ++k) {  // self term: ordered general path, one right-hand side at a time
      real* u = &acc[6 * i + 3 * k];
      const real* fk = &f[6 * i + 3 * k];
      if (wall) rbl::pair<real, true, true>(C, r[3*i], r[3*i+1], zi, r[3*i], r[3*i+1], zi, fk[0], fk[1], fk[2], 2*zi, 4*zi*zi, u[0], u[1], u[2]);
      else      rbl::pair<real, false, true>(C, r[3*i], r[3*i+1], zi, r[3*i], r[3*i+1], zi, fk[0], fk[1], fk[2], 2*zi, 4*zi*zi, u[0], u[1], u[2]);
    }
    for (int j = i + 1; j < n; ++j) {
      real zj = r[3 * j + 2];
      real fi[2][3], fj[2][3], ui[2][3], uj[2][3];
      for (int k = 0; k < 2; ++k)
        for (int c = 0; c < 3; ++c) {
          fi[k][c] = f[6 * i + 3 * k + c]; fj[k][c] = f[6 * j + 3 * k + c];
          ui[k][c] = acc[6 * i + 3 * k + c]; uj[k][c] = acc[6 * j + 3 * k + c];
        }
#define ARGS r[3*i], r[3*i+1], zi, fi, -4*zi*zi, r[3*j], r[3*j+1], zj, fj, -4*zj*zj, ui, uj
      if (wall) { if (near) rbl::pair_symR<real, true, true, 2>(C, ARGS); else rbl::pair_symR<real, true, false, 2>(C, ARGS); }
      else      { if (near) rbl::pair_symR<real, false, true, 2>(C, ARGS); else rbl::pair_symR<real, false, false, 2>(C, ARGS); }
#undef ARGS
      for (int k = 0; k < 2; ++k)
        for (int c = 0; c < 3; ++c) { acc[6 * i + 3 * k + c] = ui[k][c]; acc[6 * j + 3 * k + c] = uj[k][c]; }
    }
  }
  for (int i = 0; i < n; ++i) {
    real zi = r[3 * i + 2];
    real bi = wall ? (zi >= (real)a ? (real)1 : zi * C.inv_a) : (real)1;
    for (int c = 0; c < 3; ++c) { U1[3 * i + c] = acc[6 * i + c] * C.out_scale * bi; U2[3 * i + c] = acc[6 * i + 3 + c] * C.out_scale * bi; }
  }
}
extern "C" void pair_matvec_sym2_host_f64(const double* F1, const double* F2, const double* r, int n, double a, double eta, int wall, int near, double* U1, double* U2) { matvec_sym2<double>(F1, F2, r, n, a, eta, wall, near, U1, U2); }
extern "C" void pair_matvec_sym2_host_f32(const float* F1, const float* F2, const float* r, int n, double a, double eta, int wall, int near, float* U1, float* U2) { matvec_sym2<float>(F1, F2, r, n, a, eta, wall, near, U1, U2); }
