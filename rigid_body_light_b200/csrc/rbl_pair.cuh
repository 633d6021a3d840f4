// rbl_pair.cuh -- per-pair arithmetic of the blob-blob RPY mobility product.
//
// Replaces the two scalar pair kernels of the reference
//   mobilityUFRPY                    /root/reference/src/c_rigid_obj.cpp:31-83
//   mobilityUFSingleWallCorrection   /root/reference/src/c_rigid_obj.cpp:85-142
// and the way rotne_prager_tensor calls them (:432-445), but is NOT a transcription:
// the reference builds a 3x3 block and stores it; this computes block.f directly in
// a division-free, branch-free, FMA-dense form designed for the B200 issue slot
// budget (one FP32 warp instruction per SMSP per clock is the whole budget).
//
// Formulation (derivation in DESIGN.md section 4):
//  * coordinates stay UNSCALED (the reference differences first and scales by 1/a
//    second, :432-434,48-50; pre-scaling would lose fp32 bits on large boxes); every
//    power of the blob radius a is folded into the constants below, and one overall
//    factor a is folded into the output scale 1/(8 pi eta);
//  * the self term needs no branch: r2 carries a +tiny, the overlap ("near",
//    r < 2a) formula evaluated at d=0 is exactly 4/(3a) I, and the wall self term
//    (:98-104) is exactly the pair formula at d=0 (h_hat = 1/2, e = z);
//  * the wall term's h_hat = h_j / R_z division (:106) disappears:
//      h_hat (1-h_hat) ez^2 = z_i z_j / R^2,   ez h_hat = z_j / R,
//      (1-h_hat) ez^2 = z_i R_z / R^2,         h_hat^2 ez^2 = z_j^2 / R^2;
//  * block.f = cF f + A (dx,dy,Z) + B zhat with three scalar coefficients, so the
//    9-entry non-symmetric block is never formed.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define RBL_HD __host__ __device__ __forceinline__
#else
#define RBL_HD inline
#endif

namespace rbl {

// Constants of one (a, eta) pair; lives in the kernel parameter block so every use
// is a constant-bank operand of an FFMA/DFMA (no register, no issue slot).
template <typename real>
struct PairConsts {
  real tiny;      // added to r^2 so rsqrt(0) never happens
  real four_a2;   // near/far switch: r^2 < 4 a^2
  // free space, far:  c1 = invr + c23a2 invr^3 ; c2 = invr^3 + m2a2 invr^5
  real c23a2, m2a2;
  // free space, near: c1 = n0 + n1 r ; c2 = n2 invr
  real n0, n1, n2;
  // wall polynomials in E = Z^2 W and W = 1/R^2 (signs folded, see pair())
  real k1a, k1b, k2a, k2b;
  real m1a, m1b, m2a, m2b;
  real q1a, q1b, q2a, q2b;
  real a4c;
  real o1a, o1b, fa2;
  // epilogue
  real out_scale; // 1/(8 pi eta)
  real inv_a;
  real a;
};

template <typename real>
inline PairConsts<real> make_pair_consts(double a, double eta) {
  PairConsts<real> c;
  const double a2 = a * a, a4 = a2 * a2;
  c.tiny = sizeof(real) == 4 ? (real)1e-30 : (real)1e-280;
  c.four_a2 = (real)(4.0 * a2);
  c.c23a2 = (real)(2.0 * a2 / 3.0);
  c.m2a2 = (real)(-2.0 * a2);
  c.n0 = (real)(4.0 / (3.0 * a));
  c.n1 = (real)(-3.0 / (8.0 * a2));
  c.n2 = (real)(1.0 / (8.0 * a2));
  // a1n = -(1+2p) - W (k1 + k2 W),  k1 = (2a^2/3)(1-3E), k2 = -(2a^4/3)(1-5E); stored negated
  c.k1a = (real)(2.0 * a2);         c.k1b = (real)(-2.0 * a2 / 3.0);
  c.k2a = (real)(-10.0 * a4 / 3.0); c.k2b = (real)(2.0 * a4 / 3.0);
  // a2n = -(1-6p) - W (m1 + m2 W),  m1 = -2a^2(1-5E), m2 = (10a^4/3)(1-7E); stored negated
  c.m1a = (real)(-10.0 * a2);       c.m1b = (real)(2.0 * a2);
  c.m2a = (real)(70.0 * a4 / 3.0);  c.m2b = (real)(-10.0 * a4 / 3.0);
  // a3 = 2 zj (1 - 6 zi Z W) - Z W (q1 + q2 W),  q1 = 4a^2(1-5E), q2 = -(20a^4/3)(2-7E)
  c.q1a = (real)(-20.0 * a2);       c.q1b = (real)(4.0 * a2);
  c.q2a = (real)(140.0 * a4 / 3.0); c.q2b = (real)(-40.0 * a4 / 3.0);
  // a4 = 2 zj - (20a^4/3) Z W^2
  c.a4c = (real)(-20.0 * a4 / 3.0);
  // a5n = -(4 zj^2 + 4a^2 E) - (4a^4/3)(2-15E) W ; o1 stored negated
  c.o1a = (real)(20.0 * a4);        c.o1b = (real)(-8.0 * a4 / 3.0);
  c.fa2 = (real)(4.0 * a2);
  c.out_scale = (real)(1.0 / (8.0 * M_PI * eta));
  c.inv_a = (real)(1.0 / a);
  c.a = (real)a;
  return c;
}

// ---- fast reciprocal square roots ------------------------------------------------
RBL_HD float rsqrt_fast(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));  // one MUFU.RSQ
  return y;
#else
  return 1.0f / std::sqrt(x);
#endif
}

RBL_HD double rsqrt_fast(double x) {
#if defined(__CUDA_ARCH__)
  // MUFU.RSQ64H seed (~20 bits) + ONE third-order Householder step:
  //   e = 1 - x y^2 ;  y <- y (1 + e/2 + 3 e^2/8)   =>  rel. error ~ (5/16) e^3 < 2^-56.
  // 5 FP64-pipe ops instead of libdevice rsqrt()'s two Newton steps + special cases.
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double t = x * y;
  double e = fma(-t, y, 1.0);
  double q = fma(e, 0.375, 0.5);
  q = q * e;
  return fma(y, q, y);
#else
  return 1.0 / std::sqrt(x);
#endif
}

template <typename real>
RBL_HD real fma_(real a, real b, real c) {
#if defined(__CUDA_ARCH__)
  return fma(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}

// One ordered pair: target i at (xi,yi,zi), source j with position (xj,yj,zj), force f
// (already multiplied by the source's wall damping B_j), z2j = 2 zj, zz4j = 4 zj^2.
// Accumulates M_ij f (without the out_scale and B_i factors) into (ux,uy,uz).
//   WALL : add the Rotne-Prager-Blake single wall correction
//   NEAR : evaluate the r < 2a overlap branch too and select (also covers i == j);
//          NEAR=false is only legal when every pair of the tile has r >= 2a.
template <typename real, bool WALL, bool NEAR>
RBL_HD void pair(const PairConsts<real>& C, real xi, real yi, real zi, real xj, real yj,
                 real zj, real fx, real fy, real fz, real z2j, real zz4j, real& ux,
                 real& uy, real& uz) {
  const real dx = xi - xj, dy = yi - yj, dz = zi - zj;
  const real q = fma_(dy, dy, fma_(dx, dx, C.tiny));
  const real r2 = fma_(dz, dz, q);
  const real s = fma_(dy, fy, dx * fx);
  const real df = fma_(dz, fz, s);
  const real invr = rsqrt_fast(r2);
  const real i2 = invr * invr;
  const real i3 = invr * i2;
  real c1 = fma_(i3, C.c23a2, invr);
  real c2 = fma_(i2 * i3, C.m2a2, i3);
  if (NEAR) {
    const real r = r2 * invr;
    const real c1n = fma_(r, C.n1, C.n0);
    const real c2n = invr * C.n2;
    const bool nr = r2 < C.four_a2;
    c1 = nr ? c1n : c1;
    c2 = nr ? c2n : c2;
  }
  const real t = c2 * df;
  if (!WALL) {
    ux = fma_(c1, fx, ux); ux = fma_(t, dx, ux);
    uy = fma_(c1, fy, uy); uy = fma_(t, dy, uy);
    uz = fma_(c1, fz, uz); uz = fma_(t, dz, uz);
  } else {
    const real Z = zi + zj;
    const real R2 = fma_(Z, Z, q);
    const real w = rsqrt_fast(R2);
    const real W = w * w;
    const real g = fma_(Z, fz, s);
    const real E = fma_(-q, W, (real)1);  // Z^2 W = 1 - (dx^2 + dy^2) W: E only ever enters next to O(1) terms
    const real p = (zi * zj) * W;
    const real k1 = fma_(E, C.k1a, C.k1b);
    const real k2 = fma_(E, C.k2a, C.k2b);
    const real a1n = fma_(fma_(k2, W, k1), W, fma_(p, (real)-2, (real)-1));
    const real m1 = fma_(E, C.m1a, C.m1b);
    const real m2 = fma_(E, C.m2a, C.m2b);
    const real a2n = fma_(fma_(m2, W, m1), W, fma_(p, (real)6, (real)-1));
    const real ZW = Z * W;
    const real q1 = fma_(E, C.q1a, C.q1b);
    const real q2 = fma_(E, C.q2a, C.q2b);
    const real h3 = fma_(q2, W, q1);
    // a3 = 2 z_j (1 - 6 z_i Z W) - Z W h3 = 2 z_j - 12 p Z - Z W h3   (z_i z_j Z W = p Z)
    const real a3 = fma_(-ZW, h3, fma_(p * Z, (real)-12, z2j));
    const real a4 = fma_(ZW * W, C.a4c, z2j);
    const real o1 = fma_(E, C.o1a, C.o1b);
    const real a5n = fma_(o1, W, -fma_(E, C.fa2, zz4j));
    const real wW = w * W;
    const real cF = fma_(w, a1n, c1);
    const real A = wW * fma_(a3, fz, a2n * g);
    const real Bz = wW * fma_(a5n, fz, a4 * g);
    const real txy = t + A;
    ux = fma_(cF, fx, ux); ux = fma_(txy, dx, ux);
    uy = fma_(cF, fy, uy); uy = fma_(txy, dy, uy);
    uz = fma_(cF, fz, uz); uz = fma_(t, dz, uz); uz = fma_(A, Z, uz); uz += Bz;
  }
}


// One UNORDERED pair {i, j}: evaluates the scalar part once and applies the block in both
// directions, u_i += M_ij f_j and u_j += M_ji f_i (M_ji = M_ij^T: the reference mirrors the
// transposed block, c_rigid_obj.cpp:449-452; here M_ji is the same formula with the roles of
// z_i and z_j swapped).  Shared: distances, both rsqrt, c1, c2, a1, a2 and every polynomial
// in (W, E); per direction: the dot products with the force, a3/a4/a5 (which carry the
// SOURCE height) and the accumulation.  ~92 issue slots per unordered pair instead of
// 2 x 65 for two ordered evaluations.
//   fi*, fj* are already multiplied by the blob's wall damping B; nzz4 = -4 z^2.  (2 z_src, which
//   the ordered kernel reads from the record, enters as a literal-operand FMA on z_src here so a
//   thread need not keep it per target.)
template <typename real, bool WALL, bool NEAR>
RBL_HD void pair_sym(const PairConsts<real>& C, real xi, real yi, real zi, real fxi, real fyi,
                     real fzi, real nzz4i, real xj, real yj, real zj, real fxj, real fyj,
                     real fzj, real nzz4j, real& uxi, real& uyi, real& uzi, real& uxj,
                     real& uyj, real& uzj) {
  const real dx = xi - xj, dy = yi - yj, dz = zi - zj;
  const real q = fma_(dy, dy, fma_(dx, dx, C.tiny));
  const real r2 = fma_(dz, dz, q);
  const real sj = fma_(dy, fyj, dx * fxj);  // d_xy . f_j
  const real si = fma_(dy, fyi, dx * fxi);  // d_xy . f_i
  const real dfj = fma_(dz, fzj, sj);
  const real dfi = fma_(dz, fzi, si);
  const real invr = rsqrt_fast(r2);
  const real i2 = invr * invr;
  const real i3 = invr * i2;
  real c1 = fma_(i3, C.c23a2, invr);
  real c2 = fma_(i2 * i3, C.m2a2, i3);
  if (NEAR) {
    const real r = r2 * invr;
    const real c1n = fma_(r, C.n1, C.n0);
    const real c2n = invr * C.n2;
    const bool nr = r2 < C.four_a2;
    c1 = nr ? c1n : c1;
    c2 = nr ? c2n : c2;
  }
  const real tj = c2 * dfj;  // goes to i
  const real ti = c2 * dfi;  // goes to j  (c2 (d'.f_i) d' with d' = -d)
  if (!WALL) {
    uxi = fma_(c1, fxj, uxi); uxi = fma_(tj, dx, uxi);
    uyi = fma_(c1, fyj, uyi); uyi = fma_(tj, dy, uyi);
    uzi = fma_(c1, fzj, uzi); uzi = fma_(tj, dz, uzi);
    uxj = fma_(c1, fxi, uxj); uxj = fma_(ti, dx, uxj);
    uyj = fma_(c1, fyi, uyj); uyj = fma_(ti, dy, uyj);
    uzj = fma_(c1, fzi, uzj); uzj = fma_(ti, dz, uzj);
  } else {
    const real Z = zi + zj;
    const real R2 = fma_(Z, Z, q);
    const real w = rsqrt_fast(R2);
    const real W = w * w;
    const real gj = fma_(Z, fzj, sj);   // (dx,dy,Z) . f_j
    const real gi = fma_(Z, fzi, -si);  // (-dx,-dy,Z) . f_i
    const real E = fma_(-q, W, (real)1);  // Z^2 W = 1 - (dx^2 + dy^2) W
    const real p = (zi * zj) * W;
    const real k1 = fma_(E, C.k1a, C.k1b);
    const real k2 = fma_(E, C.k2a, C.k2b);
    const real a1n = fma_(fma_(k2, W, k1), W, fma_(p, (real)-2, (real)-1));
    const real m1 = fma_(E, C.m1a, C.m1b);
    const real m2 = fma_(E, C.m2a, C.m2b);
    const real a2n = fma_(fma_(m2, W, m1), W, fma_(p, (real)6, (real)-1));
    const real ZW = Z * W;
    const real q1 = fma_(E, C.q1a, C.q1b);
    const real q2 = fma_(E, C.q2a, C.q2b);
    const real nZWh3 = -ZW * fma_(q2, W, q1);
    // a3 of the two directions: 2 z_src (1 - 6 z_tgt Z W) - Z W h3 = 2 z_src + S3 with the shared
    // S3 = -12 p Z - Z W h3 (z_i z_j Z W = p Z): one literal-operand FMA per direction
    const real S3 = fma_(p * Z, (real)-12, nZWh3);
    const real cZW2 = (ZW * W) * C.a4c;
    const real o1 = fma_(E, C.o1a, C.o1b);
    const real nS5 = fma_(o1, W, -(E * C.fa2));  // -(4a^2 E) - (4a^4/3)(2-15E) W
    const real wW = w * W;
    const real cF = fma_(w, a1n, c1);
    // direction i <- j (source height z_j)
    {
      const real a3 = fma_(zj, (real)2, S3);
      const real a4 = fma_(zj, (real)2, cZW2);
      const real a5n = nzz4j + nS5;
      const real A = wW * fma_(a3, fzj, a2n * gj);
      const real Bz = wW * fma_(a5n, fzj, a4 * gj);
      const real txy = tj + A;
      uxi = fma_(cF, fxj, uxi); uxi = fma_(txy, dx, uxi);
      uyi = fma_(cF, fyj, uyi); uyi = fma_(txy, dy, uyi);
      uzi = fma_(cF, fzj, uzi); uzi = fma_(tj, dz, uzi); uzi = fma_(A, Z, uzi); uzi += Bz;
    }
    // direction j <- i (source height z_i, in-plane separation -d)
    {
      const real a3 = fma_(zi, (real)2, S3);
      const real a4 = fma_(zi, (real)2, cZW2);
      const real a5n = nzz4i + nS5;
      const real A = wW * fma_(a3, fzi, a2n * gi);
      const real Bz = wW * fma_(a5n, fzi, a4 * gi);
      const real txy = ti - A;
      uxj = fma_(cF, fxi, uxj); uxj = fma_(txy, dx, uxj);
      uyj = fma_(cF, fyi, uyj); uyj = fma_(txy, dy, uyj);
      uzj = fma_(cF, fzi, uzj); uzj = fma_(ti, dz, uzj); uzj = fma_(A, Z, uzj); uzj += Bz;
    }
  }
}

// pair_sym for R right-hand sides at once (block Lanczos: M^{1/2}W_1 and M^{1/2}W_2 of one BD
// step share every product).  Everything that depends only on the geometry -- ~51 of the ~93
// issue slots of pair_sym with the wall, including both rsqrt -- is computed once; only the dot
// products with the forces and the accumulation are repeated per right-hand side.
//   fi[k], fj[k]: damped forces of right-hand side k;  ui[k], uj[k]: accumulators.
template <typename real, bool WALL, bool NEAR, int R>
RBL_HD void pair_symR(const PairConsts<real>& C, real xi, real yi, real zi, const real (&fi)[R][3], real nzz4i,
                      real xj, real yj, real zj, const real (&fj)[R][3], real nzz4j, real (&ui)[R][3],
                      real (&uj)[R][3]) {
  const real dx = xi - xj, dy = yi - yj, dz = zi - zj;
  const real q = fma_(dy, dy, fma_(dx, dx, C.tiny));
  const real r2 = fma_(dz, dz, q);
  const real invr = rsqrt_fast(r2);
  const real i2 = invr * invr;
  const real i3 = invr * i2;
  real c1 = fma_(i3, C.c23a2, invr);
  real c2 = fma_(i2 * i3, C.m2a2, i3);
  if (NEAR) {
    const real r = r2 * invr;
    const real c1n = fma_(r, C.n1, C.n0);
    const real c2n = invr * C.n2;
    const bool nr = r2 < C.four_a2;
    c1 = nr ? c1n : c1;
    c2 = nr ? c2n : c2;
  }
  if (!WALL) {
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const real sj = fma_(dy, fj[k][1], dx * fj[k][0]);
      const real si = fma_(dy, fi[k][1], dx * fi[k][0]);
      const real tj = c2 * fma_(dz, fj[k][2], sj);
      const real ti = c2 * fma_(dz, fi[k][2], si);
      ui[k][0] = fma_(c1, fj[k][0], ui[k][0]); ui[k][0] = fma_(tj, dx, ui[k][0]);
      ui[k][1] = fma_(c1, fj[k][1], ui[k][1]); ui[k][1] = fma_(tj, dy, ui[k][1]);
      ui[k][2] = fma_(c1, fj[k][2], ui[k][2]); ui[k][2] = fma_(tj, dz, ui[k][2]);
      uj[k][0] = fma_(c1, fi[k][0], uj[k][0]); uj[k][0] = fma_(ti, dx, uj[k][0]);
      uj[k][1] = fma_(c1, fi[k][1], uj[k][1]); uj[k][1] = fma_(ti, dy, uj[k][1]);
      uj[k][2] = fma_(c1, fi[k][2], uj[k][2]); uj[k][2] = fma_(ti, dz, uj[k][2]);
    }
  } else {
    const real Z = zi + zj;
    const real R2 = fma_(Z, Z, q);
    const real w = rsqrt_fast(R2);
    const real W = w * w;
    const real E = fma_(-q, W, (real)1);  // Z^2 W = 1 - (dx^2 + dy^2) W
    const real p = (zi * zj) * W;
    const real k1 = fma_(E, C.k1a, C.k1b);
    const real k2 = fma_(E, C.k2a, C.k2b);
    const real a1n = fma_(fma_(k2, W, k1), W, fma_(p, (real)-2, (real)-1));
    const real m1 = fma_(E, C.m1a, C.m1b);
    const real m2 = fma_(E, C.m2a, C.m2b);
    const real wW = w * W;
    const real a2w = wW * fma_(fma_(m2, W, m1), W, fma_(p, (real)6, (real)-1));  // wW a2n
    const real ZW = Z * W;
    const real q1 = fma_(E, C.q1a, C.q1b);
    const real q2 = fma_(E, C.q2a, C.q2b);
    const real nZWh3 = -ZW * fma_(q2, W, q1);
    const real cZW2 = (ZW * W) * C.a4c;
    const real o1 = fma_(E, C.o1a, C.o1b);
    const real nS5 = fma_(o1, W, -(E * C.fa2));
    const real cF = fma_(w, a1n, c1);
    // geometry-only coefficients of the two directions, pre-multiplied by wW
    const real S3 = fma_(p * Z, (real)-12, nZWh3);  // shared part of a3 (see pair_sym)
    const real a3j = wW * fma_(zj, (real)2, S3);  // i <- j
    const real a4j = wW * fma_(zj, (real)2, cZW2);
    const real a5j = wW * (nzz4j + nS5);
    const real a3i = wW * fma_(zi, (real)2, S3);  // j <- i
    const real a4i = wW * fma_(zi, (real)2, cZW2);
    const real a5i = wW * (nzz4i + nS5);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const real sj = fma_(dy, fj[k][1], dx * fj[k][0]);
      const real si = fma_(dy, fi[k][1], dx * fi[k][0]);
      const real tj = c2 * fma_(dz, fj[k][2], sj);
      const real ti = c2 * fma_(dz, fi[k][2], si);
      const real gj = fma_(Z, fj[k][2], sj);
      const real gi = fma_(Z, fi[k][2], -si);
      {
        const real A = fma_(a3j, fj[k][2], a2w * gj);
        const real Bz = fma_(a5j, fj[k][2], a4j * gj);
        const real txy = tj + A;
        ui[k][0] = fma_(cF, fj[k][0], ui[k][0]); ui[k][0] = fma_(txy, dx, ui[k][0]);
        ui[k][1] = fma_(cF, fj[k][1], ui[k][1]); ui[k][1] = fma_(txy, dy, ui[k][1]);
        ui[k][2] = fma_(cF, fj[k][2], ui[k][2]); ui[k][2] = fma_(tj, dz, ui[k][2]);
        ui[k][2] = fma_(A, Z, ui[k][2]); ui[k][2] += Bz;
      }
      {
        const real A = fma_(a3i, fi[k][2], a2w * gi);
        const real Bz = fma_(a5i, fi[k][2], a4i * gi);
        const real txy = ti - A;
        uj[k][0] = fma_(cF, fi[k][0], uj[k][0]); uj[k][0] = fma_(txy, dx, uj[k][0]);
        uj[k][1] = fma_(cF, fi[k][1], uj[k][1]); uj[k][1] = fma_(txy, dy, uj[k][1]);
        uj[k][2] = fma_(cF, fi[k][2], uj[k][2]); uj[k][2] = fma_(ti, dz, uj[k][2]);
        uj[k][2] = fma_(A, Z, uj[k][2]); uj[k][2] += Bz;
      }
    }
  }
}

}  // namespace rbl
