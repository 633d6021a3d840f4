#!/usr/bin/env bash
# build_ref.sh -- compile the REFERENCE's own two pair kernels (mobilityUFRPY and
# mobilityUFSingleWallCorrection, /root/reference/src/c_rigid_obj.cpp:31-142) from
# the reference source WHERE IT LIES into oracle/_ref/libref_pair.so, and its member
# functions on the product path (state, placement, K/K^T/K^-1, dense assembly + apply_M,
# integrator) into oracle/_ref/libref_members.so.
#
# The whole reference translation unit cannot be built here (Eigen3 + nanobind are
# absent); these two free functions depend only on <cmath>/<iostream>/<stdexcept>,
# so the recipe streams exactly that span of the file through the compiler with a
# small prelude (the `real` typedef) and a C wrapper appended.  No reference text is
# written into the repo: the generated translation unit lives in a mktemp dir that
# is deleted, and only the .so lands in oracle/_ref/ (git-ignored, travels to the
# GPU box with the snapshot).  TEST INFRASTRUCTURE ONLY.
set -euo pipefail
REF_SRC="${REF_SRC:-/root/reference/src/c_rigid_obj.cpp}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF_SRC" ]; then
  echo "build_ref.sh: $REF_SRC not present (GPU box?) -- keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT

gen() { # $1 = real type, $2 = suffix
  {
    echo '#include <cmath>'
    echo '#include <iostream>'
    echo '#include <stdexcept>'
    echo '#include <cstdlib>'
    echo "namespace ref_$2 {"
    echo "using real = $1;"
    # span: from the signature of mobilityUFRPY up to (not including) class CManyBodies
    awk '/^void mobilityUFRPY\(/{on=1} /^class CManyBodies/{on=0} on{print}' "$REF_SRC"
    echo '}'
    cat <<WRAP
extern "C" void ref_rpy_pair_$2($1 rx, $1 ry, $1 rz, $1 *M6, int i, int j, $1 inv_a) {
  ref_$2::mobilityUFRPY(rx, ry, rz, M6[0], M6[1], M6[2], M6[3], M6[4], M6[5], i, j, inv_a);
}
extern "C" int ref_wall_pair_$2($1 rx, $1 ry, $1 rz, $1 *M9, int i, int j, $1 hj) {
  try {
    ref_$2::mobilityUFSingleWallCorrection(rx, ry, rz, M9[0], M9[1], M9[2], M9[3], M9[4], M9[5], M9[6], M9[7], M9[8], i, j, hj);
  } catch (const std::runtime_error &) { return 2; }
  return 0;
}
WRAP
  } > "$TMP/ref_pair_$2.cpp"
}
gen double f64
gen float f32
# ---- the reference's own member functions on the product path ---------------------------------
# State handling, placement, K / K^T / K^-1, the dense assembly + apply_M and the integrator are
# members of CManyBodies that use the pair kernels, a dozen data members and a subset of Eigen
# (dense, quaternion, sparse-from-triplets).  Eigen3 is not installed, so oracle/eigen_shim.inc
# (ours) supplies exactly that subset; the member functions are streamed from the reference source
# into a struct that declares the data members.  Result: libref_members.so = the reference's
#   removeMean .. KT_x_Lam (:176-410), rotne_prager_tensor (:413-459), make_damp_mat + apply_M
#   (:618-659), Block_diag_invM / diag_invM / PC_invM / get_blk_diag_lu (:461-567), apply_PC (:589-616),
#   Q_from_Om + update_X_Q(+_out) (:678-728), evolve_X_Q (:865-878)
# as written, in float and double, behind a small C API.
gen_members() { # $1 = real type, $2 = suffix
  {
    echo '#include <cmath>'
    echo '#include <cstdio>'
    echo '#include <iostream>'
    echo '#include <stdexcept>'
    echo '#include <cstdlib>'
    echo '#include <vector>'
    echo '#include <tuple>'
    echo '#include <utility>'
    echo '#include <algorithm>'
    echo '#include <initializer_list>'
    echo "namespace refm_$2 {"
    echo "using real = $1;"
    cat "$HERE/eigen_shim.inc"
    awk '/^void mobilityUFRPY\(/{on=1} /^class CManyBodies/{on=0} on{print}' "$REF_SRC"
    echo 'struct RefBody {'
    echo '  real a, dt, kBT, eta; bool PC_wall = false; bool block_diag_PC = false; double M_scale;'
    echo '  bool PC_mat_Set = false; bool cfg_set = false; int N_bod = 0; std::vector<Quat> Q_n; std::vector<Vector> X_n;'
    echo '  Matrix ref_cfg; int N_blb = 0; bool parametersSet = false; SparseM K, KT, Kinv;'
    echo '  SparseM invM; SparseM Ninv; std::vector<Eigen::LLT<Matrix>> N_lu;'
    # the reference seeds rand_vector from the wall clock (:730-741); here the test injects the noise
    echo '  std::vector<Vector> injected_noise; size_t noise_pos = 0; bool split_rand = true;'
    echo '  Vector rand_vector(int N) { (void)N; return injected_noise[noise_pos++ % injected_noise.size()]; }'
    echo '  static double timeNow() { return 0.0; }'
    awk '/^  void removeMean\(Matrix &cfg\)/{on=1} /preconditioner\/solver functions/{on=0} on{print}' "$REF_SRC"
    awk '/template <class AVector> Matrix rotne_prager_tensor\(/{on=1} /^  SparseM Block_diag_invM\(\)/{on=0} on{print}' "$REF_SRC"
    awk '/^  DiagM make_damp_mat\(/{on=1} /^  Vector M_half_W\(\)/{on=0} on{print}' "$REF_SRC"
    awk '/^  SparseM Block_diag_invM\(\)/{on=1} /template <class AVector> void test_PC\(/{on=0} on{print}' "$REF_SRC"
    awk '/^  Vector apply_PC\(const Vector &IN\)/{on=1} /^  DiagM make_damp_mat\(/{on=0} on{print}' "$REF_SRC"
    awk '/^  Vector M_half_W\(\)/{on=1} /Dynamics\/time integration/{on=0} on{print}' "$REF_SRC"
    awk '/^  Vector KTinv_RFD\(\)/{on=1} /^  Vector M_RFD\(\)/{on=0} on{print}' "$REF_SRC"
    awk '/^  Vector M_RFD\(\)/{on=1} /template <class AVector> auto M_RFD_cfgs\(/{on=0} on{print}' "$REF_SRC"
    awk '/template <class AVector> auto M_RFD_cfgs\(/{on=1} /^  void evolve_X_Q\(Vector &U\)/{on=0} on{print}' "$REF_SRC"
    awk '/^  void evolve_X_Q_RFD\(/{on=1} /^  Vector Test_Mhalf\(/{on=0} on{print}' "$REF_SRC"
    awk '/^  auto RHS_and_Midpoint\(/{on=1} /^  auto get_K\(\)/{on=0} on{print}' "$REF_SRC"
    awk '/^  Quat Q_from_Om\(/{on=1} /^  Vector rand_vector\(/{on=0} on{print}' "$REF_SRC"
    awk '/^  void evolve_X_Q\(Vector &U\)/{on=1} /^  void evolve_X_Q_RFD\(/{on=0} on{print}' "$REF_SRC"
    echo '};'
    echo '}'
    cat <<WRAP
namespace {
using B_$2 = refm_$2::RefBody;
using V_$2 = refm_$2::Vector;
V_$2 in_$2(const $1 *p, long n) { V_$2 v(n); for (long i = 0; i < n; ++i) v(i) = p[i]; return v; }
void out_$2(const refm_$2::Mat &v, $1 *p) { for (long i = 0; i < (long)v.a.size(); ++i) p[i] = v.a[(size_t)i]; }
}
extern "C" void *refm_create_$2() { return new B_$2(); }
extern "C" void refm_destroy_$2(void *h) { delete static_cast<B_$2 *>(h); }
extern "C" int refm_set_parameters_$2(void *h, double a, double dt, double kBT, double eta, const $1 *cfg, int n_blb) {
  refm_$2::Matrix c(n_blb, 3);
  for (int k = 0; k < n_blb; ++k) for (int d = 0; d < 3; ++d) c(k, d) = cfg[3 * k + d];
  static_cast<B_$2 *>(h)->setParameters(($1)a, ($1)dt, ($1)kBT, ($1)eta, c);
  return 0;
}
extern "C" int refm_set_flags_$2(void *h, int blk, int wall) {
  static_cast<B_$2 *>(h)->setBlkPC(blk != 0); static_cast<B_$2 *>(h)->setWallPC(wall != 0); return 0;
}
extern "C" int refm_set_config_$2(void *h, const $1 *X, const $1 *Q, int n_bod) {
  B_$2 *b = static_cast<B_$2 *>(h);
  b->setConfig(in_$2(X, 3L * n_bod), in_$2(Q, 4L * n_bod));
  b->set_K_mats();
  return 0;
}
extern "C" int refm_get_config_$2(void *h, $1 *X, $1 *Q) {
  auto t = static_cast<B_$2 *>(h)->getConfig();
  out_$2(t.first, X); out_$2(t.second, Q);
  return 0;
}
extern "C" int refm_positions_$2(void *h, $1 *out) {
  std::vector<$1> p = static_cast<B_$2 *>(h)->multi_body_pos();
  for (size_t i = 0; i < p.size(); ++i) out[i] = p[i];
  return 0;
}
extern "C" int refm_K_x_U_$2(void *h, const $1 *U, $1 *out) {
  B_$2 *b = static_cast<B_$2 *>(h); out_$2(b->K_x_U(in_$2(U, 6L * b->N_bod)), out); return 0;
}
extern "C" int refm_KT_x_Lam_$2(void *h, const $1 *L, $1 *out) {
  B_$2 *b = static_cast<B_$2 *>(h); out_$2(b->KT_x_Lam(in_$2(L, 3L * b->N_bod * b->N_blb)), out); return 0;
}
extern "C" int refm_Kinv_x_V_$2(void *h, const $1 *V, $1 *out) {
  B_$2 *b = static_cast<B_$2 *>(h); out_$2(b->Kinv_x_V(in_$2(V, 3L * b->N_bod * b->N_blb)), out); return 0;
}
extern "C" int refm_KTinv_x_F_$2(void *h, const $1 *F, $1 *out) {
  B_$2 *b = static_cast<B_$2 *>(h); out_$2(b->KTinv_x_F(in_$2(F, 6L * b->N_bod)), out); return 0;
}
extern "C" int refm_apply_PC_$2(void *h, const $1 *in, $1 *out) {
  B_$2 *b = static_cast<B_$2 *>(h);
  try { out_$2(b->apply_PC(in_$2(in, 3L * b->N_bod * b->N_blb + 6L * b->N_bod)), out); } catch (const std::runtime_error &) { return 2; }
  return 0;
}
extern "C" int refm_M_RFD_$2(void *h, const $1 *W, $1 *out) {
  B_$2 *b = static_cast<B_$2 *>(h);
  b->injected_noise = {in_$2(W, 3L * b->N_bod * b->N_blb)}; b->noise_pos = 0;
  try { out_$2(b->M_RFD(), out); } catch (const std::runtime_error &) { return 2; }
  return 0;
}
// the other random finite differences (:743-767, :798-863) and the RFD-sized configuration updates
// (:712-728, :880-893), noise injected where the reference draws it
extern "C" int refm_KTinv_RFD_$2(void *h, const $1 *W6, $1 *out) {
  B_$2 *b = static_cast<B_$2 *>(h);
  b->injected_noise = {in_$2(W6, 6L * b->N_bod)}; b->noise_pos = 0;
  try { out_$2(b->KTinv_RFD(), out); } catch (const std::runtime_error &) { return 2; }
  return 0;
}
extern "C" int refm_M_RFD_from_U_$2(void *h, const $1 *U, const $1 *W, $1 *out) {
  B_$2 *b = static_cast<B_$2 *>(h);
  V_$2 u = in_$2(U, 6L * b->N_bod), w = in_$2(W, 3L * b->N_bod * b->N_blb);
  try { out_$2(b->M_RFD_from_U(u, w), out); } catch (const std::runtime_error &) { return 2; }
  return 0;
}
extern "C" int refm_KT_RFD_from_U_$2(void *h, const $1 *U, const $1 *W, $1 *out) {
  B_$2 *b = static_cast<B_$2 *>(h);
  V_$2 u = in_$2(U, 6L * b->N_bod), w = in_$2(W, 3L * b->N_bod * b->N_blb);
  out_$2(b->KT_RFD_from_U(u, w), out);
  return 0;
}
extern "C" int refm_M_RFD_cfgs_$2(void *h, const $1 *U, double delta, $1 *rp, $1 *rm) {
  B_$2 *b = static_cast<B_$2 *>(h);
  V_$2 u = in_$2(U, 6L * b->N_bod);
  b->injected_noise = {V_$2(3L * b->N_bod * b->N_blb)}; b->noise_pos = 0;  // M_RFD_cfgs draws a W it never uses (:801)
  auto t = b->M_RFD_cfgs(u, delta);
  const std::vector<$1> &p = std::get<0>(t), &m = std::get<1>(t);
  for (size_t i = 0; i < p.size(); ++i) { rp[i] = p[i]; rm[i] = m[i]; }
  return 0;
}
extern "C" int refm_update_X_Q_out_$2(void *h, const $1 *U, $1 *X, $1 *Q) {
  B_$2 *b = static_cast<B_$2 *>(h);
  V_$2 u = in_$2(U, 6L * b->N_bod);
  auto t = b->update_X_Q_out(u);  // (Qout [w x y z], Xout), row-major N_bod x 4 / N_bod x 3
  const refm_$2::Matrix &Qo = std::get<0>(t), &Xo = std::get<1>(t);
  for (int j = 0; j < b->N_bod; ++j) {
    for (int d = 0; d < 4; ++d) Q[4 * j + d] = Qo(j, d);
    for (int d = 0; d < 3; ++d) X[3 * j + d] = Xo(j, d);
  }
  return 0;
}
extern "C" int refm_set_split_rand_$2(void *h, int on) { static_cast<B_$2 *>(h)->split_rand = on != 0; return 0; }
extern "C" int refm_evolve_RFD_$2(void *h, const $1 *U) {
  B_$2 *b = static_cast<B_$2 *>(h); V_$2 u = in_$2(U, 6L * b->N_bod); b->evolve_X_Q_RFD(u); return 0;
}
extern "C" int refm_M_half_W_$2(void *h, const $1 *W, $1 *out) {
  B_$2 *b = static_cast<B_$2 *>(h);
  b->injected_noise = {in_$2(W, 3L * b->N_bod * b->N_blb)}; b->noise_pos = 0;
  try { out_$2(b->M_half_W(), out); } catch (const std::runtime_error &) { return 2; }
  return 0;
}
// RHS_and_Midpoint (:917-976) with rand_vector returning W1, W2 (M_half_W twice), then Wr (M_RFD)
extern "C" int refm_RHS_$2(void *h, const $1 *slip, const $1 *force, const $1 *W1, const $1 *W2, const $1 *Wr, double kBT, $1 *out) {
  B_$2 *b = static_cast<B_$2 *>(h);
  const long n3 = 3L * b->N_bod * b->N_blb;
  b->kBT = ($1)kBT;
  b->injected_noise = {in_$2(W1, n3), in_$2(W2, n3), in_$2(Wr, n3)}; b->noise_pos = 0;
  V_$2 s = in_$2(slip, n3), f = in_$2(force, 6L * b->N_bod);
  try { out_$2(b->RHS_and_Midpoint(s, f), out); } catch (const std::runtime_error &) { return 2; }
  return 0;
}
extern "C" int refm_evolve_$2(void *h, const $1 *U) {
  B_$2 *b = static_cast<B_$2 *>(h); V_$2 u = in_$2(U, 6L * b->N_bod); b->evolve_X_Q(u); return 0;
}
extern "C" int refm_apply_M_$2(void *h, const $1 *F, const $1 *r, int n_blobs, $1 *U) {
  B_$2 *b = static_cast<B_$2 *>(h);
  std::vector<$1> rv(r, r + 3L * n_blobs);
  try { out_$2(b->apply_M(in_$2(F, 3L * n_blobs), rv), U); } catch (const std::runtime_error &) { return 2; }
  return 0;
}
// stateless convenience: apply_M for given (a, eta, wall)
extern "C" int ref_apply_M_$2(const $1 *F, const $1 *r, int n_blobs, double a, double eta, int wall, $1 *U) {
  B_$2 b;
  b.a = ($1)a; b.eta = ($1)eta; b.PC_wall = wall != 0;
  return refm_apply_M_$2(&b, F, r, n_blobs, U);
}
WRAP
  } > "$TMP/ref_members_$2.cpp"
}
gen_members double f64
gen_members float f32
${CXX_SYS:-/usr/bin/g++} -O2 -ffp-contract=off -fPIC -shared -o "$OUT/libref_members.so" "$TMP/ref_members_f64.cpp" "$TMP/ref_members_f32.cpp"
ln -sf libref_members.so "$OUT/libref_apply_M.so"
echo "built $OUT/libref_members.so from $REF_SRC"
# same flags for both sides of the bit-for-bit comparison: no FMA contraction
${CXX_SYS:-/usr/bin/g++} -O2 -ffp-contract=off -fPIC -shared -o "$OUT/libref_pair.so" "$TMP/ref_pair_f64.cpp" "$TMP/ref_pair_f32.cpp"
echo "built $OUT/libref_pair.so from $REF_SRC"
