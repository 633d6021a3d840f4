#!/usr/bin/env bash
# build_ref.sh -- compile the REFERENCE's own two pair kernels (mobilityUFRPY and
# mobilityUFSingleWallCorrection, /root/reference/src/c_rigid_obj.cpp:31-142) from
# the reference source WHERE IT LIES into oracle/_ref/libref_pair.so, and its dense
# assembly + apply_M members (:413-459, 618-659) into oracle/_ref/libref_apply_M.so.
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
# ---- the reference's own dense assembly + apply_M (c_rigid_obj.cpp:413-459, 618-659) ------------
# rotne_prager_tensor, make_damp_mat and apply_M are members of CManyBodies; they use only the
# pair kernels, the members a / eta / PC_wall and a few Eigen dense operations.  Eigen3 is not
# installed, so oracle/eigen_shim.inc (ours) supplies exactly those operations; the three member
# functions are streamed from the reference source into a struct that holds the three data
# members.  Result: libref_apply_M.so = the reference's assembly loop and its B M B F expression,
# as written, in float and double.
gen_apply() { # $1 = real type, $2 = suffix
  {
    echo '#include <cmath>'
    echo '#include <iostream>'
    echo '#include <stdexcept>'
    echo '#include <cstdlib>'
    echo '#include <vector>'
    echo '#include <algorithm>'
    echo '#include <initializer_list>'
    echo "namespace refm_$2 {"
    echo "using real = $1;"
    cat "$HERE/eigen_shim.inc"
    awk '/^void mobilityUFRPY\(/{on=1} /^class CManyBodies/{on=0} on{print}' "$REF_SRC"
    echo 'struct RefBody {'
    echo '  real a, eta; bool PC_wall;'
    awk '/template <class AVector> Matrix rotne_prager_tensor\(/{on=1} /^  SparseM Block_diag_invM\(\)/{on=0} on{print}' "$REF_SRC"
    awk '/^  DiagM make_damp_mat\(/{on=1} /^  Vector M_half_W\(\)/{on=0} on{print}' "$REF_SRC"
    echo '};'
    echo '}'
    cat <<WRAP
extern "C" int ref_apply_M_$2(const $1 *F, const $1 *r, int n_blobs, double a, double eta, int wall, $1 *U) {
  refm_$2::RefBody b;
  b.a = ($1)a; b.eta = ($1)eta; b.PC_wall = wall != 0;
  refm_$2::Vector f(3L * n_blobs);
  std::vector<$1> rv(r, r + 3L * n_blobs);
  for (long i = 0; i < 3L * n_blobs; ++i) f(i) = F[i];
  try {
    refm_$2::Vector u = b.apply_M(f, rv);
    for (long i = 0; i < 3L * n_blobs; ++i) U[i] = u(i);
  } catch (const std::runtime_error &) { return 2; }
  return 0;
}
WRAP
  } > "$TMP/ref_apply_$2.cpp"
}
gen_apply double f64
gen_apply float f32
${CXX_SYS:-/usr/bin/g++} -O2 -ffp-contract=off -fPIC -shared -o "$OUT/libref_apply_M.so" "$TMP/ref_apply_f64.cpp" "$TMP/ref_apply_f32.cpp"
echo "built $OUT/libref_apply_M.so from $REF_SRC"
# same flags for both sides of the bit-for-bit comparison: no FMA contraction
${CXX_SYS:-/usr/bin/g++} -O2 -ffp-contract=off -fPIC -shared -o "$OUT/libref_pair.so" "$TMP/ref_pair_f64.cpp" "$TMP/ref_pair_f32.cpp"
echo "built $OUT/libref_pair.so from $REF_SRC"
