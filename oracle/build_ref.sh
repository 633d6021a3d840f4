#!/usr/bin/env bash
# build_ref.sh -- compile the REFERENCE's own two pair kernels (mobilityUFRPY and
# mobilityUFSingleWallCorrection, /root/reference/src/c_rigid_obj.cpp:31-142) from
# the reference source WHERE IT LIES into oracle/_ref/libref_pair.so.
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
# same flags for both sides of the bit-for-bit comparison: no FMA contraction
${CXX_SYS:-/usr/bin/g++} -O2 -ffp-contract=off -fPIC -shared -o "$OUT/libref_pair.so" "$TMP/ref_pair_f64.cpp" "$TMP/ref_pair_f32.cpp"
echo "built $OUT/libref_pair.so from $REF_SRC"
