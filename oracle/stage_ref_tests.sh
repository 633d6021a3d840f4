#!/usr/bin/env bash
# stage_ref_tests.sh -- put the REFERENCE's own test-suite, its Python wrapper and the one blob
# model its tests load next to the prebuilt reference libraries in oracle/_ref/ (git-ignored, so
# no reference text enters the history; NOT gpurun-ignored, so it travels to the GPU box), byte
# for byte:
#   oracle/_ref/reference_tests/tests/{utils,test_import,test_interface,test_precision,test_wall}.py
#   oracle/_ref/reference_tests/structures/shell_N_12.csv        (tests/utils.py:5-6 loads ../structures/)
#   oracle/_ref/reference_tests/pkg_ref/Rigid/{__init__.py,Rigid.py}   = /root/reference/src/{__init__,Rigid}.py
#   oracle/_ref/reference_tests/pkg_ref/Rigid/c_rigid.py                = tests/host/ref_layout_c_rigid.py (ours):
#       what CMakeLists.txt:24-27 installs as c_rigid*.so, served by this repository's host class
# tests/test_gpu_reference_suite.py runs the staged suite twice on the GPU box: against this
# repository's `Rigid` package, and against the reference's own unmodified Rigid.py driving
# c_rigid.CManyBodies.  TEST INFRASTRUCTURE ONLY.
set -euo pipefail
REF="${REF_ROOT:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref/reference_tests"
if [ ! -d "$REF/tests" ]; then
  echo "stage_ref_tests.sh: $REF not present (GPU box?) -- keeping staged $OUT" >&2
  exit 0
fi
rm -rf "$OUT"
mkdir -p "$OUT/tests" "$OUT/structures" "$OUT/pkg_ref/Rigid"
for f in utils.py test_import.py test_interface.py test_precision.py test_wall.py; do
  cp "$REF/tests/$f" "$OUT/tests/$f"
done
cp "$REF/structures/shell_N_12.csv" "$OUT/structures/shell_N_12.csv"
cp "$REF/src/__init__.py" "$OUT/pkg_ref/Rigid/__init__.py"
cp "$REF/src/Rigid.py" "$OUT/pkg_ref/Rigid/Rigid.py"
cp "$HERE/../tests/host/ref_layout_c_rigid.py" "$OUT/pkg_ref/Rigid/c_rigid.py"
chmod -R u+w "$OUT"
( cd "$OUT" && find . -type f ! -name MANIFEST.sha256 | sort | xargs sha256sum > MANIFEST.sha256 )
echo "stage_ref_tests.sh: staged the reference's tests and wrapper under $OUT"
