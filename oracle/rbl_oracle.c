/*
 * rbl_oracle.c -- CPU oracle for the blob-blob RPY mobility product.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rigid_body_light_b200/ or Rigid/ may
 * import, link or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, as the checker or the
 * reported CPU baseline -- never as the product path.
 *
 * It restates /root/reference/src/c_rigid_obj.cpp:31-142,413-459,618-659 in plain
 * C (no Eigen).  The reference as a whole cannot be built here (it needs Eigen3 and
 * nanobind, neither present, no network), but oracle/build_ref.sh compiles, from the
 * reference source where it lies, (a) its two pair kernels and (b) its
 * rotne_prager_tensor + make_damp_mat + apply_M members (with oracle/eigen_shim.inc
 * supplying the few Eigen dense operations they use) into oracle/_ref/, and
 * tests/test_oracle_vs_ref.py checks this restatement against both BIT FOR BIT in
 * float and double (live, and against committed outputs under tests/golden/).
 */
#include <math.h>
#include <stdlib.h>
#include <stddef.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define REAL double
#define ACC long double
#define SQRT sqrt
#define SFX(x) x##_f64
#include "oracle_impl.inc"
#undef REAL
#undef ACC
#undef SQRT
#undef SFX

#define REAL float
#define ACC double
#define SQRT sqrtf
#define SFX(x) x##_f32
#include "oracle_impl.inc"
#undef REAL
#undef ACC
#undef SQRT
#undef SFX

int orc_num_threads(void) {
#ifdef _OPENMP
  extern int omp_get_max_threads(void);
  return omp_get_max_threads();
#else
  return 1;
#endif
}
