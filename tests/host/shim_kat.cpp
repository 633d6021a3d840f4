// Known-answer driver for oracle/eigen_shim.inc (TEST INFRASTRUCTURE): exercises every Eigen operation the
// reference's members use through the shim on fixed inputs and prints the results as JSON; the Python side
// (tests/test_eigen_shim_kat.py) recomputes them with numpy / scipy, whose conventions are Eigen's documented
// ones (column-major storage, Quaternion coefficients (x, y, z, w), Hamilton product, duplicate triplets summed).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <initializer_list>
#include <utility>
#include <vector>

namespace kat {
using real = double;
#include "../../oracle/eigen_shim.inc"

static double entry(long i, long j) { return std::sin(1.0 + 0.7 * (double)i + 1.3 * (double)j) + (i == j ? 2.0 : 0.0); }
static Matrix filled(long n, long m) {
  Matrix A(n, m);
  for (long j = 0; j < m; ++j)
    for (long i = 0; i < n; ++i) A(i, j) = entry(i, j);
  return A;
}
static void dump_raw(const char* name, const Mat& M, bool last = false) {  // storage order as it lies in memory
  std::printf("\"%s\": {\"rows\": %ld, \"cols\": %ld, \"raw\": [", name, M.rows(), M.cols());
  for (size_t k = 0; k < M.a.size(); ++k) std::printf("%s%.17g", k ? ", " : "", M.a[k]);
  std::printf("]}%s\n", last ? "" : ",");
}
}  // namespace kat

int main() {
  using namespace kat;
  std::printf("{\n");
  // storage order
  Matrix S(2, 3);
  for (long i = 0; i < 2; ++i)
    for (long j = 0; j < 3; ++j) S(i, j) = 10.0 * i + j;
  dump_raw("storage_2x3", S);
  // dense algebra
  Matrix A3 = filled(3, 3), A6 = filled(6, 6), B63 = filled(6, 3);
  dump_raw("inv3", A3.inverse());
  dump_raw("inv6", A6.inverse());
  std::printf("\"det3\": %.17g,\n", A3.determinant());
  dump_raw("transpose63", B63.transpose());
  dump_raw("prod66_63", A6 * B63);
  dump_raw("sum", A6 + A6.transpose());
  dump_raw("diff", A6 - A6.transpose());
  dump_raw("scaled", 2.5 * A3);
  dump_raw("colmean", B63.colwise().mean());
  std::printf("\"norm\": %.17g, \"sqnorm\": %.17g,\n", B63.norm(), B63.squaredNorm());
  dump_raw("block", A6.block(1, 2, 3, 2));
  // diagonal scaling (make_damp_mat: B = diag(d); B * M * B)
  Vector d(6);
  for (long i = 0; i < 6; ++i) d(i) = 0.5 + 0.1 * i;
  DiagM D(d.asDiagonal());
  dump_raw("diag_left", D * A6);
  dump_raw("diag_right", A6 * D);
  // Cholesky
  Matrix P = A6 * A6.transpose();
  for (long i = 0; i < 6; ++i) P(i, i) += 6.0;
  Eigen::LLT<Matrix> llt(P);
  dump_raw("chol_L", llt.matrixL());
  Vector rhs(6);
  for (long i = 0; i < 6; ++i) rhs(i) = std::cos(0.3 * i);
  dump_raw("chol_solve", llt.solve(rhs));
  // quaternion
  Quat q;
  q.x() = -0.5; q.y() = 0.2; q.z() = 0.7; q.w() = 0.3;
  q.normalize();
  dump_raw("quat_rot", q.toRotationMatrix());
  Quat p;
  p.x() = 0.1; p.y() = 0.9; p.z() = -0.3; p.w() = 0.4;
  p.normalize();
  Quat qp = q * p;
  std::printf("\"quat_prod_xyzw\": [%.17g, %.17g, %.17g, %.17g],\n", qp.x(), qp.y(), qp.z(), qp.w());
  Quat z0;
  z0.x() = z0.y() = z0.z() = z0.w() = 0.0;
  z0.normalize();  // Eigen leaves the zero quaternion untouched
  std::printf("\"quat_zero_xyzw\": [%.17g, %.17g, %.17g, %.17g],\n", z0.x(), z0.y(), z0.z(), z0.w());
  Quat id = Quat::Identity();
  std::printf("\"quat_identity_xyzw\": [%.17g, %.17g, %.17g, %.17g],\n", id.x(), id.y(), id.z(), id.w());
  // triplets: duplicates are summed
  std::vector<Trip> t;
  t.emplace_back(0, 0, 1.0); t.emplace_back(1, 2, 2.0); t.emplace_back(1, 2, 0.5); t.emplace_back(2, 1, -3.0); t.emplace_back(0, 0, 0.25);
  SparseM Sp(3, 3);
  Sp.setFromTriplets(t.begin(), t.end());
  dump_raw("triplets", Sp);
  dump_raw("sparse_prod", Sp * Sp.transpose());
  // stacking with the comma initialiser
  Vector st(9);
  Vector h = B63.block(0, 0, 6, 1), tl = A3.block(0, 1, 3, 1);
  st << h, tl;
  dump_raw("stacked", st, true);
  std::printf("}\n");
  return 0;
}
