// rbl_rigid.cu -- O(N) rigid-body kernels (placement, K, K^T, preconditioner, integrator).
// See rbl_rigid.cuh for the reference members each kernel replaces.
#include <algorithm>
#include <cmath>
#include <cstdint>

#include "rbl_pair.cuh"
#include "rbl_rigid.cuh"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace rbl {

namespace {

// Rotation matrix of a unit quaternion stored [w,x,y,z] (Eigen toRotationMatrix, used at
// c_rigid_obj.cpp:258,308).  Row-major R[9].
template <typename real>
__device__ __forceinline__ void quat_to_rot(const real* __restrict__ q, real* R) {
  const real w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
}

// block-wide sum of NV values per thread; result valid in thread 0 (and in `red[0..NV)`)
template <typename real, int NV>
__device__ __forceinline__ void block_sum(real (&v)[NV], real* red /* [32*NV] smem */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[wid * NV + k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      real s = red[k];
      for (int w = 1; w < nw; ++w) s += red[w * NV + k];
      v[k] = s;
      red[k] = s;
    }
  }
  __syncthreads();
}

// 6x6 inverse in place by Gauss-Jordan with partial pivoting.  The reference factors
// N^-1 = K^T Mt^-1 K with Eigen's LLT and solves (c_rigid_obj.cpp:554-567,605-608); for the
// SPD matrices of physical configurations the explicit inverse gives the same U to rounding,
// and it stays finite where LLT silently breaks down (blobs inside the wall-overlap layer make
// Mt indefinite).  Returns false only for a (numerically) singular matrix.
template <typename real>
__device__ bool inv6(real* A) {
  int perm[6];
  bool ok = true;
  for (int i = 0; i < 6; ++i) perm[i] = i;
  for (int k = 0; k < 6; ++k) {
    int p = k;
    real best = fabs(A[k * 6 + k]);
    for (int i = k + 1; i < 6; ++i)
      if (fabs(A[i * 6 + k]) > best) { best = fabs(A[i * 6 + k]); p = i; }
    if (!(best > (real)0) || !isfinite(best)) ok = false;
    if (p != k) {
      for (int j = 0; j < 6; ++j) { real t = A[k * 6 + j]; A[k * 6 + j] = A[p * 6 + j]; A[p * 6 + j] = t; }
      int t = perm[k]; perm[k] = perm[p]; perm[p] = t;
    }
    const real piv = (real)1 / A[k * 6 + k];
    A[k * 6 + k] = (real)1;
    for (int j = 0; j < 6; ++j) A[k * 6 + j] *= piv;
    for (int i = 0; i < 6; ++i) {
      if (i == k) continue;
      const real f = A[i * 6 + k];
      A[i * 6 + k] = (real)0;
      for (int j = 0; j < 6; ++j) A[i * 6 + j] -= f * A[k * 6 + j];
    }
  }
  // undo the row swaps: columns of the inverse are permuted
  real T[36];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) T[i * 6 + perm[j]] = A[i * 6 + j];
  for (int i = 0; i < 36; ++i) A[i] = T[i];
  return ok;
}

}  // namespace

// ----------------------------------------------------------------------------------
template <typename real>
__global__ void normalize_quats_kernel(real* Q, int n_bod) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bod) return;
  real w = Q[4 * b], x = Q[4 * b + 1], y = Q[4 * b + 2], z = Q[4 * b + 3];
  real n = sqrt(w * w + x * x + y * y + z * z);
  if (n > (real)0) {  // Eigen's normalize() leaves a zero quaternion untouched
    Q[4 * b] = w / n; Q[4 * b + 1] = x / n; Q[4 * b + 2] = y / n; Q[4 * b + 3] = z / n;
  }
}
template <typename real>
cudaError_t normalize_quats(real* Q, int n_bod, cudaStream_t s) {
  if (n_bod <= 0) return cudaSuccess;
  normalize_quats_kernel<real><<<(n_bod + 127) / 128, 128, 0, s>>>(Q, n_bod);
  return cudaGetLastError();
}

// r_{b,k} = R(q_b) ref_k + X_b   (get_r_vecs / single_body_pos / multi_body_pos, c_rigid_obj.cpp:257-300).
//
// A CTA owns G consecutive bodies and a 256-wide window of the 3 n_blb reals of a body.  The first G threads
// turn the G quaternions into rotation rows [R_c0 R_c1 R_c2 X_c] in shared memory (one 16/32-byte row per
// output component); every thread then owns ONE element (k, c) of the window -- its reference point ref_k sits
// in three registers -- and walks over the G bodies: one broadcast LDS.128 (three distinct rows per warp), three
// FMA, one coalesced STG per element, i.e. two memory instructions per element.
//
// History, all measured at 6.42 M blobs (profiles/r02_on_kernels.md; a fill of the same buffer takes 13.8 /
// 24.9 us in fp32 / fp64): thread per blob (~95 instructions per blob for a per-thread quaternion->rotation
// and an integer division, 12 partial sectors per store instruction, 72 % issue-busy) 27 / 47 us; CTA per body,
// thread per element re-reading ref and R per element (7 memory instructions per element: the LSU issue
// floor) 26 / 45 us; four consecutive outputs per thread with 128-bit stores 32 / 52 us; this form 18.3 /
// 31.2 us.
template <typename real>
struct alignas(16) Row4 {
  real x, y, z, w;
};
template <typename real, int G>
__global__ void place_blobs_rows_kernel(const real* __restrict__ X, const real* __restrict__ Q,
                                        const real* __restrict__ ref, unsigned n_bod, unsigned n3,
                                        real* __restrict__ r) {
  __shared__ Row4<real> RX[G][3];
  const unsigned b0 = blockIdx.x * G;
  const unsigned nb = n_bod - b0 < (unsigned)G ? n_bod - b0 : (unsigned)G;
  if (threadIdx.x < nb) {
    const size_t b = (size_t)b0 + threadIdx.x;
    real R[9];
    quat_to_rot(Q + 4 * b, R);
#pragma unroll
    for (int c = 0; c < 3; ++c) RX[threadIdx.x][c] = Row4<real>{R[3 * c], R[3 * c + 1], R[3 * c + 2], X[3 * b + c]};
  }
  __syncthreads();
  const unsigned e = blockIdx.y * blockDim.x + threadIdx.x;  // element of the body: blob k, component c
  if (e >= n3) return;
  const unsigned k = e / 3u, c = e - 3u * k;
  const real cx = ref[3 * k], cy = ref[3 * k + 1], cz = ref[3 * k + 2];
  real* __restrict__ out = r + (size_t)b0 * n3 + e;
#pragma unroll 4
  for (unsigned g = 0; g < nb; ++g) {
    const Row4<real> row = RX[g][c];
    out[(size_t)g * n3] = fma(row.x, cx, fma(row.y, cy, fma(row.z, cz, row.w)));
  }
}

template <typename real>
cudaError_t place_blobs(const real* X, const real* Q, const real* ref, int n_bod, int n_blb,
                        real* r, cudaStream_t s) {
  const long long n = (long long)n_bod * n_blb;
  if (n <= 0) return cudaSuccess;
  if (n > 0x7fffffffLL / 3) return cudaErrorInvalidValue;
  const unsigned n3 = 3u * (unsigned)n_blb;
  const unsigned threads = n3 >= 256u ? 256u : ((n3 + 31u) / 32u) * 32u;
  const unsigned wins = (n3 + threads - 1) / threads;
  if (wins > 65535u) return cudaErrorInvalidValue;  // (5.5 M blobs in ONE body)
  // 32 bodies per CTA amortise the reference point and the rotation set-up; 8 when that would leave SMs idle
  if ((unsigned long long)((n_bod + 31) / 32) * wins >= 592ull)
    place_blobs_rows_kernel<real, 32><<<dim3((unsigned)(n_bod + 31) / 32, wins), threads, 0, s>>>(X, Q, ref, (unsigned)n_bod, n3, r);
  else
    place_blobs_rows_kernel<real, 8><<<dim3((unsigned)(n_bod + 7) / 8, wins), threads, 0, s>>>(X, Q, ref, (unsigned)n_bod, n3, r);
  return cudaGetLastError();
}

template <typename real>
__global__ void k_dot_kernel(const real* __restrict__ U, const real* __restrict__ r,
                             const real* __restrict__ X, int n_bod, int n_blb, real sign,
                             const real* add, real* out) {  // add may alias out
  const unsigned iu = blockIdx.x * blockDim.x + threadIdx.x;  // 32-bit: a 64-bit division is a ~100-instruction routine; 3N < 2^31 is checked by the launcher
  if (iu >= (unsigned)n_bod * (unsigned)n_blb) return;
  const unsigned b = iu / (unsigned)n_blb;
  const size_t i = iu;
  const real* u = U + 6 * (size_t)b;
  const real px = r[3 * i] - X[3 * (size_t)b], py = r[3 * i + 1] - X[3 * (size_t)b + 1],
             pz = r[3 * i + 2] - X[3 * (size_t)b + 2];
  real vx = u[0] + (u[4] * pz - u[5] * py);
  real vy = u[1] + (u[5] * px - u[3] * pz);
  real vz = u[2] + (u[3] * py - u[4] * px);
  vx *= sign; vy *= sign; vz *= sign;
  if (add) { vx += add[3 * i]; vy += add[3 * i + 1]; vz += add[3 * i + 2]; }
  out[3 * i] = vx; out[3 * i + 1] = vy; out[3 * i + 2] = vz;
}
template <typename real>
cudaError_t k_dot(const real* U, const real* r, const real* X, int n_bod, int n_blb,
                  real sign, const real* add, real* out, cudaStream_t s) {
  const long long n = (long long)n_bod * n_blb;
  if (n <= 0) return cudaSuccess;
  if (n > 0x7fffffffLL / 3) return cudaErrorInvalidValue;
  k_dot_kernel<real><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(U, r, X, n_bod, n_blb, sign, add, out);
  return cudaGetLastError();
}

template <typename real>
__global__ void kt_dot_kernel(const real* __restrict__ lam, const real* __restrict__ r,
                              const real* __restrict__ X, int n_blb, real* __restrict__ out) {
  __shared__ real red[32 * 6];
  const int b = blockIdx.x;
  const real X0 = X[3 * (size_t)b], X1 = X[3 * (size_t)b + 1], X2 = X[3 * (size_t)b + 2];
  real v[6] = {0, 0, 0, 0, 0, 0};
  for (int k = threadIdx.x; k < n_blb; k += blockDim.x) {
    const size_t i = (size_t)b * n_blb + k;
    const real lx = lam[3 * i], ly = lam[3 * i + 1], lz = lam[3 * i + 2];
    const real px = r[3 * i] - X0, py = r[3 * i + 1] - X1, pz = r[3 * i + 2] - X2;
    v[0] += lx; v[1] += ly; v[2] += lz;
    v[3] += py * lz - pz * ly;
    v[4] += pz * lx - px * lz;
    v[5] += px * ly - py * lx;
  }
  block_sum<real, 6>(v, red);
  if (threadIdx.x < 6) out[6 * (size_t)b + threadIdx.x] = red[threadIdx.x];
}
static inline int body_block(int n_blb) { return n_blb <= 32 ? 32 : n_blb <= 64 ? 64 : n_blb <= 128 ? 128 : 256; }
template <typename real>
cudaError_t kt_dot(const real* lam, const real* r, const real* X, int n_bod, int n_blb,
                   real* out, cudaStream_t s) {
  if (n_bod <= 0) return cudaSuccess;
  kt_dot_kernel<real><<<n_bod, body_block(n_blb), 0, s>>>(lam, r, X, n_blb, out);
  return cudaGetLastError();
}

template <typename real>
__global__ void ktk_inv_blocks_kernel(const real* __restrict__ Q, const real* __restrict__ ref,
                                      int n_bod, int n_blb, real* __restrict__ S,
                                      int* __restrict__ singular) {
  // every thread recomputes the (tiny) reference-shape moments; n_blb is small
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bod) return;
  double sumr2 = 0, m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < n_blb; ++k) {
    const double x = ref[3 * k], y = ref[3 * k + 1], z = ref[3 * k + 2];
    sumr2 += x * x + y * y + z * z;
    m[0] += x * x; m[1] += x * y; m[2] += x * z;
    m[4] += y * y; m[5] += y * z; m[8] += z * z;
  }
  m[3] = m[1]; m[6] = m[2]; m[7] = m[5];
  real Rr[9];
  quat_to_rot(Q + 4 * (size_t)b, Rr);
  double R[9], T[9], D[9];
  for (int i = 0; i < 9; ++i) R[i] = Rr[i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) T[3 * i + j] = R[3 * i] * m[j] + R[3 * i + 1] * m[3 + j] + R[3 * i + 2] * m[6 + j];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      D[3 * i + j] = (i == j ? sumr2 : 0.0) - (T[3 * i] * R[3 * j] + T[3 * i + 1] * R[3 * j + 1] + T[3 * i + 2] * R[3 * j + 2]);
  const double c00 = D[4] * D[8] - D[5] * D[7], c01 = D[5] * D[6] - D[3] * D[8], c02 = D[3] * D[7] - D[4] * D[6];
  const double det = D[0] * c00 + D[1] * c01 + D[2] * c02;
  if (det < 1.0e-13) *singular = 1;  // c_rigid_obj.cpp:312-316
  const double id = 1.0 / det;
  real* o = S + 9 * (size_t)b;
  o[0] = (real)(c00 * id); o[1] = (real)((D[2] * D[7] - D[1] * D[8]) * id); o[2] = (real)((D[1] * D[5] - D[2] * D[4]) * id);
  o[3] = (real)(c01 * id); o[4] = (real)((D[0] * D[8] - D[2] * D[6]) * id); o[5] = (real)((D[2] * D[3] - D[0] * D[5]) * id);
  o[6] = (real)(c02 * id); o[7] = (real)((D[1] * D[6] - D[0] * D[7]) * id); o[8] = (real)((D[0] * D[4] - D[1] * D[3]) * id);
}
template <typename real>
cudaError_t ktk_inv_blocks(const real* Q, const real* ref, int n_bod, int n_blb, real* S,
                           int* singular, cudaStream_t s) {
  if (n_bod <= 0) return cudaSuccess;
  ktk_inv_blocks_kernel<real><<<(n_bod + 127) / 128, 128, 0, s>>>(Q, ref, n_bod, n_blb, S, singular);
  return cudaGetLastError();
}

template <typename real>
__global__ void ktk_inv_apply_kernel(const real* __restrict__ S, int n_bod, real inv_nblb,
                                     real* __restrict__ v) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bod) return;
  real* p = v + 6 * (size_t)b;
  const real* m = S + 9 * (size_t)b;
  const real t0 = p[3], t1 = p[4], t2 = p[5];
  p[0] *= inv_nblb; p[1] *= inv_nblb; p[2] *= inv_nblb;
  p[3] = m[0] * t0 + m[1] * t1 + m[2] * t2;
  p[4] = m[3] * t0 + m[4] * t1 + m[5] * t2;
  p[5] = m[6] * t0 + m[7] * t1 + m[8] * t2;
}
template <typename real>
cudaError_t ktk_inv_apply(const real* S, int n_bod, int n_blb, real* v, cudaStream_t s) {
  if (n_bod <= 0) return cudaSuccess;
  ktk_inv_apply_kernel<real><<<(n_bod + 127) / 128, 128, 0, s>>>(S, n_bod, (real)1 / (real)n_blb, v);
  return cudaGetLastError();
}

// ----------------------------------------------------------------------------------
// preconditioner
// ----------------------------------------------------------------------------------
template <typename real>
__global__ void pc_diag_build_kernel(const real* __restrict__ r, int n, real a, real scale,
                                     int wall, real* __restrict__ dinv, int* __restrict__ below) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const real f43 = (real)4 / (real)3;
  real mxx = f43, mzz = f43;
  if (wall) {
    const real h = r[3 * (size_t)i + 2] / a;
    if (h < (real)0) *below = 1;  // c_rigid_obj.cpp:95-97
    const real inv = (real)1 / h, i3 = inv * inv * inv, i5 = i3 * inv * inv;
    mxx += -(9 * inv - 2 * i3 + i5) / (real)12;  // :102-103
    mzz += -(9 * inv - 4 * i3 + i5) / (real)6;   // :104
  }
  dinv[3 * (size_t)i] = scale / mxx;  // inverse of a diagonal 3x3 (:524), times 8 pi eta a (:540)
  dinv[3 * (size_t)i + 1] = scale / mxx;
  dinv[3 * (size_t)i + 2] = scale / mzz;
}
template <typename real>
cudaError_t pc_diag_build(const real* r, int n, real a, real eta, bool wall, real* dinv,
                          int* below, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  const real scale = (real)(8.0 * M_PI * (double)eta * (double)a);
  pc_diag_build_kernel<real><<<(n + 255) / 256, 256, 0, s>>>(r, n, a, scale, wall ? 1 : 0, dinv, below);
  return cudaGetLastError();
}

// index type: 32-bit whenever the element count allows (64-bit divisions are library calls)
template <typename real, typename idx_t>
__global__ void pc_diag_mul_kernel(const real* __restrict__ dinv, const real* __restrict__ in,
                                   idx_t sz, idx_t ncols, idx_t total, real* __restrict__ out) {
  const idx_t i = (idx_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const idx_t per_body = sz * ncols;
  const idx_t b = i / per_body;
  const idx_t row = (i - b * per_body) % sz;
  out[i] = dinv[b * sz + row] * in[i];
}
template <typename real>
cudaError_t pc_diag_mul(const real* dinv, const real* in, int n_bod, int n_blb, int ncols,
                        real* out, cudaStream_t s) {
  const long long total = (long long)n_bod * 3 * n_blb * ncols;
  if (total <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  if (total < 0x7fffffffLL)
    pc_diag_mul_kernel<real, unsigned><<<blocks, 256, 0, s>>>(dinv, in, 3u * n_blb, (unsigned)ncols, (unsigned)total, out);
  else
    pc_diag_mul_kernel<real, unsigned long long><<<blocks, 256, 0, s>>>(dinv, in, 3ull * n_blb, (unsigned long long)ncols,
                                                                     (unsigned long long)total, out);
  return cudaGetLastError();
}

template <typename real>
__global__ void pc_fill_kcols_kernel(const real* __restrict__ r, const real* __restrict__ X,
                                     int n_bod, int n_blb, real* __restrict__ Kc) {
  const unsigned iu = blockIdx.x * blockDim.x + threadIdx.x;  // 32-bit: a 64-bit division is a ~100-instruction routine; 3N < 2^31 is checked by the launcher
  if (iu >= (unsigned)n_bod * (unsigned)n_blb) return;
  const unsigned b = iu / (unsigned)n_blb, k = iu - b * (unsigned)n_blb;
  const size_t i = iu;
  const int sz = 3 * n_blb;
  const real px = r[3 * i] - X[3 * (size_t)b], py = r[3 * i + 1] - X[3 * (size_t)b + 1],
             pz = r[3 * i + 2] - X[3 * (size_t)b + 2];
  real* o = Kc + (size_t)b * 6 * sz + 3 * k;
  // rows of K for blob k (c_rigid_obj.cpp:370-382): [I | -[rho]x]
  const real col[6][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, -pz, py}, {pz, 0, -px}, {-py, px, 0}};
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    o[(size_t)c * sz + 0] = col[c][0];
    o[(size_t)c * sz + 1] = col[c][1];
    o[(size_t)c * sz + 2] = col[c][2];
  }
}
template <typename real>
cudaError_t pc_fill_kcols(const real* r, const real* X, int n_bod, int n_blb, real* Kc,
                          cudaStream_t s) {
  const long long n = (long long)n_bod * n_blb;
  if (n <= 0) return cudaSuccess;
  pc_fill_kcols_kernel<real><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(r, X, n_bod, n_blb, Kc);
  return cudaGetLastError();
}

// Dense mobility block of one body, M_b (3 n_blb x 3 n_blb), from the SAME pair arithmetic as the
// product kernels (rbl::pair, rbl_pair.cuh): column q of the 3x3 block (i, j) is the velocity of blob
// i for a unit force e_q on blob j.  Upper triangle evaluated with the source's height and mirrored
// transposed, like the reference's assembly loop (c_rigid_obj.cpp:430-452).  Set-up path, not hot.
template <typename real, bool WALL>
__global__ void pc_block_assemble_kernel(const real* __restrict__ r, int n_blb, const PairConsts<real> C,
                                         real* __restrict__ M, int* __restrict__ below) {
  const int b = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_blb * n_blb) return;
  const int i = idx / n_blb, j = idx - i * n_blb;
  if (i > j) return;
  const real* rb = r + 3 * (size_t)b * n_blb;
  const real xi = rb[3 * i], yi = rb[3 * i + 1], zi = rb[3 * i + 2];
  const real xj = rb[3 * j], yj = rb[3 * j + 1], zj = rb[3 * j + 2];
  if (WALL && i == j && zi < (real)0) *below = 1;
  const int sz = 3 * n_blb;
  real* Mb = M + (size_t)b * sz * sz;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    real u[3] = {0, 0, 0};
    pair<real, WALL, true>(C, xi, yi, zi, xj, yj, zj, q == 0 ? (real)1 : (real)0, q == 1 ? (real)1 : (real)0,
                           q == 2 ? (real)1 : (real)0, (real)2 * zj, (real)4 * zj * zj, u[0], u[1], u[2]);
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const real v = u[p] * C.out_scale;
      Mb[(size_t)(3 * i + p) * sz + 3 * j + q] = v;
      if (i != j) Mb[(size_t)(3 * j + q) * sz + 3 * i + p] = v;
    }
  }
}
template <typename real>
cudaError_t pc_block_assemble(const real* r, int count, int n_blb, real a, real eta,
                              bool wall, real* M, int* below, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const PairConsts<real> C = make_pair_consts<real>((double)a, (double)eta);
  dim3 grid((n_blb * n_blb + 255) / 256, count);
  if (wall)
    pc_block_assemble_kernel<real, true><<<grid, 256, 0, s>>>(r, n_blb, C, M, below);
  else
    pc_block_assemble_kernel<real, false><<<grid, 256, 0, s>>>(r, n_blb, C, M, below);
  return cudaGetLastError();
}

// In-place Gauss-Jordan inverse WITH partial (row) pivoting, one CTA per matrix.  This is the path for
// body blocks that are NOT positive definite (blobs inside the wall-overlap layer; the reference inverts
// the same block with Eigen's pivoted Mob.inverse(), c_rigid_obj.cpp:475): at step k the row with the
// largest |A[i][k]|, i >= k, is swapped into place; the column swaps that undo the permutation are
// applied in reverse order at the end.  A zero or non-finite pivot raises the flag.
template <typename real>
__global__ void pc_block_invert_kernel(real* __restrict__ M, int sz, int* __restrict__ singular) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real* row = reinterpret_cast<real*>(smem_raw);
  real* col = row + sz;
  int* perm = reinterpret_cast<int*>(col + sz);
  __shared__ real s_val[32];
  __shared__ int s_idx[32];
  __shared__ int s_piv;
  real* A = M + (size_t)blockIdx.x * sz * sz;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = 0; k < sz; ++k) {
    // pivot search in column k
    real best = (real)-1;
    int bi = k;
    for (int i = k + threadIdx.x; i < sz; i += blockDim.x) {
      const real v = fabs(A[(size_t)i * sz + k]);
      if (v > best) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const real ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { s_val[wid] = best; s_idx[wid] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      real bv = s_val[0];
      int bx = s_idx[0];
      for (int w = 1; w < nw; ++w)
        if (s_val[w] > bv || (s_val[w] == bv && s_idx[w] < bx)) { bv = s_val[w]; bx = s_idx[w]; }
      if (!(bv > (real)0) || !isfinite(bv)) *singular = 1;
      s_piv = bx;
      perm[k] = bx;
    }
    __syncthreads();
    const int pr = s_piv;
    if (pr != k) {
      for (int j = threadIdx.x; j < sz; j += blockDim.x) {
        const real t = A[(size_t)k * sz + j];
        A[(size_t)k * sz + j] = A[(size_t)pr * sz + j];
        A[(size_t)pr * sz + j] = t;
      }
      __syncthreads();
    }
    for (int j = threadIdx.x; j < sz; j += blockDim.x) {
      row[j] = A[(size_t)k * sz + j];
      col[j] = A[(size_t)j * sz + k];
    }
    __syncthreads();
    const real p = (real)1 / row[k];
    for (int i = wid; i < sz; i += nw) {
      real* Ai = A + (size_t)i * sz;
      if (i == k) {
        for (int j = lane; j < sz; j += 32) Ai[j] = (j == k) ? p : row[j] * p;
      } else {
        const real f = col[i] * p;
        for (int j = lane; j < sz; j += 32) Ai[j] = (j == k) ? -f : Ai[j] - f * row[j];
      }
    }
    __syncthreads();
  }
  for (int k = sz - 1; k >= 0; --k) {  // A^-1 = (P A)^-1 P: undo the row swaps as column swaps, last first
    const int pr = perm[k];
    if (pr != k) {
      for (int i = threadIdx.x; i < sz; i += blockDim.x) {
        const real t = A[(size_t)i * sz + k];
        A[(size_t)i * sz + k] = A[(size_t)i * sz + pr];
        A[(size_t)i * sz + pr] = t;
      }
      __syncthreads();
    }
  }
}
template <typename real>
cudaError_t pc_block_invert(real* M, int count, int sz, int* not_spd, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  const size_t smem = 2 * (size_t)sz * sizeof(real) + (size_t)sz * sizeof(int);
  const int threads = sz <= 128 ? 256 : 1024;
  cudaError_t e = cudaFuncSetAttribute(pc_block_invert_kernel<real>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  pc_block_invert_kernel<real><<<count, threads, smem, s>>>(M, sz, not_spd);
  return cudaGetLastError();
}

// out[b][c][:] = Minv_b in[b][c][:]; 192 output rows per CTA (blob aligned), thread per row,
// symmetric matrix read column-wise so the loads coalesce.
template <typename real>
__global__ void pc_block_mul_kernel(const real* __restrict__ Minv, size_t stride,
                                    const real* __restrict__ Q, const real* __restrict__ in,
                                    int sz, int ncols, real* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real* x = reinterpret_cast<real*>(smem_raw);  // sz
  __shared__ real tile[192];
  const int b = blockIdx.y;
  const int row = blockIdx.x * 192 + threadIdx.x;
  const real* A = Minv + stride * b;
  const bool rot = (stride == 0) && (Q != nullptr);
  real R[9];
  if (rot) quat_to_rot(Q + 4 * (size_t)b, R);
  for (int c = 0; c < ncols; ++c) {
    const real* xin = in + ((size_t)b * ncols + c) * sz;
    if (rot) {  // body frame: x = R^T x_lab per blob
      for (int k = threadIdx.x; k < sz / 3; k += blockDim.x) {
        const real vx = xin[3 * k], vy = xin[3 * k + 1], vz = xin[3 * k + 2];
        x[3 * k + 0] = R[0] * vx + R[3] * vy + R[6] * vz;
        x[3 * k + 1] = R[1] * vx + R[4] * vy + R[7] * vz;
        x[3 * k + 2] = R[2] * vx + R[5] * vy + R[8] * vz;
      }
    } else {
      for (int j = threadIdx.x; j < sz; j += blockDim.x) x[j] = xin[j];
    }
    __syncthreads();
    real acc = 0;
    if (row < sz) {
#pragma unroll 8
      for (int j = 0; j < sz; ++j) acc += A[(size_t)j * sz + row] * x[j];
    }
    real* o = out + ((size_t)b * ncols + c) * sz;
    if (rot) {
      tile[threadIdx.x] = acc;
      __syncthreads();
      if (row < sz) {
        const int k3 = (threadIdx.x / 3) * 3, p = threadIdx.x - k3;
        o[row] = R[3 * p] * tile[k3] + R[3 * p + 1] * tile[k3 + 1] + R[3 * p + 2] * tile[k3 + 2];
      }
    } else if (row < sz) {
      o[row] = acc;
    }
    __syncthreads();
  }
}
template <typename real>
cudaError_t pc_block_mul(const real* Minv, size_t stride, const real* Q, const real* in,
                         int n_bod, int n_blb, int ncols, real* out, cudaStream_t s) {
  if (n_bod <= 0) return cudaSuccess;
  const int sz = 3 * n_blb;
  const size_t smem = (size_t)sz * sizeof(real);
  cudaError_t e = cudaFuncSetAttribute(pc_block_mul_kernel<real>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((sz + 191) / 192, n_bod);
  pc_block_mul_kernel<real><<<grid, 192, smem, s>>>(Minv, stride, Q, in, sz, ncols, out);
  return cudaGetLastError();
}

template <typename real>
__global__ void pc_ninv_chol_kernel(const real* __restrict__ Y, const real* __restrict__ r,
                                    const real* __restrict__ X, int n_blb, real* __restrict__ L,
                                    int* __restrict__ not_spd) {
  __shared__ real red[32 * 6];
  __shared__ real N[36];
  const int b = blockIdx.x, sz = 3 * n_blb;
  const real X0 = X[3 * (size_t)b], X1 = X[3 * (size_t)b + 1], X2 = X[3 * (size_t)b + 2];
  for (int c = 0; c < 6; ++c) {
    const real* y = Y + ((size_t)b * 6 + c) * sz;
    real v[6] = {0, 0, 0, 0, 0, 0};
    for (int k = threadIdx.x; k < n_blb; k += blockDim.x) {
      const size_t i = (size_t)b * n_blb + k;
      const real lx = y[3 * k], ly = y[3 * k + 1], lz = y[3 * k + 2];
      const real px = r[3 * i] - X0, py = r[3 * i + 1] - X1, pz = r[3 * i + 2] - X2;
      v[0] += lx; v[1] += ly; v[2] += lz;
      v[3] += py * lz - pz * ly;
      v[4] += pz * lx - px * lz;
      v[5] += px * ly - py * lx;
    }
    block_sum<real, 6>(v, red);
    if (threadIdx.x < 6) N[threadIdx.x * 6 + c] = red[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    // symmetrise (K^T Minv K is symmetric up to rounding), factor
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < i; ++j) {
        const real m = (real)0.5 * (N[i * 6 + j] + N[j * 6 + i]);
        N[i * 6 + j] = m;
        N[j * 6 + i] = m;
      }
    if (!inv6(N)) *not_spd = 1;  // N_b = (K^T Mt^-1 K)^-1, the body mobility of the PC
    for (int i = 0; i < 36; ++i) L[36 * (size_t)b + i] = N[i];
  }
}
template <typename real>
cudaError_t pc_ninv_chol(const real* Y, const real* r, const real* X, int n_bod, int n_blb,
                         real* L, int* not_spd, cudaStream_t s) {
  if (n_bod <= 0) return cudaSuccess;
  pc_ninv_chol_kernel<real><<<n_bod, body_block(n_blb), 0, s>>>(Y, r, X, n_blb, L, not_spd);
  return cudaGetLastError();
}

template <typename real>
__global__ void pc_finish_kernel(const real* __restrict__ y, const real* __restrict__ F,
                                 const real* __restrict__ Y, const real* __restrict__ L,
                                 const real* __restrict__ r, const real* __restrict__ X,
                                 int n_bod, int n_blb, real* __restrict__ out) {
  __shared__ real red[32 * 6];
  __shared__ real Ub[6];
  const int b = blockIdx.x, sz = 3 * n_blb;
  const real X0 = X[3 * (size_t)b], X1 = X[3 * (size_t)b + 1], X2 = X[3 * (size_t)b + 2];
  const real* yb = y + (size_t)b * sz;
  real v[6] = {0, 0, 0, 0, 0, 0};
  for (int k = threadIdx.x; k < n_blb; k += blockDim.x) {
    const size_t i = (size_t)b * n_blb + k;
    const real lx = yb[3 * k], ly = yb[3 * k + 1], lz = yb[3 * k + 2];
    const real px = r[3 * i] - X0, py = r[3 * i + 1] - X1, pz = r[3 * i + 2] - X2;
    v[0] += lx; v[1] += ly; v[2] += lz;
    v[3] += py * lz - pz * ly;
    v[4] += pz * lx - px * lz;
    v[5] += px * ly - py * lx;
  }
  block_sum<real, 6>(v, red);
  if (threadIdx.x == 0) {
    real rhs[6];
    for (int c = 0; c < 6; ++c) rhs[c] = -F[6 * (size_t)b + c] - red[c];  // :601
    const real* Nb = L + 36 * (size_t)b;                                   // :605-608
    for (int c = 0; c < 6; ++c) {
      real acc = 0;
      for (int d = 0; d < 6; ++d) acc += Nb[c * 6 + d] * rhs[d];
      Ub[c] = acc;
    }
  }
  __syncthreads();
  const size_t n3 = (size_t)n_bod * sz;
  const real* Yb = Y + (size_t)b * 6 * sz;
  for (int i = threadIdx.x; i < sz; i += blockDim.x) {
    real acc = yb[i];
#pragma unroll
    for (int c = 0; c < 6; ++c) acc += Yb[(size_t)c * sz + i] * Ub[c];  // Mt^-1 (slip + K U), :610
    out[(size_t)b * sz + i] = acc;
  }
  if (threadIdx.x < 6) out[n3 + 6 * (size_t)b + threadIdx.x] = Ub[threadIdx.x];
}
template <typename real>
cudaError_t pc_finish(const real* y, const real* F, const real* Y, const real* L,
                      const real* r, const real* X, int n_bod, int n_blb, real* out,
                      cudaStream_t s) {
  if (n_bod <= 0) return cudaSuccess;
  pc_finish_kernel<real><<<n_bod, body_block(n_blb), 0, s>>>(y, F, Y, L, r, X, n_bod, n_blb, out);
  return cudaGetLastError();
}


// ----------------------------------------------------------------------------------
// noise preconditioner: per-body Cholesky factor of the body's own mobility block
// ----------------------------------------------------------------------------------
// In-place lower Cholesky of `count` sz x sz matrices, blocked right-looking with 32-column panels:
// per panel (1) the 32 x 32 diagonal tile is factored in shared memory, (2) the rows below solve
// x Ld^T = a against it (a thread per row, the row's 32 entries in registers), (3) the trailing
// matrix takes a rank-32 update in 64 x 64 tiles (panel rows of the tile's i- and j-range staged
// in shared memory, a 4 x 4 register block per thread).  Every kernel is batched over the
// matrices (grid.y), so one shared 7686 x 7686 factor and 1250 per-body 1926 x 1926 factors both
// fill the machine; the trailing matrix is read and written once per PANEL instead of once per
// column (a one-CTA-per-matrix rank-1 version took 0.7 s for 200 bodies of 1926 x 1926 and 20 s for
// the single 7686 x 7686 one).  The strict upper triangle is zeroed at the end so the factor can be
// used as a plain dense matrix.  Flags a non-positive pivot.
constexpr int kCholNB = 32;

template <typename real>
__global__ void __launch_bounds__(32) chol_diag_kernel(real* __restrict__ M, int sz, int k0, int* __restrict__ not_spd) {
  __shared__ real T[kCholNB][kCholNB + 1];
  real* A = M + (size_t)blockIdx.x * sz * sz;
  const int kb = min(kCholNB, sz - k0), lane = threadIdx.x;
  for (int r = 0; r < kb; ++r)
    if (lane < kb) T[r][lane] = A[(size_t)(k0 + r) * sz + k0 + lane];
  __syncwarp();
  for (int c = 0; c < kb; ++c) {
    const real d = T[c][c];
    if (lane == 0 && (!(d > (real)0) || !isfinite(d))) *not_spd = 1;
    const real piv = sqrt(d > (real)0 ? d : (real)1);
    __syncwarp();
    if (lane >= c && lane < kb) T[lane][c] = (lane == c) ? piv : T[lane][c] / piv;  // lane = row
    __syncwarp();
    // trailing part of the tile: row = lane, columns c+1..lane
    if (lane > c && lane < kb) {
      const real l = T[lane][c];
      for (int j = c + 1; j <= lane; ++j) T[lane][j] -= l * T[j][c];
    }
    __syncwarp();
  }
  for (int r = 0; r < kb; ++r)
    if (lane < kb) A[(size_t)(k0 + r) * sz + k0 + lane] = (lane <= r) ? T[r][lane] : (real)0;
}

template <typename real>
__global__ void __launch_bounds__(128) chol_panel_kernel(real* __restrict__ M, int sz, int k0) {
  __shared__ real Ld[kCholNB][kCholNB + 1];
  real* A = M + (size_t)blockIdx.y * sz * sz;
  const int kb = min(kCholNB, sz - k0);
  for (int t = threadIdx.x; t < kCholNB * kCholNB; t += blockDim.x) {
    const int r = t / kCholNB, c = t - r * kCholNB;
    Ld[r][c] = (r < kb && c < kb) ? A[(size_t)(k0 + r) * sz + k0 + c] : (real)(r == c ? 1 : 0);
  }
  __syncthreads();
  const int i = k0 + kb + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sz) return;
  real* row = A + (size_t)i * sz + k0;
  real x[kCholNB];
#pragma unroll
  for (int c = 0; c < kCholNB; ++c) x[c] = c < kb ? row[c] : (real)0;
#pragma unroll
  for (int c = 0; c < kCholNB; ++c) {
    real v = x[c];
#pragma unroll
    for (int d = 0; d < c; ++d) v = fma(-x[d], Ld[c][d], v);
    x[c] = v / Ld[c][c];
  }
#pragma unroll
  for (int c = 0; c < kCholNB; ++c)
    if (c < kb) row[c] = x[c];
}

// trailing update, lower triangle only: tile (ti, tj), tj <= ti, of the (sz - k1) x (sz - k1) trailing matrix
template <typename real>
__global__ void __launch_bounds__(256) chol_update_kernel(real* __restrict__ M, int sz, int k0, int k1, int ntile) {
  __shared__ real Pi[64][kCholNB + 1], Pj[64][kCholNB + 1];
  real* A = M + (size_t)blockIdx.y * sz * sz;
  // linear tile index -> (ti, tj) with tj <= ti
  int ti = (int)((sqrtf(8.0f * (float)blockIdx.x + 1.0f) - 1.0f) * 0.5f);
  while ((ti + 1) * (ti + 2) / 2 <= (int)blockIdx.x) ++ti;
  while (ti * (ti + 1) / 2 > (int)blockIdx.x) --ti;
  const int tj = (int)blockIdx.x - ti * (ti + 1) / 2;
  if (ti >= ntile) return;
  const int i0 = k1 + ti * 64, j0 = k1 + tj * 64;
  const int kb = k1 - k0;
  for (int t = threadIdx.x; t < 64 * kCholNB; t += blockDim.x) {
    const int r = t / kCholNB, c = t - r * kCholNB;
    Pi[r][c] = (i0 + r < sz && c < kb) ? A[(size_t)(i0 + r) * sz + k0 + c] : (real)0;
    Pj[r][c] = (j0 + r < sz && c < kb) ? A[(size_t)(j0 + r) * sz + k0 + c] : (real)0;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x 4 block each
  real acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = (real)0;
#pragma unroll 8
  for (int c = 0; c < kCholNB; ++c) {
    real pi[4], pj[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) { pi[a] = Pi[ty + 16 * a][c]; pj[a] = Pj[tx + 16 * a][c]; }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = fma(pi[a], pj[b], acc[a][b]);
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i < sz && j <= i) A[(size_t)i * sz + j] -= acc[a][b];
    }
}

template <typename real>
__global__ void zero_upper_kernel(real* __restrict__ M, int sz) {
  real* A = M + (size_t)blockIdx.y * sz * sz;
  const size_t total = (size_t)sz * sz;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / sz), j = (int)(idx - (size_t)i * sz);
    if (j > i) A[idx] = (real)0;
  }
}

template <typename real>
cudaError_t chol_lower(real* M, int count, int sz, int* not_spd, cudaStream_t s) {
  if (count <= 0 || sz <= 0) return cudaSuccess;
  for (int k0 = 0; k0 < sz; k0 += kCholNB) {
    const int k1 = min(k0 + kCholNB, sz);
    chol_diag_kernel<real><<<count, 32, 0, s>>>(M, sz, k0, not_spd);
    const int below = sz - k1;
    if (below > 0) {
      chol_panel_kernel<real><<<dim3((below + 127) / 128, count), 128, 0, s>>>(M, sz, k0);
      const int nt = (below + 63) / 64;
      chol_update_kernel<real><<<dim3(nt * (nt + 1) / 2, count), 256, 0, s>>>(M, sz, k0, k1, nt);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  const int zb = (int)std::min<size_t>(((size_t)sz * sz + 255) / 256, 1024);
  zero_upper_kernel<real><<<dim3(zb, count), 256, 0, s>>>(M, sz);
  return cudaGetLastError();
}

// G = L^-1 for lower-triangular L (row-major).  A CTA owns 128 columns of G (a thread per column j:
// column j solves L g = e_j by forward substitution) and walks down the rows in blocks of 32:
// for a row block I the thread keeps 32 partial sums in registers and streams its finished entries
// G[k][j], k < 32 I, ONCE while the matching 32 x 32 tiles of L sit in shared memory (broadcast
// reads) -- 32 FMAs per global load.  (A row-at-a-time version re-read the finished column for
// every row: n^2/2 loads per column, L2-bound at 0.6 s for 200 bodies of 1926 x 1926.)  Then the
// 32 x 32 diagonal tile is solved in registers.  G's strict upper triangle is written as zero.
constexpr int kTriRows = 32, kTriCols = 128;
template <typename real>
__global__ void __launch_bounds__(kTriCols) tri_inverse_kernel(const real* __restrict__ Lm, real* __restrict__ Gm, int sz) {
  __shared__ real Lt[kTriRows][kTriRows + 1];
  const real* L = Lm + (size_t)blockIdx.y * sz * sz;
  real* G = Gm + (size_t)blockIdx.y * sz * sz;
  const int j_lo = blockIdx.x * kTriCols;
  const int j = j_lo + threadIdx.x;
  const bool live = j < sz;
  const int nblk = (sz + kTriRows - 1) / kTriRows;
  const int first_blk = j_lo / kTriRows;  // rows above the CTA's first column are zero for every thread
  for (int I = 0; I < first_blk; ++I)
    for (int ii = 0; ii < kTriRows; ++ii)
      if (live) G[(size_t)(I * kTriRows + ii) * sz + j] = (real)0;
  for (int I = first_blk; I < nblk; ++I) {
    const int i0 = I * kTriRows;
    real acc[kTriRows];
#pragma unroll
    for (int ii = 0; ii < kTriRows; ++ii) acc[ii] = (real)0;
    // off-diagonal tiles: acc[ii] += sum_{k < i0} L[i0+ii][k] G[k][j]   (k >= j_lo only: G is zero above)
    for (int K = first_blk; K < I; ++K) {
      const int k0 = K * kTriRows;
      __syncthreads();
      for (int t = threadIdx.x; t < kTriRows * kTriRows; t += kTriCols) {
        const int ii = t / kTriRows, kk = t - ii * kTriRows;
        const int gi = i0 + ii, gk = k0 + kk;
        Lt[ii][kk] = (gi < sz && gk < sz) ? L[(size_t)gi * sz + gk] : (real)0;
      }
      __syncthreads();
      if (live) {
#pragma unroll 4
        for (int kk = 0; kk < kTriRows; ++kk) {
          const int gk = k0 + kk;
          const real g = (gk < sz && gk >= j) ? G[(size_t)gk * sz + j] : (real)0;
#pragma unroll
          for (int ii = 0; ii < kTriRows; ++ii) acc[ii] = fma(Lt[ii][kk], g, acc[ii]);
        }
      }
    }
    // diagonal tile
    __syncthreads();
    for (int t = threadIdx.x; t < kTriRows * kTriRows; t += kTriCols) {
      const int ii = t / kTriRows, kk = t - ii * kTriRows;
      const int gi = i0 + ii, gk = i0 + kk;
      Lt[ii][kk] = (gi < sz && gk < sz) ? L[(size_t)gi * sz + gk] : (real)(ii == kk ? 1 : 0);
    }
    __syncthreads();
    if (live) {
      real x[kTriRows];
#pragma unroll
      for (int ii = 0; ii < kTriRows; ++ii) {
        const int gi = i0 + ii;
        real v = (gi == j ? (real)1 : (real)0) - acc[ii];
#pragma unroll
        for (int kk = 0; kk < ii; ++kk) v = fma(-Lt[ii][kk], x[kk], v);
        v /= Lt[ii][ii];
        x[ii] = (gi >= j) ? v : (real)0;
        if (gi < sz) G[(size_t)gi * sz + j] = x[ii];
      }
    }
  }
}
template <typename real>
cudaError_t tri_inverse(const real* L, real* G, int count, int sz, cudaStream_t s) {
  if (count <= 0) return cudaSuccess;
  dim3 grid((sz + kTriCols - 1) / kTriCols, count);
  tri_inverse_kernel<real><<<grid, kTriCols, 0, s>>>(L, G, sz);
  return cudaGetLastError();
}

// out_b = op(A_b) x_b per body: op = A (trans = 0) or A^T (trans = 1); with ONE shared matrix
// (stride = 0, Q given) body b uses the rotated factor: rot_in applies R_b^T per blob before the
// product, rot_out applies R_b after it.  Thread per output row, x staged in shared memory.
template <typename real>
__global__ void body_mat_mul_kernel(const real* __restrict__ A0, size_t stride, const real* __restrict__ Q, int rot_in,
                                    int rot_out, int trans, const real* __restrict__ in, int sz, int ncols,
                                    real* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real* x = reinterpret_cast<real*>(smem_raw);  // sz
  __shared__ real tile[192];
  const int v = blockIdx.y;     // vector index: layout [body][column][sz]
  const int b = v / ncols;      // its body
  const int row = blockIdx.x * 192 + threadIdx.x;
  const real* A = A0 + stride * b;
  real R[9];
  if (rot_in || rot_out) quat_to_rot(Q + 4 * (size_t)b, R);
  const real* xin = in + (size_t)v * sz;
  if (rot_in) {
    for (int k = threadIdx.x; k < sz / 3; k += blockDim.x) {
      const real vx = xin[3 * k], vy = xin[3 * k + 1], vz = xin[3 * k + 2];
      x[3 * k + 0] = R[0] * vx + R[3] * vy + R[6] * vz;
      x[3 * k + 1] = R[1] * vx + R[4] * vy + R[7] * vz;
      x[3 * k + 2] = R[2] * vx + R[5] * vy + R[8] * vz;
    }
  } else {
    for (int j = threadIdx.x; j < sz; j += blockDim.x) x[j] = xin[j];
  }
  __syncthreads();
  real acc = 0;
  if (row < sz) {
    if (trans) {
#pragma unroll 8
      for (int j = 0; j < sz; ++j) acc += A[(size_t)j * sz + row] * x[j];
    } else {
      const real* Ar = A + (size_t)row * sz;
#pragma unroll 8
      for (int j = 0; j < sz; ++j) acc += Ar[j] * x[j];
    }
  }
  real* o = out + (size_t)v * sz;
  if (rot_out) {
    tile[threadIdx.x] = acc;
    __syncthreads();
    if (row < sz) {
      const int k3 = (threadIdx.x / 3) * 3, p = threadIdx.x - k3;
      o[row] = R[3 * p] * tile[k3] + R[3 * p + 1] * tile[k3 + 1] + R[3 * p + 2] * tile[k3 + 2];
    }
  } else if (row < sz) {
    o[row] = acc;
  }
}
template <typename real>
cudaError_t body_mat_mul(const real* A, size_t stride, const real* Q, bool rot_in, bool rot_out, bool trans,
                         const real* in, int n_bod, int n_blb, real* out, cudaStream_t s, int ncols) {
  if (n_bod <= 0 || ncols <= 0) return cudaSuccess;
  const int sz = 3 * n_blb;
  const size_t smem = (size_t)sz * sizeof(real);
  cudaError_t e = cudaFuncSetAttribute(body_mat_mul_kernel<real>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid((sz + 191) / 192, (unsigned)n_bod * ncols);
  body_mat_mul_kernel<real><<<grid, 192, smem, s>>>(A, stride, Q, rot_in ? 1 : 0, rot_out ? 1 : 0, trans ? 1 : 0, in, sz, ncols, out);
  return cudaGetLastError();
}

// ----------------------------------------------------------------------------------
// integrator (Q_from_Om :679-689, update_X_Q :691-710)
// ----------------------------------------------------------------------------------
template <typename real>
__global__ void integrate_kernel(const real* __restrict__ U, real scale, int n_bod,
                                 const real* X, const real* Q, real* Xo, real* Qo) {
  // X/Xo and Q/Qo may alias (in-place evolve): each thread reads its body, then writes it
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_bod) return;
  const real* u = U + 6 * (size_t)b;
  const real ox = u[3] * scale, oy = u[4] * scale, oz = u[5] * scale;
  const double th = (double)sqrt(ox * ox + oy * oy + oz * oz);  // Om.norm() in `real`, then double
  double rw = cos(th / 2.0), rx = 0, ry = 0, rz = 0;
  if (th > 1.0e-10) {
    const double sc = sin(th / 2.0) / th;
    rx = sc * ox; ry = sc * oy; rz = sc * oz;
  }
  real pw = (real)rw, px = (real)rx, py = (real)ry, pz = (real)rz;
  real n = sqrt(pw * pw + px * px + py * py + pz * pz);
  pw /= n; px /= n; py /= n; pz /= n;
  const real qw = Q[4 * (size_t)b], qx = Q[4 * (size_t)b + 1], qy = Q[4 * (size_t)b + 2], qz = Q[4 * (size_t)b + 3];
  real w = pw * qw - px * qx - py * qy - pz * qz;
  real x = pw * qx + px * qw + py * qz - pz * qy;
  real y = pw * qy - px * qz + py * qw + pz * qx;
  real z = pw * qz + px * qy - py * qx + pz * qw;
  n = sqrt(w * w + x * x + y * y + z * z);
  Qo[4 * (size_t)b] = w / n; Qo[4 * (size_t)b + 1] = x / n; Qo[4 * (size_t)b + 2] = y / n; Qo[4 * (size_t)b + 3] = z / n;
  Xo[3 * (size_t)b] = X[3 * (size_t)b] + u[0] * scale;
  Xo[3 * (size_t)b + 1] = X[3 * (size_t)b + 1] + u[1] * scale;
  Xo[3 * (size_t)b + 2] = X[3 * (size_t)b + 2] + u[2] * scale;
}
template <typename real>
cudaError_t integrate(const real* U, real scale, int n_bod, const real* X, const real* Q,
                      real* Xo, real* Qo, cudaStream_t s) {
  if (n_bod <= 0) return cudaSuccess;
  integrate_kernel<real><<<(n_bod + 127) / 128, 128, 0, s>>>(U, scale, n_bod, X, Q, Xo, Qo);
  return cudaGetLastError();
}

#define INST(real)                                                                                  \
  template cudaError_t normalize_quats<real>(real*, int, cudaStream_t);                             \
  template cudaError_t place_blobs<real>(const real*, const real*, const real*, int, int, real*,    \
                                         cudaStream_t);                                             \
  template cudaError_t k_dot<real>(const real*, const real*, const real*, int, int, real,           \
                                   const real*, real*, cudaStream_t);                               \
  template cudaError_t kt_dot<real>(const real*, const real*, const real*, int, int, real*,         \
                                    cudaStream_t);                                                  \
  template cudaError_t ktk_inv_blocks<real>(const real*, const real*, int, int, real*, int*,        \
                                            cudaStream_t);                                          \
  template cudaError_t ktk_inv_apply<real>(const real*, int, int, real*, cudaStream_t);             \
  template cudaError_t pc_diag_build<real>(const real*, int, real, real, bool, real*, int*,         \
                                           cudaStream_t);                                           \
  template cudaError_t pc_diag_mul<real>(const real*, const real*, int, int, int, real*,            \
                                         cudaStream_t);                                             \
  template cudaError_t pc_fill_kcols<real>(const real*, const real*, int, int, real*,               \
                                           cudaStream_t);                                           \
  template cudaError_t pc_block_assemble<real>(const real*, int, int, real, real, bool, real*,      \
                                               int*, cudaStream_t);                                 \
  template cudaError_t pc_block_invert<real>(real*, int, int, int*, cudaStream_t);                  \
  template cudaError_t pc_block_mul<real>(const real*, size_t, const real*, const real*, int, int,  \
                                          int, real*, cudaStream_t);                                \
  template cudaError_t pc_ninv_chol<real>(const real*, const real*, const real*, int, int, real*,   \
                                          int*, cudaStream_t);                                      \
  template cudaError_t pc_finish<real>(const real*, const real*, const real*, const real*,          \
                                       const real*, const real*, int, int, real*, cudaStream_t);    \
  template cudaError_t chol_lower<real>(real*, int, int, int*, cudaStream_t);                       \
  template cudaError_t tri_inverse<real>(const real*, real*, int, int, cudaStream_t);               \
  template cudaError_t body_mat_mul<real>(const real*, size_t, const real*, bool, bool, bool,       \
                                          const real*, int, int, real*, cudaStream_t, int);         \
  template cudaError_t integrate<real>(const real*, real, int, const real*, const real*, real*,     \
                                       real*, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace rbl
