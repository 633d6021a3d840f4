// rbl_krylov.cuh -- vector kernels for the device Krylov drivers (GMRES for the saddle
// system, Lanczos for (B M B)^{1/2} W).  The reference ships neither (SURVEY.md F2/F3:
// it expects scipy/pyamg to drive apply_saddle/apply_PC, and M_half_W is a dense Cholesky,
// /root/reference/src/c_rigid_obj.cpp:661-675); these are the device-side replacements.
#pragma once
#include <cuda_runtime.h>

namespace rbl {

// out[i] = sum_k V[i*ld + k] * w[k], i < m  (deterministic two-stage reduction;
// `partial` must hold m * kDotBlocks reals)
constexpr int kDotBlocks = 296;
template <typename real>
cudaError_t multi_dot(const real* V, size_t ld, int m, const real* w, size_t n, real* partial,
                      real* out, cudaStream_t s);
// w[k] += sum_i coef[i] * V[i*ld + k]   (coef on the DEVICE)
template <typename real>
cudaError_t multi_axpy(const real* V, size_t ld, int m, const real* coef, real sign, real* w,
                       size_t n, cudaStream_t s);
// y = a*x (+ y if accumulate)
template <typename real>
cudaError_t scale_copy(const real* x, real a, real* y, size_t n, bool accumulate, cudaStream_t s);
// y = x / sqrt(*s2)  (s2 on the DEVICE: a squared norm; y = 0 if it is not positive) -- lets a Krylov
// driver normalise the next basis vector without reading the norm back first
template <typename real>
cudaError_t scale_by_inv_sqrt(const real* x, const real* s2, real* y, size_t n, cudaStream_t s);
// y = (To)(a * x) (+ y if accumulate) between precisions (mixed-precision refinement: the float mirror
// of a double context)
template <typename From, typename To>
cudaError_t cast_scale(const From* x, double a, To* y, size_t n, bool accumulate, cudaStream_t s);
// y[k] = x[k] for k < n_head, -x[k] after (the sign flip between apply_saddle's and
// apply_PC's conventions)
template <typename real>
cudaError_t flip_tail(const real* x, size_t n_head, size_t n, real* y, cudaStream_t s);

// w1[k], w2[k], wr[k] = three independent standard normals that are pure functions of
// (seed, step, first + k): Philox4x32-10 per element + Box-Muller.  `first` is the GLOBAL index of
// element 0 (partitioned suspensions pass their offset, so the noise is partition invariant).
template <typename real>
cudaError_t normal_triplet(unsigned long long seed, unsigned long long step, unsigned long long first, size_t n,
                           real* w1, real* w2, real* wr, cudaStream_t s);

}  // namespace rbl
