// rbl_rigid.cuh -- O(N) rigid-body kernels around the mobility product.
//
// Reference members replaced (all /root/reference/src/c_rigid_obj.cpp):
//   setConfig normalisation :201-233      get_r_vecs..multi_body_pos :257-300
//   Make_K_Kinv/K_x_U/KT_x_Lam/Kinv_x_V/KTinv_x_F :302-410 (matrix-free here)
//   diag_invM :489-543, Block_diag_invM :461-487, get_blk_diag_lu :554-567,
//   apply_PC :589-616      Q_from_Om/update_X_Q/evolve_X_Q :679-710,865-878
// All are HBM/launch-latency bound: one pass over 3N reals.
#pragma once
#include <cuda_runtime.h>

namespace rbl {

template <typename real>
cudaError_t normalize_quats(real* Q, int n_bod, cudaStream_t s);

// r_{b,k} = R(q_b) ref_k + X_b ; Q stored [w,x,y,z] (unit).  r: 3 * n_bod * n_blb.
template <typename real>
cudaError_t place_blobs(const real* X, const real* Q, const real* ref, int n_bod, int n_blb,
                        real* r, cudaStream_t s);

// out_k = sign * (u_b + omega_b x (r_k - X_b)) + (add ? add_k : 0)
template <typename real>
cudaError_t k_dot(const real* U, const real* r, const real* X, int n_bod, int n_blb,
                  real sign, const real* add, real* out, cudaStream_t s);

// out_b = [sum_k lam_k ; sum_k (r_k - X_b) x lam_k]   (6 per body)
template <typename real>
cudaError_t kt_dot(const real* lam, const real* r, const real* X, int n_bod, int n_blb,
                   real* out, cudaStream_t s);

// (K^T K)^-1 per body (closed form of block_KTKinv, :302-326): G_b = diag(I/n_blb, S_b),
// S_b = (sum|ref|^2 I - R_b (sum ref ref^T) R_b^T)^-1.  Writes S_b (9 reals per body,
// row-major) and sets *singular = 1 where the reference exit()s (det < 1e-13).
template <typename real>
cudaError_t ktk_inv_blocks(const real* Q, const real* ref, int n_bod, int n_blb, real* S,
                           int* singular, cudaStream_t s);
// in-place G_b on 6-vectors: v_b <- diag(I/n_blb, S_b) v_b
template <typename real>
cudaError_t ktk_inv_apply(const real* S, int n_bod, int n_blb, real* v, cudaStream_t s);

// ---- preconditioner -----------------------------------------------------------------
// Common structure (apply_PC :589-616):  y = Mt^-1 slip ;  U_b = N_b (-F_b - K_b^T y_b) ;
// Lambda = y + (Mt^-1 K)_b U_b.   Y = Mt^-1 K (sz x 6 per body, layout [b][c][sz]) and the
// 6x6 Cholesky factors L of N^-1 = K^T Y are built once per configuration.

// diag PC: per blob the (diagonal) inverse self mobility, 3 reals per blob (:489-543)
template <typename real>
cudaError_t pc_diag_build(const real* r, int n, real a, real eta, bool wall, real* dinv,
                          int* below_wall, cudaStream_t s);
// out = dinv .* in over ncols column vectors of 3N reals; `in` layout [b][c][sz]
template <typename real>
cudaError_t pc_diag_mul(const real* dinv, const real* in, int n_bod, int n_blb, int ncols,
                        real* out, cudaStream_t s);

// columns of K per body, layout [b][c][sz]
template <typename real>
cudaError_t pc_fill_kcols(const real* r, const real* X, int n_bod, int n_blb, real* Kc,
                          cudaStream_t s);

// block PC (:461-487): dense (3 n_blb)^2 RPY(+wall) matrix of each of `count` bodies
// (blob positions r, body-major), symmetric, scaled by 1/(8 pi eta a) like :456.
template <typename real>
cudaError_t pc_block_assemble(const real* r, int count, int n_blb, real a, real eta,
                              bool wall, real* M, int* below_wall, cudaStream_t s);
// in-place batched inverse of symmetric (SPD for physical configurations) matrices, one
// CTA per matrix (Gauss-Jordan without pivoting; flags only exactly singular pivots)
template <typename real>
cudaError_t pc_block_invert(real* M, int count, int sz, int* not_spd, cudaStream_t s);
// out[b][c][:] = Minv_b in[b][c][:], c < ncols (1 or 6).  stride = sz*sz for per-body
// matrices or 0 for ONE shared reference-shape matrix, in which case body b uses
// Rb Minv Rb^T (free-space RPY is rotation covariant) with Rb from Q.
template <typename real>
cudaError_t pc_block_mul(const real* Minv, size_t stride, const real* Q, const real* in,
                         int n_bod, int n_blb, int ncols, real* out, cudaStream_t s);

// N^-1_b = K_b^T Y_b, inverted: L holds N_b (row-major 6x6, 36 reals per body)
template <typename real>
cudaError_t pc_ninv_chol(const real* Y, const real* r, const real* X, int n_bod, int n_blb,
                         real* L, int* not_spd, cudaStream_t s);
// out = [y + Y U ; U],  U_b = N_b (-F_b - K_b^T y_b)
template <typename real>
cudaError_t pc_finish(const real* y, const real* F, const real* Y, const real* L,
                      const real* r, const real* X, int n_bod, int n_blb, real* out,
                      cudaStream_t s);

// ---- noise preconditioner (block-Cholesky preconditioned Lanczos) -----------------------
// The Brownian increment only needs a vector of covariance A = B M B, not the symmetric square
// root: with L_b the Cholesky factor of body b's own mobility block Mt_b (the matrix
// Block_diag_invM assembles, :461-487) and G = L^-1,   g = L (G A G^T)^{1/2} W   has covariance A,
// and G A G^T is ~15x better conditioned than A, so Lanczos needs ~3x fewer products.
// In-place lower Cholesky of `count` sz x sz matrices (upper triangle zeroed), one CTA each.
template <typename real>
cudaError_t chol_lower(real* M, int count, int sz, int* not_spd, cudaStream_t s);
// G = L^-1 (both lower triangular, row-major, `count` matrices)
template <typename real>
cudaError_t tri_inverse(const real* L, real* G, int count, int sz, cudaStream_t s);
// out_b = A_b x_b (trans = false) or A_b^T x_b (trans = true) for every body; stride = sz*sz, or 0
// for ONE shared matrix of the reference shape used through the body rotation (rot_in: R_b^T per
// blob before the product, rot_out: R_b after it)
template <typename real>
cudaError_t body_mat_mul(const real* A, size_t stride, const real* Q, bool rot_in, bool rot_out, bool trans,
                         const real* in, int n_bod, int n_blb, real* out, cudaStream_t s, int ncols = 1);

// ---- integrator ---------------------------------------------------------------------
// Qo <- exp(omega * scale) Q (normalised), Xo <- X + u * scale  (scale = dt for
// evolve_X_Q; in-place allowed)
template <typename real>
cudaError_t integrate(const real* U, real scale, int n_bod, const real* X, const real* Q,
                      real* Xo, real* Qo, cudaStream_t s);

}  // namespace rbl
