/*
 * rbl.h -- C ABI of the B200-native hot path of Rigid_Body_Light (librbl.so).
 *
 * The reference has no C ABI: its boundary is the nanobind class
 * c_rigid.CManyBodies (/root/reference/src/c_rigid_obj.cpp:997-1027) consumed by
 * /root/reference/src/Rigid.py.  This header is what a host class in ANY language
 * binds instead of that C++ class; each entry point cites the reference member it
 * replaces.  rigid_body_light_b200/csrc/c_rigid.cpp (pybind11) is the host class
 * shipped here; INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *  - plain pointers and sizes only; no C++ or torch types cross this boundary;
 *  - `real` arrays are in the context's precision (RBL_F32 -> float, RBL_F64 -> double);
 *  - host-pointer calls are synchronous: results are in the output buffer on return;
 *  - rbl_dev_* calls take DEVICE pointers, enqueue on the context stream and return
 *    immediately; rbl_sync() waits and reports deferred errors (blob below the wall);
 *  - every call returns an rbl_status; rbl_last_error(ctx) holds the message;
 *  - a context is bound to one CUDA device and is not thread-safe;
 *  - there is no CPU fallback: without a CUDA device rbl_create fails.
 */
#ifndef RBL_H
#define RBL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rbl_ctx rbl_ctx;

typedef enum {
  RBL_OK = 0,
  RBL_ERR_INVALID = 1,    /* bad argument / size */
  RBL_ERR_CUDA = 2,       /* CUDA runtime error (message has the cudaError string) */
  RBL_ERR_BELOW_WALL = 3, /* a blob centre has z < 0 with the wall on (c_rigid_obj.cpp:95-97) */
  RBL_ERR_SINGULAR = 4,   /* K^T K singular (:313-316) or a PC block not SPD */
  RBL_ERR_STATE = 5,      /* parameters / configuration not set yet */
  RBL_ERR_NOMEM = 6
} rbl_status;

enum { RBL_F32 = 4, RBL_F64 = 8 };

/* ---- lifetime ------------------------------------------------------------------ */
/* precision: RBL_F32 or RBL_F64 (the reference picks it at compile time,
 * eigen_defines.h:5-37); device: CUDA ordinal, or -1 for the current device. */
int rbl_create(int precision, int device, rbl_ctx** out);
void rbl_destroy(rbl_ctx* ctx);
const char* rbl_last_error(const rbl_ctx* ctx); /* ctx may be NULL: last create error */
int rbl_precision(const rbl_ctx* ctx);
int rbl_sm_count(const rbl_ctx* ctx);
const char* rbl_version(void);

/* ---- state: CManyBodies setters/getters (host pointers) --------------------------- */
/* setParameters (:183-195): stores a, dt, kBT, eta and the MEAN-REMOVED ref config
 * (n_blb x 3, row-major). */
int rbl_set_parameters(rbl_ctx* ctx, double a, double dt, double kBT, double eta,
                       const void* ref_cfg, int n_blb);
/* setBlkPC / setWallPC (:197-199).  `wall` gates the wall correction of BOTH the
 * preconditioner and apply_M, like PC_wall in the reference. */
int rbl_set_flags(rbl_ctx* ctx, int block_pc, int wall);
/* setConfig (:201-233): X 3*n_bod, Q 4*n_bod as [w,x,y,z]; quaternions are normalised.
 * DEVIATION: the reference leaves PC_mat_Set alone here (only evolve_X_Q resets it, :877), so its next
 * apply_PC mixes the old invM / N_lu with the new K.  Here the preconditioner is rebuilt at its next
 * use: apply_PC after rbl_set_config is the preconditioner of that configuration. */
int rbl_set_config(rbl_ctx* ctx, const void* X, const void* Q, int n_bod);
/* getConfig (:235-255) */
int rbl_get_config(rbl_ctx* ctx, void* X, void* Q);
/* set_K_mats (:395-402): refreshes the cached blob positions / lever arms K is made of */
int rbl_set_K_mats(rbl_ctx* ctx);
int rbl_n_bodies(const rbl_ctx* ctx);
int rbl_blobs_per_body(const rbl_ctx* ctx);

/* ---- operators (host pointers) ---------------------------------------------------- */
/* multi_body_pos (:295-300): out 3*N, blob-major xyz */
int rbl_blob_positions(rbl_ctx* ctx, void* out);
/* K_x_U (:404): U 6*n_bod -> out 3*N */
int rbl_K_dot(rbl_ctx* ctx, const void* U, void* out);
/* KT_x_Lam (:410): lambda 3*N -> out 6*n_bod */
int rbl_KT_dot(rbl_ctx* ctx, const void* lambda, void* out);
/* Kinv_x_V (:406) and KTinv_x_F (:408) */
int rbl_Kinv_dot(rbl_ctx* ctx, const void* V, void* out);
int rbl_KTinv_dot(rbl_ctx* ctx, const void* F, void* out);
/* apply_M (:641-659): U = M F, or B M B F with the wall; n_blobs is free (it need not
 * equal n_bod*n_blb, tests/test_interface.py:171-177). F, r, out: 3*n_blobs.
 * Reproducibility: the default (symmetric) product evaluates every unordered pair once and adds the
 * two reactions with floating-point atomics (RED.ADD), so two runs agree to rounding (~1e-7 / 1e-16
 * relative), not bit for bit; the sum order INSIDE a warp is fixed.  rbl_set_matvec_mode(ctx, 1) selects
 * the ordered kernel, which is bit-reproducible (per-CTA partial sums added in CTA order, no atomics)
 * at ~0.72x the speed -- the deterministic path for a bitwise-repeatable trajectory.
 * DEVIATION: two DISTINCT blobs closer than 1e-12 a abort the reference ("TWO BLOBS ARE OVERLAPPING",
 * exit(), :53-58); here such a pair is evaluated with the overlap formula at r -> 0 (the self-mobility
 * block), finite and continuous, and no error is raised. */
int rbl_apply_M(rbl_ctx* ctx, const void* F, const void* r, int n_blobs, void* out);
/* apply_PC (:589-616): in/out 3*N + 6*n_bod, exact inverse of [Mt -K; -K^T 0] */
int rbl_apply_PC(rbl_ctx* ctx, const void* in, void* out);
/* RigidBody.apply_saddle (Rigid.py:73-80): out = [M lam - K U ; K^T lam], fused on device */
int rbl_apply_saddle(rbl_ctx* ctx, const void* x, void* out);
/* evolve_X_Q (:865-878): U 6*n_bod (velocities; multiplied by dt inside), rebuilds K,
 * invalidates the preconditioner */
int rbl_evolve(rbl_ctx* ctx, const void* U);
/* get_K / get_Kinv (:978-992) as CSC arrays.  K is 3N x 6n_bod with 9 stored entries per
 * blob (indptr 6*n_bod+1, indices/data 9*N); Kinv is 6n_bod x 3N with 4 stored entries
 * per column (indptr 3*N+1, indices/data 12*N; a computed entry that happens to be
 * exactly zero is kept, the reference's .pruned() would drop it). */
int rbl_export_K_csc(rbl_ctx* ctx, int64_t* indptr, int32_t* indices, void* data);
int rbl_export_Kinv_csc(rbl_ctx* ctx, int64_t* indptr, int32_t* indices, void* data);

/* ---- Krylov drivers (absent from the reference) --------------------- */
/* Right-preconditioned restarted GMRES on the saddle system  A x = rhs  with
 * A = apply_saddle and preconditioner P^-1 = apply_PC composed with the sign flip that
 * maps the reference's two conventions onto each other.
 * rhs/x: 3N+6n_bod.  Returns iterations in *iters and ||r||/||rhs|| in *relres. */
int rbl_gmres(rbl_ctx* ctx, const void* rhs, void* x, double tol, int restart, int max_iter,
              int* iters, double* relres);
/* out = (B M B)^{1/2} W by Lanczos (replaces the dense Cholesky of M_half_W, :661-675).
 * W, out: 3*N.  tol is the relative change of the iterate. */
int rbl_lanczos_sqrt(rbl_ctx* ctx, const void* W, void* out, double tol, int max_iter,
                     int* iters);

/* Two right-hand sides per pass: out1 = M F1 (or B M B F1), out2 = M F2, same semantics as
 * rbl_apply_M.  The geometry of every blob pair is evaluated once for both vectors
 * (rpy_matvec_sym2_kernel); this is the product behind rbl_lanczos_sqrt2 and the BD step. */
int rbl_apply_M2(rbl_ctx* ctx, const void* F1, const void* F2, const void* r, int n_blobs, void* out1, void* out2);
/* (B M B)^{1/2} W1 and (B M B)^{1/2} W2 by two Lanczos recurrences in lockstep that share every
 * mobility product (the two M_half_W calls of one step, :930-935).  iters2: two ints. */
int rbl_lanczos_sqrt2(rbl_ctx* ctx, const void* W1, const void* W2, void* out1, void* out2, double tol,
                      int max_iter, int* iters2);
/* Block-Cholesky preconditioned noise.  A Brownian increment needs covariance A = B M B, not the
 * symmetric square root: with L_b = chol(Mt_b) of each body's own mobility block (the matrix
 * Block_diag_invM assembles, :461-487) and G = L^-1,  g = L (G A G^T)^{1/2} W  has covariance A
 * (the reference's own M_half_W returns yet another root, the Cholesky factor times W, :661-675),
 * and Lanczos on G A G^T needs ~3x fewer products.  mode 0: symmetric square root everywhere;
 * 1 (default): preconditioned inside rbl_bd_step only; 2: rbl_lanczos_sqrt / rbl_lanczos_sqrt2
 * return the preconditioned vector too.  Falls back to the plain recurrence when a body block is
 * not positive definite (blobs inside the wall-overlap layer) or the factors do not fit. */
int rbl_set_noise_preconditioner(rbl_ctx* ctx, int mode);
/* Self-check of the factors at the current configuration, like the reference's unbound test_PC /
 * Test_Mhalf (:569-587, :895-915): factor_err = |L L^T x - Mt x| / |Mt x| against freshly assembled
 * body blocks, inverse_err = |G L x - x| / |x| for a fixed pseudo-random x; active = 0 if the
 * context fell back to the plain recurrence.  Rank-local. */
int rbl_noise_selfcheck(rbl_ctx* ctx, double* factor_err, double* inverse_err, int* active);
/* ---- random finite differences (unbound members of the reference, :743-863), noise supplied ------
 * The reference draws W inside each function from a wall-clock seeded generator (:730-741); here the
 * caller passes it, so every function is deterministic and can be compared with the reference's own
 * code given the same noise (tests/test_gpu_rfd.py).  delta <= 0 selects the reference's value: 1e-4
 * for M_RFD / KTinv_RFD (:745,771; 4e-3 in a float context, see rbl_set_rfd_delta), 1e-3 for the
 * *_from_U forms (:820,844).  n = blobs of the context, n_bod = its bodies.
 *   rbl_M_RFD          out[3n]    = (M(q+) - M(q-)) W / delta,  q+- = q +- (delta/2) K^-1 W     (:769-796)
 *   rbl_M_RFD_from_U   out[3n]    = the same with q+- = q +- (delta/2) U, U[6 n_bod]             (:818-840)
 *   rbl_KT_RFD_from_U  out[6nbod] = (K(q+)^T - K(q-)^T) W / delta, W[3n]                         (:842-863)
 *   rbl_KTinv_RFD      out[6nbod] = K^T (Kinv(q+)^T - Kinv(q-)^T) W / delta, q+- = q +- (delta/2) W, W[6 n_bod] (:743-767)
 *   rbl_M_RFD_cfgs     r_plus, r_minus [3n] = blob positions of q +- (delta/2) U                 (:798-816)
 *   rbl_update_X_Q_out X_out[3 n_bod], Q_out[4 n_bod] = q displaced by U (a displacement), state untouched (:712-728)
 *   rbl_evolve_RFD     installs q displaced by U and rebuilds K; a built preconditioner is kept   (:880-893)
 * rbl_M_RFD / rbl_M_RFD_from_U are collective on a partitioned suspension; the others are rank-local. */
int rbl_M_RFD(rbl_ctx* ctx, const void* W, double delta, void* out);
int rbl_M_RFD_from_U(rbl_ctx* ctx, const void* U, const void* W, double delta, void* out);
int rbl_KT_RFD_from_U(rbl_ctx* ctx, const void* U, const void* W, double delta, void* out);
int rbl_KTinv_RFD(rbl_ctx* ctx, const void* W6, double delta, void* out);
int rbl_M_RFD_cfgs(rbl_ctx* ctx, const void* U, double delta, void* r_plus, void* r_minus);
int rbl_update_X_Q_out(rbl_ctx* ctx, const void* U, void* X_out, void* Q_out);
int rbl_evolve_RFD(rbl_ctx* ctx, const void* U);
/* RFD step of rbl_bd_step and default of rbl_M_RFD / rbl_KTinv_RFD.  0 restores the defaults: 1e-4 in a
 * double context (the reference's value), 4e-3 in a float context -- a centred difference quotient in
 * float is most accurate near eps^(1/3) x (length over which M varies ~ the body radius): rounding
 * 3e-7 |M W| / delta against truncation ~ (delta/a)^2 / 6. */
int rbl_set_rfd_delta(rbl_ctx* ctx, double delta);
/* Mixed precision for DOUBLE contexts.  The context keeps a float mirror of itself (same parameters, flags,
 * configuration; same stream; on a partitioned suspension the mirror shares the context's communicator and
 * sets up its own float-sized peer buffers -- collective: every rank the same mode at the same point):
 *   0 (default) everything in double;
 *   1 rbl_gmres / the solve of rbl_bd_step: float GMRES corrections inside an iterative refinement whose
 *     residual b - apply_saddle(x) is evaluated in DOUBLE; stops on the same ||b - A x|| / ||b|| <= tol;
 *   2 additionally the mobility products inside the Lanczos square roots run in float (vectors, recurrence
 *     and block-Cholesky factors stay double): the increment is the square root of the float-rounded
 *     operator, M (1 + O(1e-7)) -- below any Lanczos tolerance >= 1e-6, NOT bitwise the double increment. */
int rbl_set_mixed_precision(rbl_ctx* ctx, int mode);
/* 1 (default, like the reference's split_rand = true, :150): two Brownian increments per step, c1 = 2 sqrt(kBT/dt),
 * c2 = sqrt(kBT/dt), BI = c2 (M^{1/2}W1 - M^{1/2}W2) (:943-948).  0: one increment, c1 = c2 = sqrt(2 kBT/dt),
 * BI = c2 M^{1/2}W1 (:949-953); rbl_bd_step then ignores W2 (may be NULL). */
int rbl_set_split_rand(rbl_ctx* ctx, int enable);
/* 1 (default): rbl_bd_step uses the paired Lanczos; 0: two separate single-vector runs */
int rbl_set_lanczos_pairing(rbl_ctx* ctx, int enable);

/* One rigid-body Brownian-dynamics step: the trapezoidal-slip midpoint scheme
 * that RHS_and_Midpoint (:917-976) sets up but never finishes, composed on the device:
 *   M^{1/2}W_1, M^{1/2}W_2 by Lanczos;  RFD drift (M(q+) - M(q-)) W_r / delta with
 *   q+- = q +- (delta/2) K^-1 W_r (:769-796);  BI = sqrt(kBT/dt) (M^{1/2}W_1 - M^{1/2}W_2) (:945-948);
 *   midpoint q' = q + (dt/2) K^-1 (2 sqrt(kBT/dt) M^{1/2}W_1) (:954-958), K and PC rebuilt THERE;
 *   GMRES solve of  apply_saddle([lambda;U]) = [slip - kBT*RFD - BI ; F_ext]  at q';
 *   q <- q + dt*U from the ORIGINAL configuration (evolve_X_Q, :865-878).
 * F_ext: 6*n_bod.  slip: 3*N or NULL (zero).  W1, W2, Wr: 3*N standard-normal vectors supplied by
 * the caller (the reference seeds its generator from the wall clock, :731); all three NULL or
 * kBT == 0 gives the deterministic step.  U_out: 6*n_bod rigid velocities of the step. */
int rbl_bd_step(rbl_ctx* ctx, const void* F_ext, const void* slip, const void* W1, const void* W2,
                const void* Wr, double kBT, double gmres_tol, int restart, int max_iter,
                double lanczos_tol, int lanczos_max_iter, void* U_out, int* gmres_iters, double* relres);

/* rbl_bd_step with the three noise vectors drawn ON THE DEVICE: element e of (W1, W2, Wr) is a pure
 * function of (seed, step, GLOBAL element index e) -- one Philox4x32-10 block per element, Box-Muller --
 * so a trajectory is reproducible from (seed, first step) and does not depend on how the suspension
 * is partitioned over GPUs (the reference seeds std::normal_distribution from the wall clock,
 * :730-741).  rbl_normals / rbl_dev_normals expose the generator: n elements starting at global
 * element `first` (host / device pointers). */
int rbl_bd_step_seeded(rbl_ctx* ctx, const void* F_ext, const void* slip, uint64_t seed, uint64_t step, double kBT,
                       double gmres_tol, int restart, int max_iter, double lanczos_tol, int lanczos_max_iter,
                       void* U_out, int* gmres_iters, double* relres);
int rbl_normals(rbl_ctx* ctx, uint64_t seed, uint64_t step, uint64_t first, size_t n, void* W1, void* W2, void* Wr);
int rbl_dev_normals(rbl_ctx* ctx, uint64_t seed, uint64_t step, uint64_t first, size_t n, void* dW1, void* dW2,
                    void* dWr);

/* ---- device-resident API (device pointers, asynchronous on the context stream) ---- */
/* Targets [tgt_first, tgt_first+n_tgt) of the n_blobs sources: out has 3*n_tgt reals.
 * This is the entry a multi-GPU host shards by body range. */
int rbl_dev_apply_M(rbl_ctx* ctx, const void* dF, const void* dr, int n_blobs, int tgt_first,
                    int n_tgt, void* dout);
/* device-pointer form of rbl_apply_M2 (all n_blobs blobs are targets and sources) */
int rbl_dev_apply_M2(rbl_ctx* ctx, const void* dF1, const void* dF2, const void* dr, int n_blobs, void* dout1,
                     void* dout2);
int rbl_dev_blob_positions(rbl_ctx* ctx, void* dout);
int rbl_dev_K_dot(rbl_ctx* ctx, const void* dU, void* dout);
int rbl_dev_KT_dot(rbl_ctx* ctx, const void* dlambda, void* dout);
int rbl_dev_apply_PC(rbl_ctx* ctx, const void* din, void* dout);
int rbl_dev_apply_saddle(rbl_ctx* ctx, const void* dx, void* dout);
/* One rank's share of apply_saddle when bodies are partitioned across GPUs: the context
 * holds this rank's bodies (blobs [tgt_first, tgt_first + N_local) of the global n_blobs);
 * dlambda_all / dr_all are the all-gathered forces and positions (3*n_blobs).
 * dout_local = [ (M lambda)_local - K_local U_local ; K_local^T lambda_local ]. */
int rbl_dev_apply_saddle_shard(rbl_ctx* ctx, const void* dlambda_all, const void* dr_all,
                               int n_blobs, int tgt_first, const void* dU_local, void* dout_local);
/* Partial product for multi-GPU hosts using the symmetric kernel: dout (3*n_blobs, already scaled)
 * holds the contribution of share `part` of `n_parts` of the unordered-pair work; the full product
 * is the SUM of the n_parts partial outputs (all-reduce). */
int rbl_dev_apply_M_part(rbl_ctx* ctx, const void* dF, const void* dr, int n_blobs, int part, int n_parts,
                         void* dout);
/* dout_local = [ dMlambda_local - K_local U_local ; K_local^T lambda_local ] for the context's bodies */
int rbl_dev_saddle_finish(rbl_ctx* ctx, const void* dMlambda_local, const void* dlambda_local,
                          const void* dU_local, void* dout_local);
int rbl_sync(rbl_ctx* ctx);

/* ---- partitioned suspensions: one context per GPU, NCCL inside the library ---------------
 * The reference is single process / single thread; this is the multi-GPU path of SURVEY.md
 * section 8e.  Each rank creates a context, sets ITS bodies (a contiguous range of the global
 * body list) with rbl_set_config, then all ranks call rbl_comm_init together.  From then on the
 * SAME entry points work on rank-local slices -- vectors are [lambda_local (3 N_local) ;
 * U_local (6 n_bod_local)] -- and are collective: rbl_apply_saddle, rbl_dev_apply_saddle,
 * rbl_gmres, rbl_lanczos_sqrt, rbl_bd_step.  Per mobility product: ncclAllGather of the forces
 * (positions once per configuration), this rank's share of the unordered-pair work over all
 * blobs, ncclReduceScatter of the partial products; dot products of the Krylov drivers are
 * batched and all-reduced; status codes are max-reduced so every rank returns the same one.
 * K, K^T, the preconditioner and the integrator are rank-local (whole bodies per rank).
 * rbl_apply_M / rbl_dev_apply_M stay local, non-collective pure functions of their arguments.
 * NCCL is bound with dlopen("libnccl.so.2") at rbl_comm_init time; a single-GPU host never needs it.
 *
 * The two exchanges around every product.  rbl_comm_init also tries to set up a PEER-MEMORY exchange
 * (csrc/rbl_peer.cuh): every rank exports one device buffer through CUDA IPC and maps every other
 * rank's; the all-gather of the forces becomes NVLink stores into the peers' buffers plus an epoch
 * hand-shake, the reduce-scatter becomes a hand-shake plus a kernel that sums this rank's rows of the
 * world's partial products in RANK ORDER straight from the peers' memory (so, unlike ncclReduceScatter,
 * the order of that sum is fixed).  It is on only if EVERY rank could set it up (one process per GPU on
 * one node, peer access between the devices); otherwise, or after rbl_comm_set_exchange(ctx, 0), or with
 * RBL_PEER_EXCHANGE=0 in the environment, the NCCL collectives above are used.  Waits inside the
 * hand-shakes are bounded (RBL_PEER_TIMEOUT_S, default 30 s): a rank that never arrives turns into
 * RBL_ERR_CUDA at the next synchronisation instead of a hung GPU.  Positions (once per configuration),
 * dot products and status codes stay on NCCL. */
/* rank 0: 128 bytes of ncclUniqueId for the host to broadcast (any transport) */
int rbl_comm_unique_id(void* out128);
/* blobs_per_rank: `world` entries (blobs, not bodies); entry `rank` must equal this context's N */
int rbl_comm_init(rbl_ctx* ctx, const void* uid128, int rank, int world, const int* blobs_per_rank);
/* 1: peer-memory exchange in use, 0: NCCL collectives; _why: what kept the peer exchange off ("" if on) */
int rbl_comm_exchange(const rbl_ctx* ctx);
const char* rbl_comm_exchange_why(const rbl_ctx* ctx);
/* collective (every rank, same mode): 0 = NCCL collectives, 1 = peer memory (RBL_ERR_STATE if unavailable) */
int rbl_comm_set_exchange(rbl_ctx* ctx, int mode);
/* while rbl_profile_matvec is on, CUDA events bracket the phases of every partitioned product: ms3 = average
 * milliseconds of {gather of the forces, pack + product + scale, reduce of the partial products}, n = products
 * seen; reset != 0 clears */
int rbl_comm_profile(rbl_ctx* ctx, double* ms3, int* n, int reset);
int rbl_comm_world(const rbl_ctx* ctx); /* 1 without a communicator */
int rbl_comm_rank(const rbl_ctx* ctx);
/* the context's cudaStream_t (as void*); set_stream lets a host share its own stream */
void* rbl_stream(rbl_ctx* ctx);
int rbl_set_stream(rbl_ctx* ctx, void* cuda_stream);

/* ---- memory + timing helpers (so a C/C++ host needs nothing but this header) ------- */
int rbl_dev_alloc(rbl_ctx* ctx, size_t bytes, void** dptr);
int rbl_dev_free(rbl_ctx* ctx, void* dptr);
int rbl_pinned_alloc(rbl_ctx* ctx, size_t bytes, void** hptr);
int rbl_pinned_free(rbl_ctx* ctx, void* hptr);
int rbl_memcpy_h2d(rbl_ctx* ctx, void* dst, const void* src, size_t bytes);
int rbl_memcpy_d2h(rbl_ctx* ctx, void* dst, const void* src, size_t bytes);
/* CUDA-event timer on the context stream */
int rbl_timer_start(rbl_ctx* ctx);
int rbl_timer_stop(rbl_ctx* ctx, double* elapsed_ms);
/* writes more than L2 (126 MB) so the next timed call starts cold */
int rbl_flush_l2(rbl_ctx* ctx);

/* ---- tuning + measurement ------------------------------------------------------------ */
int rbl_num_matvec_variants(const rbl_ctx* ctx);
/* targets per thread and threads per CTA of variant idx */
int rbl_matvec_variant_info(const rbl_ctx* ctx, int idx, int* targets_per_thread, int* threads);
int rbl_set_matvec_variant(rbl_ctx* ctx, int idx); /* -1 = automatic */
/* Product kernel selection.  0 (default): the symmetric kernel (one evaluation per unordered pair,
 * floating-point atomics: reproducible to rounding) whenever targets == sources, the ordered kernel
 * otherwise.  1: always the ordered kernel (bit-reproducible, every ordered pair evaluated). */
int rbl_set_matvec_mode(rbl_ctx* ctx, int mode);
int rbl_num_sym_variants(const rbl_ctx* ctx);
int rbl_sym_variant_info(const rbl_ctx* ctx, int idx, int* targets_per_thread, int* threads);
int rbl_set_sym_variant(rbl_ctx* ctx, int idx); /* -1 = automatic */
/* Experiment knob.  The unordered-pair tile triangle is cut into shares of equal chunk COUNT (per CTA, and per
 * GPU on a partitioned suspension).  w in (0, 1) cuts by COST instead, a diagonal unit -- the ordered loop on
 * the tiles that hold the self pairs -- weighing w of a symmetric one (by instruction count 0.77 with the wall,
 * 0.80 in free space).  Measured at cfg2: no gain beyond noise (profiles/r02_part_balance.md), hence the default
 * w <= 0 or 1: equal counts.  Results do not depend on w beyond summation order. */
int rbl_set_split_weight(rbl_ctx* ctx, double w);
/* The cut points themselves (host arithmetic only, no device): bounds[0..grid] = 32-source-chunk indices that
 * delimit the `grid` CTAs of share `part` of `n_parts` for n_blobs blobs and a target tile of tgt_tile blobs
 * (a multiple of 256); *total_chunks = chunks of the whole triangle. */
int rbl_plan_cost_bounds(int n_blobs, int tgt_tile, int grid, int part, int n_parts, double w, long long* bounds,
                         long long* total_chunks);
/* sources per warp-private reaction-reduction chunk of symmetric variant idx */
int rbl_sym_variant_chunk(const rbl_ctx* ctx, int idx);
int rbl_num_sym2_variants(const rbl_ctx* ctx);
int rbl_sym2_variant_info(const rbl_ctx* ctx, int idx, int* targets_per_thread, int* threads);
int rbl_set_sym2_variant(rbl_ctx* ctx, int idx); /* two-right-hand-side kernel; -1 = automatic */
/* kernels launched by this context since creation (bench.py's gpu_launches) */
int64_t rbl_launch_count(const rbl_ctx* ctx);
/* mobility products (whole, or one rank's share) launched since creation */
int64_t rbl_product_count(const rbl_ctx* ctx);
/* Lanczos iterations of the two M^{1/2}W solves of the last rbl_bd_step */
int rbl_bd_stats(const rbl_ctx* ctx, int* lanczos_iters_1, int* lanczos_iters_2);
/* average duration (ms, CUDA events around the kernel alone) of the matvec kernel over
 * the launches since the last reset; enable/disable with rbl_profile_matvec */
int rbl_profile_matvec(rbl_ctx* ctx, int enable);
int rbl_matvec_profile(rbl_ctx* ctx, double* avg_ms, int64_t* launches, int reset);
/* wall clock (ms, accumulated over the rbl_bd_step calls made while rbl_profile_matvec is on; each phase
 * ends with a stream synchronisation, so never enable it in a timed run) of the six phases of a BD step:
 * [0] inputs + noise, [1] Lanczos M^{1/2}W (c_rigid_obj.cpp:661-675), [2] random finite difference
 * (:769-796), [3] midpoint configuration (:954-958), [4] GMRES incl. the preconditioner build,
 * [5] evolve (:865-878) + output */
int rbl_bd_phase_ms(rbl_ctx* ctx, double* out6, int reset);
/* sustained FMA-pipe peak of this precision (TFLOP/s), the roofline denominator */
int rbl_fma_peak(rbl_ctx* ctx, int iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* RBL_H */
