// rbl_capi.cu -- context object + the C ABI of include/rbl.h.
//
// The context owns what CManyBodies owns in the reference
// (/root/reference/src/c_rigid_obj.cpp:144-168): parameters, reference configuration,
// X_n/Q_n, the K data (here: cached blob positions) and the lazily built preconditioner
// -- but resident in HBM.  Host-pointer entry points stage through device buffers and
// call the same device path the rbl_dev_* entry points expose.
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/rbl.h"
#include "rbl_comm.h"
#include "rbl_krylov.cuh"
#include "rbl_matvec.cuh"
#include "rbl_rigid.cuh"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {
thread_local std::string g_create_error;
}

// ---------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  ~DevBuf() { release(); }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

struct rbl_ctx {
  int precision = 0;
  int device = 0;
  int sm_count = 0;
  std::string err;
  virtual ~rbl_ctx() { delete comm; }
  int fail(int code, const std::string& msg) {
    err = msg;
    return code;
  }
  // virtual interface (void* = real* of the context's precision)
  virtual int set_parameters(double a, double dt, double kBT, double eta, const void* cfg, int n_blb) = 0;
  virtual int set_flags(int block_pc, int wall) = 0;
  virtual int set_config(const void* X, const void* Q, int n_bod) = 0;
  virtual int get_config(void* X, void* Q) = 0;
  virtual int set_K_mats() = 0;
  virtual int n_bodies() const = 0;
  virtual int blobs_per_body() const = 0;
  virtual int blob_positions(void* out, bool dev) = 0;
  virtual int K_dot(const void* U, void* out, bool dev) = 0;
  virtual int KT_dot(const void* lam, void* out, bool dev) = 0;
  virtual int Kinv_dot(const void* V, void* out) = 0;
  virtual int KTinv_dot(const void* F, void* out) = 0;
  virtual int apply_M(const void* F, const void* r, int n, void* out) = 0;
  virtual int dev_apply_M(const void* F, const void* r, int n, int t0, int nt, void* out) = 0;
  virtual int apply_PC(const void* in, void* out, bool dev) = 0;
  virtual int apply_saddle(const void* x, void* out, bool dev) = 0;
  virtual int saddle_shard(const void* lam_all, const void* r_all, int n_all, int t0, const void* U, void* out) = 0;
  virtual int evolve(const void* U) = 0;
  virtual int export_K(int64_t* indptr, int32_t* indices, void* data) = 0;
  virtual int export_Kinv(int64_t* indptr, int32_t* indices, void* data) = 0;
  virtual int gmres(const void* rhs, void* x, double tol, int restart, int max_iter, int* iters, double* relres) = 0;
  virtual int lanczos(const void* W, void* out, double tol, int max_iter, int* iters) = 0;
  virtual int bd_step(const void* F_ext, const void* slip, const void* W1, const void* W2, const void* Wr, double kBT,
                      double tol, int restart, int max_iter, double ltol, int lmax, void* U_out, int* iters,
                      double* relres) = 0;
  virtual int bd_step_seeded(const void* F_ext, const void* slip, unsigned long long seed, unsigned long long step,
                             double kBT, double tol, int restart, int max_iter, double ltol, int lmax, void* U_out,
                             int* iters, double* relres) = 0;
  virtual int M_RFD(const void* U6, const void* W, double delta, void* out) = 0;
  virtual int KT_RFD(const void* U6, const void* W, double delta, void* out6) = 0;
  virtual int KTinv_RFD(const void* W6, double delta, void* out6) = 0;
  virtual int RFD_cfgs(const void* U6, double delta, void* r_plus, void* r_minus) = 0;
  virtual int displaced_config(const void* U6, void* X_out, void* Q_out) = 0;
  virtual int evolve_RFD(const void* U6) = 0;
  virtual int normals(unsigned long long seed, unsigned long long step, unsigned long long first, size_t n, void* W1,
                      void* W2, void* Wr, bool dev) = 0;
  virtual int set_mixed(int mode) = 0;
  virtual int sync() = 0;
  virtual int fma_peak(int iters, double* tflops) = 0;
  virtual int num_variants() const = 0;
  virtual int variant_info(int idx, int* T, int* threads) const = 0;
  virtual int num_sym_variants() const = 0;
  virtual int sym_variant_info(int idx, int* T, int* threads) const = 0;
  virtual int sym_variant_chunk(int idx) const = 0;
  virtual int dev_apply_M_part(const void* F, const void* r, int n, int part, int n_parts, void* out) = 0;
  virtual int saddle_finish(const void* Mlam, const void* lam, const void* U, void* out) = 0;
  virtual int comm_init(const void* uid128, int rank, int world, const int* blobs_per_rank) = 0;
  virtual int comm_exchange(const char** why) const = 0;
  virtual int comm_set_exchange(int mode) = 0;
  virtual int comm_profile(double* ms3, int* n, int reset) = 0;
  virtual int dev_apply_M2(const void* F1, const void* F2, const void* r, int n, void* out1, void* out2) = 0;
  virtual int apply_M2(const void* F1, const void* F2, const void* r, int n, void* out1, void* out2) = 0;
  virtual int lanczos2(const void* W1, const void* W2, void* out1, void* out2, double tol, int max_iter, int* iters2) = 0;
  virtual int noise_selfcheck(double* factor_err, double* inverse_err, int* active) = 0;
  virtual int num_sym2_variants() const = 0;
  virtual int sym2_variant_info(int idx, int* T, int* threads) const = 0;

  // shared plumbing
  rbl::Comm* comm = nullptr;  // non-null: this context holds one rank's bodies of a partitioned suspension
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaEvent_t t0 = nullptr, t1 = nullptr;
  int64_t launches = 0;
  int64_t products = 0;        // mobility products (full or one rank's share) since creation
  int last_lanczos[2] = {0, 0};  // Lanczos iterations of the last rbl_bd_step
  int variant = -1;
  int sym_variant = -1;
  int sym2_variant = -1;
  // 0: symmetric square root everywhere; 1 (default): block-Cholesky preconditioned noise inside
  // rbl_bd_step; 2: also in rbl_lanczos_sqrt / rbl_lanczos_sqrt2 (they then return L (G A G^T)^{1/2} W)
  int noise_mode = 1;
  double split_weight = 0;   // cost of a diagonal unit relative to a symmetric one; <= 0: per-mode estimate, 1: equal counts
  bool split_rand = true;    // BD step: two Brownian increments W1, W2 (c_rigid_obj.cpp:150,943-953); false: one
  double rfd_delta = 0;      // <= 0: default of the precision (1e-4 double like :771, 4e-3 float)
  bool pair_lanczos = true;  // BD step: M^{1/2}W_1 and M^{1/2}W_2 in lockstep over the two-right-hand-side product
  int mode = 0;  // 0: symmetric kernel when targets == sources; 1: ordered kernel always
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  // wall-clock per phase of the last rbl_bd_step calls while profiling is on (ms, accumulated):
  // 0 inputs+noise, 1 Lanczos, 2 RFD, 3 midpoint configuration, 4 GMRES (with the PC build), 5 evolve+output
  double bd_phase_ms[6] = {0, 0, 0, 0, 0, 0};
  DevBuf flush_buf;
};

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return this->fail(e_ == cudaErrorMemoryAllocation ? RBL_ERR_NOMEM : RBL_ERR_CUDA,  \
                        std::string(#call) + ": " + cudaGetErrorString(e_));             \
  } while (0)
// a launcher that enqueued `n` kernels
#define LAUNCH(n, call) \
  do {                  \
    CK(call);           \
    this->launches += (n); \
  } while (0)
#define RET(call)                 \
  do {                            \
    int s_ = (call);              \
    if (s_ != RBL_OK) return s_;  \
  } while (0)

// ---------------------------------------------------------------------------------------
template <typename real>
struct Ctx;

template <typename real>
struct Ctx final : rbl_ctx {
  // parameters (setParameters, c_rigid_obj.cpp:183-195)
  double a = 0, dt = 0, kBT = 0, eta = 0;
  bool params_set = false, cfg_set = false, wall = false, block_pc = false, pc_set = false;
  bool r_valid = false;
  int n_blb = 0, n_bod = 0;
  int pc_n_bod = 0;
  // device state
  DevBuf d_ref, d_X, d_Q, d_r, d_S;
  // matvec workspace
  DevBuf d_rec, d_box_src, d_box_tgt, d_scratch, d_flags, d_raw;
  // staging for the host-pointer API
  DevBuf d_in0, d_in1, d_out0;
  // preconditioner
  DevBuf d_dinv, d_Minv, d_Kc, d_Y, d_L, d_y;
  bool pc_shared = false;
  bool pc_chol = false;  // block PC with the wall: Mt^-1 = G^T G from the bodies' Cholesky factors (d_NG)
  DevBuf d_pt;           // scratch for the two-stage G^T (G x) products
  // krylov
  DevBuf d_V, d_w, d_z, d_tmp, d_partial, d_coef, d_dots;
  // BD step
  DevBuf d_rhs, d_sol, d_mh1, d_mh2, d_rfd, d_noise, d_uom, d_Xs, d_Qs, d_Xp, d_Qp, d_rp, d_t1, d_t2;
  // partitioned mode (comm != nullptr): all-gathered positions / forces, partial product, agreed status
  DevBuf d_r_all, d_lam_all, d_mbuf, d_status;
  // two-right-hand-side product / paired Lanczos
  DevBuf d_rec2, d_raw2, d_V2, d_w2, d_in2, d_out2, d_wr;
  // noise preconditioner: Cholesky factors L of the bodies' own mobility blocks and G = L^-1
  DevBuf d_NL, d_NG, d_nt1, d_nt2, d_nu1, d_nu2;
  bool noise_set = false, noise_ok = false, noise_shared = false, noise_shared_ready = false;
  bool r_all_valid = false;
  // Packed records and tile boxes are a function of the POSITIONS only: inside the Krylov drivers the
  // configuration is fixed while the force vector changes every product, so a product on positions
  // the records were already packed from only rewrites the three force words of each record
  // (repack_forces) and skips both tile-box launches.  `gen` identifies a position set: cfg_gen for the
  // context's own blob positions (d_r), r_all_gen for the all-gathered ones; 0 = unknown (always repack).
  unsigned long long gen_counter = 0, cfg_gen = 0, r_all_gen = 0;
  struct RecState {
    unsigned long long gen = 0;
    int n = 0, tile = 0;
    const void* buf = nullptr;
    bool matches(unsigned long long g, int n_, int tile_, const void* b) const {
      return g != 0 && g == gen && n_ == n && tile_ == tile && b == buf;
    }
  } rec1_state, rec2_state;
  // equal-cost cut points of the unit triangle (sym_cost_bounds), cached per (n, variant, share)
  struct BoundsCache {
    std::vector<long long> host;
    DevBuf dev;
    long long key[6] = {-1, -1, -1, -1, -1, -1};
  } bounds1, bounds2;
  // Default: equal COUNTS (w = 1).  By instruction count a diagonal unit costs 0.77 (wall: 72 / 93) or
  // 0.80 (free space: 27 / 33) of a symmetric one, but cutting by that cost changed neither the kernel
  // time nor the spread between the shares of an 8-way partition beyond run-to-run noise at cfg2
  // (profiles/r02_part_balance.md: the spread follows the near-tile pattern of the geometry and the
  // per-launch ramp, not the diagonal units), so the knob stays an experiment (rbl_set_split_weight).
  double diag_weight() const { return split_weight > 0 ? split_weight : 1.0; }
  int cost_bounds(BoundsCache& bc, rbl::SymPlan* plan, int variant, int part, int n_parts, const long long** out) {
    const double w = diag_weight();
    if (w == 1.0) {  // equal counts: the kernels cut [u0, u1) themselves
      *out = nullptr;
      return RBL_OK;
    }
    const long long key[6] = {plan->n, variant, part, n_parts, plan->grid, (long long)(w * 1e6)};
    bool same = bc.dev.p != nullptr;
    for (int i = 0; i < 6; ++i) same = same && bc.key[i] == key[i];
    if (!same) {
      bc.host.assign((size_t)plan->grid + 1, 0);
      rbl::sym_cost_bounds(plan, part, n_parts, w, bc.host.data());
      CK(bc.dev.ensure(bc.host.size() * sizeof(long long)));
      CK(cudaMemcpyAsync(bc.dev.p, bc.host.data(), bc.host.size() * sizeof(long long), cudaMemcpyHostToDevice, stream));
      for (int i = 0; i < 6; ++i) bc.key[i] = key[i];
    } else {
      plan->u0 = bc.host.front();
      plan->u1 = bc.host.back();
    }
    *out = bc.dev.template as<long long>();
    return RBL_OK;
  }

  // mixed precision (double contexts, single GPU): a float mirror of this context on the same stream.
  //  1: GMRES solves run in float inside an iterative refinement whose residual is the DOUBLE operator
  //     (same stopping rule ||b - A x|| / ||b|| <= tol on the double residual);
  //  2: additionally the mobility products of the Lanczos noise run in float (the vectors, the
  //     recurrence and the block-Cholesky factors stay double): the increment is (M~)^{1/2} W for the
  //     float-rounded operator M~ = M (1 + O(1e-7)), below the Lanczos tolerances in use (>= 1e-6).
  int mixed = 0;
  Ctx<float>* shadow = nullptr;
  unsigned long long shadow_gen = 0;
  std::vector<double> ref_host;  // mean-removed reference configuration, for the mirror
  DevBuf d_c32a, d_c32b, d_c32c, d_c32d, d_mr, d_me;
  int mixed_outer = 0;           // refinement cycles of the last mixed solve

  enum { FLAG_BELOW = 0, FLAG_SINGULAR = 1, FLAG_NOT_SPD = 2, FLAG_NOISE = 3, FLAG_PEER = 4, N_FLAGS = 5 };

  ~Ctx() override {
    if (comm) comm->peer_release(true, d_status.as<int>(), stream);
    for (auto& ev : prof_comm)
      for (cudaEvent_t e : ev) cudaEventDestroy(e);
    delete shadow;
    for (auto& pr : prof_events) {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    if (t0) cudaEventDestroy(t0);
    if (t1) cudaEventDestroy(t1);
    if (own_stream && stream) cudaStreamDestroy(stream);
  }

  long long N() const { return (long long)n_bod * n_blb; }
  size_t sys_size() const { return (size_t)(3 * N() + 6 * (long long)n_bod); }

  int init() {
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    own_stream = true;
    CK(cudaEventCreate(&t0));
    CK(cudaEventCreate(&t1));
    CK(d_flags.ensure(N_FLAGS * sizeof(int)));
    CK(cudaMemsetAsync(d_flags.p, 0, N_FLAGS * sizeof(int), stream));
    return RBL_OK;
  }

  int h2d(void* dst, const void* src, size_t bytes) {
    if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    return RBL_OK;
  }
  int d2h(void* dst, const void* src, size_t bytes) {
    if (bytes) CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    return RBL_OK;
  }

  // waits for the stream and turns device flags into status codes
  int sync() override {
    int flags[N_FLAGS] = {0, 0, 0, 0, 0};
    CK(cudaMemcpyAsync(flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (flags[0] || flags[1] || flags[2] || flags[FLAG_PEER]) {
      CK(cudaMemsetAsync(d_flags.p, 0, sizeof(flags), stream));
      if (flags[FLAG_PEER])
        return fail(RBL_ERR_CUDA, "peer-memory exchange: a rank of the partitioned suspension did not arrive within the time limit "
                                  "(RBL_PEER_TIMEOUT_S); results of this call are invalid");
      if (flags[FLAG_BELOW])
        return fail(RBL_ERR_BELOW_WALL,
                    "A blob has its center below the wall (z<0). Cannot compute mobility- check "
                    "your configuration.");
      if (flags[FLAG_SINGULAR])
        return fail(RBL_ERR_SINGULAR, "K^T*K is singular (is your rigid body a dimer?)");
      pc_set = false;
      return fail(RBL_ERR_SINGULAR, "preconditioner block is singular");
    }
    return RBL_OK;
  }

  // ---- state --------------------------------------------------------------------------
  int set_parameters(double a_, double dt_, double kBT_, double eta_, const void* cfg,
                     int n_blb_) override {
    if (n_blb_ <= 0 || !cfg) return fail(RBL_ERR_INVALID, "setParameters: empty rigid configuration");
    if (!(a_ > 0) || !(eta_ > 0)) return fail(RBL_ERR_INVALID, "setParameters: a and eta must be positive");
    a = a_; dt = dt_; kBT = kBT_; eta = eta_;
    n_blb = n_blb_;
    // removeMean (c_rigid_obj.cpp:176-181), in `real` like the reference
    const real* c = static_cast<const real*>(cfg);
    std::vector<real> ref(c, c + 3 * (size_t)n_blb);
    real mean[3] = {0, 0, 0};
    for (int k = 0; k < n_blb; ++k)
      for (int d = 0; d < 3; ++d) mean[d] += ref[3 * k + d];
    for (int d = 0; d < 3; ++d) mean[d] /= (real)n_blb;
    for (int k = 0; k < n_blb; ++k)
      for (int d = 0; d < 3; ++d) ref[3 * k + d] -= mean[d];
    CK(d_ref.ensure(ref.size() * sizeof(real)));
    RET(h2d(d_ref.p, ref.data(), ref.size() * sizeof(real)));
    ref_host.assign(ref.begin(), ref.end());
    delete shadow;  // the mirror follows the parameters: rebuilt at its next use
    shadow = nullptr;
    CK(cudaStreamSynchronize(stream));
    params_set = true;
    r_valid = false;
    pc_set = false;
    noise_set = false;
    noise_shared_ready = false;
    return RBL_OK;
  }

  int set_flags(int block_pc_, int wall_) override {
    if (block_pc_ >= 0 && (block_pc_ != 0) != block_pc) { block_pc = block_pc_ != 0; pc_set = false; }
    if (wall_ >= 0 && (wall_ != 0) != wall) { wall = wall_ != 0; pc_set = false; noise_set = false; }
    return RBL_OK;
  }

  int set_config(const void* X, const void* Q, int n_bod_) override {
    if (!params_set) return fail(RBL_ERR_STATE, "setConfig before setParameters");
    if (n_bod_ <= 0 || !X || !Q) return fail(RBL_ERR_INVALID, "setConfig: empty configuration");
    n_bod = n_bod_;
    CK(d_X.ensure(3 * (size_t)n_bod * sizeof(real)));
    CK(d_Q.ensure(4 * (size_t)n_bod * sizeof(real)));
    RET(h2d(d_X.p, X, 3 * (size_t)n_bod * sizeof(real)));
    RET(h2d(d_Q.p, Q, 4 * (size_t)n_bod * sizeof(real)));
    LAUNCH(1, rbl::normalize_quats<real>(d_Q.as<real>(), n_bod, stream));  // :216
    CK(cudaStreamSynchronize(stream));
    cfg_set = true;
    r_valid = false;
    // DELIBERATE DEVIATION.  The reference keeps a built preconditioner across setConfig (only
    // evolve_X_Q resets PC_mat_Set, :877): apply_PC then mixes the OLD invM / N_lu with the NEW K
    // (:601-610).  Here the factors are device-resident and partly expressed through the current
    // configuration (rotated shared free-space factor, K products in pc_finish), so a kept PC would be
    // neither the reference's stale one nor a fresh one.  The preconditioner is therefore rebuilt at its
    // next use: apply_PC after set_config is the preconditioner OF THAT CONFIGURATION (what a user of
    // the reference gets by calling evolve / constructing anew).  tests/test_gpu_rigid.py pins this.
    pc_set = false;
    return RBL_OK;
  }

  int get_config(void* X, void* Q) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    RET(d2h(X, d_X.p, 3 * (size_t)n_bod * sizeof(real)));
    RET(d2h(Q, d_Q.p, 4 * (size_t)n_bod * sizeof(real)));
    CK(cudaStreamSynchronize(stream));
    return RBL_OK;
  }

  int set_K_mats() override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    CK(d_r.ensure(3 * (size_t)N() * sizeof(real)));
    CK(d_S.ensure(9 * (size_t)n_bod * sizeof(real)));
    LAUNCH(1, rbl::place_blobs<real>(d_X.as<real>(), d_Q.as<real>(), d_ref.as<real>(), n_bod, n_blb,
                                     d_r.as<real>(), stream));
    LAUNCH(1, rbl::ktk_inv_blocks<real>(d_Q.as<real>(), d_ref.as<real>(), n_bod, n_blb, d_S.as<real>(),
                                        d_flags.as<int>() + FLAG_SINGULAR, stream));
    r_valid = true;
    r_all_valid = false;
    cfg_gen = ++gen_counter;
    noise_set = false;  // the per-body factors follow the blob positions
    return RBL_OK;
  }
  int need_K() {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    if (!r_valid) RET(set_K_mats());
    return RBL_OK;
  }
  int n_bodies() const override { return n_bod; }
  int blobs_per_body() const override { return n_blb; }

  // ---- O(N) operators -------------------------------------------------------------------
  int blob_positions(void* out, bool dev) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    const size_t bytes = 3 * (size_t)N() * sizeof(real);
    if (dev) {
      LAUNCH(1, rbl::place_blobs<real>(d_X.as<real>(), d_Q.as<real>(), d_ref.as<real>(), n_bod, n_blb,
                                       static_cast<real*>(out), stream));
      return RBL_OK;
    }
    CK(d_out0.ensure(bytes));
    LAUNCH(1, rbl::place_blobs<real>(d_X.as<real>(), d_Q.as<real>(), d_ref.as<real>(), n_bod, n_blb,
                                     d_out0.as<real>(), stream));
    RET(d2h(out, d_out0.p, bytes));
    CK(cudaStreamSynchronize(stream));
    return RBL_OK;
  }

  int K_dot(const void* U, void* out, bool dev) override {
    RET(need_K());
    const size_t nin = 6 * (size_t)n_bod * sizeof(real), nout = 3 * (size_t)N() * sizeof(real);
    const real* dU = static_cast<const real*>(U);
    real* dO = static_cast<real*>(out);
    if (!dev) {
      CK(d_in0.ensure(nin));
      CK(d_out0.ensure(nout));
      RET(h2d(d_in0.p, U, nin));
      dU = d_in0.as<real>();
      dO = d_out0.as<real>();
    }
    LAUNCH(1, rbl::k_dot<real>(dU, d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, (real)1, nullptr, dO, stream));
    if (!dev) {
      RET(d2h(out, dO, nout));
      RET(sync());
    }
    return RBL_OK;
  }

  int KT_dot(const void* lam, void* out, bool dev) override {
    RET(need_K());
    const size_t nin = 3 * (size_t)N() * sizeof(real), nout = 6 * (size_t)n_bod * sizeof(real);
    const real* dL = static_cast<const real*>(lam);
    real* dO = static_cast<real*>(out);
    if (!dev) {
      CK(d_in0.ensure(nin));
      CK(d_out0.ensure(nout));
      RET(h2d(d_in0.p, lam, nin));
      dL = d_in0.as<real>();
      dO = d_out0.as<real>();
    }
    LAUNCH(1, rbl::kt_dot<real>(dL, d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, dO, stream));
    if (!dev) {
      RET(d2h(out, dO, nout));
      RET(sync());
    }
    return RBL_OK;
  }

  int Kinv_dot(const void* V, void* out) override {  // (K^T K)^-1 K^T V  (:390,406)
    RET(need_K());
    const size_t nin = 3 * (size_t)N() * sizeof(real), nout = 6 * (size_t)n_bod * sizeof(real);
    CK(d_in0.ensure(nin));
    CK(d_out0.ensure(nout));
    RET(h2d(d_in0.p, V, nin));
    LAUNCH(1, rbl::kt_dot<real>(d_in0.as<real>(), d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, d_out0.as<real>(), stream));
    LAUNCH(1, rbl::ktk_inv_apply<real>(d_S.as<real>(), n_bod, n_blb, d_out0.as<real>(), stream));
    RET(d2h(out, d_out0.p, nout));
    return sync();
  }
  int KTinv_dot(const void* F, void* out) override {  // K (K^T K)^-1 F  (:408)
    RET(need_K());
    const size_t nin = 6 * (size_t)n_bod * sizeof(real), nout = 3 * (size_t)N() * sizeof(real);
    CK(d_in0.ensure(nin));
    CK(d_out0.ensure(nout));
    RET(h2d(d_in0.p, F, nin));
    LAUNCH(1, rbl::ktk_inv_apply<real>(d_S.as<real>(), n_bod, n_blb, d_in0.as<real>(), stream));
    LAUNCH(1, rbl::k_dot<real>(d_in0.as<real>(), d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, (real)1, nullptr, d_out0.as<real>(), stream));
    RET(d2h(out, d_out0.p, nout));
    return sync();
  }

  // ---- the mobility product ---------------------------------------------------------------
  int pick_variant(int n_tgt) const {
    if (variant >= 0) return variant;
    // big problems: the widest register tile; small ones: the smallest target tile so the
    // stream-K unit grid still covers the SMs
    return n_tgt >= 16384 ? 0 : rbl::matvec_num_variants<real>() - 1;
  }

  int pick_sym_variant(int n) const {
    if (sym_variant >= 0) return sym_variant;
    return rbl::matvec_sym_default_variant<real>(wall, n);
  }

  // symmetric kernel: share `part` of `n_parts` of the unordered-pair work; out = partial product
  int dev_apply_M_part(const void* F, const void* r, int n, int part, int n_parts, void* out) override {
    return apply_M_part_gen(F, r, n, part, n_parts, out, 0);
  }
  int apply_M_part_gen(const void* F, const void* r, int n, int part, int n_parts, void* out, unsigned long long gen) {
    if (!params_set) return fail(RBL_ERR_STATE, "apply_M before setParameters");
    if (n < 0 || n_parts < 1 || part < 0 || part >= n_parts) return fail(RBL_ERR_INVALID, "apply_M_part: bad share");
    if (n == 0) return RBL_OK;
    const int v = pick_sym_variant(n);
    rbl::SymArgs<real> A;
    CK(rbl::matvec_sym_plan<real>(v, wall, n, part, n_parts, sm_count, &A.plan));
    RET(cost_bounds(bounds1, &A.plan, v, part, n_parts, &A.bounds));
    const size_t n_pad = (size_t)A.plan.n_src_tiles * rbl::kSrcTile;
    CK(d_rec.ensure(n_pad * rbl::kRecReals * sizeof(real)));
    CK(d_box_src.ensure(6 * (size_t)A.plan.n_src_tiles * sizeof(float)));
    CK(d_box_tgt.ensure(6 * (size_t)A.plan.n_tgt_tiles * sizeof(float)));
    CK(d_raw.ensure(3 * n_pad * sizeof(real)));
    if (rec1_state.matches(gen, n, A.plan.tgt_tile, d_rec.p)) {
      LAUNCH(1, rbl::repack_forces<real>(static_cast<const real*>(F), n, wall, (real)a, d_rec.as<real>(),
                                         d_flags.as<int>() + FLAG_BELOW, stream));
    } else {
      LAUNCH(1, rbl::pack_records<real>(static_cast<const real*>(r), static_cast<const real*>(F), n, (int)n_pad, wall,
                                        (real)a, d_rec.as<real>(), d_flags.as<int>() + FLAG_BELOW, stream));
      LAUNCH(1, rbl::tile_boxes<real>(d_rec.as<real>(), 0, (int)n_pad, rbl::kSrcTile, d_box_src.as<float>(), stream));
      LAUNCH(1, rbl::tile_boxes<real>(d_rec.as<real>(), 0, n, A.plan.tgt_tile, d_box_tgt.as<float>(), stream));
      rec1_state = {gen, n, A.plan.tgt_tile, d_rec.p};
      rec2_state.gen = 0;  // the two kernels share the tile-box buffers
    }
    A.rec = d_rec.as<real>();
    A.box_src = d_box_src.as<float>();
    A.box_tgt = d_box_tgt.as<float>();
    A.raw = d_raw.as<real>();
    A.out = static_cast<real*>(out);
    A.C = rbl::make_pair_consts<real>(a, eta);
    A.wall = wall ? 1 : 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (profile) {
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      prof_events.emplace_back(e0, e1);
    }
    LAUNCH(2, rbl::matvec_sym_launch<real>(v, A, stream, e0, e1));
    ++products;
    return RBL_OK;
  }

  int saddle_finish(const void* Mlam, const void* lam, const void* U, void* out) override {
    RET(need_K());
    const size_t n3 = 3 * (size_t)N();
    real* dout = static_cast<real*>(out);
    LAUNCH(1, rbl::k_dot<real>(static_cast<const real*>(U), d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, (real)-1,
                               static_cast<const real*>(Mlam), dout, stream));
    LAUNCH(1, rbl::kt_dot<real>(static_cast<const real*>(lam), d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, dout + n3, stream));
    return RBL_OK;
  }

  int dev_apply_M(const void* F, const void* r, int n, int t0, int nt, void* out) override {
    if (!params_set) return fail(RBL_ERR_STATE, "apply_M before setParameters");
    if (n < 0 || t0 < 0 || nt < 0 || t0 + nt > n) return fail(RBL_ERR_INVALID, "apply_M: bad target range");
    if (n == 0 || nt == 0) return RBL_OK;
    if (mode == 0 && t0 == 0 && nt == n) return apply_M_part_gen(F, r, n, 0, 1, out, 0);
    const int v = pick_variant(nt);
    rbl::MatvecArgs<real> A;
    CK(rbl::matvec_plan<real>(v, wall, n, t0, nt, sm_count, &A.plan));
    const size_t n_pad = (size_t)A.plan.n_src_tiles * rbl::kSrcTile;
    CK(d_rec.ensure(n_pad * rbl::kRecReals * sizeof(real)));
    CK(d_box_src.ensure(6 * (size_t)A.plan.n_src_tiles * sizeof(float)));
    CK(d_box_tgt.ensure(6 * (size_t)A.plan.n_tgt_tiles * sizeof(float)));
    CK(d_scratch.ensure(2 * (size_t)A.plan.grid * 3 * A.plan.tgt_tile * sizeof(real)));
    rec1_state.gen = rec2_state.gen = 0;  // this path repacks the shared record and box buffers
    LAUNCH(1, rbl::pack_records<real>(static_cast<const real*>(r), static_cast<const real*>(F), n, (int)n_pad, wall,
                                      (real)a, d_rec.as<real>(), d_flags.as<int>() + FLAG_BELOW, stream));
    LAUNCH(1, rbl::tile_boxes<real>(d_rec.as<real>(), 0, (int)n_pad, rbl::kSrcTile, d_box_src.as<float>(), stream));
    LAUNCH(1, rbl::tile_boxes<real>(d_rec.as<real>(), t0, nt, A.plan.tgt_tile, d_box_tgt.as<float>(), stream));
    A.rec = d_rec.as<real>();
    A.box_src = d_box_src.as<float>();
    A.box_tgt = d_box_tgt.as<float>();
    A.out = static_cast<real*>(out);
    A.scratch = d_scratch.as<real>();
    A.C = rbl::make_pair_consts<real>(a, eta);
    A.wall = wall ? 1 : 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (profile) {
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      prof_events.emplace_back(e0, e1);
    }
    LAUNCH(2, rbl::matvec_launch<real>(v, A, stream, e0, e1));
    ++products;
    return RBL_OK;
  }


  // ---- partitioned mode ------------------------------------------------
  // The context holds this rank's bodies; vectors are rank-local slices [lambda_local ; U_local].
  int comm_init(const void* uid128, int rank, int world, const int* blobs_per_rank) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "rbl_comm_init: set this rank's configuration first");
    if (!uid128 || !blobs_per_rank || world < 1 || rank < 0 || rank >= world)
      return fail(RBL_ERR_INVALID, "rbl_comm_init: bad rank / world / arguments");
    if ((long long)blobs_per_rank[rank] != N())
      return fail(RBL_ERR_INVALID, "rbl_comm_init: blobs_per_rank[rank] differs from this context's blob count");
    long long tot = 0;
    for (int r = 0; r < world; ++r) tot += blobs_per_rank[r];
    if (tot > 0x7fffffffLL) return fail(RBL_ERR_INVALID, "rbl_comm_init: more than 2^31-1 blobs");
    if (comm) comm->peer_release(true, d_status.as<int>(), stream);
    delete comm;
    comm = nullptr;
    CK(cudaStreamSynchronize(stream));
    auto* c = new rbl::Comm();
    if (!c->init(uid128, rank, world, blobs_per_rank)) {
      std::string m = c->err;
      delete c;
      return fail(RBL_ERR_CUDA, "rbl_comm_init: " + m);
    }
    comm = c;
    r_all_valid = false;
    CK(d_status.ensure(4 * sizeof(int)));
    // the exchanges around every product go over peer memory when every rank can map every other rank's
    // buffer (rbl_peer.cuh); otherwise NCCL collectives (rbl_comm_exchange tells which, and why)
    comm->peer_setup(sizeof(real), d_status.as<int>(), stream);
    return RBL_OK;
  }
  int comm_exchange(const char** why) const override {
    if (why) *why = comm ? comm->peer_why.c_str() : "no communicator";
    return comm && comm->peer_active() ? 1 : 0;
  }
  int comm_set_exchange(int mode) override {
    if (!comm) return fail(RBL_ERR_STATE, "rbl_comm_set_exchange: no communicator");
    if (mode != 0 && mode != 1) return fail(RBL_ERR_INVALID, "rbl_comm_set_exchange: mode must be 0 (NCCL) or 1 (peer memory)");
    if (mode == 1 && !comm->peer.on) return fail(RBL_ERR_STATE, "rbl_comm_set_exchange: the peer-memory exchange is not available: " + comm->peer_why);
    RET(csync());  // a quiescent point on every rank: the two protocols never interleave
    comm->peer.use = mode == 1;
    return RBL_OK;
  }
  // events around the two exchanges of every product while profiling is on: [before gather, after gather,
  // before reduce, after reduce]
  std::vector<std::array<cudaEvent_t, 4>> prof_comm;
  int comm_mark(int k) {
    if (!profile || !comm) return RBL_OK;
    if (k == 0) {
      std::array<cudaEvent_t, 4> ev{};
      for (auto& e : ev) CK(cudaEventCreate(&e));
      prof_comm.push_back(ev);
    }
    CK(cudaEventRecord(prof_comm.back()[k], stream));
    return RBL_OK;
  }
  int comm_profile(double* ms3, int* n, int reset) override {
    CK(cudaStreamSynchronize(stream));
    double acc[3] = {0, 0, 0};
    for (auto& ev : prof_comm)
      for (int k = 0; k < 3; ++k) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, ev[k], ev[k + 1]));
        acc[k] += t;
      }
    const int m = (int)prof_comm.size();
    if (ms3)
      for (int k = 0; k < 3; ++k) ms3[k] = m ? acc[k] / m : 0;
    if (n) *n = m;
    if (reset) {
      for (auto& ev : prof_comm)
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
      prof_comm.clear();
    }
    return RBL_OK;
  }
#define NK(call)                                                        \
  do {                                                                  \
    if (!(call)) return this->fail(RBL_ERR_CUDA, "NCCL: " + comm->err); \
  } while (0)

  // every rank returns the same status: the largest one seen anywhere (a rank that failed alone
  // would otherwise leave the others waiting inside the next collective)
  int agree(int st) {
    if (!comm) return st;
    int h = st;
    CK(cudaMemcpyAsync(d_status.p, &h, sizeof(int), cudaMemcpyHostToDevice, stream));
    NK(comm->allreduce_max_int(d_status.as<int>(), 1, stream));
    CK(cudaMemcpyAsync(&h, d_status.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    if (h != RBL_OK && st == RBL_OK) return fail(h, "another rank of the partitioned suspension reported an error");
    return st != RBL_OK ? st : h;
  }
  int csync() { return agree(sync()); }

  // Partitioned mode: every allocation and plan a product needs, made BEFORE the first collective of a
  // driver and agreed between the ranks (agree()), so that an out-of-memory or a failed plan on one rank
  // surfaces as the same error everywhere instead of leaving the other ranks inside the next collective.
  int reserve_comm_workspace(bool two_rhs) {
    if (!comm) return RBL_OK;
    const int n = (int)comm->n_all;
    const size_t n3 = 3 * (size_t)n;
    rbl::SymPlan plan;
    CK(rbl::matvec_sym_plan<real>(pick_sym_variant(n), wall, n, comm->rank, comm->world, sm_count, &plan));
    size_t n_pad = (size_t)plan.n_src_tiles * rbl::kSrcTile;
    const void* before = d_r_all.p;
    CK(d_r_all.ensure(2 * n3 * sizeof(real)));
    if (d_r_all.p != before) r_all_valid = false;
    if (!comm->peer_active()) {  // (the peer exchange keeps both inside its symmetric buffer)
      CK(d_lam_all.ensure((two_rhs ? 2 : 1) * n3 * sizeof(real)));
      CK(d_mbuf.ensure((two_rhs ? 2 : 1) * n3 * sizeof(real)));
    }
    CK(d_rec.ensure(n_pad * rbl::kRecReals * sizeof(real)));
    CK(d_raw.ensure(3 * n_pad * sizeof(real)));
    size_t tgt_tiles = (size_t)plan.n_tgt_tiles;
    if (two_rhs) {
      const int v2 = sym2_variant >= 0 ? sym2_variant : rbl::matvec_sym2_default_variant<real>(wall, n);
      CK(rbl::matvec_sym2_plan<real>(v2, wall, n, comm->rank, comm->world, sm_count, &plan));
      n_pad = (size_t)plan.n_src_tiles * rbl::kSrcTile;
      CK(d_rec2.ensure(n_pad * rbl::kRec2Reals * sizeof(real)));
      CK(d_raw2.ensure(2 * 3 * n_pad * sizeof(real)));
      tgt_tiles = std::max(tgt_tiles, (size_t)plan.n_tgt_tiles);
    }
    CK(d_box_src.ensure(6 * (size_t)plan.n_src_tiles * sizeof(float)));
    CK(d_box_tgt.ensure(6 * tgt_tiles * sizeof(float)));
    return RBL_OK;
  }
  int agree_workspace(bool two_rhs) { return comm ? agree(reserve_comm_workspace(two_rhs)) : (int)RBL_OK; }

  // out_local = rows of this rank of  B M B F  where F_local / r_local are this rank's slices.
  // Single context: the plain product.  Partitioned: all-gather F (and the positions, unless they
  // are the cached configuration), this rank's share of the unordered-pair work over ALL blobs,
  // reduce-scatter of the partial products (each rank keeps the sum of its own rows).
  int prod_M(const real* F_local, const real* r_local, bool r_is_config, real* out_local) {
    const int nl = (int)N();
    if (!comm) {
      if (mode == 0) return apply_M_part_gen(F_local, r_local, nl, 0, 1, out_local, r_is_config ? cfg_gen : 0);
      return dev_apply_M(F_local, r_local, nl, 0, nl, out_local);
    }
    const size_t n3_all = 3 * (size_t)comm->n_all;
    const bool peer = comm->peer_active();
    const void* before = d_r_all.p;
    CK(d_r_all.ensure(2 * n3_all * sizeof(real)));  // [configuration ; scratch positions (RFD)]
    if (d_r_all.p != before) r_all_valid = false;
    if (!peer) {
      CK(d_lam_all.ensure(n3_all * sizeof(real)));
      CK(d_mbuf.ensure(n3_all * sizeof(real)));
    }
    real* lam = peer ? comm->peer_lam<real>() : d_lam_all.as<real>();
    real* mb = peer ? comm->peer_mbuf<real>() : d_mbuf.as<real>();
    real* r_all = d_r_all.as<real>();
    if (r_is_config) {
      if (!r_all_valid) {
        NK(comm->allgatherv<real>(r_local, r_all, 3, stream));
        r_all_valid = true;
        r_all_gen = ++gen_counter;
      }
    } else {
      r_all += n3_all;
      NK(comm->allgatherv<real>(r_local, r_all, 3, stream));
    }
    int* pf = d_flags.as<int>() + FLAG_PEER;
    RET(comm_mark(0));
    if (peer) {
      NK(comm->peer_allgather<real>(&F_local, 1, pf, stream));
      launches += 2;
    } else {
      NK(comm->allgatherv<real>(F_local, lam, 3, stream));
    }
    RET(comm_mark(1));
    RET(apply_M_part_gen(lam, r_all, (int)comm->n_all, comm->rank, comm->world, mb, r_is_config ? r_all_gen : 0));
    RET(comm_mark(2));
    if (peer) {
      NK(comm->peer_reduce_scatter<real>(&out_local, 1, pf, stream));
      launches += 2;
    } else {
      NK(comm->reduce_scatterv<real>(mb, out_local, 3, stream));
    }
    RET(comm_mark(3));
    return RBL_OK;
  }
  // d_dots[slot .. slot+m) = <V_i, w> summed over the ranks (results stay on the device; the Krylov
  // drivers read every scalar of an iteration back with ONE copy, read_slots)
  int gdots(const real* V, size_t ld, int m, const real* w, size_t n, int slot = 0) {
    LAUNCH(2, rbl::multi_dot<real>(V, ld, m, w, n, d_partial.as<real>(), d_dots.as<real>() + slot, stream));
    if (comm) NK(comm->allreduce_sum<real>(d_dots.as<real>() + slot, (size_t)m, stream));
    return RBL_OK;
  }


  // ---- two right-hand sides per pass (rpy_matvec_sym2_kernel) ----------------------------------
  int dev_apply_M2_part(const void* F1, const void* F2, const void* r, int n, int part, int n_parts, void* out1,
                        void* out2, unsigned long long gen = 0) {
    if (!params_set) return fail(RBL_ERR_STATE, "apply_M before setParameters");
    if (n < 0 || n_parts < 1 || part < 0 || part >= n_parts) return fail(RBL_ERR_INVALID, "apply_M2_part: bad share");
    if (n == 0) return RBL_OK;
    const int v = sym2_variant >= 0 ? sym2_variant : rbl::matvec_sym2_default_variant<real>(wall, n);
    rbl::Sym2Args<real> A;
    CK(rbl::matvec_sym2_plan<real>(v, wall, n, part, n_parts, sm_count, &A.plan));
    RET(cost_bounds(bounds2, &A.plan, v, part, n_parts, &A.bounds));
    const size_t n_pad = (size_t)A.plan.n_src_tiles * rbl::kSrcTile;
    CK(d_rec2.ensure(n_pad * rbl::kRec2Reals * sizeof(real)));
    CK(d_box_src.ensure(6 * (size_t)A.plan.n_src_tiles * sizeof(float)));
    CK(d_box_tgt.ensure(6 * (size_t)A.plan.n_tgt_tiles * sizeof(float)));
    CK(d_raw2.ensure(2 * 3 * n_pad * sizeof(real)));
    if (rec2_state.matches(gen, n, A.plan.tgt_tile, d_rec2.p)) {
      LAUNCH(1, rbl::repack_forces2<real>(static_cast<const real*>(F1), static_cast<const real*>(F2), n, wall, (real)a,
                                          d_rec2.as<real>(), d_flags.as<int>() + FLAG_BELOW, stream));
    } else {
      LAUNCH(1, rbl::pack_records2<real>(static_cast<const real*>(r), static_cast<const real*>(F1), static_cast<const real*>(F2),
                                         n, (int)n_pad, wall, (real)a, d_rec2.as<real>(), d_flags.as<int>() + FLAG_BELOW, stream));
      LAUNCH(1, rbl::tile_boxes<real>(d_rec2.as<real>(), 0, (int)n_pad, rbl::kSrcTile, d_box_src.as<float>(), stream, rbl::kRec2Reals));
      LAUNCH(1, rbl::tile_boxes<real>(d_rec2.as<real>(), 0, n, A.plan.tgt_tile, d_box_tgt.as<float>(), stream, rbl::kRec2Reals));
      rec2_state = {gen, n, A.plan.tgt_tile, d_rec2.p};
      rec1_state.gen = 0;  // the two kernels share the tile-box buffers
    }
    A.rec = d_rec2.as<real>();
    A.box_src = d_box_src.as<float>();
    A.box_tgt = d_box_tgt.as<float>();
    A.raw = d_raw2.as<real>();
    A.out1 = static_cast<real*>(out1);
    A.out2 = static_cast<real*>(out2);
    A.C = rbl::make_pair_consts<real>(a, eta);
    A.wall = wall ? 1 : 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (profile) {
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      prof_events.emplace_back(e0, e1);
    }
    LAUNCH(2, rbl::matvec_sym2_launch<real>(v, A, stream, e0, e1));
    products += 2;
    return RBL_OK;
  }
  // out_k_local = rows of this rank of B M B F_k (k = 1, 2) at the cached configuration
  int prod_M2(const real* F1_local, const real* F2_local, real* out1_local, real* out2_local) {
    const int nl = (int)N();
    if (!comm) return dev_apply_M2_part(F1_local, F2_local, d_r.p, nl, 0, 1, out1_local, out2_local, cfg_gen);
    const size_t n3_all = 3 * (size_t)comm->n_all;
    const bool peer = comm->peer_active();
    const void* before = d_r_all.p;
    CK(d_r_all.ensure(2 * n3_all * sizeof(real)));
    if (d_r_all.p != before) r_all_valid = false;
    if (!peer) {
      CK(d_lam_all.ensure(2 * n3_all * sizeof(real)));
      CK(d_mbuf.ensure(2 * n3_all * sizeof(real)));
    }
    if (!r_all_valid) {
      NK(comm->allgatherv<real>(d_r.as<real>(), d_r_all.as<real>(), 3, stream));
      r_all_valid = true;
      r_all_gen = ++gen_counter;
    }
    real* lam1 = peer ? comm->peer_lam<real>(0) : d_lam_all.as<real>();
    real* lam2 = peer ? comm->peer_lam<real>(1) : lam1 + n3_all;
    real* mb1 = peer ? comm->peer_mbuf<real>(0) : d_mbuf.as<real>();
    real* mb2 = peer ? comm->peer_mbuf<real>(1) : mb1 + n3_all;
    int* pf = d_flags.as<int>() + FLAG_PEER;
    RET(comm_mark(0));
    if (peer) {
      const real* send[2] = {F1_local, F2_local};
      NK(comm->peer_allgather<real>(send, 2, pf, stream));
      launches += 3;
    } else {
      NK(comm->allgatherv<real>(F1_local, lam1, 3, stream));
      NK(comm->allgatherv<real>(F2_local, lam2, 3, stream));
    }
    RET(comm_mark(1));
    RET(dev_apply_M2_part(lam1, lam2, d_r_all.p, (int)comm->n_all, comm->rank, comm->world, mb1, mb2, r_all_gen));
    RET(comm_mark(2));
    if (peer) {
      real* recv[2] = {out1_local, out2_local};
      NK(comm->peer_reduce_scatter<real>(recv, 2, pf, stream));
      launches += 3;
    } else {
      NK(comm->reduce_scatterv<real>(mb1, out1_local, 3, stream));
      NK(comm->reduce_scatterv<real>(mb2, out2_local, 3, stream));
    }
    RET(comm_mark(3));
    return RBL_OK;
  }
  int dev_apply_M2(const void* F1, const void* F2, const void* r, int n, void* out1, void* out2) override {
    return dev_apply_M2_part(F1, F2, r, n, 0, 1, out1, out2);
  }
  int apply_M2(const void* F1, const void* F2, const void* r, int n, void* out1, void* out2) override {
    if (n < 0) return fail(RBL_ERR_INVALID, "apply_M2: negative size");
    if (n == 0) return RBL_OK;
    const size_t bytes = 3 * (size_t)n * sizeof(real);
    CK(d_in0.ensure(bytes));
    CK(d_in1.ensure(bytes));
    CK(d_in2.ensure(bytes));
    CK(d_out0.ensure(bytes));
    CK(d_out2.ensure(bytes));
    RET(h2d(d_in0.p, F1, bytes));
    RET(h2d(d_in2.p, F2, bytes));
    RET(h2d(d_in1.p, r, bytes));
    RET(dev_apply_M2_part(d_in0.p, d_in2.p, d_in1.p, n, 0, 1, d_out0.p, d_out2.p));
    RET(d2h(out1, d_out0.p, bytes));
    RET(d2h(out2, d_out2.p, bytes));
    return sync();
  }
  int num_sym2_variants() const override { return rbl::matvec_sym2_num_variants<real>(); }
  int sym2_variant_info(int idx, int* T, int* threads) const override {
    if (idx < 0 || idx >= rbl::matvec_sym2_num_variants<real>()) return RBL_ERR_INVALID;
    const rbl::MatvecVariant v = rbl::matvec_sym2_variant<real>(idx);
    *T = v.T;
    *threads = v.threads;
    return RBL_OK;
  }

  int apply_M(const void* F, const void* r, int n, void* out) override {
    if (n < 0) return fail(RBL_ERR_INVALID, "apply_M: negative size");
    if (n == 0) return RBL_OK;
    const size_t bytes = 3 * (size_t)n * sizeof(real);
    CK(d_in0.ensure(bytes));
    CK(d_in1.ensure(bytes));
    CK(d_out0.ensure(bytes));
    RET(h2d(d_in0.p, F, bytes));
    RET(h2d(d_in1.p, r, bytes));
    RET(dev_apply_M(d_in0.p, d_in1.p, n, 0, n, d_out0.p));
    RET(d2h(out, d_out0.p, bytes));
    return sync();
  }

  int dev_saddle(const real* dx, real* dout) {
    RET(need_K());
    const int n = (int)N();
    // slip = M lam - K U ; F = K^T lam   (Rigid.py:73-80)
    RET(prod_M(dx, d_r.as<real>(), true, dout));
    LAUNCH(1, rbl::k_dot<real>(dx + 3 * (size_t)n, d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, (real)-1, dout, dout, stream));
    LAUNCH(1, rbl::kt_dot<real>(dx, d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, dout + 3 * (size_t)n, stream));
    return RBL_OK;
  }
  int saddle_shard(const void* lam_all, const void* r_all, int n_all, int t0, const void* U, void* out) override {
    RET(need_K());
    const int nl = (int)N();
    if (t0 < 0 || t0 + nl > n_all) return fail(RBL_ERR_INVALID, "saddle shard: local blob range outside the global range");
    real* dout = static_cast<real*>(out);
    const real* lam_local = static_cast<const real*>(lam_all) + 3 * (size_t)t0;
    RET(dev_apply_M(lam_all, r_all, n_all, t0, nl, dout));
    LAUNCH(1, rbl::k_dot<real>(static_cast<const real*>(U), d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, (real)-1, dout, dout, stream));
    LAUNCH(1, rbl::kt_dot<real>(lam_local, d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, dout + 3 * (size_t)nl, stream));
    return RBL_OK;
  }
  int apply_saddle(const void* x, void* out, bool dev) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    if (dev) return dev_saddle(static_cast<const real*>(x), static_cast<real*>(out));
    const size_t bytes = sys_size() * sizeof(real);
    if (comm) {  // allocations agreed before the first collective
      int st = reserve_comm_workspace(false);
      if (st == RBL_OK && (d_in0.ensure(bytes) != cudaSuccess || d_out0.ensure(bytes) != cudaSuccess)) {
        cudaGetLastError();
        st = fail(RBL_ERR_NOMEM, "apply_saddle: out of device memory");
      }
      st = agree(st);
      if (st != RBL_OK) return st;
    }
    CK(d_in0.ensure(bytes));
    CK(d_out0.ensure(bytes));
    RET(h2d(d_in0.p, x, bytes));
    RET(dev_saddle(d_in0.as<real>(), d_out0.as<real>()));
    RET(d2h(out, d_out0.p, bytes));
    return csync();
  }

  // ---- preconditioner -----------------------------------------------------------------------
  int build_pc() {
    RET(need_K());
    const int sz = 3 * n_blb;
    const size_t n3 = 3 * (size_t)N();
    CK(d_Kc.ensure(6 * n3 * sizeof(real)));
    CK(d_Y.ensure(6 * n3 * sizeof(real)));
    CK(d_L.ensure(36 * (size_t)n_bod * sizeof(real)));
    CK(d_y.ensure(n3 * sizeof(real)));
    int* fl = d_flags.as<int>();
    LAUNCH(1, rbl::pc_fill_kcols<real>(d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, d_Kc.as<real>(), stream));
    pc_chol = false;
    if (!block_pc) {
      CK(d_dinv.ensure(n3 * sizeof(real)));
      LAUNCH(1, rbl::pc_diag_build<real>(d_r.as<real>(), (int)N(), (real)a, (real)eta, wall, d_dinv.as<real>(), fl + FLAG_BELOW, stream));
      LAUNCH(1, rbl::pc_diag_mul<real>(d_dinv.as<real>(), d_Kc.as<real>(), n_bod, n_blb, 6, d_Y.as<real>(), stream));
    } else {
      // Block PC.  Mt_b^-1 = G_b^T G_b from the Cholesky factors the noise preconditioner uses (the same
      // matrices; 1/6 of the memory traffic of an explicit Gauss-Jordan inverse, and blocked kernels).
      // Free space: ONE factor of the reference shape, M_b^-1 = R G_ref^T G_ref R^T (rotation covariance),
      // where the reference inverts one dense matrix per body per configuration (:461-487).
      // Gauss-Jordan remains the path for blocks that are not positive definite (blobs inside the
      // wall-overlap layer, overlapping a = 1 test geometries).
      pc_shared = !wall;
      if (!noise_set) RET(build_noise_pc());
      pc_chol = noise_ok;
      const size_t st2 = pc_shared ? 0 : (size_t)sz * sz;
      if (pc_chol) {
        CK(d_pt.ensure(6 * n3 * sizeof(real)));
        LAUNCH(1, rbl::body_mat_mul<real>(d_NG.as<real>(), st2, d_Q.as<real>(), pc_shared, false, false, d_Kc.as<real>(), n_bod, n_blb,
                                          d_pt.as<real>(), stream, 6));
        LAUNCH(1, rbl::body_mat_mul<real>(d_NG.as<real>(), st2, d_Q.as<real>(), false, pc_shared, true, d_pt.as<real>(), n_bod, n_blb,
                                          d_Y.as<real>(), stream, 6));
      } else if (pc_shared) {
        CK(d_Minv.ensure((size_t)sz * sz * sizeof(real)));
        LAUNCH(1, rbl::pc_block_assemble<real>(d_ref.as<real>(), 1, n_blb, (real)a, (real)eta, false, d_Minv.as<real>(), fl + FLAG_BELOW, stream));
        LAUNCH(1, rbl::pc_block_invert<real>(d_Minv.as<real>(), 1, sz, fl + FLAG_NOT_SPD, stream));
        LAUNCH(1, rbl::pc_block_mul<real>(d_Minv.as<real>(), 0, d_Q.as<real>(), d_Kc.as<real>(), n_bod, n_blb, 6, d_Y.as<real>(), stream));
      } else {
        const size_t bytes = (size_t)n_bod * sz * sz * sizeof(real);
        cudaError_t e = d_Minv.ensure(bytes);
        if (e != cudaSuccess) {
          cudaGetLastError();
          char msg[256];
          snprintf(msg, sizeof(msg),
                   "block preconditioner with wall needs %.1f GB for %d dense %dx%d body blocks; use the "
                   "diagonal preconditioner (block_PC=False) at this size",
                   bytes / 1e9, n_bod, sz, sz);
          return fail(RBL_ERR_NOMEM, msg);
        }
        LAUNCH(1, rbl::pc_block_assemble<real>(d_r.as<real>(), n_bod, n_blb, (real)a, (real)eta, true, d_Minv.as<real>(), fl + FLAG_BELOW, stream));
        LAUNCH(1, rbl::pc_block_invert<real>(d_Minv.as<real>(), n_bod, sz, fl + FLAG_NOT_SPD, stream));
        LAUNCH(1, rbl::pc_block_mul<real>(d_Minv.as<real>(), (size_t)sz * sz, nullptr, d_Kc.as<real>(), n_bod, n_blb, 6, d_Y.as<real>(), stream));
      }
    }
    LAUNCH(1, rbl::pc_ninv_chol<real>(d_Y.as<real>(), d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, d_L.as<real>(), fl + FLAG_NOT_SPD, stream));
    RET(sync());  // below-wall / not-SPD surface here, like the reference's lazy build (:591-596)
    pc_set = true;
    pc_n_bod = n_bod;
    return RBL_OK;
  }

  int dev_pc(const real* din, real* dout) {
    if (!pc_set) RET(build_pc());
    const int sz = 3 * n_blb;
    const size_t n3 = 3 * (size_t)N();
    if (!block_pc) {
      LAUNCH(1, rbl::pc_diag_mul<real>(d_dinv.as<real>(), din, n_bod, n_blb, 1, d_y.as<real>(), stream));
    } else if (pc_chol) {
      const size_t st2 = pc_shared ? 0 : (size_t)sz * sz;
      LAUNCH(1, rbl::body_mat_mul<real>(d_NG.as<real>(), st2, d_Q.as<real>(), pc_shared, false, false, din, n_bod, n_blb, d_pt.as<real>(), stream));
      LAUNCH(1, rbl::body_mat_mul<real>(d_NG.as<real>(), st2, d_Q.as<real>(), false, pc_shared, true, d_pt.as<real>(), n_bod, n_blb, d_y.as<real>(), stream));
    } else
      LAUNCH(1, rbl::pc_block_mul<real>(d_Minv.as<real>(), pc_shared ? 0 : (size_t)sz * sz, pc_shared ? d_Q.as<real>() : nullptr,
                                        din, n_bod, n_blb, 1, d_y.as<real>(), stream));
    LAUNCH(1, rbl::pc_finish<real>(d_y.as<real>(), din + n3, d_Y.as<real>(), d_L.as<real>(), d_r.as<real>(), d_X.as<real>(),
                                   n_bod, n_blb, dout, stream));
    return RBL_OK;
  }
  int apply_PC(const void* in, void* out, bool dev) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    if (dev) return dev_pc(static_cast<const real*>(in), static_cast<real*>(out));
    const size_t bytes = sys_size() * sizeof(real);
    CK(d_in0.ensure(bytes));
    CK(d_out0.ensure(bytes));
    RET(h2d(d_in0.p, in, bytes));
    RET(dev_pc(d_in0.as<real>(), d_out0.as<real>()));
    RET(d2h(out, d_out0.p, bytes));
    return sync();
  }

  // ---- integrator ---------------------------------------------------------------------------
  int evolve(const void* U) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    const size_t bytes = 6 * (size_t)n_bod * sizeof(real);
    CK(d_in0.ensure(bytes));
    RET(h2d(d_in0.p, U, bytes));
    LAUNCH(1, rbl::integrate<real>(d_in0.as<real>(), (real)dt, n_bod, d_X.as<real>(), d_Q.as<real>(), d_X.as<real>(), d_Q.as<real>(), stream));
    RET(set_K_mats());  // :876
    pc_set = false;     // :877
    return sync();
  }

  // ---- CSC export of K and Kinv (host side; :368-390, :978-992) -------------------------------
  int fetch_geometry(std::vector<real>& r, std::vector<real>& X, std::vector<real>& S) {
    RET(need_K());
    r.resize(3 * (size_t)N());
    X.resize(3 * (size_t)n_bod);
    S.resize(9 * (size_t)n_bod);
    RET(d2h(r.data(), d_r.p, r.size() * sizeof(real)));
    RET(d2h(X.data(), d_X.p, X.size() * sizeof(real)));
    RET(d2h(S.data(), d_S.p, S.size() * sizeof(real)));
    return sync();
  }
  int export_K(int64_t* indptr, int32_t* indices, void* data_) override {
    std::vector<real> r, X, S;
    RET(fetch_geometry(r, X, S));
    real* data = static_cast<real*>(data_);
    int64_t nnz = 0;
    for (int b = 0; b < n_bod; ++b) {
      const size_t base = (size_t)b * n_blb;
      for (int c = 0; c < 6; ++c) {
        indptr[6 * b + c] = nnz;
        for (int k = 0; k < n_blb; ++k) {
          const size_t i = base + k;
          const real px = r[3 * i] - X[3 * b], py = r[3 * i + 1] - X[3 * b + 1], pz = r[3 * i + 2] - X[3 * b + 2];
          const int32_t row = (int32_t)(3 * i);
          switch (c) {
            case 0: case 1: case 2:
              indices[nnz] = row + c; data[nnz++] = 1; break;
            case 3:  // omega_x: rows y,z = (-pz, +py)   (:380-381)
              indices[nnz] = row + 1; data[nnz++] = -pz;
              indices[nnz] = row + 2; data[nnz++] = py; break;
            case 4:  // omega_y: rows x,z = (+pz, -px)   (:376,382)
              indices[nnz] = row + 0; data[nnz++] = pz;
              indices[nnz] = row + 2; data[nnz++] = -px; break;
            default: // omega_z: rows x,y = (-py, +px)   (:377-378)
              indices[nnz] = row + 0; data[nnz++] = -py;
              indices[nnz] = row + 1; data[nnz++] = px; break;
          }
        }
      }
    }
    indptr[6 * (size_t)n_bod] = nnz;
    return RBL_OK;
  }
  int export_Kinv(int64_t* indptr, int32_t* indices, void* data_) override {
    std::vector<real> r, X, S;
    RET(fetch_geometry(r, X, S));
    real* data = static_cast<real*>(data_);
    int64_t nnz = 0;
    const real inv_n = (real)1 / (real)n_blb;
    for (int b = 0; b < n_bod; ++b) {
      const real* Sb = &S[9 * (size_t)b];
      for (int k = 0; k < n_blb; ++k) {
        const size_t i = (size_t)b * n_blb + k;
        const real px = r[3 * i] - X[3 * b], py = r[3 * i + 1] - X[3 * b + 1], pz = r[3 * i + 2] - X[3 * b + 2];
        // rows of -[rho]x : the rotational part of K's row p
        const real kr[3][3] = {{0, pz, -py}, {-pz, 0, px}, {py, -px, 0}};
        for (int p = 0; p < 3; ++p) {
          indptr[3 * i + p] = nnz;
          indices[nnz] = 6 * b + p; data[nnz++] = inv_n;
          for (int q = 0; q < 3; ++q) {
            indices[nnz] = 6 * b + 3 + q;
            data[nnz++] = Sb[3 * q] * kr[p][0] + Sb[3 * q + 1] * kr[p][1] + Sb[3 * q + 2] * kr[p][2];
          }
        }
      }
    }
    indptr[3 * (size_t)N()] = nnz;
    return RBL_OK;
  }

  // ---- Krylov drivers -------------------------------------------------------------------------
  int read_scalars(const real* d, int m, std::vector<double>& out) {
    std::vector<real> h(m);
    RET(d2h(h.data(), d, m * sizeof(real)));
    CK(cudaStreamSynchronize(stream));
    out.resize(m);
    for (int i = 0; i < m; ++i) out[i] = (double)h[i];
    return RBL_OK;
  }
  int dev_norm(const real* v, size_t n, double* out) {
    RET(gdots(v, n, 1, v, n));
    std::vector<double> s;
    RET(read_scalars(d_dots.as<real>(), 1, s));
    *out = std::sqrt(std::max(s[0], 0.0));
    return RBL_OK;
  }

  // ---- mixed precision -------------------------------------------------------------------------------
  int set_mixed(int mode) override {
    if (mode < 0 || mode > 2) return fail(RBL_ERR_INVALID, "mixed precision mode must be 0, 1 or 2");
    if (mode != 0 && sizeof(real) != 8) return fail(RBL_ERR_INVALID, "mixed precision applies to double contexts");
    mixed = mode;
    return RBL_OK;
  }
  bool use_mixed(int level) const { return sizeof(real) == 8 && mixed >= level; }
  // bring the float mirror to this context's parameters, flags and CURRENT configuration
  int sync_shadow() {
    if constexpr (std::is_same<real, double>::value) {
      RET(need_K());
      if (!shadow) {
        auto* sh = new Ctx<float>();
        sh->precision = RBL_F32; sh->device = device; sh->sm_count = sm_count;
        int st = sh->init();
        if (st != RBL_OK) { err = sh->err; delete sh; return st; }
        cudaStreamSynchronize(sh->stream);
        cudaStreamDestroy(sh->stream);
        sh->stream = stream;
        sh->own_stream = false;
        std::vector<float> rf(ref_host.begin(), ref_host.end());
        st = sh->set_parameters(a, dt, kBT, eta, rf.data(), n_blb);
        if (st != RBL_OK) { err = sh->err; delete sh; return st; }
        if (comm) {
          // partitioned suspension: the mirror shares this context's communicator (same ranks, same stream
          // order -- every rank builds its mirror at the same point of the same driver) and sets up its own,
          // float-sized, peer buffers; it follows this context's choice of exchange
          if (cudaSuccess != sh->d_status.ensure(4 * sizeof(int))) { cudaGetLastError(); delete sh; return fail(RBL_ERR_NOMEM, "mixed precision: out of device memory"); }
          sh->comm = comm->share();
          sh->comm->peer_setup(sizeof(float), sh->d_status.template as<int>(), stream);
        }
        shadow = sh;
        shadow_gen = 0;
      }
      if (comm && shadow->comm) shadow->comm->peer.use = shadow->comm->peer.on && comm->peer_active();
      shadow->stream = stream;  // (rbl_set_stream may have moved this context since the mirror was built)
      shadow->set_flags(block_pc ? 1 : 0, wall ? 1 : 0);
      shadow->noise_mode = 0;
      if (shadow_gen != cfg_gen || shadow->n_bod != n_bod) {
        shadow->n_bod = n_bod;
        CK(shadow->d_X.ensure(3 * (size_t)n_bod * sizeof(float)));
        CK(shadow->d_Q.ensure(4 * (size_t)n_bod * sizeof(float)));
        LAUNCH(1, (rbl::cast_scale<double, float>(d_X.template as<double>(), 1.0, shadow->d_X.template as<float>(), 3 * (size_t)n_bod, false, stream)));
        LAUNCH(1, (rbl::cast_scale<double, float>(d_Q.template as<double>(), 1.0, shadow->d_Q.template as<float>(), 4 * (size_t)n_bod, false, stream)));
        shadow->cfg_set = true;
        int st = shadow->set_K_mats();
        if (st != RBL_OK) return fail(st, shadow->err);
        shadow->pc_set = false;
        shadow_gen = cfg_gen;
      }
    }
    return RBL_OK;
  }
  // x = A^-1 b to ||b - A x|| <= tol ||b|| (DOUBLE residual): float GMRES corrections inside an
  // iterative refinement.  A cycle that fails to halve the residual hands over to the double solver.
  int dev_gmres_mixed(const real* b, real* xs, double tol, int restart, int max_iter, int* iters, double* relres) {
    if constexpr (std::is_same<real, double>::value) {
      RET(sync_shadow());
      const size_t n = sys_size();
      CK(d_c32a.ensure(n * sizeof(float)));
      CK(d_c32b.ensure(n * sizeof(float)));
      CK(d_mr.ensure(n * sizeof(real)));
      CK(d_me.ensure(n * sizeof(real)));
      CK(d_partial.ensure((size_t)4 * rbl::kDotBlocks * sizeof(real)));
      CK(d_dots.ensure((size_t)8 * sizeof(real)));
      real* r = d_mr.template as<real>();
      real* Ax = d_me.template as<real>();
      float* r32 = d_c32a.template as<float>();
      float* e32 = d_c32b.template as<float>();
      CK(cudaMemsetAsync(xs, 0, n * sizeof(real), stream));
      LAUNCH(1, rbl::scale_copy<real>(b, (real)1, r, n, false, stream));
      double bnorm = 0;
      RET(dev_norm(b, n, &bnorm));
      *iters = 0;
      *relres = 0;
      mixed_outer = 0;
      if (bnorm == 0) return RBL_OK;
      double res = bnorm;
      int total = 0;
      for (int outer = 0; outer < 12 && total < max_iter; ++outer) {
        ++mixed_outer;
        const double need = tol * bnorm / res;                        // reduction still missing
        const double inner_tol = std::min(1e-3, std::max(0.3 * need, 2e-6));
        LAUNCH(1, (rbl::cast_scale<double, float>(r, 1.0 / res, r32, n, false, stream)));
        int it = 0;
        double rr = 0;
        int st = shadow->dev_gmres(r32, e32, inner_tol, restart, max_iter - total, &it, &rr);
        if (st != RBL_OK) return fail(st, shadow->err);
        total += it;
        LAUNCH(1, (rbl::cast_scale<float, double>(e32, res, xs, n, true, stream)));
        RET(dev_saddle(xs, Ax));
        LAUNCH(1, rbl::scale_copy<real>(b, (real)1, r, n, false, stream));
        LAUNCH(1, rbl::scale_copy<real>(Ax, (real)-1, r, n, true, stream));
        double nres = 0;
        RET(dev_norm(r, n, &nres));
        const bool stalled = nres > 0.5 * res;
        res = nres;
        if (res / bnorm <= tol) break;
        if (stalled) {  // float corrections no longer help: finish in double on the residual equation
          int it2 = 0;
          double rr2 = 0;
          RET(dev_gmres(r, Ax, std::min(0.5, tol * bnorm / res), restart, std::max(1, max_iter - total), &it2, &rr2));
          total += it2;
          LAUNCH(1, rbl::scale_copy<real>(Ax, (real)1, xs, n, true, stream));
          res = rr2 * res;
          break;
        }
      }
      *iters = total;
      *relres = res / bnorm;
      return RBL_OK;
    } else {
      return dev_gmres(b, xs, tol, restart, max_iter, iters, relres);
    }
  }
  // the two products of a paired Lanczos iteration through the float mirror (mode 2)
  int prod_M2_mirror(const real* F1, const real* F2, real* out1, real* out2) {
    if constexpr (std::is_same<real, double>::value) {
      RET(sync_shadow());
      const size_t n3 = 3 * (size_t)N();
      for (DevBuf* bf : {&d_c32a, &d_c32b, &d_c32c, &d_c32d}) CK(bf->ensure(std::max(n3, sys_size()) * sizeof(float)));
      LAUNCH(1, (rbl::cast_scale<double, float>(F1, 1.0, d_c32a.template as<float>(), n3, false, stream)));
      LAUNCH(1, (rbl::cast_scale<double, float>(F2, 1.0, d_c32b.template as<float>(), n3, false, stream)));
      int st = shadow->prod_M2(d_c32a.template as<float>(), d_c32b.template as<float>(), d_c32c.template as<float>(),
                               d_c32d.template as<float>());
      if (st != RBL_OK) return fail(st, shadow->err);
      products += 2;
      LAUNCH(1, (rbl::cast_scale<float, double>(d_c32c.template as<float>(), 1.0, out1, n3, false, stream)));
      LAUNCH(1, (rbl::cast_scale<float, double>(d_c32d.template as<float>(), 1.0, out2, n3, false, stream)));
      return RBL_OK;
    } else {
      return prod_M2(F1, F2, out1, out2);
    }
  }
  int prod_M_mirror(const real* F, real* out) {
    if constexpr (std::is_same<real, double>::value) {
      RET(sync_shadow());
      const size_t n3 = 3 * (size_t)N();
      for (DevBuf* bf : {&d_c32a, &d_c32c}) CK(bf->ensure(std::max(n3, sys_size()) * sizeof(float)));
      LAUNCH(1, (rbl::cast_scale<double, float>(F, 1.0, d_c32a.template as<float>(), n3, false, stream)));
      int st = shadow->prod_M(d_c32a.template as<float>(), shadow->d_r.template as<float>(), true, d_c32c.template as<float>());
      if (st != RBL_OK) return fail(st, shadow->err);
      ++products;
      LAUNCH(1, (rbl::cast_scale<float, double>(d_c32c.template as<float>(), 1.0, out, n3, false, stream)));
      return RBL_OK;
    } else {
      return prod_M(F, d_r.template as<real>(), true, out);
    }
  }

  int gmres(const void* rhs, void* x, double tol, int restart, int max_iter, int* iters, double* relres) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    const size_t n = sys_size();
    CK(d_in1.ensure(n * sizeof(real)));   // rhs
    CK(d_out0.ensure(n * sizeof(real)));  // x
    RET(h2d(d_in1.p, rhs, n * sizeof(real)));
    if (use_mixed(1))
      RET(dev_gmres_mixed(d_in1.as<real>(), d_out0.as<real>(), tol, restart, max_iter, iters, relres));
    else
      RET(dev_gmres(d_in1.as<real>(), d_out0.as<real>(), tol, restart, max_iter, iters, relres));
    RET(d2h(x, d_out0.p, n * sizeof(real)));
    return csync();
  }

  // device-resident core: b and xs are device vectors of sys_size() reals (distinct from the
  // solver's own work buffers)
  int dev_gmres(const real* b, real* xs, double tol, int restart, int max_iter, int* iters, double* relres) {
    if (restart < 1 || max_iter < 1) return fail(RBL_ERR_INVALID, "gmres: restart and max_iter must be >= 1");
    RET(need_K());
    if (!pc_set || comm) {  // lazily built like the reference (:591-596); every rank must agree it exists
      int st = reserve_comm_workspace(false);
      if (st == RBL_OK && !pc_set) st = build_pc();
      st = agree(st);
      if (st != RBL_OK) return st;
    }
    const size_t n = sys_size(), n_head = 3 * (size_t)N();
    const int m = restart;
    CK(d_V.ensure((size_t)(m + 1) * n * sizeof(real)));
    CK(d_w.ensure(n * sizeof(real)));
    CK(d_z.ensure(n * sizeof(real)));
    CK(d_tmp.ensure(n * sizeof(real)));
    CK(d_partial.ensure((size_t)(m + 2) * rbl::kDotBlocks * sizeof(real)));
    CK(d_coef.ensure((size_t)(m + 2) * sizeof(real)));
    CK(d_dots.ensure((size_t)(3 * (m + 2)) * sizeof(real)));
    real* V = d_V.as<real>();
    real* w = d_w.as<real>();
    real* z = d_z.as<real>();
    real* tmp = d_tmp.as<real>();
    const int slot2 = m + 2, slotn = 2 * (m + 2);  // second Gram-Schmidt pass, squared norm
    CK(cudaMemsetAsync(xs, 0, n * sizeof(real), stream));
    double bnorm = 0;
    RET(dev_norm(b, n, &bnorm));
    int total = 0;
    double res = bnorm;
    if (bnorm == 0) {
      *iters = 0;
      *relres = 0;
      return RBL_OK;
    }
    std::vector<double> H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), hcol;
    std::vector<real> coef(m + 1);
    bool first = true;
    while (total < max_iter) {
      // r0 = b - A x  (x = 0 on the first cycle)
      if (first) {
        LAUNCH(1, rbl::scale_copy<real>(b, (real)1, w, n, false, stream));
      } else {
        RET(dev_saddle(xs, w));
        LAUNCH(1, rbl::scale_copy<real>(w, (real)-1, w, n, false, stream));
        LAUNCH(1, rbl::scale_copy<real>(b, (real)1, w, n, true, stream));
      }
      first = false;
      double beta = 0;
      RET(dev_norm(w, n, &beta));
      res = beta;
      if (beta / bnorm <= tol) break;
      LAUNCH(1, rbl::scale_copy<real>(w, (real)(1.0 / beta), V, n, false, stream));
      std::fill(g.begin(), g.end(), 0.0);
      g[0] = beta;
      int j = 0;
      for (; j < m && total < max_iter; ++j, ++total) {
        // z = P S v_j ; w = A z
        LAUNCH(1, rbl::flip_tail<real>(V + (size_t)j * n, n_head, n, tmp, stream));
        RET(dev_pc(tmp, z));
        RET(dev_saddle(z, w));
        // classical Gram-Schmidt, twice; coefficients, the norm and the normalisation of v_{j+1} stay
        // on the device: ONE device->host read per iteration (for the Givens update and the stopping test)
        std::fill(H.begin() + (size_t)j * (m + 1), H.begin() + (size_t)(j + 1) * (m + 1), 0.0);
        RET(gdots(V, n, j + 1, w, n, 0));
        LAUNCH(1, rbl::multi_axpy<real>(V, n, j + 1, d_dots.as<real>(), (real)-1, w, n, stream));
        RET(gdots(V, n, j + 1, w, n, slot2));
        LAUNCH(1, rbl::multi_axpy<real>(V, n, j + 1, d_dots.as<real>() + slot2, (real)-1, w, n, stream));
        RET(gdots(w, n, 1, w, n, slotn));
        LAUNCH(1, rbl::scale_by_inv_sqrt<real>(w, d_dots.as<real>() + slotn, V + (size_t)(j + 1) * n, n, stream));
        RET(read_scalars(d_dots.as<real>(), slotn + 1, hcol));
        for (int i = 0; i <= j; ++i) H[(size_t)j * (m + 1) + i] = hcol[i] + hcol[slot2 + i];
        const double hn = std::sqrt(std::max(hcol[slotn], 0.0));
        H[(size_t)j * (m + 1) + j + 1] = hn;
        // Givens
        double* h = &H[(size_t)j * (m + 1)];
        for (int i = 0; i < j; ++i) {
          const double t = cs[i] * h[i] + sn[i] * h[i + 1];
          h[i + 1] = -sn[i] * h[i] + cs[i] * h[i + 1];
          h[i] = t;
        }
        const double den = std::hypot(h[j], h[j + 1]);
        cs[j] = den > 0 ? h[j] / den : 1.0;
        sn[j] = den > 0 ? h[j + 1] / den : 0.0;
        h[j] = den;
        h[j + 1] = 0;
        g[j + 1] = -sn[j] * g[j];
        g[j] = cs[j] * g[j];
        res = std::fabs(g[j + 1]);
        if (res / bnorm <= tol || hn == 0) {
          ++j;
          ++total;
          break;
        }
      }
      // y = H^-1 g ; x += P S (V y)
      std::vector<double> y(j);
      for (int i = j - 1; i >= 0; --i) {
        double s = g[i];
        for (int k = i + 1; k < j; ++k) s -= H[(size_t)k * (m + 1) + i] * y[k];
        y[i] = s / H[(size_t)i * (m + 1) + i];
      }
      for (int i = 0; i < j; ++i) coef[i] = (real)y[i];
      RET(h2d(d_coef.p, coef.data(), j * sizeof(real)));
      CK(cudaMemsetAsync(w, 0, n * sizeof(real), stream));
      LAUNCH(1, rbl::multi_axpy<real>(V, n, j, d_coef.as<real>(), (real)1, w, n, stream));
      LAUNCH(1, rbl::flip_tail<real>(w, n_head, n, tmp, stream));
      RET(dev_pc(tmp, z));
      LAUNCH(1, rbl::scale_copy<real>(z, (real)1, xs, n, true, stream));
      CK(cudaStreamSynchronize(stream));  // coef (host) must outlive the copy
      if (res / bnorm <= tol) break;
    }
    *iters = total;
    *relres = res / bnorm;
    return RBL_OK;
  }

  // symmetric tridiagonal eigen-decomposition by cyclic Jacobi on the dense k x k matrix
  static void jacobi_eig(std::vector<double>& A, int k, std::vector<double>& evec) {
    evec.assign((size_t)k * k, 0.0);
    for (int i = 0; i < k; ++i) evec[(size_t)i * k + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
      double off = 0;
      for (int p = 0; p < k; ++p)
        for (int q = p + 1; q < k; ++q) off += A[(size_t)p * k + q] * A[(size_t)p * k + q];
      if (off < 1e-30) break;
      for (int p = 0; p < k; ++p)
        for (int q = p + 1; q < k; ++q) {
          const double apq = A[(size_t)p * k + q];
          if (std::fabs(apq) < 1e-300) continue;
          const double th = (A[(size_t)q * k + q] - A[(size_t)p * k + p]) / (2 * apq);
          const double t = (th >= 0 ? 1.0 : -1.0) / (std::fabs(th) + std::sqrt(th * th + 1));
          const double c = 1 / std::sqrt(t * t + 1), s = t * c;
          for (int i = 0; i < k; ++i) {
            const double aip = A[(size_t)i * k + p], aiq = A[(size_t)i * k + q];
            A[(size_t)i * k + p] = c * aip - s * aiq;
            A[(size_t)i * k + q] = s * aip + c * aiq;
          }
          for (int i = 0; i < k; ++i) {
            const double api = A[(size_t)p * k + i], aqi = A[(size_t)q * k + i];
            A[(size_t)p * k + i] = c * api - s * aqi;
            A[(size_t)q * k + i] = s * api + c * aqi;
          }
          for (int i = 0; i < k; ++i) {
            const double vip = evec[(size_t)i * k + p], viq = evec[(size_t)i * k + q];
            evec[(size_t)i * k + p] = c * vip - s * viq;
            evec[(size_t)i * k + q] = s * vip + c * viq;
          }
        }
    }
  }

  int lanczos(const void* W, void* out, double tol, int max_iter, int* iters) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    const size_t n = 3 * (size_t)N();
    CK(d_in1.ensure(n * sizeof(real)));
    CK(d_out0.ensure(n * sizeof(real)));
    RET(h2d(d_in1.p, W, n * sizeof(real)));
    RET(dev_lanczos(d_in1.as<real>(), d_out0.as<real>(), tol, max_iter, iters, noise_mode == 2));
    RET(d2h(out, d_out0.p, n * sizeof(real)));
    return csync();
  }

  // device-resident core: dout = (B M B)^{1/2} dW at the CURRENT configuration (d_r)
  int dev_lanczos(const real* dW, real* dout, double tol, int max_iter, int* iters, bool precondition = false) {
    if (max_iter < 1) return fail(RBL_ERR_INVALID, "lanczos: max_iter must be >= 1");
    RET(need_K());
    bool pc = false;
    RET(agree_workspace(false));
    RET(want_noise_pc(precondition, &pc));
    const size_t n = 3 * (size_t)N();
    const int nb = (int)N();
    const int m = max_iter;
    CK(d_V.ensure((size_t)(m + 1) * n * sizeof(real)));
    CK(d_w.ensure(n * sizeof(real)));
    CK(d_partial.ensure((size_t)(m + 2) * rbl::kDotBlocks * sizeof(real)));
    CK(d_coef.ensure((size_t)(m + 2) * sizeof(real)));
    CK(d_dots.ensure((size_t)(3 * (m + 4)) * sizeof(real)));
    real* V = d_V.as<real>();
    real* w = d_w.as<real>();
    LAUNCH(1, rbl::scale_copy<real>(dW, (real)1, w, n, false, stream));
    double wnorm = 0;
    RET(dev_norm(w, n, &wnorm));
    if (wnorm == 0) {
      CK(cudaMemsetAsync(dout, 0, n * sizeof(real), stream));
      *iters = 0;
      return RBL_OK;
    }
    LAUNCH(1, rbl::scale_copy<real>(w, (real)(1.0 / wnorm), V, n, false, stream));
    std::vector<double> alpha, beta, y_prev, y, s;
    int k = 0;
    for (; k < m;) {
      // w = M v_k - beta_{k-1} v_{k-1}
      if (pc) {  // w = G A G^T v_k
        RET(noise_GT(V + (size_t)k * n, d_nt1.as<real>()));
        if (use_mixed(2)) RET(prod_M_mirror(d_nt1.as<real>(), d_nu1.as<real>()));
        else RET(prod_M(d_nt1.as<real>(), d_r.as<real>(), true, d_nu1.as<real>()));
        RET(noise_G(d_nu1.as<real>(), w));
      } else {
        if (use_mixed(2)) RET(prod_M_mirror(V + (size_t)k * n, w));
        else RET(prod_M(V + (size_t)k * n, d_r.as<real>(), true, w));
      }
      if (k > 0) LAUNCH(1, rbl::scale_copy<real>(V + (size_t)(k - 1) * n, (real)(-beta[k - 1]), w, n, true, stream));
      // alpha_k, the full reorthogonalisation (keeps the basis orthonormal in fp32 too), beta_k^2 and the
      // normalised v_{k+1} are all computed on the device: ONE device->host read per iteration
      RET(gdots(V + (size_t)k * n, n, 1, w, n, 0));
      LAUNCH(1, rbl::multi_axpy<real>(V + (size_t)k * n, n, 1, d_dots.as<real>(), (real)-1, w, n, stream));
      RET(gdots(V, n, k + 1, w, n, 2));
      LAUNCH(1, rbl::multi_axpy<real>(V, n, k + 1, d_dots.as<real>() + 2, (real)-1, w, n, stream));
      RET(gdots(w, n, 1, w, n, 1));
      LAUNCH(1, rbl::scale_by_inv_sqrt<real>(w, d_dots.as<real>() + 1, V + (size_t)(k + 1) * n, n, stream));
      RET(read_scalars(d_dots.as<real>(), 2, s));
      alpha.push_back(s[0]);
      const double bn = std::sqrt(std::max(s[1], 0.0));
      ++k;
      // y = ||W|| T_k^{1/2} e_1
      std::vector<double> T((size_t)k * k, 0.0), Z;
      for (int i = 0; i < k; ++i) {
        T[(size_t)i * k + i] = alpha[i];
        if (i + 1 < k) T[(size_t)i * k + i + 1] = T[(size_t)(i + 1) * k + i] = beta[i];
      }
      jacobi_eig(T, k, Z);
      y.assign(k, 0.0);
      for (int e = 0; e < k; ++e) {
        const double lam = std::max(T[(size_t)e * k + e], 0.0);
        const double f = std::sqrt(lam) * Z[(size_t)0 * k + e] * wnorm;
        for (int i = 0; i < k; ++i) y[i] += Z[(size_t)i * k + e] * f;
      }
      double diff = 0, nrm = 0;
      for (int i = 0; i < k; ++i) {
        const double d = y[i] - (i < (int)y_prev.size() ? y_prev[i] : 0.0);
        diff += d * d;
        nrm += y[i] * y[i];
      }
      y_prev = y;
      const bool converged = k > 1 && std::sqrt(diff) <= tol * std::sqrt(nrm);
      if (converged || bn <= 1e-14 * wnorm || k == m) break;
      beta.push_back(bn);  // v_k = w / beta is already in place
    }
    std::vector<real> coef(k);
    for (int i = 0; i < k; ++i) coef[i] = (real)y[i];
    RET(h2d(d_coef.p, coef.data(), k * sizeof(real)));
    real* acc = pc ? d_nt1.as<real>() : dout;
    CK(cudaMemsetAsync(acc, 0, n * sizeof(real), stream));
    LAUNCH(1, rbl::multi_axpy<real>(V, n, k, d_coef.as<real>(), (real)1, acc, n, stream));
    if (pc) RET(noise_L(acc, dout));
    CK(cudaStreamSynchronize(stream));  // coef (host) must outlive the copy
    *iters = k;
    return RBL_OK;
  }




  // ---- noise preconditioner ---------------------------------------------------------------------
  // L_b = chol(Mt_b), G_b = L_b^-1 for the bodies of this context.  Free space: ONE factor of the
  // reference shape (M_b = R M_ref R^T  =>  L_b = R L_ref), computed once per parameter set.  With the
  // wall: one factor per body and configuration.  noise_ok = false (L = I for this context's bodies)
  // when a block is not positive definite (blobs inside the wall-overlap layer) or the factors do not fit.
  int build_noise_pc() {
    RET(need_K());
    // a block PC built on these factors (pc_chol) goes with them: it is rebuilt at its next use
    if (pc_set && pc_chol && wall) pc_set = false;  // (the shared free-space factor never changes)
    const int sz = 3 * n_blb;
    int* fl = d_flags.as<int>();
    int bad = 0;
    if (!wall) {
      noise_shared = true;
      if (!noise_shared_ready) {
        const size_t bytes = (size_t)sz * sz * sizeof(real);
        if (d_NL.ensure(bytes) != cudaSuccess || d_NG.ensure(bytes) != cudaSuccess) {
          cudaGetLastError();
          bad = 1;
        } else {
          LAUNCH(1, rbl::pc_block_assemble<real>(d_ref.as<real>(), 1, n_blb, (real)a, (real)eta, false, d_NL.as<real>(), fl + FLAG_NOISE, stream));
          LAUNCH(1, rbl::chol_lower<real>(d_NL.as<real>(), 1, sz, fl + FLAG_NOISE, stream));
          LAUNCH(1, rbl::tri_inverse<real>(d_NL.as<real>(), d_NG.as<real>(), 1, sz, stream));
        }
      }
    } else {
      noise_shared = false;
      const size_t bytes = (size_t)n_bod * sz * sz * sizeof(real);
      if (d_NL.ensure(bytes) != cudaSuccess || d_NG.ensure(bytes) != cudaSuccess) {
        cudaGetLastError();
        d_NL.release();
        d_NG.release();
        bad = 1;
      } else {
        LAUNCH(1, rbl::pc_block_assemble<real>(d_r.as<real>(), n_bod, n_blb, (real)a, (real)eta, true, d_NL.as<real>(), fl + FLAG_NOISE, stream));
        LAUNCH(1, rbl::chol_lower<real>(d_NL.as<real>(), n_bod, sz, fl + FLAG_NOISE, stream));
        LAUNCH(1, rbl::tri_inverse<real>(d_NL.as<real>(), d_NG.as<real>(), n_bod, sz, stream));
      }
    }
    if (!bad) {
      int f = 0;
      CK(cudaMemcpyAsync(&f, fl + FLAG_NOISE, sizeof(int), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      if (f) {
        CK(cudaMemsetAsync(fl + FLAG_NOISE, 0, sizeof(int), stream));
        bad = 1;
      }
    }
    // rank-local on purpose (no collective here): a rank that falls back simply uses L = I for ITS
    // bodies, which is still a valid block-diagonal preconditioner of the global operator
    noise_ok = !bad;
    if (noise_ok && !wall) noise_shared_ready = true;
    noise_set = true;
    const size_t n3 = 3 * (size_t)N();
    if (noise_ok)
      for (DevBuf* b : {&d_nt1, &d_nt2, &d_nu1, &d_nu2}) CK(b->ensure(n3 * sizeof(real)));
    return RBL_OK;
  }
  // out = G^T v, G u, L y for every body of the context
  int noise_GT(const real* v, real* out) {
    LAUNCH(1, rbl::body_mat_mul<real>(d_NG.as<real>(), noise_shared ? 0 : (size_t)9 * n_blb * n_blb, d_Q.as<real>(), false,
                                      noise_shared, true, v, n_bod, n_blb, out, stream));
    return RBL_OK;
  }
  int noise_G(const real* u, real* out) {
    LAUNCH(1, rbl::body_mat_mul<real>(d_NG.as<real>(), noise_shared ? 0 : (size_t)9 * n_blb * n_blb, d_Q.as<real>(), noise_shared,
                                      false, false, u, n_bod, n_blb, out, stream));
    return RBL_OK;
  }
  int noise_L(const real* y, real* out) {
    LAUNCH(1, rbl::body_mat_mul<real>(d_NL.as<real>(), noise_shared ? 0 : (size_t)9 * n_blb * n_blb, d_Q.as<real>(), false,
                                      noise_shared, false, y, n_bod, n_blb, out, stream));
    return RBL_OK;
  }
  // decides whether this Lanczos run is preconditioned (builds the factors if needed)
  int want_noise_pc(bool requested, bool* use) {
    *use = false;
    if (!requested) return RBL_OK;
    if (!noise_set) RET(build_noise_pc());
    *use = noise_ok;
    return RBL_OK;
  }


  // Self-check of the noise factors in the spirit of the reference's unbound test_PC / Test_Mhalf
  // (:569-587, :895-915): for a fixed pseudo-random x,  |L L^T x - Mt x| / |Mt x|  with Mt the
  // freshly assembled body blocks, and  |G L x - x| / |x|.  active = 0 when the context fell back
  // to the unpreconditioned recurrence (blocks not positive definite / factors do not fit).
  int noise_selfcheck(double* factor_err, double* inverse_err, int* active) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    RET(build_noise_pc());
    *active = noise_ok ? 1 : 0;
    *factor_err = *inverse_err = 0;
    if (!noise_ok) return RBL_OK;
    const int sz = 3 * n_blb;
    const size_t n3 = 3 * (size_t)N();
    const size_t stride = noise_shared ? 0 : (size_t)sz * sz;
    std::vector<real> hx(n3);
    unsigned long long st = 88172645463325252ull;
    for (auto& v : hx) {  // xorshift: deterministic, no library
      st ^= st << 13; st ^= st >> 7; st ^= st << 17;
      v = (real)((double)(st >> 11) / 9007199254740992.0 - 0.5);
    }
    for (DevBuf* b : {&d_nt1, &d_nt2, &d_nu1, &d_nu2}) CK(b->ensure(n3 * sizeof(real)));
    real* x = d_nt1.as<real>();
    real* t = d_nt2.as<real>();
    real* y = d_nu1.as<real>();
    real* z = d_nu2.as<real>();
    RET(h2d(x, hx.data(), n3 * sizeof(real)));
    // y = L (L^T x)
    LAUNCH(1, rbl::body_mat_mul<real>(d_NL.as<real>(), stride, d_Q.as<real>(), noise_shared, false, true, x, n_bod, n_blb, t, stream));
    LAUNCH(1, rbl::body_mat_mul<real>(d_NL.as<real>(), stride, d_Q.as<real>(), false, noise_shared, false, t, n_bod, n_blb, y, stream));
    // z = G (L x) - x  -> kept in t after the subtraction below
    LAUNCH(1, rbl::body_mat_mul<real>(d_NL.as<real>(), stride, d_Q.as<real>(), false, noise_shared, false, x, n_bod, n_blb, t, stream));
    RET(noise_G(t, z));
    LAUNCH(1, rbl::scale_copy<real>(x, (real)-1, z, n3, true, stream));
    CK(d_partial.ensure(4 * rbl::kDotBlocks * sizeof(real)));
    CK(d_dots.ensure(4 * sizeof(real)));
    double zn = 0, xn = 0;
    RET(dev_norm(z, n3, &zn));
    RET(dev_norm(x, n3, &xn));
    *inverse_err = zn / xn;
    // Mt x with freshly assembled blocks (scratch: the G buffer, rebuilt afterwards)
    int* fl = d_flags.as<int>();
    if (noise_shared) {
      LAUNCH(1, rbl::pc_block_assemble<real>(d_ref.as<real>(), 1, n_blb, (real)a, (real)eta, false, d_NG.as<real>(), fl + FLAG_NOISE, stream));
      // M_b = R M_ref R^T: rotate in, multiply (symmetric), rotate out
      LAUNCH(1, rbl::body_mat_mul<real>(d_NG.as<real>(), 0, d_Q.as<real>(), true, true, false, x, n_bod, n_blb, t, stream));
    } else {
      LAUNCH(1, rbl::pc_block_assemble<real>(d_r.as<real>(), n_bod, n_blb, (real)a, (real)eta, true, d_NG.as<real>(), fl + FLAG_NOISE, stream));
      LAUNCH(1, rbl::body_mat_mul<real>(d_NG.as<real>(), stride, d_Q.as<real>(), false, false, false, x, n_bod, n_blb, t, stream));
    }
    LAUNCH(1, rbl::scale_copy<real>(t, (real)-1, y, n3, true, stream));
    double yn = 0, tn = 0;
    RET(dev_norm(y, n3, &yn));
    RET(dev_norm(t, n3, &tn));
    *factor_err = yn / tn;
    CK(cudaMemsetAsync(fl + FLAG_NOISE, 0, sizeof(int), stream));
    LAUNCH(1, rbl::tri_inverse<real>(d_NL.as<real>(), d_NG.as<real>(), noise_shared ? 1 : n_bod, sz, stream));  // restore G
    return sync();
  }

  // y = ||W|| T_k^{1/2} e_1 for the Lanczos tridiagonal (alpha, beta)
  static void lanczos_coeffs(const std::vector<double>& alpha, const std::vector<double>& beta, int k, double wnorm,
                             std::vector<double>& y) {
    std::vector<double> T((size_t)k * k, 0.0), Z;
    for (int i = 0; i < k; ++i) {
      T[(size_t)i * k + i] = alpha[i];
      if (i + 1 < k) T[(size_t)i * k + i + 1] = T[(size_t)(i + 1) * k + i] = beta[i];
    }
    jacobi_eig(T, k, Z);
    y.assign(k, 0.0);
    for (int e = 0; e < k; ++e) {
      const double lam = std::max(T[(size_t)e * k + e], 0.0);
      const double f = std::sqrt(lam) * Z[(size_t)0 * k + e] * wnorm;
      for (int i = 0; i < k; ++i) y[i] += Z[(size_t)i * k + e] * f;
    }
  }

  int lanczos2(const void* W1, const void* W2, void* out1, void* out2, double tol, int max_iter, int* iters2) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    const size_t n = 3 * (size_t)N();
    CK(d_in1.ensure(n * sizeof(real)));
    CK(d_in2.ensure(n * sizeof(real)));
    CK(d_out0.ensure(n * sizeof(real)));
    CK(d_out2.ensure(n * sizeof(real)));
    RET(h2d(d_in1.p, W1, n * sizeof(real)));
    RET(h2d(d_in2.p, W2, n * sizeof(real)));
    RET(dev_lanczos2(d_in1.as<real>(), d_in2.as<real>(), d_out0.as<real>(), d_out2.as<real>(), tol, max_iter, iters2,
                     noise_mode == 2));
    RET(d2h(out1, d_out0.p, n * sizeof(real)));
    RET(d2h(out2, d_out2.p, n * sizeof(real)));
    return csync();
  }

  // Two Lanczos recurrences in lockstep, (B M B)^{1/2} W_1 and (B M B)^{1/2} W_2, sharing every
  // mobility product through the two-right-hand-side kernel.  Each recurrence is exactly
  // dev_lanczos (same stopping rule); one that has converged is frozen while the other finishes.
  int dev_lanczos2(const real* dW1, const real* dW2, real* dout1, real* dout2, double tol, int max_iter, int* iters2,
                   bool precondition) {
    if (max_iter < 1) return fail(RBL_ERR_INVALID, "lanczos: max_iter must be >= 1");
    RET(need_K());
    bool pc = false;
    RET(agree_workspace(true));
    RET(want_noise_pc(precondition, &pc));
    const size_t n = 3 * (size_t)N();
    const int m = max_iter;
    CK(d_V.ensure((size_t)(m + 1) * n * sizeof(real)));
    CK(d_V2.ensure((size_t)(m + 1) * n * sizeof(real)));
    CK(d_w.ensure(n * sizeof(real)));
    CK(d_w2.ensure(n * sizeof(real)));
    CK(d_partial.ensure((size_t)(m + 2) * rbl::kDotBlocks * sizeof(real)));
    CK(d_coef.ensure((size_t)(m + 2) * sizeof(real)));
    CK(d_dots.ensure((size_t)(3 * (m + 4)) * sizeof(real)));
    const int slot_stride = m + 4;  // per recurrence: [alpha, beta^2, reorthogonalisation coefficients ...]
    struct Rec {
      real* V; real* w; const real* W; real* out;
      std::vector<double> alpha, beta, y_prev, y;
      double wnorm = 0; int k = 0; bool done = false;
    } R[2];
    R[0].V = d_V.as<real>();  R[0].w = d_w.as<real>();  R[0].W = dW1; R[0].out = dout1;
    R[1].V = d_V2.as<real>(); R[1].w = d_w2.as<real>(); R[1].W = dW2; R[1].out = dout2;
    std::vector<double> s;
    for (auto& q : R) {
      LAUNCH(1, rbl::scale_copy<real>(q.W, (real)1, q.w, n, false, stream));
      RET(dev_norm(q.w, n, &q.wnorm));
      if (q.wnorm == 0) {
        CK(cudaMemsetAsync(q.out, 0, n * sizeof(real), stream));
        CK(cudaMemsetAsync(q.V, 0, n * sizeof(real), stream));  // a harmless input for the shared products
        q.done = true;
      } else {
        LAUNCH(1, rbl::scale_copy<real>(q.w, (real)(1.0 / q.wnorm), q.V, n, false, stream));
      }
    }
    for (int step = 0; step < m && !(R[0].done && R[1].done); ++step) {
      // a frozen recurrence feeds its last basis vector (the result is ignored)
      const real* v0 = R[0].V + (size_t)(R[0].done ? std::max(R[0].k - 1, 0) : R[0].k) * n;
      const real* v1 = R[1].V + (size_t)(R[1].done ? std::max(R[1].k - 1, 0) : R[1].k) * n;
      if (pc) {  // w = G A G^T v
        RET(noise_GT(v0, d_nt1.as<real>()));
        RET(noise_GT(v1, d_nt2.as<real>()));
        if (use_mixed(2)) RET(prod_M2_mirror(d_nt1.as<real>(), d_nt2.as<real>(), d_nu1.as<real>(), d_nu2.as<real>()));
        else RET(prod_M2(d_nt1.as<real>(), d_nt2.as<real>(), d_nu1.as<real>(), d_nu2.as<real>()));
        RET(noise_G(d_nu1.as<real>(), R[0].w));
        RET(noise_G(d_nu2.as<real>(), R[1].w));
      } else {
        if (use_mixed(2)) RET(prod_M2_mirror(v0, v1, R[0].w, R[1].w));
        else RET(prod_M2(v0, v1, R[0].w, R[1].w));
      }
      // device side of both recurrences first (alpha, reorthogonalisation, beta^2, normalised next
      // basis vector), then ONE device->host read for the two of them
      for (int r = 0; r < 2; ++r) {
        auto& q = R[r];
        if (q.done) continue;
        const int k = q.k, base = r * slot_stride;
        real* V = q.V;
        real* w = q.w;
        real* dd = d_dots.as<real>() + base;
        if (k > 0) LAUNCH(1, rbl::scale_copy<real>(V + (size_t)(k - 1) * n, (real)(-q.beta[k - 1]), w, n, true, stream));
        RET(gdots(V + (size_t)k * n, n, 1, w, n, base));
        LAUNCH(1, rbl::multi_axpy<real>(V + (size_t)k * n, n, 1, dd, (real)-1, w, n, stream));
        RET(gdots(V, n, k + 1, w, n, base + 2));
        LAUNCH(1, rbl::multi_axpy<real>(V, n, k + 1, dd + 2, (real)-1, w, n, stream));
        RET(gdots(w, n, 1, w, n, base + 1));
        LAUNCH(1, rbl::scale_by_inv_sqrt<real>(w, dd + 1, V + (size_t)(k + 1) * n, n, stream));
      }
      RET(read_scalars(d_dots.as<real>(), slot_stride + 2, s));
      for (int r = 0; r < 2; ++r) {
        auto& q = R[r];
        if (q.done) continue;
        const int k = q.k;
        q.alpha.push_back(s[r * slot_stride]);
        const double bn = std::sqrt(std::max(s[r * slot_stride + 1], 0.0));
        q.k = k + 1;
        lanczos_coeffs(q.alpha, q.beta, q.k, q.wnorm, q.y);
        double diff = 0, nrm = 0;
        for (int i = 0; i < q.k; ++i) {
          const double d = q.y[i] - (i < (int)q.y_prev.size() ? q.y_prev[i] : 0.0);
          diff += d * d;
          nrm += q.y[i] * q.y[i];
        }
        q.y_prev = q.y;
        const bool converged = q.k > 1 && std::sqrt(diff) <= tol * std::sqrt(nrm);
        if (converged || bn <= 1e-14 * q.wnorm || q.k == m) {
          q.done = true;
          continue;
        }
        q.beta.push_back(bn);  // v_k = w / beta is already in place
      }
    }
    for (int r = 0; r < 2; ++r) {
      auto& q = R[r];
      iters2[r] = q.k;
      if (q.wnorm == 0) continue;
      std::vector<real> coef(q.k);
      for (int i = 0; i < q.k; ++i) coef[i] = (real)q.y[i];
      RET(h2d(d_coef.p, coef.data(), q.k * sizeof(real)));
      real* acc = pc ? d_nt1.as<real>() : q.out;
      CK(cudaMemsetAsync(acc, 0, n * sizeof(real), stream));
      LAUNCH(1, rbl::multi_axpy<real>(q.V, n, q.k, d_coef.as<real>(), (real)1, acc, n, stream));
      if (pc) RET(noise_L(acc, q.out));  // g = L (G A G^T)^{1/2} W
      CK(cudaStreamSynchronize(stream));  // coef (host) must outlive the copy
    }
    return RBL_OK;
  }

  // ---- Brownian-dynamics step (see rbl_bd_step in include/rbl.h) -------------------------------
  int kinv_dev(const real* v3n, real* out6) {  // out = (K^T K)^-1 K^T v   (:390,406)
    LAUNCH(1, rbl::kt_dot<real>(v3n, d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, out6, stream));
    LAUNCH(1, rbl::ktk_inv_apply<real>(d_S.as<real>(), n_bod, n_blb, out6, stream));
    return RBL_OK;
  }
  // ---- random finite differences (c_rigid_obj.cpp:743-863), deterministic given the noise -------
  // delta: the reference hard-wires 1e-4 (M_RFD, KTinv_RFD) and 1e-3 (the *_from_U forms) in whatever
  // precision it was compiled for.  In float a centred difference quotient of M(q) W is best near
  // eps^(1/3) * (length scale of M's variation ~ body radius 1) ~ 4e-3: rounding 3e-7 |M W| / delta
  // against truncation (delta / a)^2 / 6; that is the float default.  rbl_set_rfd_delta overrides both.
  double default_delta() const { return rfd_delta > 0 ? rfd_delta : (sizeof(real) == 8 ? 1.0e-4 : 4.0e-3); }
  int rfd_buffers() {
    const size_t n3 = 3 * (size_t)N(), n6 = 6 * (size_t)n_bod;
    for (DevBuf* b : {&d_rp, &d_t1, &d_t2, &d_rfd, &d_wr}) CK(b->ensure(n3 * sizeof(real)));
    CK(d_uom.ensure(n6 * sizeof(real)));
    CK(d_Xp.ensure(3 * (size_t)n_bod * sizeof(real)));
    CK(d_Qp.ensure(4 * (size_t)n_bod * sizeof(real)));
    return RBL_OK;
  }
  // (X', Q') = configuration displaced by scale * U (update_X_Q, :691-710) and its blob positions
  int displace(const real* U6, double scale, real* rp) {
    LAUNCH(1, rbl::integrate<real>(U6, (real)scale, n_bod, d_X.as<real>(), d_Q.as<real>(), d_Xp.as<real>(),
                                   d_Qp.as<real>(), stream));
    if (rp)
      LAUNCH(1, rbl::place_blobs<real>(d_Xp.as<real>(), d_Qp.as<real>(), d_ref.as<real>(), n_bod, n_blb, rp, stream));
    return RBL_OK;
  }
  // out = (M(q+) - M(q-)) W / delta with q+- = q +- (delta/2) U   (M_RFD_from_U :818-840; M_RFD :769-796 with
  // U = K^-1 W).  All pointers on the device; collective in partitioned mode.
  int dev_M_RFD(const real* U6, const real* W, double delta, real* out) {
    const size_t n3 = 3 * (size_t)N();
    real* Mpm[2] = {d_t1.as<real>(), d_t2.as<real>()};
    for (int sgn = 0; sgn < 2; ++sgn) {
      RET(displace(U6, (sgn ? -0.5 : 0.5) * delta, d_rp.as<real>()));
      RET(prod_M(W, d_rp.as<real>(), false, Mpm[sgn]));
    }
    LAUNCH(1, rbl::scale_copy<real>(d_t1.as<real>(), (real)(1.0 / delta), out, n3, false, stream));
    LAUNCH(1, rbl::scale_copy<real>(d_t2.as<real>(), (real)(-1.0 / delta), out, n3, true, stream));
    return RBL_OK;
  }
  int M_RFD(const void* U6, const void* W, double delta, void* out) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    if (!W || !out) return fail(RBL_ERR_INVALID, "M_RFD: W and out are required");
    RET(need_K());
    RET(rfd_buffers());
    const size_t n3 = 3 * (size_t)N(), n6 = 6 * (size_t)n_bod;
    if (!(delta > 0)) delta = U6 ? (rfd_delta > 0 ? rfd_delta : 1.0e-3) : default_delta();  // :771 / :820
    RET(h2d(d_wr.p, W, n3 * sizeof(real)));
    if (U6) RET(h2d(d_uom.p, U6, n6 * sizeof(real)));
    else RET(kinv_dev(d_wr.as<real>(), d_uom.as<real>()));  // UOM = Kinv * W  (:776)
    RET(dev_M_RFD(d_uom.as<real>(), d_wr.as<real>(), delta, d_rfd.as<real>()));
    RET(d2h(out, d_rfd.p, n3 * sizeof(real)));
    return csync();
  }
  // out = (K(q+)^T - K(q-)^T) W / delta   (KT_RFD_from_U, :842-863)
  int KT_RFD(const void* U6, const void* W, double delta, void* out6) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    if (!U6 || !W || !out6) return fail(RBL_ERR_INVALID, "KT_RFD_from_U: U, W and out are required");
    RET(need_K());
    RET(rfd_buffers());
    const size_t n3 = 3 * (size_t)N(), n6 = 6 * (size_t)n_bod;
    if (!(delta > 0)) delta = rfd_delta > 0 ? rfd_delta : 1.0e-3;  // :844
    CK(d_in0.ensure(2 * n6 * sizeof(real)));
    RET(h2d(d_wr.p, W, n3 * sizeof(real)));
    RET(h2d(d_uom.p, U6, n6 * sizeof(real)));
    real* pm = d_in0.as<real>();
    for (int sgn = 0; sgn < 2; ++sgn) {
      RET(displace(d_uom.as<real>(), (sgn ? -0.5 : 0.5) * delta, d_rp.as<real>()));
      LAUNCH(1, rbl::kt_dot<real>(d_wr.as<real>(), d_rp.as<real>(), d_Xp.as<real>(), n_bod, n_blb, pm + sgn * n6, stream));
    }
    LAUNCH(1, rbl::scale_copy<real>(pm, (real)(1.0 / delta), pm, n6, false, stream));
    LAUNCH(1, rbl::scale_copy<real>(pm + n6, (real)(-1.0 / delta), pm, n6, true, stream));
    RET(d2h(out6, pm, n6 * sizeof(real)));
    return sync();
  }
  // out = K^T (Kinv(q+)^T - Kinv(q-)^T) W / delta with q+- = q +- (delta/2) W   (KTinv_RFD, :743-767)
  int KTinv_RFD(const void* W6, double delta, void* out6) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    if (!W6 || !out6) return fail(RBL_ERR_INVALID, "KTinv_RFD: W and out are required");
    RET(need_K());
    RET(rfd_buffers());
    const size_t n3 = 3 * (size_t)N(), n6 = 6 * (size_t)n_bod;
    if (!(delta > 0)) delta = default_delta();  // :745
    CK(d_in0.ensure(2 * n6 * sizeof(real)));
    CK(d_in1.ensure(9 * (size_t)n_bod * sizeof(real)));
    RET(h2d(d_uom.p, W6, n6 * sizeof(real)));
    real* tmp6 = d_in0.as<real>();
    real* Sp = d_in1.as<real>();
    real* v[2] = {d_t1.as<real>(), d_t2.as<real>()};
    for (int sgn = 0; sgn < 2; ++sgn) {
      RET(displace(d_uom.as<real>(), (sgn ? -0.5 : 0.5) * delta, d_rp.as<real>()));
      LAUNCH(1, rbl::ktk_inv_blocks<real>(d_Qp.as<real>(), d_ref.as<real>(), n_bod, n_blb, Sp, d_flags.as<int>() + FLAG_SINGULAR, stream));
      CK(cudaMemcpyAsync(tmp6, d_uom.p, n6 * sizeof(real), cudaMemcpyDeviceToDevice, stream));
      LAUNCH(1, rbl::ktk_inv_apply<real>(Sp, n_bod, n_blb, tmp6, stream));   // (K^T K)^-1 W
      LAUNCH(1, rbl::k_dot<real>(tmp6, d_rp.as<real>(), d_Xp.as<real>(), n_bod, n_blb, (real)1, nullptr, v[sgn], stream));  // Kinv^T W
    }
    LAUNCH(1, rbl::scale_copy<real>(v[0], (real)(1.0 / delta), d_rfd.as<real>(), n3, false, stream));
    LAUNCH(1, rbl::scale_copy<real>(v[1], (real)(-1.0 / delta), d_rfd.as<real>(), n3, true, stream));
    LAUNCH(1, rbl::kt_dot<real>(d_rfd.as<real>(), d_r.as<real>(), d_X.as<real>(), n_bod, n_blb, tmp6 + n6, stream));  // KT * out (:766)
    RET(d2h(out6, tmp6 + n6, n6 * sizeof(real)));
    return sync();
  }
  // blob positions of q +- (delta/2) U   (M_RFD_cfgs, :798-816)
  int RFD_cfgs(const void* U6, double delta, void* r_plus, void* r_minus) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    if (!U6 || !r_plus || !r_minus || !(delta > 0)) return fail(RBL_ERR_INVALID, "M_RFD_cfgs: U, delta > 0 and two outputs are required");
    RET(rfd_buffers());
    const size_t n3 = 3 * (size_t)N(), n6 = 6 * (size_t)n_bod;
    RET(h2d(d_uom.p, U6, n6 * sizeof(real)));
    RET(displace(d_uom.as<real>(), 0.5 * delta, d_t1.as<real>()));
    RET(displace(d_uom.as<real>(), -0.5 * delta, d_t2.as<real>()));
    RET(d2h(r_plus, d_t1.p, n3 * sizeof(real)));
    RET(d2h(r_minus, d_t2.p, n3 * sizeof(real)));
    return sync();
  }
  // (X, Q) displaced by U (units of displacement), state untouched   (update_X_Q_out, :712-728)
  int displaced_config(const void* U6, void* X_out, void* Q_out) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    if (!U6 || !X_out || !Q_out) return fail(RBL_ERR_INVALID, "update_X_Q_out: U and two outputs are required");
    RET(rfd_buffers());
    RET(h2d(d_uom.p, U6, 6 * (size_t)n_bod * sizeof(real)));
    RET(displace(d_uom.as<real>(), 1.0, nullptr));
    RET(d2h(X_out, d_Xp.p, 3 * (size_t)n_bod * sizeof(real)));
    RET(d2h(Q_out, d_Qp.p, 4 * (size_t)n_bod * sizeof(real)));
    return sync();
  }
  // install the configuration displaced by U and rebuild K; a built preconditioner is KEPT, stale on
  // purpose (evolve_X_Q_RFD, :880-893: the displacement is an RFD-sized one)
  int evolve_RFD(const void* U6) override {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    if (!U6) return fail(RBL_ERR_INVALID, "evolve_X_Q_RFD: U is required");
    RET(rfd_buffers());
    RET(h2d(d_uom.p, U6, 6 * (size_t)n_bod * sizeof(real)));
    LAUNCH(1, rbl::integrate<real>(d_uom.as<real>(), (real)1, n_bod, d_X.as<real>(), d_Q.as<real>(), d_X.as<real>(), d_Q.as<real>(), stream));
    const bool keep = pc_set;
    RET(set_K_mats());
    pc_set = keep;
    return sync();
  }

  int bd_step(const void* F_ext, const void* slip, const void* W1, const void* W2, const void* Wr, double kBT_,
              double tol, int restart, int max_iter, double ltol, int lmax, void* U_out, int* iters,
              double* relres) override {
    const bool brownian = kBT_ > 0 && W1 && (W2 || !split_rand) && Wr;
    if (kBT_ > 0 && !brownian)
      return fail(RBL_ERR_INVALID, split_rand ? "bd_step: kBT > 0 needs the three noise vectors W1, W2, Wr"
                                              : "bd_step: kBT > 0 needs the noise vectors W1 and Wr");
    return bd_step_core(F_ext, slip, W1, W2, Wr, false, 0, 0, kBT_, tol, restart, max_iter, ltol, lmax, U_out, iters, relres);
  }
  int bd_step_seeded(const void* F_ext, const void* slip, unsigned long long seed, unsigned long long step, double kBT_,
                     double tol, int restart, int max_iter, double ltol, int lmax, void* U_out, int* iters,
                     double* relres) override {
    return bd_step_core(F_ext, slip, nullptr, nullptr, nullptr, true, seed, step, kBT_, tol, restart, max_iter, ltol, lmax,
                        U_out, iters, relres);
  }
  // global index of this context's first vector element (partitioned: after the lower ranks' blobs)
  unsigned long long first_element() const { return comm ? 3ull * (unsigned long long)comm->first[comm->rank] : 0ull; }
  int normals(unsigned long long seed, unsigned long long step, unsigned long long first, size_t n, void* W1, void* W2,
              void* Wr, bool dev) override {
    if (!W1 || !W2 || !Wr) return fail(RBL_ERR_INVALID, "normals: three output vectors are required");
    if (dev) {
      LAUNCH(1, rbl::normal_triplet<real>(seed, step, first, n, static_cast<real*>(W1), static_cast<real*>(W2),
                                          static_cast<real*>(Wr), stream));
      return RBL_OK;
    }
    for (DevBuf* b : {&d_in0, &d_in1, &d_in2}) CK(b->ensure(n * sizeof(real)));
    LAUNCH(1, rbl::normal_triplet<real>(seed, step, first, n, d_in0.as<real>(), d_in1.as<real>(), d_in2.as<real>(), stream));
    RET(d2h(W1, d_in0.p, n * sizeof(real)));
    RET(d2h(W2, d_in1.p, n * sizeof(real)));
    RET(d2h(Wr, d_in2.p, n * sizeof(real)));
    return sync();
  }
  int bd_step_core(const void* F_ext, const void* slip, const void* W1, const void* W2, const void* Wr, bool rng,
                   unsigned long long seed, unsigned long long step, double kBT_, double tol, int restart, int max_iter,
                   double ltol, int lmax, void* U_out, int* iters, double* relres) {
    if (!cfg_set) return fail(RBL_ERR_STATE, "ERROR CONFIG NOT INITIALIZED YET!!");
    if (!F_ext || !U_out) return fail(RBL_ERR_INVALID, "bd_step: F_ext and U_out are required");
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](int phase) {  // profiling only: a stream sync per phase, never in a timed run
      if (!profile) return;
      cudaStreamSynchronize(stream);
      const auto now = std::chrono::steady_clock::now();
      bd_phase_ms[phase] += std::chrono::duration<double, std::milli>(now - t_last).count();
      t_last = now;
    };
    const bool brownian = kBT_ > 0 && (rng || (W1 && (W2 || !split_rand) && Wr));
    if (brownian && !(dt > 0)) return fail(RBL_ERR_INVALID, "bd_step: dt must be positive");
    RET(need_K());
    const size_t n3 = 3 * (size_t)N(), n6 = 6 * (size_t)n_bod, n = n3 + n6;
    const int nb = (int)N();
    {  // every buffer of the step, allocated and (partitioned mode) agreed before the first collective
      int st_alloc = [&]() -> int {
        CK(d_rhs.ensure(n * sizeof(real)));
        CK(d_sol.ensure(n * sizeof(real)));
        if (brownian) {
          for (DevBuf* b : {&d_mh1, &d_mh2, &d_rfd, &d_noise, &d_rp, &d_t1, &d_t2, &d_wr}) CK(b->ensure(n3 * sizeof(real)));
          CK(d_uom.ensure(n6 * sizeof(real)));
          for (DevBuf* b : {&d_Xs, &d_Xp}) CK(b->ensure(3 * (size_t)n_bod * sizeof(real)));
          for (DevBuf* b : {&d_Qs, &d_Qp}) CK(b->ensure(4 * (size_t)n_bod * sizeof(real)));
        }
        return reserve_comm_workspace(brownian && split_rand && pair_lanczos);
      }();
      if (comm) st_alloc = agree(st_alloc);
      if (st_alloc != RBL_OK) return st_alloc;
    }
    real* rhs = d_rhs.as<real>();
    real* sol = d_sol.as<real>();
    if (slip) RET(h2d(rhs, slip, n3 * sizeof(real)));
    else CK(cudaMemsetAsync(rhs, 0, n3 * sizeof(real), stream));
    RET(h2d(rhs + n3, F_ext, n6 * sizeof(real)));
    if (brownian) {
      real* w1 = d_noise.as<real>();
      real* w2 = d_rfd.as<real>();  // d_rfd is free until the RFD below
      real* noise = d_wr.as<real>();  // W_r of the random finite difference
      if (rng) {  // counter-based normals on the device, a pure function of (seed, step, global element)
        LAUNCH(1, rbl::normal_triplet<real>(seed, step, first_element(), n3, w1, w2, noise, stream));
      } else {
        RET(h2d(w1, W1, n3 * sizeof(real)));
        if (split_rand) RET(h2d(w2, W2, n3 * sizeof(real)));
        RET(h2d(noise, Wr, n3 * sizeof(real)));
      }
      int it = 0;
      mark(0);
      // Brownian increments at q^n  (M_half_W, :661-675, via Lanczos)
      if (!split_rand) {  // one Brownian increment (:949-953)
        RET(dev_lanczos(w1, d_mh1.as<real>(), ltol, lmax, &it, noise_mode >= 1));
        last_lanczos[0] = it;
        last_lanczos[1] = 0;
      } else if (pair_lanczos) {
        RET(dev_lanczos2(w1, w2, d_mh1.as<real>(), d_mh2.as<real>(), ltol, lmax, last_lanczos, noise_mode >= 1));
      } else {
        RET(dev_lanczos(w1, d_mh1.as<real>(), ltol, lmax, &it, noise_mode >= 1));
        last_lanczos[0] = it;
        RET(dev_lanczos(w2, d_mh2.as<real>(), ltol, lmax, &it, noise_mode >= 1));
        last_lanczos[1] = it;
      }
      mark(1);
      // random finite difference  (M_RFD, :769-796)
      RET(kinv_dev(noise, d_uom.as<real>()));
      RET(dev_M_RFD(d_uom.as<real>(), noise, default_delta(), d_rfd.as<real>()));
      mark(2);
      // (no K^T finite difference in the force row: with q+- = q +- (delta/2) K^-1 W_r its expectation
      // (d_k K^T) K^-T e_k vanishes identically -- tests/test_oracle_bd_drift.py)
      // RHS slip -= kBT * RFD + c2 (M^{1/2}W1 - M^{1/2}W2)   (:945-948,963)
      // split_rand: c1 = 2 sqrt(kBT/dt), c2 = sqrt(kBT/dt), BI = c2 (M^{1/2}W1 - M^{1/2}W2)  (:943-948);
      // single increment: c1 = c2 = sqrt(2 kBT/dt), BI = c2 M^{1/2}W1  (:949-953)
      const double c1 = split_rand ? 2.0 * std::sqrt(kBT_ / dt) : std::sqrt(2.0 * kBT_ / dt);
      const double c2 = split_rand ? std::sqrt(kBT_ / dt) : std::sqrt(2.0 * kBT_ / dt);
      LAUNCH(1, rbl::scale_copy<real>(d_rfd.as<real>(), (real)(-kBT_), rhs, n3, true, stream));
      LAUNCH(1, rbl::scale_copy<real>(d_mh1.as<real>(), (real)(-c2), rhs, n3, true, stream));
      if (split_rand) LAUNCH(1, rbl::scale_copy<real>(d_mh2.as<real>(), (real)(c2), rhs, n3, true, stream));
      // midpoint configuration q' = q + (dt/2) K^-1 (c1 M^{1/2}W1)   (:954-958), installed
      RET(kinv_dev(d_mh1.as<real>(), d_uom.as<real>()));
      CK(cudaMemcpyAsync(d_Xs.p, d_X.p, 3 * (size_t)n_bod * sizeof(real), cudaMemcpyDeviceToDevice, stream));
      CK(cudaMemcpyAsync(d_Qs.p, d_Q.p, 4 * (size_t)n_bod * sizeof(real), cudaMemcpyDeviceToDevice, stream));
      LAUNCH(1, rbl::integrate<real>(d_uom.as<real>(), (real)(0.5 * dt * c1), n_bod, d_X.as<real>(), d_Q.as<real>(),
                                     d_X.as<real>(), d_Q.as<real>(), stream));
      RET(set_K_mats());
      pc_set = false;
      mark(3);
    }
    int st = use_mixed(1) ? dev_gmres_mixed(rhs, sol, tol, restart, max_iter, iters, relres)
                          : dev_gmres(rhs, sol, tol, restart, max_iter, iters, relres);
    mark(4);
    if (brownian) {  // back to q^n whatever the solver said
      CK(cudaMemcpyAsync(d_X.p, d_Xs.p, 3 * (size_t)n_bod * sizeof(real), cudaMemcpyDeviceToDevice, stream));
      CK(cudaMemcpyAsync(d_Q.p, d_Qs.p, 4 * (size_t)n_bod * sizeof(real), cudaMemcpyDeviceToDevice, stream));
      r_valid = false;
      pc_set = false;
    }
    if (st != RBL_OK) return st;
    // q^{n+1} = q^n + dt U   (evolve_X_Q, :865-878)
    LAUNCH(1, rbl::integrate<real>(sol + n3, (real)dt, n_bod, d_X.as<real>(), d_Q.as<real>(), d_X.as<real>(),
                                   d_Q.as<real>(), stream));
    RET(set_K_mats());
    pc_set = false;
    RET(d2h(U_out, sol + n3, n6 * sizeof(real)));
    const int st_end = csync();
    mark(5);
    return st_end;
  }

  // ---- measurement ----------------------------------------------------------------------------
  int fma_peak(int iters, double* tflops) override {
    CK(d_out0.ensure(256));
    double flops = 0;
    // warm-up + timed
    LAUNCH(1, rbl::fma_peak_launch<real>(sm_count, std::max(iters / 8, 1), d_out0.as<real>(), &flops, stream));
    CK(cudaEventRecord(t0, stream));
    LAUNCH(1, rbl::fma_peak_launch<real>(sm_count, iters, d_out0.as<real>(), &flops, stream));
    CK(cudaEventRecord(t1, stream));
    CK(cudaEventSynchronize(t1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    *tflops = flops / (ms * 1e-3) / 1e12;
    return RBL_OK;
  }
  int num_sym_variants() const override { return rbl::matvec_sym_num_variants<real>(); }
  int sym_variant_info(int idx, int* T, int* threads) const override {
    if (idx < 0 || idx >= rbl::matvec_sym_num_variants<real>()) return RBL_ERR_INVALID;
    const rbl::MatvecVariant v = rbl::matvec_sym_variant<real>(idx);
    *T = v.T;
    *threads = v.threads;
    return RBL_OK;
  }
  int sym_variant_chunk(int idx) const override {
    if (idx < 0 || idx >= rbl::matvec_sym_num_variants<real>()) return -1;
    return rbl::matvec_sym_variant<real>(idx).rc;
  }
  int num_variants() const override { return rbl::matvec_num_variants<real>(); }
  int variant_info(int idx, int* T, int* threads) const override {
    if (idx < 0 || idx >= rbl::matvec_num_variants<real>()) return RBL_ERR_INVALID;
    const rbl::MatvecVariant v = rbl::matvec_variant<real>(idx);
    *T = v.T;
    *threads = v.threads;
    return RBL_OK;
  }
};

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
#define CTX_OR_FAIL(ctx)          \
  if (!(ctx)) return RBL_ERR_INVALID

extern "C" {

const char* rbl_version(void) { return "rigid_body_light_b200 0.1.0 (sm_100a)"; }

int rbl_create(int precision, int device, rbl_ctx** out) {
  if (!out) return RBL_ERR_INVALID;
  *out = nullptr;
  if (precision != RBL_F32 && precision != RBL_F64) {
    g_create_error = "rbl_create: precision must be RBL_F32 (4) or RBL_F64 (8)";
    return RBL_ERR_INVALID;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    g_create_error = std::string("rbl_create: no CUDA device (there is no CPU fallback): ") +
                     (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return RBL_ERR_CUDA;
  }
  if (device < 0) {
    e = cudaGetDevice(&device);
    if (e != cudaSuccess) {
      g_create_error = std::string("cudaGetDevice: ") + cudaGetErrorString(e);
      return RBL_ERR_CUDA;
    }
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return RBL_ERR_CUDA;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    g_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
    return RBL_ERR_CUDA;
  }
  if (prop.major != 10) {
    g_create_error = std::string("rbl_create: kernels are built for sm_100a only; device is ") + prop.name;
    return RBL_ERR_CUDA;
  }
  rbl_ctx* c = nullptr;
  int st;
  if (precision == RBL_F32) {
    auto* p = new Ctx<float>();
    p->precision = precision; p->device = device; p->sm_count = prop.multiProcessorCount;
    st = p->init();
    c = p;
  } else {
    auto* p = new Ctx<double>();
    p->precision = precision; p->device = device; p->sm_count = prop.multiProcessorCount;
    st = p->init();
    c = p;
  }
  if (st != RBL_OK) {
    g_create_error = c->err;
    delete c;
    return st;
  }
  *out = c;
  return RBL_OK;
}

void rbl_destroy(rbl_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  delete ctx;
}

const char* rbl_last_error(const rbl_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int rbl_precision(const rbl_ctx* ctx) { return ctx ? ctx->precision : 0; }
int rbl_sm_count(const rbl_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

#define BIND_DEVICE(ctx) cudaSetDevice((ctx)->device)

int rbl_set_parameters(rbl_ctx* ctx, double a, double dt, double kBT, double eta, const void* ref_cfg, int n_blb) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->set_parameters(a, dt, kBT, eta, ref_cfg, n_blb);
}
int rbl_set_flags(rbl_ctx* ctx, int block_pc, int wall) { CTX_OR_FAIL(ctx); return ctx->set_flags(block_pc, wall); }
int rbl_set_config(rbl_ctx* ctx, const void* X, const void* Q, int n_bod) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->set_config(X, Q, n_bod); }
int rbl_get_config(rbl_ctx* ctx, void* X, void* Q) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->get_config(X, Q); }
int rbl_set_K_mats(rbl_ctx* ctx) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); int s = ctx->set_K_mats(); return s != RBL_OK ? s : ctx->sync(); }
int rbl_n_bodies(const rbl_ctx* ctx) { return ctx ? ctx->n_bodies() : 0; }
int rbl_blobs_per_body(const rbl_ctx* ctx) { return ctx ? ctx->blobs_per_body() : 0; }

int rbl_blob_positions(rbl_ctx* ctx, void* out) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->blob_positions(out, false); }
int rbl_K_dot(rbl_ctx* ctx, const void* U, void* out) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->K_dot(U, out, false); }
int rbl_KT_dot(rbl_ctx* ctx, const void* lam, void* out) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->KT_dot(lam, out, false); }
int rbl_Kinv_dot(rbl_ctx* ctx, const void* V, void* out) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->Kinv_dot(V, out); }
int rbl_KTinv_dot(rbl_ctx* ctx, const void* F, void* out) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->KTinv_dot(F, out); }
int rbl_apply_M(rbl_ctx* ctx, const void* F, const void* r, int n, void* out) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->apply_M(F, r, n, out); }
int rbl_apply_PC(rbl_ctx* ctx, const void* in, void* out) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->apply_PC(in, out, false); }
int rbl_apply_saddle(rbl_ctx* ctx, const void* x, void* out) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->apply_saddle(x, out, false); }
int rbl_evolve(rbl_ctx* ctx, const void* U) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->evolve(U); }
int rbl_export_K_csc(rbl_ctx* ctx, int64_t* indptr, int32_t* indices, void* data) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->export_K(indptr, indices, data); }
int rbl_export_Kinv_csc(rbl_ctx* ctx, int64_t* indptr, int32_t* indices, void* data) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->export_Kinv(indptr, indices, data); }

int rbl_gmres(rbl_ctx* ctx, const void* rhs, void* x, double tol, int restart, int max_iter, int* iters, double* relres) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  int it = 0; double rr = 0;
  int s = ctx->gmres(rhs, x, tol, restart, max_iter, &it, &rr);
  if (iters) *iters = it;
  if (relres) *relres = rr;
  return s;
}
int rbl_lanczos_sqrt(rbl_ctx* ctx, const void* W, void* out, double tol, int max_iter, int* iters) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  int it = 0;
  int s = ctx->lanczos(W, out, tol, max_iter, &it);
  if (iters) *iters = it;
  return s;
}

int rbl_bd_step(rbl_ctx* ctx, const void* F_ext, const void* slip, const void* W1, const void* W2, const void* Wr,
                double kBT, double tol, int restart, int max_iter, double ltol, int lmax, void* U_out, int* iters,
                double* relres) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  int it = 0; double rr = 0;
  int s = ctx->bd_step(F_ext, slip, W1, W2, Wr, kBT, tol, restart, max_iter, ltol, lmax, U_out, &it, &rr);
  if (iters) *iters = it;
  if (relres) *relres = rr;
  return s;
}

int rbl_bd_step_seeded(rbl_ctx* ctx, const void* F_ext, const void* slip, uint64_t seed, uint64_t step, double kBT,
                       double tol, int restart, int max_iter, double ltol, int lmax, void* U_out, int* iters,
                       double* relres) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  int it = 0; double rr = 0;
  int s = ctx->bd_step_seeded(F_ext, slip, seed, step, kBT, tol, restart, max_iter, ltol, lmax, U_out, &it, &rr);
  if (iters) *iters = it;
  if (relres) *relres = rr;
  return s;
}
int rbl_normals(rbl_ctx* ctx, uint64_t seed, uint64_t step, uint64_t first, size_t n, void* W1, void* W2, void* Wr) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->normals(seed, step, first, n, W1, W2, Wr, false);
}
int rbl_dev_normals(rbl_ctx* ctx, uint64_t seed, uint64_t step, uint64_t first, size_t n, void* dW1, void* dW2, void* dWr) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->normals(seed, step, first, n, dW1, dW2, dWr, true);
}
int rbl_apply_M2(rbl_ctx* ctx, const void* F1, const void* F2, const void* r, int n, void* out1, void* out2) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->apply_M2(F1, F2, r, n, out1, out2);
}
int rbl_dev_apply_M2(rbl_ctx* ctx, const void* dF1, const void* dF2, const void* dr, int n, void* dout1, void* dout2) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->dev_apply_M2(dF1, dF2, dr, n, dout1, dout2);
}
int rbl_lanczos_sqrt2(rbl_ctx* ctx, const void* W1, const void* W2, void* out1, void* out2, double tol, int max_iter,
                      int* iters2) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  int it[2] = {0, 0};
  int s = ctx->lanczos2(W1, W2, out1, out2, tol, max_iter, it);
  if (iters2) { iters2[0] = it[0]; iters2[1] = it[1]; }
  return s;
}
int rbl_set_noise_preconditioner(rbl_ctx* ctx, int mode) {
  CTX_OR_FAIL(ctx);
  if (mode < 0 || mode > 2) return ctx->fail(RBL_ERR_INVALID, "noise preconditioner mode must be 0, 1 or 2");
  ctx->noise_mode = mode;
  return RBL_OK;
}
int rbl_noise_selfcheck(rbl_ctx* ctx, double* factor_err, double* inverse_err, int* active) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  double f = 0, g = 0;
  int a = 0;
  int s = ctx->noise_selfcheck(&f, &g, &a);
  if (factor_err) *factor_err = f;
  if (inverse_err) *inverse_err = g;
  if (active) *active = a;
  return s;
}
int rbl_M_RFD(rbl_ctx* ctx, const void* W, double delta, void* out) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->M_RFD(nullptr, W, delta, out);
}
int rbl_M_RFD_from_U(rbl_ctx* ctx, const void* U, const void* W, double delta, void* out) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  if (!U) return ctx->fail(RBL_ERR_INVALID, "M_RFD_from_U: U is required");
  return ctx->M_RFD(U, W, delta, out);
}
int rbl_KT_RFD_from_U(rbl_ctx* ctx, const void* U, const void* W, double delta, void* out) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->KT_RFD(U, W, delta, out);
}
int rbl_KTinv_RFD(rbl_ctx* ctx, const void* W6, double delta, void* out) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->KTinv_RFD(W6, delta, out);
}
int rbl_M_RFD_cfgs(rbl_ctx* ctx, const void* U, double delta, void* r_plus, void* r_minus) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->RFD_cfgs(U, delta, r_plus, r_minus);
}
int rbl_update_X_Q_out(rbl_ctx* ctx, const void* U, void* X_out, void* Q_out) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->displaced_config(U, X_out, Q_out);
}
int rbl_evolve_RFD(rbl_ctx* ctx, const void* U) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->evolve_RFD(U); }
int rbl_set_rfd_delta(rbl_ctx* ctx, double delta) { CTX_OR_FAIL(ctx); ctx->rfd_delta = delta > 0 ? delta : 0; return RBL_OK; }
int rbl_set_mixed_precision(rbl_ctx* ctx, int mode) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->set_mixed(mode); }
int rbl_plan_cost_bounds(int n_blobs, int tgt_tile, int grid, int part, int n_parts, double w, long long* bounds,
                         long long* total_chunks) {
  if (n_blobs <= 0 || tgt_tile <= 0 || tgt_tile % rbl::kSrcTile != 0 || grid < 1 || n_parts < 1 || part < 0 ||
      part >= n_parts || !bounds)
    return RBL_ERR_INVALID;
  rbl::SymPlan plan{};
  plan.n = n_blobs;
  plan.n_src_tiles = (n_blobs + rbl::kSrcTile - 1) / rbl::kSrcTile;
  plan.tgt_tile = tgt_tile;
  plan.diag = tgt_tile / rbl::kSrcTile;
  plan.n_tgt_tiles = (n_blobs + tgt_tile - 1) / tgt_tile;
  const long long I = plan.n_tgt_tiles;
  plan.units = I * plan.n_src_tiles - (long long)plan.diag * (I * (I - 1) / 2);
  plan.grid = grid;
  rbl::sym_cost_bounds(&plan, part, n_parts, w, bounds);
  if (total_chunks) *total_chunks = plan.units * (rbl::kSrcTile / 32);
  return RBL_OK;
}
int rbl_set_split_weight(rbl_ctx* ctx, double w) { CTX_OR_FAIL(ctx); ctx->split_weight = w > 0 ? w : 0; return RBL_OK; }
int rbl_set_split_rand(rbl_ctx* ctx, int enable) { CTX_OR_FAIL(ctx); ctx->split_rand = enable != 0; return RBL_OK; }
int rbl_set_lanczos_pairing(rbl_ctx* ctx, int enable) { CTX_OR_FAIL(ctx); ctx->pair_lanczos = enable != 0; return RBL_OK; }
int rbl_num_sym2_variants(const rbl_ctx* ctx) { return ctx ? ctx->num_sym2_variants() : 0; }
int rbl_sym2_variant_info(const rbl_ctx* ctx, int idx, int* T, int* threads) {
  if (!ctx || !T || !threads) return RBL_ERR_INVALID;
  return ctx->sym2_variant_info(idx, T, threads);
}
int rbl_set_sym2_variant(rbl_ctx* ctx, int idx) {
  CTX_OR_FAIL(ctx);
  if (idx >= ctx->num_sym2_variants()) return ctx->fail(RBL_ERR_INVALID, "no such two-right-hand-side variant");
  ctx->sym2_variant = idx < 0 ? -1 : idx;
  return RBL_OK;
}
int rbl_dev_apply_M(rbl_ctx* ctx, const void* dF, const void* dr, int n, int t0, int nt, void* dout) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->dev_apply_M(dF, dr, n, t0, nt, dout); }
int rbl_dev_blob_positions(rbl_ctx* ctx, void* dout) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->blob_positions(dout, true); }
int rbl_dev_K_dot(rbl_ctx* ctx, const void* dU, void* dout) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->K_dot(dU, dout, true); }
int rbl_dev_KT_dot(rbl_ctx* ctx, const void* dl, void* dout) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->KT_dot(dl, dout, true); }
int rbl_dev_apply_PC(rbl_ctx* ctx, const void* din, void* dout) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->apply_PC(din, dout, true); }
int rbl_dev_apply_saddle(rbl_ctx* ctx, const void* dx, void* dout) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->apply_saddle(dx, dout, true); }
int rbl_dev_apply_saddle_shard(rbl_ctx* ctx, const void* dl, const void* dr, int n, int t0, const void* dU, void* dout) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->saddle_shard(dl, dr, n, t0, dU, dout);
}
int rbl_dev_apply_M_part(rbl_ctx* ctx, const void* dF, const void* dr, int n, int part, int n_parts, void* dout) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->dev_apply_M_part(dF, dr, n, part, n_parts, dout);
}
int rbl_dev_saddle_finish(rbl_ctx* ctx, const void* dM, const void* dl, const void* dU, void* dout) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->saddle_finish(dM, dl, dU, dout);
}
int rbl_comm_unique_id(void* out128) {
  if (!out128) return RBL_ERR_INVALID;
  std::string why;
  if (!rbl::Comm::unique_id(out128, &why)) {
    g_create_error = "rbl_comm_unique_id: " + why;
    return RBL_ERR_CUDA;
  }
  return RBL_OK;
}
int rbl_comm_init(rbl_ctx* ctx, const void* uid128, int rank, int world, const int* blobs_per_rank) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->comm_init(uid128, rank, world, blobs_per_rank);
}
int rbl_comm_exchange(const rbl_ctx* ctx) { return ctx ? ctx->comm_exchange(nullptr) : 0; }
const char* rbl_comm_exchange_why(const rbl_ctx* ctx) {
  const char* why = "";
  if (ctx) ctx->comm_exchange(&why);
  return why;
}
int rbl_comm_set_exchange(rbl_ctx* ctx, int mode) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->comm_set_exchange(mode); }
int rbl_comm_profile(rbl_ctx* ctx, double* ms3, int* n, int reset) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  return ctx->comm_profile(ms3, n, reset);
}
int rbl_comm_world(const rbl_ctx* ctx) { return ctx && ctx->comm ? ctx->comm->world : 1; }
int rbl_comm_rank(const rbl_ctx* ctx) { return ctx && ctx->comm ? ctx->comm->rank : 0; }
int rbl_sync(rbl_ctx* ctx) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return ctx->sync(); }
void* rbl_stream(rbl_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int rbl_set_stream(rbl_ctx* ctx, void* s) {
  CTX_OR_FAIL(ctx);
  if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
  ctx->stream = (cudaStream_t)s;
  ctx->own_stream = false;
  return RBL_OK;
}

static int cuda_status(rbl_ctx* ctx, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return RBL_OK;
  return ctx->fail(e == cudaErrorMemoryAllocation ? RBL_ERR_NOMEM : RBL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
int rbl_dev_alloc(rbl_ctx* ctx, size_t bytes, void** p) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return cuda_status(ctx, cudaMalloc(p, bytes), "cudaMalloc"); }
int rbl_dev_free(rbl_ctx* ctx, void* p) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return cuda_status(ctx, cudaFree(p), "cudaFree"); }
int rbl_pinned_alloc(rbl_ctx* ctx, size_t bytes, void** p) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return cuda_status(ctx, cudaMallocHost(p, bytes), "cudaMallocHost"); }
int rbl_pinned_free(rbl_ctx* ctx, void* p) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return cuda_status(ctx, cudaFreeHost(p), "cudaFreeHost"); }
int rbl_memcpy_h2d(rbl_ctx* ctx, void* dst, const void* src, size_t bytes) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  int s = cuda_status(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync");
  return s != RBL_OK ? s : cuda_status(ctx, cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
}
int rbl_memcpy_d2h(rbl_ctx* ctx, void* dst, const void* src, size_t bytes) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  int s = cuda_status(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync");
  return s != RBL_OK ? s : cuda_status(ctx, cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
}
int rbl_timer_start(rbl_ctx* ctx) { CTX_OR_FAIL(ctx); BIND_DEVICE(ctx); return cuda_status(ctx, cudaEventRecord(ctx->t0, ctx->stream), "cudaEventRecord"); }
int rbl_timer_stop(rbl_ctx* ctx, double* ms) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  int s = cuda_status(ctx, cudaEventRecord(ctx->t1, ctx->stream), "cudaEventRecord");
  if (s != RBL_OK) return s;
  s = cuda_status(ctx, cudaEventSynchronize(ctx->t1), "cudaEventSynchronize");
  if (s != RBL_OK) return s;
  float f = 0;
  s = cuda_status(ctx, cudaEventElapsedTime(&f, ctx->t0, ctx->t1), "cudaEventElapsedTime");
  if (ms) *ms = f;
  return s;
}
int rbl_flush_l2(rbl_ctx* ctx) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  const size_t bytes = (size_t)256 << 20;  // 2x the 126 MB L2
  int s = cuda_status(ctx, ctx->flush_buf.ensure(bytes), "cudaMalloc(flush)");
  if (s != RBL_OK) return s;
  return cuda_status(ctx, cudaMemsetAsync(ctx->flush_buf.p, 1, bytes, ctx->stream), "cudaMemsetAsync");
}

int rbl_num_matvec_variants(const rbl_ctx* ctx) { return ctx ? ctx->num_variants() : 0; }
int rbl_matvec_variant_info(const rbl_ctx* ctx, int idx, int* T, int* threads) {
  if (!ctx || !T || !threads) return RBL_ERR_INVALID;
  return ctx->variant_info(idx, T, threads);
}
int rbl_set_matvec_variant(rbl_ctx* ctx, int idx) {
  CTX_OR_FAIL(ctx);
  if (idx >= ctx->num_variants()) return ctx->fail(RBL_ERR_INVALID, "no such matvec variant");
  ctx->variant = idx < 0 ? -1 : idx;
  return RBL_OK;
}
int rbl_set_matvec_mode(rbl_ctx* ctx, int mode) {
  CTX_OR_FAIL(ctx);
  if (mode != 0 && mode != 1) return ctx->fail(RBL_ERR_INVALID, "matvec mode must be 0 (symmetric) or 1 (ordered)");
  ctx->mode = mode;
  return RBL_OK;
}
int rbl_num_sym_variants(const rbl_ctx* ctx) { return ctx ? ctx->num_sym_variants() : 0; }
int rbl_sym_variant_info(const rbl_ctx* ctx, int idx, int* T, int* threads) {
  if (!ctx || !T || !threads) return RBL_ERR_INVALID;
  return ctx->sym_variant_info(idx, T, threads);
}
int rbl_sym_variant_chunk(const rbl_ctx* ctx, int idx) { return ctx ? ctx->sym_variant_chunk(idx) : -1; }
int rbl_set_sym_variant(rbl_ctx* ctx, int idx) {
  CTX_OR_FAIL(ctx);
  if (idx >= ctx->num_sym_variants()) return ctx->fail(RBL_ERR_INVALID, "no such symmetric matvec variant");
  ctx->sym_variant = idx < 0 ? -1 : idx;
  return RBL_OK;
}
int64_t rbl_launch_count(const rbl_ctx* ctx) { return ctx ? ctx->launches : 0; }
int64_t rbl_product_count(const rbl_ctx* ctx) { return ctx ? ctx->products : 0; }
int rbl_bd_stats(const rbl_ctx* ctx, int* lanczos_iters_1, int* lanczos_iters_2) {
  if (!ctx) return RBL_ERR_INVALID;
  if (lanczos_iters_1) *lanczos_iters_1 = ctx->last_lanczos[0];
  if (lanczos_iters_2) *lanczos_iters_2 = ctx->last_lanczos[1];
  return RBL_OK;
}
int rbl_profile_matvec(rbl_ctx* ctx, int enable) { CTX_OR_FAIL(ctx); ctx->profile = enable != 0; return RBL_OK; }
int rbl_bd_phase_ms(rbl_ctx* ctx, double* out6, int reset) {
  CTX_OR_FAIL(ctx);
  if (!out6) return RBL_ERR_INVALID;
  for (int i = 0; i < 6; ++i) out6[i] = ctx->bd_phase_ms[i];
  if (reset) for (int i = 0; i < 6; ++i) ctx->bd_phase_ms[i] = 0;
  return RBL_OK;
}
int rbl_matvec_profile(rbl_ctx* ctx, double* avg_ms, int64_t* launches, int reset) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  int s = cuda_status(ctx, cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
  if (s != RBL_OK) return s;
  double tot = 0;
  for (auto& pr : ctx->prof_events) {
    float f = 0;
    s = cuda_status(ctx, cudaEventElapsedTime(&f, pr.first, pr.second), "cudaEventElapsedTime");
    if (s != RBL_OK) return s;
    tot += f;
  }
  const int64_t nl = (int64_t)ctx->prof_events.size();
  if (avg_ms) *avg_ms = nl ? tot / nl : 0.0;
  if (launches) *launches = nl;
  if (reset) {
    for (auto& pr : ctx->prof_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    ctx->prof_events.clear();
  }
  return RBL_OK;
}
int rbl_fma_peak(rbl_ctx* ctx, int iters, double* tflops) {
  CTX_OR_FAIL(ctx); BIND_DEVICE(ctx);
  if (!tflops || iters < 1) return RBL_ERR_INVALID;
  return ctx->fma_peak(iters, tflops);
}

}  // extern "C"
