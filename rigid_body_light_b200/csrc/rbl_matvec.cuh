// rbl_matvec.cuh -- device-side interface of the RPY mobility product U = B M B F.
//
// Replaces  CManyBodies::rotne_prager_tensor + apply_M + make_damp_mat
//           (/root/reference/src/c_rigid_obj.cpp:413-459, 618-659)
// which assemble a dense 3N x 3N matrix and run a GEMV.  Here the product is
// matrix-free: a persistent, stream-K scheduled, TMA-staged tiled n-body kernel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "rbl_pair.cuh"

namespace rbl {

// One packed source/target record: 8 reals, 32 B (fp32) / 64 B (fp64), so a record
// is exactly two (fp32) / four (fp64) 128-bit shared-memory loads and a source tile
// is one contiguous 16-byte-aligned span for a 1-D TMA bulk copy.
//   [0..2] position (unscaled)   [3..5] B_j * force   [6] 2 z   [7] 4 z^2
constexpr int kRecReals = 8;

// Source tile (records per TMA stage).  Compile-time so the inner loop unrolls with
// immediate shared-memory offsets.
constexpr int kSrcTile = 256;

struct MatvecPlan {
  int n_src;        // real sources
  int n_src_tiles;  // ceil(n_src / kSrcTile); the record array is padded to this
  int tgt_first;    // index (into the record array) of the first target
  int n_tgt;        // number of targets
  int n_tgt_tiles;  // ceil(n_tgt / targets-per-CTA-tile)
  int tgt_tile;     // targets per CTA tile (= threads * T of the chosen variant)
  int grid;         // persistent CTAs (multiple of the SM count)
};

template <typename real>
struct MatvecArgs {
  const real* rec;       // packed records, (n_src_tiles * kSrcTile) x 8
  const float* box_src;  // per source tile: min xyz, max xyz
  const float* box_tgt;  // per target tile
  real* out;             // 3 * n_tgt, local target order
  real* scratch;         // 2 * grid partial tiles, [slot][3][tgt_tile]
  MatvecPlan plan;
  PairConsts<real> C;
  int wall;
};

// Which kernel variant to launch (targets per thread x threads per CTA).  0 = default.
// Variants exist so the tile shape can be tuned on hardware without recompiling.
struct MatvecVariant {
  int T;
  int threads;
  int rc = 0;  // symmetric kernel: sources per warp-private reaction-reduction chunk (0 = warp butterfly)
};

template <typename real>
int matvec_num_variants();
template <typename real>
MatvecVariant matvec_variant(int idx);

// Plans a launch: picks tile sizes and the persistent grid from the occupancy of
// the chosen variant on the current device.
template <typename real>
cudaError_t matvec_plan(int variant, bool wall, int n_src, int tgt_first, int n_tgt,
                        int sm_count, MatvecPlan* plan);

// Packs positions + forces into records (and flags blobs below the wall).
//   r, F: 3 n reals each (blob-major xyz).  rec must hold n_src_tiles*kSrcTile records.
//   below_wall_flag: device int, set to 1 if wall && any z < 0.
template <typename real>
cudaError_t pack_records(const real* r, const real* F, int n, int n_padded, bool wall,
                         real a, real* rec, int* below_wall_flag, cudaStream_t s);

// Only refreshes the force part of already packed records (positions unchanged).
template <typename real>
cudaError_t repack_forces(const real* F, int n, bool wall, real a, real* rec, int* below_wall_flag,
                          cudaStream_t s);

// Axis-aligned boxes of consecutive groups of `tile` records starting at `first`.
template <typename real>
cudaError_t tile_boxes(const real* rec, int first, int count, int tile, float* boxes,
                       cudaStream_t s, int stride = kRecReals);

// The product itself: matvec kernel + deterministic fix-up of split target tiles.
// If ev0/ev1 are non-null they are recorded immediately around the main kernel (the
// roofline's "dominant kernel" duration, measured live on the launching stream).
template <typename real>
cudaError_t matvec_launch(int variant, const MatvecArgs<real>& args, cudaStream_t s,
                          cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr);

// ---- symmetric product (targets == sources) ----------------------------------------------
// Evaluates every UNORDERED blob pair once and applies the block in both directions
// (pair_sym); 1.25-1.4x fewer instructions per ordered pair than the ordered kernel.  The
// reaction on the source side is reduced across the warp and accumulated with global
// floating-point atomics, so results are reproducible to rounding, not bit for bit (the
// ordered kernel stays available for that).  Work = the upper triangle of the
// (target tile x 256-source tile) grid, linearised row by row; [part, n_parts) selects a
// contiguous share of it (multi-GPU: the partial products are summed with an all-reduce).
struct SymPlan {
  int n;            // blobs (targets == sources)
  int n_src_tiles;  // ceil(n / kSrcTile)
  int n_tgt_tiles;  // ceil(n / tgt_tile)
  int tgt_tile;     // T * threads, a multiple of kSrcTile
  int diag;         // tgt_tile / kSrcTile: source tiles per row that overlap the target tile
  long long units;  // units of the whole triangle
  long long u0, u1; // this launch's share, in 32-source chunks (8 per tile unit)
  int grid;
};

template <typename real>
struct SymArgs {
  const real* rec;
  const float* box_src;
  const float* box_tgt;
  real* raw;   // 3 * n_src_tiles * kSrcTile accumulators, zeroed by the launcher
  real* out;   // 3 * n
  const long long* bounds;  // grid + 1 cut points of [u0, u1) (equal-COST shares, sym_cost_bounds); null: equal counts
  SymPlan plan;
  PairConsts<real> C;
  int wall;
};

// Equal-COST cuts of the unit triangle.  A diagonal unit (source tile inside the target tile) runs the
// ordered loop on T x 256 ordered pairs and costs w_diag (< 1) of a symmetric unit, and short rows hold
// relatively more of them: equal COUNTS would let the CTAs -- and, on a partitioned suspension, the GPUs
// -- that own the end of the triangle finish ~1-2 % early.  Fills bounds[0..grid] with the cut points of
// share `part` of `n_parts` (32-source chunk indices, non-decreasing) and sets plan->u0 / plan->u1 to
// the share's ends.  Pure host arithmetic, O(rows + grid).
void sym_cost_bounds(SymPlan* plan, int part, int n_parts, double w_diag, long long* bounds);

template <typename real>
int matvec_sym_num_variants();
template <typename real>
MatvecVariant matvec_sym_variant(int idx);
template <typename real>
int matvec_sym_default_variant(bool wall, int n);
template <typename real>
cudaError_t matvec_sym_plan(int variant, bool wall, int n, int part, int n_parts, int sm_count,
                            SymPlan* plan);
// memset(raw) + symmetric kernel + scale kernel (out = raw * B_i / (8 pi eta))
template <typename real>
cudaError_t matvec_sym_launch(int variant, const SymArgs<real>& args, cudaStream_t s,
                              cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr);

// ---- symmetric product with TWO right-hand sides ------------------------------------------
// U1 = B M B F1 and U2 = B M B F2 in one pass over the unordered pairs: the geometry of a pair
// (the larger half of its cost) is evaluated once for both.  Used by the paired Lanczos of a
// BD step.  Record: 12 reals  x y z f1x | f1y f1z f2x f2y | f2z -4z^2 0 0.
constexpr int kRec2Reals = 12;

template <typename real>
struct Sym2Args {
  const real* rec;   // (n_src_tiles * kSrcTile) x 12
  const float* box_src;
  const float* box_tgt;
  real* raw;         // 2 x (3 * n_src_tiles * kSrcTile) accumulators, zeroed by the launcher
  real* out1;        // 3 * n
  real* out2;        // 3 * n
  const long long* bounds;  // see SymArgs
  SymPlan plan;
  PairConsts<real> C;
  int wall;
};

template <typename real>
cudaError_t pack_records2(const real* r, const real* F1, const real* F2, int n, int n_padded, bool wall, real a,
                          real* rec, int* below_wall_flag, cudaStream_t s);
template <typename real>
cudaError_t repack_forces2(const real* F1, const real* F2, int n, bool wall, real a, real* rec,
                           int* below_wall_flag, cudaStream_t s);
template <typename real>
int matvec_sym2_num_variants();
template <typename real>
MatvecVariant matvec_sym2_variant(int idx);
template <typename real>
int matvec_sym2_default_variant(bool wall, int n);
template <typename real>
cudaError_t matvec_sym2_plan(int variant, bool wall, int n, int part, int n_parts, int sm_count, SymPlan* plan);
template <typename real>
cudaError_t matvec_sym2_launch(int variant, const Sym2Args<real>& args, cudaStream_t s,
                               cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr);

// FMA-pipe peak microbenchmark (the roofline denominator SURVEY.md section 8d asks for).
// Returns flop executed; time it with events around the call.
template <typename real>
cudaError_t fma_peak_launch(int sm_count, int iters, real* sink, double* flops,
                            cudaStream_t s);

}  // namespace rbl
