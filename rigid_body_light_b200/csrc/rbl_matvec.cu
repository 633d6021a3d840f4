// rbl_matvec.cu -- the O(N^2) hot path: U = B M B F, matrix-free, for sm_100a.
//
// Reference being replaced: CManyBodies::apply_M / rotne_prager_tensor / make_damp_mat,
// /root/reference/src/c_rigid_obj.cpp:413-459,618-659 (dense assembly + GEMV).
//
// Design:
//  * pack_records: positions + damped forces -> 8-real records (one TMA-able array).
//  * rpy_matvec_kernel: persistent CTAs, one per resident slot of the 148 SMs.  The
//    (target tile x source tile) unit grid is cut into `grid` equal contiguous unit
//    ranges (stream-K), so every SM gets the same number of pair evaluations to
//    within one 256-source tile regardless of N.  Each thread owns T targets in
//    registers; source tiles are streamed through a 2-stage shared-memory ring by
//    1-D TMA bulk copies (cp.async.bulk + mbarrier) issued by one thread; every
//    thread reads the same source record per step (shared-memory broadcast, two
//    LDS.128 per source in fp32 amortised over T targets).
//  * Far-field fast path: per-tile bounding boxes decide, per unit and CTA-uniformly,
//    whether any pair of the unit can have r < 2a; if not, the overlap branch and its
//    selects are compiled out of the loop.
//  * rpy_fixup_kernel: target tiles whose source range was split between CTAs are
//    summed from per-CTA scratch slots in CTA order -- deterministic, no atomics.
#include <cstdio>

#include "rbl_matvec.cuh"

namespace rbl {

// ----------------------------------------------------------------------------------
// small PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: SYNCS.*, UBLKCP)
// ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src,
                                            uint32_t bytes, unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ----------------------------------------------------------------------------------
// record packing
// ----------------------------------------------------------------------------------
template <typename real>
struct Vec;
template <>
struct Vec<float> {
  using v4 = float4;
};
template <>
struct Vec<double> {
  using v2 = double2;
};

__device__ __forceinline__ void store_rec(float* rec, size_t k, float x, float y, float z,
                                          float fx, float fy, float fz) {
  float4* p = reinterpret_cast<float4*>(rec + k * kRecReals);
  p[0] = make_float4(x, y, z, fx);
  p[1] = make_float4(fy, fz, 2.0f * z, 4.0f * z * z);
}
__device__ __forceinline__ void store_rec(double* rec, size_t k, double x, double y,
                                          double z, double fx, double fy, double fz) {
  double2* p = reinterpret_cast<double2*>(rec + k * kRecReals);
  p[0] = make_double2(x, y);
  p[1] = make_double2(z, fx);
  p[2] = make_double2(fy, fz);
  p[3] = make_double2(2.0 * z, 4.0 * z * z);
}

// B_j of make_damp_mat (c_rigid_obj.cpp:618-639): 1 if z >= a else z/a.
template <typename real>
__device__ __forceinline__ real damp(real z, real a, real inv_a) {
  return z >= a ? (real)1 : z * inv_a;
}

template <typename real>
__global__ void pack_records_kernel(const real* __restrict__ r, const real* __restrict__ F,
                                    int n, int n_padded, int wall, real a, real inv_a,
                                    real* __restrict__ rec, int* __restrict__ below) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_padded) return;
  int src = k < n ? k : n - 1;  // padding = copy of the last blob with zero force
  real x = r[3 * (size_t)src], y = r[3 * (size_t)src + 1], z = r[3 * (size_t)src + 2];
  real fx = 0, fy = 0, fz = 0;
  if (k < n) {
    real b = wall ? damp(z, a, inv_a) : (real)1;
    fx = b * F[3 * (size_t)k];
    fy = b * F[3 * (size_t)k + 1];
    fz = b * F[3 * (size_t)k + 2];
    if (wall && z < (real)0) *below = 1;  // reference throws here (c_rigid_obj.cpp:95-97)
  }
  store_rec(rec, (size_t)k, x, y, z, fx, fy, fz);
}

template <typename real>
__global__ void repack_forces_kernel(const real* __restrict__ F, int n, int wall, real a,
                                     real inv_a, real* __restrict__ rec, int* __restrict__ below) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  real* p = rec + (size_t)k * kRecReals;
  const real z = p[2];
  real b = wall ? damp(z, a, inv_a) : (real)1;
  p[3] = b * F[3 * (size_t)k];
  p[4] = b * F[3 * (size_t)k + 1];
  p[5] = b * F[3 * (size_t)k + 2];
  if (wall && z < (real)0) *below = 1;  // raised on every product, like the reference (c_rigid_obj.cpp:95-97)
}

template <typename real>
cudaError_t pack_records(const real* r, const real* F, int n, int n_padded, bool wall,
                         real a, real* rec, int* below, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  int threads = 256, blocks = (n_padded + threads - 1) / threads;
  pack_records_kernel<real><<<blocks, threads, 0, s>>>(r, F, n, n_padded, wall ? 1 : 0, a,
                                                       (real)1 / a, rec, below);
  return cudaGetLastError();
}
template <typename real>
cudaError_t repack_forces(const real* F, int n, bool wall, real a, real* rec, int* below,
                          cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  int threads = 256, blocks = (n + threads - 1) / threads;
  repack_forces_kernel<real><<<blocks, threads, 0, s>>>(F, n, wall ? 1 : 0, a, (real)1 / a,
                                                        rec, below);
  return cudaGetLastError();
}

// ----------------------------------------------------------------------------------
// tile bounding boxes (outward-rounded floats so the far test is conservative)
// ----------------------------------------------------------------------------------
__device__ __forceinline__ float to_float_down(float v) { return v; }
__device__ __forceinline__ float to_float_up(float v) { return v; }
__device__ __forceinline__ float to_float_down(double v) { return __double2float_rd(v); }
__device__ __forceinline__ float to_float_up(double v) { return __double2float_ru(v); }

template <typename real>
__global__ void tile_boxes_kernel(const real* __restrict__ rec, int first, int count,
                                  int tile, float* __restrict__ boxes, int stride) {
  const int t = blockIdx.x;
  const int lo = t * tile;
  const int hi = min(lo + tile, count);
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int k = lo + threadIdx.x; k < hi; k += blockDim.x) {
    const real* p = rec + (size_t)(first + k) * stride;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      real v = p[c];
      mn[c] = fminf(mn[c], to_float_down(v));
      mx[c] = fmaxf(mx[c], to_float_up(v));
    }
  }
  __shared__ float smn[3][32], smx[3][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
      mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
    }
    if (lane == 0) {
      smn[c][wid] = mn[c];
      smx[c][wid] = mx[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    float a = smn[c][0], b = smx[c][0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      a = fminf(a, smn[c][w]);
      b = fmaxf(b, smx[c][w]);
    }
    boxes[6 * (size_t)t + c] = a;
    boxes[6 * (size_t)t + 3 + c] = b;
  }
}

template <typename real>
cudaError_t tile_boxes(const real* rec, int first, int count, int tile, float* boxes,
                       cudaStream_t s, int stride) {
  if (count <= 0) return cudaSuccess;
  int tiles = (count + tile - 1) / tile;
  tile_boxes_kernel<real><<<tiles, 128, 0, s>>>(rec, first, count, tile, boxes, stride);
  return cudaGetLastError();
}

// squared gap between two boxes, evaluated in double from (outward-rounded) floats
__device__ __forceinline__ double box_gap2(const float* __restrict__ A,
                                           const float* __restrict__ B) {
  double d2 = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double g = fmax(fmax((double)__ldg(A + c) - (double)__ldg(B + 3 + c),
                         (double)__ldg(B + c) - (double)__ldg(A + 3 + c)),
                    0.0);
    d2 += g * g;
  }
  return d2;
}

// ----------------------------------------------------------------------------------
// inner loop over one staged source tile
// ----------------------------------------------------------------------------------
template <bool WALL, bool NEAR, int T>
__device__ __forceinline__ void tile_compute(const float* __restrict__ sb,
                                             const PairConsts<float>& C, const float (&xi)[T],
                                             const float (&yi)[T], const float (&zi)[T],
                                             float (&ux)[T], float (&uy)[T], float (&uz)[T],
                                             int jb = 0, int je = kSrcTile) {
  const float4* __restrict__ s4 = reinterpret_cast<const float4*>(sb);
  // two-level summation: a fresh accumulator per 256-source tile, added to the running sum
  // once per tile.  Keeps the fp32 rounding error at ~sqrt(256)+sqrt(N/256) ulps instead of
  // sqrt(N) (N = 162 000 sequential fp32 adds would sit right at the 1e-5 parity bound).
  float lx[T], ly[T], lz[T];
#pragma unroll
  for (int t = 0; t < T; ++t) lx[t] = ly[t] = lz[t] = 0.0f;
#pragma unroll 2
  for (int j = jb; j < je; ++j) {
    const float4 p = s4[2 * j];      // x y z fx
    const float4 q = s4[2 * j + 1];  // fy fz 2z 4z^2
#pragma unroll
    for (int t = 0; t < T; ++t)
      pair<float, WALL, NEAR>(C, xi[t], yi[t], zi[t], p.x, p.y, p.z, p.w, q.x, q.y, q.z, q.w,
                              lx[t], ly[t], lz[t]);
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    ux[t] += lx[t];
    uy[t] += ly[t];
    uz[t] += lz[t];
  }
}

template <bool WALL, bool NEAR, int T>
__device__ __forceinline__ void tile_compute(const double* __restrict__ sb,
                                             const PairConsts<double>& C,
                                             const double (&xi)[T], const double (&yi)[T],
                                             const double (&zi)[T], double (&ux)[T],
                                             double (&uy)[T], double (&uz)[T], int jb = 0,
                                             int je = kSrcTile) {
  const double2* __restrict__ s2 = reinterpret_cast<const double2*>(sb);
#pragma unroll 2
  for (int j = jb; j < je; ++j) {
    const double2 p0 = s2[4 * j];      // x y
    const double2 p1 = s2[4 * j + 1];  // z fx
    const double2 p2 = s2[4 * j + 2];  // fy fz
    double2 p3 = make_double2(0.0, 0.0);
    if (WALL) p3 = s2[4 * j + 3];      // 2z 4z^2
#pragma unroll
    for (int t = 0; t < T; ++t)
      pair<double, WALL, NEAR>(C, xi[t], yi[t], zi[t], p0.x, p0.y, p1.x, p1.y, p2.x, p2.y,
                               p3.x, p3.y, ux[t], uy[t], uz[t]);
  }
}

// ----------------------------------------------------------------------------------
// the persistent stream-K matvec kernel
// ----------------------------------------------------------------------------------
template <typename real, bool WALL, int T, int NT>
__global__ void __launch_bounds__(NT) rpy_matvec_kernel(const MatvecArgs<real> A) {
  constexpr int TT = T * NT;
  constexpr uint32_t kTileBytes = kSrcTile * kRecReals * sizeof(real);
  __shared__ __align__(128) real sbuf[2][kSrcTile * kRecReals];
  __shared__ __align__(8) unsigned long long mbar[2];

  const int tid = threadIdx.x;
  const int ns = A.plan.n_src_tiles;
  const long long U = (long long)A.plan.n_tgt_tiles * ns;
  const long long g0 = U * blockIdx.x / gridDim.x;
  const long long g1 = U * (blockIdx.x + 1) / gridDim.x;
  if (g0 >= g1) return;  // CTA-uniform: more CTAs than units

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  int ttile = (int)(g0 / ns);
  int stile = (int)(g0 - (long long)ttile * ns);
  const int first_ttile = ttile;
  if (tid == 0) {
    mbar_expect_tx(&mbar[0], kTileBytes);
    tma_load_1d(sbuf[0], A.rec + (size_t)stile * kSrcTile * kRecReals, kTileBytes, &mbar[0]);
  }

  real xi[T], yi[T], zi[T], ux[T], uy[T], uz[T];
  bool fresh = true;       // accumulators must be (re)initialised for `ttile`
  bool seg_from_start = (stile == 0);
  const double near2 = (double)A.C.four_a2 * (1.0 + 1e-6);

  for (long long g = g0; g < g1; ++g) {
    const int it = (int)(g - g0);
    const int buf = it & 1;
    const uint32_t parity = (uint32_t)(it >> 1) & 1u;

    // prefetch the next unit's source tile into the other stage (it was released by
    // the __syncthreads at the end of the previous iteration)
    if (tid == 0 && g + 1 < g1) {
      int nst = stile + 1 == ns ? 0 : stile + 1;
      mbar_expect_tx(&mbar[buf ^ 1], kTileBytes);
      tma_load_1d(sbuf[buf ^ 1], A.rec + (size_t)nst * kSrcTile * kRecReals, kTileBytes,
                  &mbar[buf ^ 1]);
    }

    if (fresh) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        int li = ttile * TT + tid + t * NT;
        if (li >= A.plan.n_tgt) li = A.plan.n_tgt - 1;  // padding lanes recompute the last target
        const real* p = A.rec + (size_t)(A.plan.tgt_first + li) * kRecReals;
        xi[t] = p[0];
        yi[t] = p[1];
        zi[t] = p[2];
        ux[t] = uy[t] = uz[t] = (real)0;
      }
      fresh = false;
    }

    const bool far =
        box_gap2(A.box_tgt + 6 * (size_t)ttile, A.box_src + 6 * (size_t)stile) > near2;

    mbar_wait(&mbar[buf], parity);
    if (far)
      tile_compute<WALL, false, T>(sbuf[buf], A.C, xi, yi, zi, ux, uy, uz);
    else
      tile_compute<WALL, true, T>(sbuf[buf], A.C, xi, yi, zi, ux, uy, uz);
    __syncthreads();  // every thread is done with sbuf[buf] before it is refilled

    const bool tile_end = (stile + 1 == ns);
    if (tile_end || g + 1 == g1) {
      const bool complete = seg_from_start && tile_end;
      if (complete) {
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int li = ttile * TT + tid + t * NT;
          if (li < A.plan.n_tgt) {
            real sc = A.C.out_scale;
            if (WALL) sc *= damp(zi[t], A.C.a, A.C.inv_a);
            A.out[3 * (size_t)li + 0] = ux[t] * sc;
            A.out[3 * (size_t)li + 1] = uy[t] * sc;
            A.out[3 * (size_t)li + 2] = uz[t] * sc;
          }
        }
      } else {
        const size_t slot = 2 * (size_t)blockIdx.x + (ttile == first_ttile ? 0 : 1);
        real* sp = A.scratch + slot * 3 * TT;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int l = tid + t * NT;
          sp[l] = ux[t];
          sp[TT + l] = uy[t];
          sp[2 * TT + l] = uz[t];
        }
      }
      fresh = true;
      seg_from_start = true;
    }
    if (tile_end) {
      stile = 0;
      ++ttile;
    } else {
      ++stile;
    }
  }
}

// Sums the scratch slots of target tiles that were split between CTAs, in CTA order.
template <typename real, bool WALL>
__global__ void rpy_fixup_kernel(const MatvecArgs<real> A) {
  const int t = blockIdx.x;
  const int TT = A.plan.tgt_tile;
  const int ns = A.plan.n_src_tiles;
  const long long G = A.plan.grid;
  const long long U = (long long)A.plan.n_tgt_tiles * ns;
  const long long b = (long long)t * ns, e = b + ns;
  long long c = b * G / U;
  while (c > 0 && U * c / G > b) --c;
  while (U * (c + 1) / G <= b) ++c;
  {
    const long long g0 = U * c / G, g1 = U * (c + 1) / G;
    if (g0 <= b && g1 >= e) return;  // one CTA owned the whole tile and wrote it itself
  }
  for (int l = threadIdx.x; l < TT; l += blockDim.x) {
    const int li = t * TT + l;
    if (li >= A.plan.n_tgt) break;
    real sx = 0, sy = 0, sz = 0;
    for (long long cc = c; cc < G; ++cc) {
      const long long g0 = U * cc / G;
      if (g0 >= e) break;
      const long long g1 = U * (cc + 1) / G;
      if (g0 >= g1) continue;
      const int first_tt = (int)(g0 / ns);
      const size_t slot = 2 * (size_t)cc + (t == first_tt ? 0 : 1);
      const real* sp = A.scratch + slot * 3 * TT;
      sx += sp[l];
      sy += sp[TT + l];
      sz += sp[2 * TT + l];
    }
    real sc = A.C.out_scale;
    if (WALL) {
      const real z = A.rec[(size_t)(A.plan.tgt_first + li) * kRecReals + 2];
      sc *= damp(z, A.C.a, A.C.inv_a);
    }
    A.out[3 * (size_t)li + 0] = sx * sc;
    A.out[3 * (size_t)li + 1] = sy * sc;
    A.out[3 * (size_t)li + 2] = sz * sc;
  }
}

// ----------------------------------------------------------------------------------
// variants, planning, launch
// ----------------------------------------------------------------------------------
#define RBL_F32_VARIANTS(X) X(4, 256) X(8, 128) X(4, 128) X(2, 256) X(8, 256) X(1, 128)
#define RBL_F64_VARIANTS(X) X(2, 256) X(4, 128) X(2, 128) X(4, 256) X(1, 256) X(1, 128)

template <>
int matvec_num_variants<float>() { return 6; }
template <>
int matvec_num_variants<double>() { return 6; }

template <>
MatvecVariant matvec_variant<float>(int idx) {
  static const MatvecVariant v[] = {
#define X(T, NT) {T, NT},
      RBL_F32_VARIANTS(X)
#undef X
  };
  return v[idx];
}
template <>
MatvecVariant matvec_variant<double>(int idx) {
  static const MatvecVariant v[] = {
#define X(T, NT) {T, NT},
      RBL_F64_VARIANTS(X)
#undef X
  };
  return v[idx];
}

template <typename real, bool WALL, int T, int NT>
static cudaError_t occupancy_of(int* blocks_per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      blocks_per_sm, rpy_matvec_kernel<real, WALL, T, NT>, NT, 0);
}

template <typename real>
static cudaError_t variant_occupancy(int variant, bool wall, int* bps);

template <>
cudaError_t variant_occupancy<float>(int variant, bool wall, int* bps) {
  int k = 0;
#define X(T, NT)                                                              \
  if (variant == k++)                                                         \
    return wall ? occupancy_of<float, true, T, NT>(bps) : occupancy_of<float, false, T, NT>(bps);
  RBL_F32_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}
template <>
cudaError_t variant_occupancy<double>(int variant, bool wall, int* bps) {
  int k = 0;
#define X(T, NT)                                                              \
  if (variant == k++)                                                         \
    return wall ? occupancy_of<double, true, T, NT>(bps) : occupancy_of<double, false, T, NT>(bps);
  RBL_F64_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

template <typename real>
cudaError_t matvec_plan(int variant, bool wall, int n_src, int tgt_first, int n_tgt,
                        int sm_count, MatvecPlan* plan) {
  if (variant < 0 || variant >= matvec_num_variants<real>()) return cudaErrorInvalidValue;
  const MatvecVariant v = matvec_variant<real>(variant);
  int bps = 0;
  cudaError_t e = variant_occupancy<real>(variant, wall, &bps);
  if (e != cudaSuccess) return e;
  if (bps < 1) return cudaErrorLaunchOutOfResources;
  plan->n_src = n_src;
  plan->n_src_tiles = (n_src + kSrcTile - 1) / kSrcTile;
  plan->tgt_first = tgt_first;
  plan->n_tgt = n_tgt;
  plan->tgt_tile = v.T * v.threads;
  plan->n_tgt_tiles = (n_tgt + plan->tgt_tile - 1) / plan->tgt_tile;
  plan->grid = sm_count * bps;
  return cudaSuccess;
}

template <typename real, bool WALL, int T, int NT>
static cudaError_t launch_one(const MatvecArgs<real>& a, cudaStream_t s, cudaEvent_t ev0,
                              cudaEvent_t ev1) {
  if (ev0) cudaEventRecord(ev0, s);
  rpy_matvec_kernel<real, WALL, T, NT><<<a.plan.grid, NT, 0, s>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (ev1) cudaEventRecord(ev1, s);
  rpy_fixup_kernel<real, WALL><<<a.plan.n_tgt_tiles, 256, 0, s>>>(a);
  return cudaGetLastError();
}

template <>
cudaError_t matvec_launch<float>(int variant, const MatvecArgs<float>& a, cudaStream_t s,
                                 cudaEvent_t ev0, cudaEvent_t ev1) {
  if (a.plan.n_tgt <= 0 || a.plan.n_src <= 0) return cudaSuccess;
  int k = 0;
#define X(T, NT)                                                   \
  if (variant == k++)                                              \
    return a.wall ? launch_one<float, true, T, NT>(a, s, ev0, ev1) : launch_one<float, false, T, NT>(a, s, ev0, ev1);
  RBL_F32_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}
template <>
cudaError_t matvec_launch<double>(int variant, const MatvecArgs<double>& a, cudaStream_t s,
                                  cudaEvent_t ev0, cudaEvent_t ev1) {
  if (a.plan.n_tgt <= 0 || a.plan.n_src <= 0) return cudaSuccess;
  int k = 0;
#define X(T, NT)                                                   \
  if (variant == k++)                                              \
    return a.wall ? launch_one<double, true, T, NT>(a, s, ev0, ev1) : launch_one<double, false, T, NT>(a, s, ev0, ev1);
  RBL_F64_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}


// ----------------------------------------------------------------------------------
// symmetric kernel: one evaluation per unordered pair
// ----------------------------------------------------------------------------------
__device__ __forceinline__ void load_full_rec(const float* p, float& x, float& y, float& z, float& fx,
                                              float& fy, float& fz, float& nz4) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  x = a.x; y = a.y; z = a.z; fx = a.w; fy = b.x; fz = b.y; nz4 = -b.w;
}
__device__ __forceinline__ void load_full_rec(const double* p, double& x, double& y, double& z, double& fx,
                                              double& fy, double& fz, double& nz4) {
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = q[0], b = q[1], c = q[2], d = q[3];
  x = a.x; y = a.y; z = b.x; fx = b.y; fy = c.x; fz = c.y; nz4 = -d.y;
}

// ---- reaction sums through a warp-private shared-memory tile ---------------------------------
// A warp butterfly per source costs ~19.5 issue slots per source and warp (5 shuffle stages x 3 values).
// Here every lane parks its partial reaction on source j (summed over its T targets) in row
// (j, component) of a warp-private tile, and after RC sources the tile is read back TRANSPOSED:
// 32 / RC lanes per source, each sums RC consecutive lane-partials with 128-bit loads and adds its
// share; the 32 / RC partial sums of a source are combined with log2(32 / RC) shuffles and ONE lane per
// source adds the total to the global accumulator (RED.ADD).  3 STS per source + (3 RC/4 LDS.128 +
// 3 (RC-1) adds + 3 RED) per RC sources  ~=  7 issue slots per source and warp, and the sum order inside
// the warp is fixed.  The global accumulators are stored component-major (x[N] y[N] z[N]), so one RED
// instruction of RC lanes touches RC consecutive words (2 sectors for RC = 16) instead of RC words with
// stride 3 (6 sectors): the L2 atomic units see a third of the sector operations.  Row stride: 3 * stride words = 12 (mod 32) makes the transposed 128-bit reads of
// a quarter-warp hit 8 distinct 4-bank groups; the writes are lane-contiguous.
template <typename real>
struct RedLayout;
template <>
struct RedLayout<float> {
  static constexpr int kStride = 36;  // 3 * 36 words = 108 = 12 (mod 32)
};
template <>
struct RedLayout<double> {
  static constexpr int kStride = 34;  // 3 * 68 words = 204 = 12 (mod 32)
};
template <typename real, int RC>
__host__ __device__ constexpr size_t red_tile_reals() {
  return RC > 0 ? (size_t)RC * 3 * RedLayout<real>::kStride : 0;
}

template <int RC>
__device__ __forceinline__ float row_sum(const float* __restrict__ row) {
  const float4* q = reinterpret_cast<const float4*>(row);
  float4 v = q[0];
#pragma unroll
  for (int k = 1; k < RC / 4; ++k) {
    const float4 w = q[k];
    v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
  }
  return (v.x + v.y) + (v.z + v.w);
}
template <int RC>
__device__ __forceinline__ double row_sum(const double* __restrict__ row) {
  const double2* q = reinterpret_cast<const double2*>(row);
  double2 v = q[0], u = q[1];
#pragma unroll
  for (int k = 1; k < RC / 4; ++k) {
    const double2 w0 = q[2 * k], w1 = q[2 * k + 1];
    v.x += w0.x; v.y += w0.y; u.x += w1.x; u.y += w1.y;
  }
  return (v.x + v.y) + (u.x + u.y);
}

template <typename real, bool WALL, bool NEAR, int T, int RC>
__device__ __forceinline__ void tile_compute_symt(const real* __restrict__ sb, const PairConsts<real>& C,
                                                  const real (&xi)[T], const real (&yi)[T], const real (&zi)[T],
                                                  const real (&fxi)[T], const real (&fyi)[T], const real (&fzi)[T],
                                                  const real (&nz4i)[T], real (&lx)[T], real (&ly)[T],
                                                  real (&lz)[T], real* __restrict__ raw_tile, size_t raw_ld,
                                                  real* __restrict__ red, int jb, int je) {
  static_assert(RC == 8 || RC == 16 || RC == 32, "reduction chunk");
  constexpr int STR = RedLayout<real>::kStride;
  const int lane = threadIdx.x & 31;
  const int src = lane % RC, part = lane / RC;
  real* __restrict__ wr = red + lane;
  const real* __restrict__ rd = red + (size_t)(src * 3) * STR + part * RC;
  for (int j0 = jb; j0 < je; j0 += RC) {
#pragma unroll 1
    for (int jj = 0; jj < RC / 2; ++jj) {
      real xa, ya, za, fxa, fya, fza, nz4a, xb, yb, zb, fxb, fyb, fzb, nz4b;
      load_full_rec(sb + (size_t)(j0 + jj) * kRecReals, xa, ya, za, fxa, fya, fza, nz4a);
      load_full_rec(sb + (size_t)(j0 + jj + RC / 2) * kRecReals, xb, yb, zb, fxb, fyb, fzb, nz4b);
      real ax = 0, ay = 0, az = 0, bx = 0, by = 0, bz = 0;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        pair_sym<real, WALL, NEAR>(C, xi[t], yi[t], zi[t], fxi[t], fyi[t], fzi[t], nz4i[t], xa, ya, za, fxa, fya,
                                   fza, nz4a, lx[t], ly[t], lz[t], ax, ay, az);
        pair_sym<real, WALL, NEAR>(C, xi[t], yi[t], zi[t], fxi[t], fyi[t], fzi[t], nz4i[t], xb, yb, zb, fxb, fyb,
                                   fzb, nz4b, lx[t], ly[t], lz[t], bx, by, bz);
      }
      real* __restrict__ wa = wr + (size_t)(jj * 3) * STR;
      real* __restrict__ wb = wr + (size_t)((jj + RC / 2) * 3) * STR;
      wa[0] = ax; wa[STR] = ay; wa[2 * STR] = az;
      wb[0] = bx; wb[STR] = by; wb[2 * STR] = bz;
    }
    __syncwarp();
    real sx = row_sum<RC>(rd), sy = row_sum<RC>(rd + STR), sz = row_sum<RC>(rd + 2 * STR);
#pragma unroll
    for (int o = RC; o < 32; o <<= 1) {  // the 32 / RC lanes of a source
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sy += __shfl_xor_sync(0xffffffffu, sy, o);
      sz += __shfl_xor_sync(0xffffffffu, sz, o);
    }
    if (part == 0) {
      real* o = raw_tile + (j0 + src);
      atomicAdd(o, sx);
      atomicAdd(o + raw_ld, sy);
      atomicAdd(o + 2 * raw_ld, sz);
    }
    __syncwarp();
  }
}

constexpr int kSymChunks = kSrcTile / 32;  // 32-source chunks per tile unit

// first unit index of row I of the triangle: rows have n_src_tiles - I*diag units
__device__ __host__ __forceinline__ long long sym_row_offset(long long I, int ns, int diag) {
  return I * ns - (long long)diag * (I * (I - 1) / 2);
}

// Reaction sums go through the warp-private shared-memory tile (tile_compute_symt, RC sources per
// chunk).  Target sums are flushed into the global
// accumulators (RED.ADD) after every tile unit -- that IS the second summation level of the fp32
// path, and it frees the 3T registers a running sum per target would hold.
// source tile of the unit at position `pos` of row I: J = I D (the diagonal units), I D + 1, ..., ns - 1.
// (Walking odd rows from the far end, so that CTAs on neighbouring rows do not add reaction sums to the
// same accumulators in lockstep, was measured: no effect beyond run-to-run noise, profiles/r02_part_balance.md.)
__device__ __host__ __forceinline__ int sym_tile_of(int I, int pos, int ns, int D) {
  (void)ns;
  return I * D + pos;
}

void sym_cost_bounds(SymPlan* plan, int part, int n_parts, double w_diag, long long* bounds) {
  const int ns = plan->n_src_tiles, D = plan->diag, ntt = plan->n_tgt_tiles, grid = plan->grid;
  if (!(w_diag > 0.0)) w_diag = 1.0;
  auto row_units = [&](int I, int* nd) {
    const int len = ns - I * D;
    *nd = len < D ? len : D;
    return len;
  };
  double total = 0.0;
  for (int I = 0; I < ntt; ++I) {
    int nd;
    const int len = row_units(I, &nd);
    total += kSymChunks * (nd * w_diag + (len - nd));
  }
  const long long pieces = (long long)n_parts * grid;
  int I = 0;
  double before = 0.0;  // cost of the rows above I
  for (int c = 0; c <= grid; ++c) {
    const long long q = (long long)part * grid + c;
    long long f;
    if (q >= pieces) {
      f = plan->units * kSymChunks;
    } else {
      const double target = total * ((double)q / (double)pieces);
      for (;;) {
        int nd;
        const int len = row_units(I, &nd);
        const double rc = kSymChunks * (nd * w_diag + (len - nd));
        if (I + 1 < ntt && before + rc <= target) { before += rc; ++I; } else break;
      }
      int nd;
      const int len = row_units(I, &nd);
      const double rem = target - before;
      const double diag_cost = kSymChunks * nd * w_diag;
      long long in_row = rem < diag_cost ? (long long)(rem / w_diag) : (long long)kSymChunks * nd + (long long)(rem - diag_cost);
      const long long row_chunks = (long long)kSymChunks * len;
      if (in_row > row_chunks) in_row = row_chunks;
      if (in_row < 0) in_row = 0;
      f = sym_row_offset(I, ns, D) * kSymChunks + in_row;
    }
    bounds[c] = f;
    if (c > 0 && bounds[c] < bounds[c - 1]) bounds[c] = bounds[c - 1];
  }
  plan->u0 = bounds[0];
  plan->u1 = bounds[grid];
}

template <typename real, int T, int NT, int RC>
__host__ __device__ constexpr size_t sym_smem_bytes() {
  return (2 * (size_t)kSrcTile * kRecReals + (size_t)(NT / 32) * red_tile_reals<real, RC>()) * sizeof(real);
}

template <typename real, bool WALL, int T, int NT, int RC>
__global__ void __launch_bounds__(NT) rpy_matvec_sym_kernel(const SymArgs<real> A) {
  constexpr int TT = T * NT;
  constexpr uint32_t kTileBytes = kSrcTile * kRecReals * sizeof(real);
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  real* const sbuf0 = reinterpret_cast<real*>(smem_dyn);
  real* const sbuf1 = sbuf0 + kSrcTile * kRecReals;
  real* const red = sbuf1 + kSrcTile * kRecReals + (size_t)(threadIdx.x >> 5) * red_tile_reals<real, RC>();
  __shared__ __align__(8) unsigned long long mbar[2];

  const int tid = threadIdx.x;
  const int ns = A.plan.n_src_tiles, D = A.plan.diag, ntt = A.plan.n_tgt_tiles;
  // work is cut at the granularity of 32-source chunks (kSymChunks per tile unit) so that the
  // shares of different CTAs -- and of different GPUs -- differ by at most one chunk
  const long long span = A.plan.u1 - A.plan.u0;
  const long long f0 = A.bounds ? A.bounds[blockIdx.x] : A.plan.u0 + span * blockIdx.x / gridDim.x;
  const long long f1 = A.bounds ? A.bounds[blockIdx.x + 1] : A.plan.u0 + span * (blockIdx.x + 1) / gridDim.x;
  if (f0 >= f1) return;
  const long long g0 = f0 / kSymChunks, g_last = (f1 - 1) / kSymChunks, g1 = g_last + 1;
  const int jb_first = (int)(f0 - g0 * kSymChunks) * 32;
  const int je_last = (int)(f1 - 1 - g_last * kSymChunks + 1) * 32;

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  // locate the row of g0: largest I with offset(I) <= g0
  int I;
  {
    int lo = 0, hi = ntt - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (sym_row_offset(mid, ns, D) <= g0) lo = mid; else hi = mid - 1;
    }
    I = lo;
  }
  // position of the first unit inside its row
  int pos = (int)(g0 - sym_row_offset(I, ns, D));
  int J = sym_tile_of(I, pos, ns, D);
  if (tid == 0) {
    mbar_expect_tx(&mbar[0], kTileBytes);
    tma_load_1d(sbuf0, A.rec + (size_t)J * kSrcTile * kRecReals, kTileBytes, &mbar[0]);
  }

  real xi[T], yi[T], zi[T], fxi[T], fyi[T], fzi[T], nz4i[T];
  bool fresh = true;
  const double near2 = (double)A.C.four_a2 * (1.0 + 1e-6);
  const size_t raw_ld = (size_t)ns * kSrcTile;  // accumulators: x[raw_ld] y[raw_ld] z[raw_ld]

  for (long long g = g0; g < g1; ++g) {
    const int it = (int)(g - g0);
    const int buf = it & 1;
    const uint32_t parity = (uint32_t)(it >> 1) & 1u;
    const bool row_end = (pos + 1 == ns - I * D);
    const int jb = (g == g0) ? jb_first : 0;
    const int je = (g == g_last) ? je_last : kSrcTile;
    const real* cur = buf ? sbuf1 : sbuf0;
    real* nxt = buf ? sbuf0 : sbuf1;

    if (tid == 0 && g + 1 < g1) {
      const int nJ = row_end ? sym_tile_of(I + 1, 0, ns, D) : sym_tile_of(I, pos + 1, ns, D);
      mbar_expect_tx(&mbar[buf ^ 1], kTileBytes);
      tma_load_1d(nxt, A.rec + (size_t)nJ * kSrcTile * kRecReals, kTileBytes, &mbar[buf ^ 1]);
    }

    if (fresh) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        int li = I * TT + tid + t * NT;
        const bool pad = li >= A.plan.n;
        if (pad) li = A.plan.n - 1;
        load_full_rec(A.rec + (size_t)li * kRecReals, xi[t], yi[t], zi[t], fxi[t], fyi[t], fzi[t], nz4i[t]);
        if (pad) fxi[t] = fyi[t] = fzi[t] = (real)0;  // padding lanes exert nothing
      }
      fresh = false;
    }

    const bool far = box_gap2(A.box_tgt + 6 * (size_t)I, A.box_src + 6 * (size_t)J) > near2;
    const bool diagonal = J < (I + 1) * D;  // source tile lies inside this target tile: ordered

    real ux[T], uy[T], uz[T];
#pragma unroll
    for (int t = 0; t < T; ++t) ux[t] = uy[t] = uz[t] = (real)0;

    mbar_wait(&mbar[buf], parity);
    if (diagonal) {
      tile_compute<WALL, true, T>(cur, A.C, xi, yi, zi, ux, uy, uz, jb, je);
    } else {
      real* raw_tile = A.raw + (size_t)J * kSrcTile;
      if (far)
        tile_compute_symt<real, WALL, false, T, RC>(cur, A.C, xi, yi, zi, fxi, fyi, fzi, nz4i, ux, uy, uz, raw_tile, raw_ld, red, jb, je);
      else
        tile_compute_symt<real, WALL, true, T, RC>(cur, A.C, xi, yi, zi, fxi, fyi, fzi, nz4i, ux, uy, uz, raw_tile, raw_ld, red, jb, je);
    }
    // flush this unit's target sums (per-tile accumulators: the first summation level)
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int li = I * TT + tid + t * NT;
      if (li < A.plan.n) {  // component-major accumulators: a warp adds 32 consecutive words per instruction
        atomicAdd(A.raw + li, ux[t]);
        atomicAdd(A.raw + raw_ld + li, uy[t]);
        atomicAdd(A.raw + 2 * raw_ld + li, uz[t]);
      }
    }
    __syncthreads();  // every thread is done with `cur` before it is refilled

    if (row_end) {
      ++I;
      pos = 0;
      fresh = true;
    } else {
      ++pos;
    }
    J = sym_tile_of(I, pos, ns, D);
  }
}

template <typename real, bool WALL>
__global__ void rpy_sym_scale_kernel(const SymArgs<real> A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.plan.n) return;
  real sc = A.C.out_scale;
  if (WALL) sc *= damp(A.rec[(size_t)i * kRecReals + 2], A.C.a, A.C.inv_a);
  const size_t ld = (size_t)A.plan.n_src_tiles * kSrcTile;
  A.out[3 * (size_t)i + 0] = A.raw[i] * sc;
  A.out[3 * (size_t)i + 1] = A.raw[ld + i] * sc;
  A.out[3 * (size_t)i + 2] = A.raw[2 * ld + i] * sc;
}

// (T targets per thread, NT threads per CTA, RC sources per reaction-reduction chunk)
#define RBL_F32_SYM_VARIANTS(X) \
  X(6, 256, 16) X(4, 128, 16) X(4, 256, 16) X(5, 256, 16) X(6, 256, 8) X(2, 256, 8) X(1, 256, 8)
#define RBL_F64_SYM_VARIANTS(X) \
  X(4, 256, 16) X(3, 256, 16) X(4, 128, 16) X(3, 256, 8) X(2, 128, 8) X(1, 256, 8)

template <>
MatvecVariant matvec_sym_variant<float>(int idx) {
  static const MatvecVariant v[] = {
#define X(T, NT, RC) {T, NT, RC},
      RBL_F32_SYM_VARIANTS(X)
#undef X
  };
  return v[idx];
}
template <>
MatvecVariant matvec_sym_variant<double>(int idx) {
  static const MatvecVariant v[] = {
#define X(T, NT, RC) {T, NT, RC},
      RBL_F64_SYM_VARIANTS(X)
#undef X
  };
  return v[idx];
}
template <>
int matvec_sym_num_variants<float>() {
  int n = 0;
#define X(T, NT, RC) ++n;
  RBL_F32_SYM_VARIANTS(X)
#undef X
  return n;
}
template <>
int matvec_sym_num_variants<double>() {
  int n = 0;
#define X(T, NT, RC) ++n;
  RBL_F64_SYM_VARIANTS(X)
#undef X
  return n;
}

// defaults measured on B200 at 162 000 blobs (profiles/r02_probe_variants_cfg2.jsonl, profiles/r02_ncu_variants.md): fp32 (6,256,16) with and
// without the wall; fp64 (4,256,16) with the wall, (3,256,16) in free space; small problems take the
// smallest target tile (the last variant) so that the unit triangle still covers the SMs
template <>
int matvec_sym_default_variant<float>(bool, int n) {
  return n < 16384 ? matvec_sym_num_variants<float>() - 1 : 0;
}
template <>
int matvec_sym_default_variant<double>(bool wall, int n) {
  return n < 16384 ? matvec_sym_num_variants<double>() - 1 : (wall ? 0 : 1);
}

template <typename real, bool WALL, int T, int NT, int RC>
static cudaError_t sym_occupancy_of(int* bps) {
  constexpr size_t smem = sym_smem_bytes<real, T, NT, RC>();
  cudaError_t e = cudaFuncSetAttribute(rpy_matvec_sym_kernel<real, WALL, T, NT, RC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, rpy_matvec_sym_kernel<real, WALL, T, NT, RC>, NT, smem);
}
template <typename real>
static cudaError_t sym_variant_occupancy(int variant, bool wall, int* bps);
template <>
cudaError_t sym_variant_occupancy<float>(int variant, bool wall, int* bps) {
  int k = 0;
#define X(T, NT, RC) \
  if (variant == k++) return wall ? sym_occupancy_of<float, true, T, NT, RC>(bps) : sym_occupancy_of<float, false, T, NT, RC>(bps);
  RBL_F32_SYM_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}
template <>
cudaError_t sym_variant_occupancy<double>(int variant, bool wall, int* bps) {
  int k = 0;
#define X(T, NT, RC) \
  if (variant == k++) return wall ? sym_occupancy_of<double, true, T, NT, RC>(bps) : sym_occupancy_of<double, false, T, NT, RC>(bps);
  RBL_F64_SYM_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

template <typename real>
cudaError_t matvec_sym_plan(int variant, bool wall, int n, int part, int n_parts, int sm_count, SymPlan* plan) {
  if (variant < 0 || variant >= matvec_sym_num_variants<real>() || n_parts < 1 || part < 0 || part >= n_parts)
    return cudaErrorInvalidValue;
  const MatvecVariant v = matvec_sym_variant<real>(variant);
  int bps = 0;
  cudaError_t e = sym_variant_occupancy<real>(variant, wall, &bps);
  if (e != cudaSuccess) return e;
  if (bps < 1) return cudaErrorLaunchOutOfResources;
  plan->n = n;
  plan->n_src_tiles = (n + kSrcTile - 1) / kSrcTile;
  plan->tgt_tile = v.T * v.threads;
  plan->diag = plan->tgt_tile / kSrcTile;
  plan->n_tgt_tiles = (n + plan->tgt_tile - 1) / plan->tgt_tile;
  plan->units = sym_row_offset(plan->n_tgt_tiles, plan->n_src_tiles, plan->diag);
  const long long fine = plan->units * kSymChunks;  // shares are cut at 32-source granularity
  plan->u0 = fine * part / n_parts;
  plan->u1 = fine * (part + 1) / n_parts;
  plan->grid = sm_count * bps;
  return cudaSuccess;
}

template <typename real, bool WALL, int T, int NT, int RC>
static cudaError_t sym_launch_one(const SymArgs<real>& a, cudaStream_t s, cudaEvent_t ev0, cudaEvent_t ev1) {
  static_assert((T * NT) % kSrcTile == 0, "target tile must be a multiple of the source tile");
  cudaError_t e = cudaMemsetAsync(a.raw, 0, 3 * (size_t)a.plan.n_src_tiles * kSrcTile * sizeof(real), s);
  if (e != cudaSuccess) return e;
  if (ev0) cudaEventRecord(ev0, s);
  if (a.plan.u1 > a.plan.u0)
    rpy_matvec_sym_kernel<real, WALL, T, NT, RC><<<a.plan.grid, NT, sym_smem_bytes<real, T, NT, RC>(), s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (ev1) cudaEventRecord(ev1, s);
  rpy_sym_scale_kernel<real, WALL><<<(a.plan.n + 255) / 256, 256, 0, s>>>(a);
  return cudaGetLastError();
}

template <>
cudaError_t matvec_sym_launch<float>(int variant, const SymArgs<float>& a, cudaStream_t s, cudaEvent_t ev0,
                                     cudaEvent_t ev1) {
  if (a.plan.n <= 0) return cudaSuccess;
  int k = 0;
#define X(T, NT, RC) \
  if (variant == k++) return a.wall ? sym_launch_one<float, true, T, NT, RC>(a, s, ev0, ev1) : sym_launch_one<float, false, T, NT, RC>(a, s, ev0, ev1);
  RBL_F32_SYM_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}
template <>
cudaError_t matvec_sym_launch<double>(int variant, const SymArgs<double>& a, cudaStream_t s, cudaEvent_t ev0,
                                      cudaEvent_t ev1) {
  if (a.plan.n <= 0) return cudaSuccess;
  int k = 0;
#define X(T, NT, RC) \
  if (variant == k++) return a.wall ? sym_launch_one<double, true, T, NT, RC>(a, s, ev0, ev1) : sym_launch_one<double, false, T, NT, RC>(a, s, ev0, ev1);
  RBL_F64_SYM_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}


// ----------------------------------------------------------------------------------
// symmetric kernel, TWO right-hand sides: U1 = B M B F1 and U2 = B M B F2 in one pass.
// The paired Lanczos of a BD step (M^{1/2}W_1 and M^{1/2}W_2, c_rigid_obj.cpp:930-935) applies
// M to two vectors per iteration; the geometry of a pair -- distances, both rsqrt, every wall
// polynomial -- is evaluated once for both (pair_symR).  Same organisation as
// rpy_matvec_sym_kernel: stream-K over the unordered tile triangle, TMA-staged source tiles,
// warp-transposed reaction sums, RED.ADD into zeroed accumulators.
//   record (12 reals): x y z f1x | f1y f1z f2x f2y | f2z -4z^2 0 0
// ----------------------------------------------------------------------------------
template <typename real>
struct Rec2 {
  real x, y, z, nz4;
  real f[2][3];
};
__device__ __forceinline__ void load_rec2(const float* p, Rec2<float>& r) {
  const float4* q = reinterpret_cast<const float4*>(p);
  const float4 a = q[0], b = q[1], c = q[2];
  r.x = a.x; r.y = a.y; r.z = a.z; r.f[0][0] = a.w;
  r.f[0][1] = b.x; r.f[0][2] = b.y; r.f[1][0] = b.z; r.f[1][1] = b.w;
  r.f[1][2] = c.x; r.nz4 = c.y;
}
__device__ __forceinline__ void load_rec2(const double* p, Rec2<double>& r) {
  const double2* q = reinterpret_cast<const double2*>(p);
  const double2 a = q[0], b = q[1], c = q[2], d = q[3], e = q[4];
  r.x = a.x; r.y = a.y; r.z = b.x; r.f[0][0] = b.y;
  r.f[0][1] = c.x; r.f[0][2] = c.y; r.f[1][0] = d.x; r.f[1][1] = d.y;
  r.f[1][2] = e.x; r.nz4 = e.y;
}

template <typename real>
__global__ void pack_records2_kernel(const real* __restrict__ r, const real* __restrict__ F1,
                                     const real* __restrict__ F2, int n, int n_padded, int wall, real a,
                                     real inv_a, real* __restrict__ rec, int* __restrict__ below) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_padded) return;
  int src = k < n ? k : n - 1;
  real x = r[3 * (size_t)src], y = r[3 * (size_t)src + 1], z = r[3 * (size_t)src + 2];
  real f[6] = {0, 0, 0, 0, 0, 0};
  if (k < n) {
    real b = wall ? damp(z, a, inv_a) : (real)1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      f[c] = b * F1[3 * (size_t)k + c];
      f[3 + c] = b * F2[3 * (size_t)k + c];
    }
    if (wall && z < (real)0) *below = 1;
  }
  real* p = rec + (size_t)k * kRec2Reals;
  p[0] = x; p[1] = y; p[2] = z;
#pragma unroll
  for (int c = 0; c < 6; ++c) p[3 + c] = f[c];
  p[9] = (real)-4 * z * z;
  p[10] = p[11] = (real)0;
}
template <typename real>
cudaError_t pack_records2(const real* r, const real* F1, const real* F2, int n, int n_padded, bool wall, real a,
                          real* rec, int* below, cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  int threads = 256, blocks = (n_padded + threads - 1) / threads;
  pack_records2_kernel<real><<<blocks, threads, 0, s>>>(r, F1, F2, n, n_padded, wall ? 1 : 0, a, (real)1 / a, rec, below);
  return cudaGetLastError();
}

// positions unchanged since pack_records2: only the six force words of each record
template <typename real>
__global__ void repack_forces2_kernel(const real* __restrict__ F1, const real* __restrict__ F2, int n, int wall,
                                      real a, real inv_a, real* __restrict__ rec, int* __restrict__ below) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  real* p = rec + (size_t)k * kRec2Reals;
  const real z = p[2];
  const real b = wall ? damp(z, a, inv_a) : (real)1;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    p[3 + c] = b * F1[3 * (size_t)k + c];
    p[6 + c] = b * F2[3 * (size_t)k + c];
  }
  if (wall && z < (real)0) *below = 1;
}
template <typename real>
cudaError_t repack_forces2(const real* F1, const real* F2, int n, bool wall, real a, real* rec, int* below,
                           cudaStream_t s) {
  if (n <= 0) return cudaSuccess;
  int threads = 256, blocks = (n + threads - 1) / threads;
  repack_forces2_kernel<real><<<blocks, threads, 0, s>>>(F1, F2, n, wall ? 1 : 0, a, (real)1 / a, rec, below);
  return cudaGetLastError();
}

// diagonal tiles (they hold the self pairs): ordered general path, one right-hand side at a time
template <typename real, bool WALL, int T>
__device__ __forceinline__ void tile_compute2_ordered(const real* __restrict__ sb, const PairConsts<real>& C,
                                                      const real (&xi)[T], const real (&yi)[T], const real (&zi)[T],
                                                      real (&u)[T][2][3], int jb, int je) {
#pragma unroll 1
  for (int j = jb; j < je; ++j) {
    Rec2<real> s;
    load_rec2(sb + (size_t)j * kRec2Reals, s);
    const real z2 = (real)2 * s.z, zz4 = -s.nz4;
#pragma unroll
    for (int t = 0; t < T; ++t) {
#pragma unroll
      for (int k = 0; k < 2; ++k)
        pair<real, WALL, true>(C, xi[t], yi[t], zi[t], s.x, s.y, s.z, s.f[k][0], s.f[k][1], s.f[k][2], z2, zz4,
                               u[t][k][0], u[t][k][1], u[t][k][2]);
    }
  }
}

// Two right-hand sides through the warp-private reduction tile (see tile_compute_symt): 6 rows per
// source (2 right-hand sides x 3 components), rows ordered component-major, row = (k*3 + c) * RC +
// source, so that consecutive sources are one row stride apart (36 words = 4 mod 32 in fp32, 68
// words = 4 mod 32 in fp64: the transposed 128-bit reads of a quarter-warp hit distinct bank groups).
template <typename real, int RC>
__host__ __device__ constexpr size_t red2_tile_reals() {
  return RC > 0 ? (size_t)RC * 6 * RedLayout<real>::kStride : 0;
}

template <typename real, bool WALL, bool NEAR, int T, int RC>
__device__ __forceinline__ void tile_compute_sym2t(const real* __restrict__ sb, const PairConsts<real>& C,
                                                   const real (&xi)[T], const real (&yi)[T], const real (&zi)[T],
                                                   const real (&nz4i)[T], const real (&fi)[T][2][3],
                                                   real (&u)[T][2][3], real* __restrict__ raw1,
                                                   real* __restrict__ raw2, size_t raw_ld, real* __restrict__ red,
                                                   int jb, int je) {
  static_assert(RC == 8 || RC == 16 || RC == 32, "reduction chunk");
  constexpr int STR = RedLayout<real>::kStride;
  const int lane = threadIdx.x & 31;
  const int src = lane % RC, part = lane / RC;
  real* __restrict__ wr = red + lane;
  const real* __restrict__ rd = red + (size_t)src * STR + part * RC;
  for (int j0 = jb; j0 < je; j0 += RC) {
#pragma unroll 1
    for (int jj = 0; jj < RC / 2; ++jj) {
      Rec2<real> sa, sb2;
      load_rec2(sb + (size_t)(j0 + jj) * kRec2Reals, sa);
      load_rec2(sb + (size_t)(j0 + jj + RC / 2) * kRec2Reals, sb2);
      real ra[2][3] = {{0, 0, 0}, {0, 0, 0}}, rb[2][3] = {{0, 0, 0}, {0, 0, 0}};
#pragma unroll
      for (int t = 0; t < T; ++t) {
        pair_symR<real, WALL, NEAR, 2>(C, xi[t], yi[t], zi[t], fi[t], nz4i[t], sa.x, sa.y, sa.z, sa.f, sa.nz4, u[t], ra);
        pair_symR<real, WALL, NEAR, 2>(C, xi[t], yi[t], zi[t], fi[t], nz4i[t], sb2.x, sb2.y, sb2.z, sb2.f, sb2.nz4, u[t], rb);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          wr[(size_t)((k * 3 + c) * RC + jj) * STR] = ra[k][c];
          wr[(size_t)((k * 3 + c) * RC + jj + RC / 2) * STR] = rb[k][c];
        }
    }
    __syncwarp();
    real s6[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) s6[q] = row_sum<RC>(rd + (size_t)(q * RC) * STR);
#pragma unroll
    for (int o = RC; o < 32; o <<= 1)
#pragma unroll
      for (int q = 0; q < 6; ++q) s6[q] += __shfl_xor_sync(0xffffffffu, s6[q], o);
    if (part == 0) {
      real* o1 = raw1 + (j0 + src);
      real* o2 = raw2 + (j0 + src);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        atomicAdd(o1 + c * raw_ld, s6[c]);
        atomicAdd(o2 + c * raw_ld, s6[3 + c]);
      }
    }
    __syncwarp();
  }
}

template <typename real, int T, int NT, int RC>
__host__ __device__ constexpr size_t sym2_smem_bytes() {
  return (2 * (size_t)kSrcTile * kRec2Reals + (size_t)(NT / 32) * red2_tile_reals<real, RC>()) * sizeof(real);
}

template <typename real, bool WALL, int T, int NT, int RC>
__global__ void __launch_bounds__(NT) rpy_matvec_sym2_kernel(const Sym2Args<real> A) {
  constexpr int TT = T * NT;
  constexpr uint32_t kTileBytes = kSrcTile * kRec2Reals * sizeof(real);
  extern __shared__ __align__(128) unsigned char smem_dyn[];
  real* sbuf0 = reinterpret_cast<real*>(smem_dyn);
  real* sbuf1 = sbuf0 + kSrcTile * kRec2Reals;
  real* const red = sbuf1 + kSrcTile * kRec2Reals + (size_t)(threadIdx.x >> 5) * red2_tile_reals<real, RC>();
  __shared__ __align__(8) unsigned long long mbar[2];

  const int tid = threadIdx.x;
  const int ns = A.plan.n_src_tiles, D = A.plan.diag, ntt = A.plan.n_tgt_tiles;
  const long long span = A.plan.u1 - A.plan.u0;
  const long long f0 = A.bounds ? A.bounds[blockIdx.x] : A.plan.u0 + span * blockIdx.x / gridDim.x;
  const long long f1 = A.bounds ? A.bounds[blockIdx.x + 1] : A.plan.u0 + span * (blockIdx.x + 1) / gridDim.x;
  if (f0 >= f1) return;
  const long long g0 = f0 / kSymChunks, g_last = (f1 - 1) / kSymChunks, g1 = g_last + 1;
  const int jb_first = (int)(f0 - g0 * kSymChunks) * 32;
  const int je_last = (int)(f1 - 1 - g_last * kSymChunks + 1) * 32;

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  int I;
  {
    int lo = 0, hi = ntt - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (sym_row_offset(mid, ns, D) <= g0) lo = mid; else hi = mid - 1;
    }
    I = lo;
  }
  int pos = (int)(g0 - sym_row_offset(I, ns, D));
  int J = sym_tile_of(I, pos, ns, D);
  if (tid == 0) {
    mbar_expect_tx(&mbar[0], kTileBytes);
    tma_load_1d(sbuf0, A.rec + (size_t)J * kSrcTile * kRec2Reals, kTileBytes, &mbar[0]);
  }

  real xi[T], yi[T], zi[T], nz4i[T], fi[T][2][3];
  bool fresh = true;
  const double near2 = (double)A.C.four_a2 * (1.0 + 1e-6);
  const size_t raw_ld = (size_t)ns * kSrcTile;  // accumulators: [right-hand side][x y z][raw_ld]

  for (long long g = g0; g < g1; ++g) {
    const int it = (int)(g - g0);
    const int buf = it & 1;
    const uint32_t parity = (uint32_t)(it >> 1) & 1u;
    const bool row_end = (pos + 1 == ns - I * D);
    const int jb = (g == g0) ? jb_first : 0;
    const int je = (g == g_last) ? je_last : kSrcTile;
    real* cur = buf ? sbuf1 : sbuf0;
    real* nxt = buf ? sbuf0 : sbuf1;

    if (tid == 0 && g + 1 < g1) {
      const int nJ = row_end ? sym_tile_of(I + 1, 0, ns, D) : sym_tile_of(I, pos + 1, ns, D);
      mbar_expect_tx(&mbar[buf ^ 1], kTileBytes);
      tma_load_1d(nxt, A.rec + (size_t)nJ * kSrcTile * kRec2Reals, kTileBytes, &mbar[buf ^ 1]);
    }

    if (fresh) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        int li = I * TT + tid + t * NT;
        const bool pad = li >= A.plan.n;
        if (pad) li = A.plan.n - 1;
        Rec2<real> me;
        load_rec2(A.rec + (size_t)li * kRec2Reals, me);
        xi[t] = me.x; yi[t] = me.y; zi[t] = me.z; nz4i[t] = me.nz4;
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int c = 0; c < 3; ++c) fi[t][k][c] = pad ? (real)0 : me.f[k][c];
      }
      fresh = false;
    }

    const bool far = box_gap2(A.box_tgt + 6 * (size_t)I, A.box_src + 6 * (size_t)J) > near2;
    const bool diagonal = J < (I + 1) * D;

    real u[T][2][3];
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int k = 0; k < 2; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) u[t][k][c] = (real)0;

    mbar_wait(&mbar[buf], parity);
    if (diagonal) {
      tile_compute2_ordered<real, WALL, T>(cur, A.C, xi, yi, zi, u, jb, je);
    } else {
      real* raw1 = A.raw + (size_t)J * kSrcTile;  // [right-hand side][component][blob]
      real* raw2 = raw1 + 3 * raw_ld;
      if (far)
        tile_compute_sym2t<real, WALL, false, T, RC>(cur, A.C, xi, yi, zi, nz4i, fi, u, raw1, raw2, raw_ld, red, jb, je);
      else
        tile_compute_sym2t<real, WALL, true, T, RC>(cur, A.C, xi, yi, zi, nz4i, fi, u, raw1, raw2, raw_ld, red, jb, je);
    }
    // flush this unit's target sums (the per-tile accumulators are the first summation level)
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int li = I * TT + tid + t * NT;
      if (li < A.plan.n) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
          for (int c = 0; c < 3; ++c) atomicAdd(A.raw + (size_t)(3 * k + c) * raw_ld + li, u[t][k][c]);
      }
    }
    __syncthreads();

    if (row_end) {
      ++I;
      pos = 0;
      fresh = true;
    } else {
      ++pos;
    }
    J = sym_tile_of(I, pos, ns, D);
  }
}

template <typename real, bool WALL>
__global__ void rpy_sym2_scale_kernel(const Sym2Args<real> A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.plan.n) return;
  real sc = A.C.out_scale;
  if (WALL) sc *= damp(A.rec[(size_t)i * kRec2Reals + 2], A.C.a, A.C.inv_a);
  const size_t raw_ld = (size_t)A.plan.n_src_tiles * kSrcTile;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    A.out1[3 * (size_t)i + c] = A.raw[(size_t)c * raw_ld + i] * sc;
    A.out2[3 * (size_t)i + c] = A.raw[(size_t)(3 + c) * raw_ld + i] * sc;
  }
}

#define RBL_F32_SYM2_VARIANTS(X) X(3, 256, 8) X(4, 128, 8) X(4, 256, 8) X(2, 256, 8) X(1, 256, 8)
#define RBL_F64_SYM2_VARIANTS(X) X(3, 256, 8) X(2, 256, 8) X(2, 128, 8) X(1, 256, 8)

template <>
int matvec_sym2_num_variants<float>() {
  int n = 0;
#define X(T, NT, RC) ++n;
  RBL_F32_SYM2_VARIANTS(X)
#undef X
  return n;
}
template <>
int matvec_sym2_num_variants<double>() {
  int n = 0;
#define X(T, NT, RC) ++n;
  RBL_F64_SYM2_VARIANTS(X)
#undef X
  return n;
}
template <>
MatvecVariant matvec_sym2_variant<float>(int idx) {
  static const MatvecVariant v[] = {
#define X(T, NT, RC) {T, NT, RC},
      RBL_F32_SYM2_VARIANTS(X)
#undef X
  };
  return v[idx];
}
template <>
MatvecVariant matvec_sym2_variant<double>(int idx) {
  static const MatvecVariant v[] = {
#define X(T, NT, RC) {T, NT, RC},
      RBL_F64_SYM2_VARIANTS(X)
#undef X
  };
  return v[idx];
}
// index 0 = the default measured on B200 at 172 032 blobs with the wall (profiles/r02_sym2_sweep_cfg3.jsonl:
// fp32 (3,256,8) 70.6 ms against 96.7 ms for two single passes, fp64 (3,256,8) 136.6 against 179.5 ms); small
// problems take the smallest target tile (the last variant)
template <>
int matvec_sym2_default_variant<float>(bool, int n) { return n < 16384 ? matvec_sym2_num_variants<float>() - 1 : 0; }
template <>
int matvec_sym2_default_variant<double>(bool, int n) { return n < 16384 ? matvec_sym2_num_variants<double>() - 1 : 0; }

template <typename real, bool WALL, int T, int NT, int RC>
static cudaError_t sym2_occupancy_of(int* bps) {
  constexpr size_t smem = sym2_smem_bytes<real, T, NT, RC>();
  cudaError_t e = cudaFuncSetAttribute(rpy_matvec_sym2_kernel<real, WALL, T, NT, RC>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(bps, rpy_matvec_sym2_kernel<real, WALL, T, NT, RC>, NT, smem);
}
template <typename real>
static cudaError_t sym2_variant_occupancy(int variant, bool wall, int* bps);
template <>
cudaError_t sym2_variant_occupancy<float>(int variant, bool wall, int* bps) {
  int k = 0;
#define X(T, NT, RC) \
  if (variant == k++) return wall ? sym2_occupancy_of<float, true, T, NT, RC>(bps) : sym2_occupancy_of<float, false, T, NT, RC>(bps);
  RBL_F32_SYM2_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}
template <>
cudaError_t sym2_variant_occupancy<double>(int variant, bool wall, int* bps) {
  int k = 0;
#define X(T, NT, RC) \
  if (variant == k++) return wall ? sym2_occupancy_of<double, true, T, NT, RC>(bps) : sym2_occupancy_of<double, false, T, NT, RC>(bps);
  RBL_F64_SYM2_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

template <typename real>
cudaError_t matvec_sym2_plan(int variant, bool wall, int n, int part, int n_parts, int sm_count, SymPlan* plan) {
  if (variant < 0 || variant >= matvec_sym2_num_variants<real>() || n_parts < 1 || part < 0 || part >= n_parts)
    return cudaErrorInvalidValue;
  const MatvecVariant v = matvec_sym2_variant<real>(variant);
  int bps = 0;
  cudaError_t e = sym2_variant_occupancy<real>(variant, wall, &bps);
  if (e != cudaSuccess) return e;
  if (bps < 1) return cudaErrorLaunchOutOfResources;
  plan->n = n;
  plan->n_src_tiles = (n + kSrcTile - 1) / kSrcTile;
  plan->tgt_tile = v.T * v.threads;
  plan->diag = plan->tgt_tile / kSrcTile;
  plan->n_tgt_tiles = (n + plan->tgt_tile - 1) / plan->tgt_tile;
  plan->units = sym_row_offset(plan->n_tgt_tiles, plan->n_src_tiles, plan->diag);
  const long long fine = plan->units * kSymChunks;
  plan->u0 = fine * part / n_parts;
  plan->u1 = fine * (part + 1) / n_parts;
  plan->grid = sm_count * bps;
  return cudaSuccess;
}

template <typename real, bool WALL, int T, int NT, int RC>
static cudaError_t sym2_launch_one(const Sym2Args<real>& a, cudaStream_t s, cudaEvent_t ev0, cudaEvent_t ev1) {
  static_assert((T * NT) % kSrcTile == 0, "target tile must be a multiple of the source tile");
  cudaError_t e = cudaMemsetAsync(a.raw, 0, 2 * 3 * (size_t)a.plan.n_src_tiles * kSrcTile * sizeof(real), s);
  if (e != cudaSuccess) return e;
  if (ev0) cudaEventRecord(ev0, s);
  if (a.plan.u1 > a.plan.u0)
    rpy_matvec_sym2_kernel<real, WALL, T, NT, RC><<<a.plan.grid, NT, sym2_smem_bytes<real, T, NT, RC>(), s>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (ev1) cudaEventRecord(ev1, s);
  rpy_sym2_scale_kernel<real, WALL><<<(a.plan.n + 255) / 256, 256, 0, s>>>(a);
  return cudaGetLastError();
}

template <>
cudaError_t matvec_sym2_launch<float>(int variant, const Sym2Args<float>& a, cudaStream_t s, cudaEvent_t ev0,
                                      cudaEvent_t ev1) {
  if (a.plan.n <= 0) return cudaSuccess;
  int k = 0;
#define X(T, NT, RC) \
  if (variant == k++) return a.wall ? sym2_launch_one<float, true, T, NT, RC>(a, s, ev0, ev1) : sym2_launch_one<float, false, T, NT, RC>(a, s, ev0, ev1);
  RBL_F32_SYM2_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}
template <>
cudaError_t matvec_sym2_launch<double>(int variant, const Sym2Args<double>& a, cudaStream_t s, cudaEvent_t ev0,
                                       cudaEvent_t ev1) {
  if (a.plan.n <= 0) return cudaSuccess;
  int k = 0;
#define X(T, NT, RC) \
  if (variant == k++) return a.wall ? sym2_launch_one<double, true, T, NT, RC>(a, s, ev0, ev1) : sym2_launch_one<double, false, T, NT, RC>(a, s, ev0, ev1);
  RBL_F64_SYM2_VARIANTS(X)
#undef X
  return cudaErrorInvalidValue;
}

// ----------------------------------------------------------------------------------
// FMA-pipe peak microbenchmark: 16 independent FMA chains per thread, register operands
// ----------------------------------------------------------------------------------
template <typename real>
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, real a, real b,
                                                       real* __restrict__ sink) {
  real x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = (real)(threadIdx.x + k) * (real)1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int k = 0; k < 16; ++k) x[k] = fma(x[k], a, b);
    }
  }
  real s = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) s += x[k];
  if (s == (real)123.456) sink[0] = s;  // never true; keeps the chains alive
}

template <typename real>
cudaError_t fma_peak_launch(int sm_count, int iters, real* sink, double* flops,
                            cudaStream_t s) {
  const int blocks = sm_count * 8, threads = 256;
  fma_peak_kernel<real><<<blocks, threads, 0, s>>>(iters, (real)0.999, (real)1e-3, sink);
  *flops = 2.0 * 16 * 8 * (double)iters * threads * blocks;
  return cudaGetLastError();
}

// explicit instantiations
#define INST(real)                                                                         \
  template cudaError_t matvec_plan<real>(int, bool, int, int, int, int, MatvecPlan*);      \
  template cudaError_t pack_records<real>(const real*, const real*, int, int, bool, real,  \
                                          real*, int*, cudaStream_t);                      \
  template cudaError_t repack_forces<real>(const real*, int, bool, real, real*, int*,      \
                                           cudaStream_t);                                  \
  template cudaError_t repack_forces2<real>(const real*, const real*, int, bool, real, real*, int*, cudaStream_t); \
  template cudaError_t tile_boxes<real>(const real*, int, int, int, float*, cudaStream_t, int); \
  template cudaError_t pack_records2<real>(const real*, const real*, const real*, int, int, bool, real, real*, int*, cudaStream_t); \
  template cudaError_t matvec_sym2_plan<real>(int, bool, int, int, int, int, SymPlan*);      \
  template cudaError_t matvec_sym_plan<real>(int, bool, int, int, int, int, SymPlan*);           \
  template cudaError_t fma_peak_launch<real>(int, int, real*, double*, cudaStream_t);
INST(float)
INST(double)
#undef INST

}  // namespace rbl
