// rbl_krylov.cu -- see rbl_krylov.cuh
#include <cstdint>

#include "rbl_krylov.cuh"

namespace rbl {

template <typename real>
__global__ void __launch_bounds__(256) multi_dot_stage1(const real* __restrict__ V, size_t ld,
                                                        int m, const real* __restrict__ w,
                                                        size_t n, real* __restrict__ partial) {
  __shared__ real red[8];
  for (int i = 0; i < m; ++i) {
    const real* v = V + (size_t)i * ld;
    real acc = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (size_t)gridDim.x * blockDim.x)
      acc += v[k] * w[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      real s = 0;
      for (int q = 0; q < 8; ++q) s += red[q];
      partial[(size_t)i * gridDim.x + blockIdx.x] = s;
    }
    __syncthreads();
  }
}
template <typename real>
__global__ void multi_dot_stage2(const real* __restrict__ partial, int nblocks, int m,
                                 real* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  real s = 0;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)i * nblocks + b];
  out[i] = s;
}
template <typename real>
cudaError_t multi_dot(const real* V, size_t ld, int m, const real* w, size_t n, real* partial,
                      real* out, cudaStream_t s) {
  if (m <= 0) return cudaSuccess;
  multi_dot_stage1<real><<<kDotBlocks, 256, 0, s>>>(V, ld, m, w, n, partial);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  multi_dot_stage2<real><<<(m + 63) / 64, 64, 0, s>>>(partial, kDotBlocks, m, out);
  return cudaGetLastError();
}

template <typename real>
__global__ void multi_axpy_kernel(const real* __restrict__ V, size_t ld, int m,
                                  const real* __restrict__ coef, real sign,
                                  real* __restrict__ w, size_t n) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n;
       k += (size_t)gridDim.x * blockDim.x) {
    real acc = 0;
    for (int i = 0; i < m; ++i) acc += coef[i] * V[(size_t)i * ld + k];
    w[k] += sign * acc;
  }
}
template <typename real>
cudaError_t multi_axpy(const real* V, size_t ld, int m, const real* coef, real sign, real* w,
                       size_t n, cudaStream_t s) {
  if (m <= 0 || n == 0) return cudaSuccess;
  multi_axpy_kernel<real><<<592, 256, 0, s>>>(V, ld, m, coef, sign, w, n);
  return cudaGetLastError();
}

template <typename real>
__global__ void scale_copy_kernel(const real* __restrict__ x, real a, real* y, size_t n,
                                  int accumulate) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n;
       k += (size_t)gridDim.x * blockDim.x)
    y[k] = accumulate ? y[k] + a * x[k] : a * x[k];
}
template <typename real>
cudaError_t scale_copy(const real* x, real a, real* y, size_t n, bool accumulate,
                       cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  scale_copy_kernel<real><<<592, 256, 0, s>>>(x, a, y, n, accumulate ? 1 : 0);
  return cudaGetLastError();
}

template <typename real>
__global__ void scale_by_inv_sqrt_kernel(const real* __restrict__ x, const real* __restrict__ s2, real* y, size_t n) {
  const real v = *s2;
  const real a = v > (real)0 ? (real)1 / sqrt(v) : (real)0;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x)
    y[k] = a * x[k];
}
template <typename real>
cudaError_t scale_by_inv_sqrt(const real* x, const real* s2, real* y, size_t n, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  scale_by_inv_sqrt_kernel<real><<<592, 256, 0, s>>>(x, s2, y, n);
  return cudaGetLastError();
}

template <typename From, typename To>
__global__ void cast_scale_kernel(const From* __restrict__ x, double a, To* y, size_t n, int accumulate) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    const double v = a * (double)x[k];
    y[k] = accumulate ? (To)((double)y[k] + v) : (To)v;
  }
}
template <typename From, typename To>
cudaError_t cast_scale(const From* x, double a, To* y, size_t n, bool accumulate, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  cast_scale_kernel<From, To><<<592, 256, 0, s>>>(x, a, y, n, accumulate ? 1 : 0);
  return cudaGetLastError();
}
template cudaError_t cast_scale<double, float>(const double*, double, float*, size_t, bool, cudaStream_t);
template cudaError_t cast_scale<float, double>(const float*, double, double*, size_t, bool, cudaStream_t);

template <typename real>
__global__ void flip_tail_kernel(const real* __restrict__ x, size_t n_head, size_t n, real* y) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n;
       k += (size_t)gridDim.x * blockDim.x)
    y[k] = k < n_head ? x[k] : -x[k];
}
template <typename real>
cudaError_t flip_tail(const real* x, size_t n_head, size_t n, real* y, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  flip_tail_kernel<real><<<592, 256, 0, s>>>(x, n_head, n, y);
  return cudaGetLastError();
}


// ---------------------------------------------------------------------------------------
// Counter-based normal deviates (Philox4x32-10, Salmon et al. SC'11) for the Brownian step.
// One Philox block per vector ELEMENT, keyed by (seed) and counted by (global element index,
// step): element e of the three noise vectors of a step is a pure function of
// (seed, step, e_global), so the noise -- and with it the trajectory -- does not depend on how the
// suspension is partitioned over GPUs.  (The reference draws from std::normal_distribution seeded
// with the wall clock, c_rigid_obj.cpp:730-741.)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

template <typename real>
__global__ void normal_triplet_kernel(unsigned long long seed, unsigned long long step, unsigned long long first,
                                      size_t n, real* __restrict__ w1, real* __restrict__ w2, real* __restrict__ wr) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long e = first + k;
    uint32_t r[4];
    philox4x32_10((uint32_t)e, (uint32_t)(e >> 32), (uint32_t)step, (uint32_t)(step >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    // Box-Muller in double (the uniforms have 32 bits; (r + 0.5) / 2^32 is never 0 or 1)
    const double u0 = ((double)r[0] + 0.5) * 2.3283064365386963e-10, u1 = ((double)r[1] + 0.5) * 2.3283064365386963e-10;
    const double u2 = ((double)r[2] + 0.5) * 2.3283064365386963e-10, u3 = ((double)r[3] + 0.5) * 2.3283064365386963e-10;
    const double ra = sqrt(-2.0 * log(u0)), rb = sqrt(-2.0 * log(u2));
    double sa, ca, sb, cb;
    sincospi(2.0 * u1, &sa, &ca);
    sincospi(2.0 * u3, &sb, &cb);
    (void)sb;
    w1[k] = (real)(ra * ca);
    w2[k] = (real)(ra * sa);
    wr[k] = (real)(rb * cb);
  }
}
template <typename real>
cudaError_t normal_triplet(unsigned long long seed, unsigned long long step, unsigned long long first, size_t n,
                           real* w1, real* w2, real* wr, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  normal_triplet_kernel<real><<<592, 256, 0, s>>>(seed, step, first, n, w1, w2, wr);
  return cudaGetLastError();
}

#define INST(real)                                                                              \
  template cudaError_t multi_dot<real>(const real*, size_t, int, const real*, size_t, real*,    \
                                       real*, cudaStream_t);                                    \
  template cudaError_t multi_axpy<real>(const real*, size_t, int, const real*, real, real*,     \
                                        size_t, cudaStream_t);                                  \
  template cudaError_t scale_copy<real>(const real*, real, real*, size_t, bool, cudaStream_t);  \
  template cudaError_t flip_tail<real>(const real*, size_t, size_t, real*, cudaStream_t);       \
  template cudaError_t scale_by_inv_sqrt<real>(const real*, const real*, real*, size_t, cudaStream_t); \
  template cudaError_t normal_triplet<real>(unsigned long long, unsigned long long,             \
                                            unsigned long long, size_t, real*, real*, real*,    \
                                            cudaStream_t);
INST(float)
INST(double)
#undef INST
}  // namespace rbl
