// rbl_krylov.cu -- see rbl_krylov.cuh
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

#define INST(real)                                                                              \
  template cudaError_t multi_dot<real>(const real*, size_t, int, const real*, size_t, real*,    \
                                       real*, cudaStream_t);                                    \
  template cudaError_t multi_axpy<real>(const real*, size_t, int, const real*, real, real*,     \
                                        size_t, cudaStream_t);                                  \
  template cudaError_t scale_copy<real>(const real*, real, real*, size_t, bool, cudaStream_t);  \
  template cudaError_t flip_tail<real>(const real*, size_t, size_t, real*, cudaStream_t);
INST(float)
INST(double)
#undef INST
}  // namespace rbl
