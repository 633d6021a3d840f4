// rbl_peer.cu -- kernels of the peer-memory exchange (see rbl_peer.cuh for the protocol).
#include "rbl_peer.cuh"

namespace rbl {
namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <typename real>
__global__ void peer_push_kernel(PeerTable T, size_t dst_off_bytes, const real* __restrict__ src, size_t n) {
  real* dst = reinterpret_cast<real*>(static_cast<char*>(T.base[blockIdx.y]) + dst_off_bytes);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

__global__ void peer_signal_wait_kernel(PeerTable T, int world, int rank, int kind, unsigned long long epoch,
                                        unsigned long long timeout_ns, int* err_flag) {
  const int r = threadIdx.x;
  if (r >= world) return;
  // everything this stream wrote before (the pushed slices, the partial product) is complete: kernels of a
  // stream run in order; the fence + release store publish it at system scope before the epoch
  __threadfence_system();
  st_release_sys(static_cast<unsigned long long*>(T.base[r]) + kind * kMaxPeers + rank, epoch);
  const unsigned long long* mine = static_cast<const unsigned long long*>(T.base[rank]) + kind * kMaxPeers + r;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(mine) < epoch) {
    __nanosleep(100);
    if (global_ns() - t0 > timeout_ns) {
      atomicExch(err_flag, 1);
      break;
    }
  }
}

template <typename real>
__global__ void peer_reduce_kernel(PeerTable T, int world, size_t src_off_bytes, size_t n, real* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    real v[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)  // all loads in flight before the first add: one NVLink round trip
      v[r] = r < world ? __ldcv(reinterpret_cast<const real*>(static_cast<const char*>(T.base[r]) + src_off_bytes) + i) : (real)0;
    real acc = v[0];
#pragma unroll
    for (int r = 1; r < kMaxPeers; ++r) acc += v[r];  // rank order: reproducible given the partial products
    out[i] = acc;
  }
}

}  // namespace

template <typename real>
cudaError_t peer_push(const PeerTable& T, int world, size_t dst_off_bytes, const real* src, size_t n, cudaStream_t s) {
  if (n == 0 || world < 1) return cudaSuccess;
  const unsigned bx = (unsigned)((n + 1023) / 1024 < 128 ? (n + 1023) / 1024 : 128);
  peer_push_kernel<real><<<dim3(bx, (unsigned)world), 256, 0, s>>>(T, dst_off_bytes, src, n);
  return cudaGetLastError();
}

cudaError_t peer_signal_wait(const PeerTable& T, int world, int rank, int kind, unsigned long long epoch,
                             unsigned long long timeout_ns, int* err_flag, cudaStream_t s) {
  peer_signal_wait_kernel<<<1, 32, 0, s>>>(T, world, rank, kind, epoch, timeout_ns, err_flag);
  return cudaGetLastError();
}

template <typename real>
cudaError_t peer_reduce(const PeerTable& T, int world, size_t src_off_bytes, size_t n, real* out, cudaStream_t s) {
  if (n == 0) return cudaSuccess;
  const unsigned bx = (unsigned)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
  peer_reduce_kernel<real><<<bx, 256, 0, s>>>(T, world, src_off_bytes, n, out);
  return cudaGetLastError();
}

template cudaError_t peer_push<float>(const PeerTable&, int, size_t, const float*, size_t, cudaStream_t);
template cudaError_t peer_push<double>(const PeerTable&, int, size_t, const double*, size_t, cudaStream_t);
template cudaError_t peer_reduce<float>(const PeerTable&, int, size_t, size_t, float*, cudaStream_t);
template cudaError_t peer_reduce<double>(const PeerTable&, int, size_t, size_t, double*, cudaStream_t);

}  // namespace rbl
