// rbl_peer.cuh -- the two exchanges around a partitioned mobility product, done by this library's own
// kernels over NVLink / NVSwitch peer memory instead of NCCL collectives (the
// reference has no distributed path, /root/reference/src/c_rigid_obj.cpp is single process).
//
// Every rank owns one "symmetric" device buffer with the same layout
//     [ epochs: 3 kinds x kMaxPeers x u64 | lambda of ALL blobs (2 right-hand sides) | partial products (2) ]
// exported with cudaIpcGetMemHandle and mapped by every other rank of the node (one process per GPU).
//   * all-gather of lambda  = PUSH: each rank stores its slice into every peer's lambda region
//     (fire-and-forget NVLink stores), then raises its epoch in every peer's "lambda ready" row and waits
//     until all ranks have raised theirs in its own;
//   * reduce-scatter of the partial products = PULL: after its share of the product a rank raises its
//     epoch in every peer's "partial ready" row, waits for all of them, then sums ITS rows of the world's
//     partial products in rank order (deterministic) straight from the peers' buffers.
// Buffer reuse needs no extra barrier: a peer can only push lambda k+1 after its reduce k, which waited
// for this rank's "partial ready k", which this rank raised after the kernel that consumed lambda k; and
// this rank only overwrites its partial product k after it has seen every peer's "lambda ready k+1",
// which a peer raises after its reduce k has read that partial product.
// Waits are bounded (a wall-clock budget on %globaltimer): a peer that never arrives raises an error
// flag that the next host synchronisation reports, instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace rbl {

constexpr int kMaxPeers = 16;
constexpr size_t kPeerFlagBytes = 1024;  // 3 kinds x kMaxPeers x 8 = 384 B of epochs, padded

struct PeerTable {
  void* base[kMaxPeers];  // symmetric buffer of every rank as mapped into this process (own rank: the allocation)
};

enum PeerKind { kPeerLambdaReady = 0, kPeerPartialReady = 1, kPeerClosing = 2 };

// dst_r[i] = src[i], i < n, for every rank r, dst_r = (real*)(base[r] + dst_off_bytes)
template <typename real>
cudaError_t peer_push(const PeerTable& T, int world, size_t dst_off_bytes, const real* src, size_t n, cudaStream_t s);

// raise epochs[kind][rank] = epoch on every rank, then wait until epochs[kind][r] >= epoch for every r on
// this rank; on timeout *err_flag = 1
cudaError_t peer_signal_wait(const PeerTable& T, int world, int rank, int kind, unsigned long long epoch,
                             unsigned long long timeout_ns, int* err_flag, cudaStream_t s);

// out[i] = sum_{r = 0 .. world-1} src_r[i], src_r = (const real*)(base[r] + src_off_bytes), fixed rank order
template <typename real>
cudaError_t peer_reduce(const PeerTable& T, int world, size_t src_off_bytes, size_t n, real* out, cudaStream_t s);

}  // namespace rbl
