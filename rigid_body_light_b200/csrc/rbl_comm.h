// rbl_comm.h -- NCCL plumbing for a context that holds ONE RANK's bodies of a suspension that
// is partitioned over the GPUs of a node (the reference has no
// distributed path at all, /root/reference/src/c_rigid_obj.cpp is single process, single thread).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a single-GPU host never needs it, and a
// process that already imported torch gets torch's own copy of the library (same soname), so
// there is one NCCL per process.  Only the handful of entry points below are used.
//
// Collectives here are "v" shaped because body ranges need not be equal: all-gather-v is a
// group of broadcasts, reduce-scatter-v a group of reduces.  Everything is enqueued on the
// context stream, so kernels and collectives are ordered without host synchronisation.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rbl_peer.cuh"

namespace rbl {

// One rank's view of the peer-memory exchange (rbl_peer.cuh): its own symmetric buffer, the other
// ranks' buffers as mapped through CUDA IPC, and the epochs of the two hand-shakes.
struct PeerExchange {
  PeerTable tab{};
  void* own = nullptr;
  size_t bytes = 0;
  size_t region = 0;    // bytes of ONE right-hand side: 3 n_all reals, rounded up to 256
  size_t lam_off = 0;   // lambda of all blobs, two right-hand sides
  size_t mbuf_off = 0;  // partial products, two right-hand sides
  unsigned long long epoch[2] = {0, 0};
  unsigned long long timeout_ns = 30ull * 1000000000ull;
  bool mapped[kMaxPeers] = {};
  bool on = false;      // set up on every rank
  bool use = false;     // selected (rbl_comm_set_exchange)
};

struct Comm {
  void* lib = nullptr;
  ncclComm_t comm = nullptr;
  bool owns_comm = true;  // false: a view on another context's communicator (share())
  int rank = 0, world = 1;
  std::vector<long long> count;  // blobs held by each rank
  std::vector<long long> first;  // first global blob of each rank
  long long n_all = 0;
  bool even = true;
  std::string err;
  PeerExchange peer;
  std::string peer_why;  // why the peer exchange is off (empty when it is on)

  // entry points
  ncclResult_t (*pCommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*pCommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*pAllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*pAllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*pReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*pBroadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*pReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*pGroupStart)() = nullptr;
  ncclResult_t (*pGroupEnd)() = nullptr;
  const char* (*pGetErrorString)(ncclResult_t) = nullptr;

  ~Comm() {
    if (comm && pCommDestroy && owns_comm) pCommDestroy(comm);
    // the library handle is left open on purpose: torch may share it
  }
  // A second view on the same NCCL communicator for the float mirror of a double context (mixed
  // precision): same ranks and blob ranges, same stream order, its own (float-sized) peer buffers.
  Comm* share() const {
    Comm* c = new Comm();
    c->lib = lib;
    c->comm = comm;
    c->owns_comm = false;
    c->rank = rank;
    c->world = world;
    c->count = count;
    c->first = first;
    c->n_all = n_all;
    c->even = even;
    c->pCommInitRank = pCommInitRank;
    c->pCommDestroy = pCommDestroy;
    c->pAllGather = pAllGather;
    c->pAllReduce = pAllReduce;
    c->pReduceScatter = pReduceScatter;
    c->pBroadcast = pBroadcast;
    c->pReduce = pReduce;
    c->pGroupStart = pGroupStart;
    c->pGroupEnd = pGroupEnd;
    c->pGetErrorString = pGetErrorString;
    return c;
  }

  // ---- peer-memory exchange ------------------------------------------------------------------
  // max over the ranks of v (a host integer), through the device word d_int; < 0 on an NCCL / CUDA error
  int agree_max(int v, int* d_int, cudaStream_t s) {
    if (cudaMemcpyAsync(d_int, &v, sizeof(int), cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
    if (!allreduce_max_int(d_int, 1, s)) return -1;
    int h = 0;
    if (cudaMemcpyAsync(&h, d_int, sizeof(int), cudaMemcpyDeviceToHost, s) != cudaSuccess) return -1;
    if (cudaStreamSynchronize(s) != cudaSuccess) return -1;
    return h;
  }

  // Collective.  Allocates this rank's symmetric buffer, hands its CUDA IPC handle to every other rank
  // (one ncclAllGather of 64 bytes per rank) and maps theirs.  Every rank ends with the same answer:
  // the exchange is on everywhere or nowhere (then peer_why says what failed HERE, or that another rank
  // failed), and the NCCL collectives stay in charge.  `d_int`: a device word for the agreements.
  bool peer_setup(size_t real_size, int* d_int, cudaStream_t s) {
    peer = PeerExchange();
    peer_why.clear();
    const char* env = getenv("RBL_PEER_EXCHANGE");
    bool ok_local = !(env && env[0] == '0');
    if (!ok_local) peer_why = "disabled by RBL_PEER_EXCHANGE=0";
    if (ok_local && world > kMaxPeers) {
      ok_local = false;
      peer_why = "more ranks than the peer table holds";
    }
    if (const char* t = getenv("RBL_PEER_TIMEOUT_S")) {
      const double sec = atof(t);
      if (sec > 0) peer.timeout_ns = (unsigned long long)(sec * 1e9);
    }
    peer.region = ((3 * (size_t)n_all * real_size + 255) / 256) * 256;
    peer.lam_off = kPeerFlagBytes;
    peer.mbuf_off = peer.lam_off + 2 * peer.region;
    peer.bytes = peer.mbuf_off + 2 * peer.region;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok_local) {
      cudaError_t e = cudaMalloc(&peer.own, peer.bytes);
      if (e == cudaSuccess) e = cudaMemsetAsync(peer.own, 0, kPeerFlagBytes, s);
      if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine, peer.own);
      if (e != cudaSuccess) {
        ok_local = false;
        peer_why = std::string("symmetric buffer: ") + cudaGetErrorString(e);
        cudaGetLastError();
      }
    }
    // handles of all ranks (always executed, so the collective sequence never depends on a local failure)
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    char* d_h = nullptr;
    std::vector<cudaIpcMemHandle_t> all(world);
    bool xch = cudaMalloc(&d_h, 64 * (size_t)world) == cudaSuccess;
    if (xch) xch = cudaMemcpyAsync(d_h + 64 * (size_t)rank, &mine, 64, cudaMemcpyHostToDevice, s) == cudaSuccess;
    if (xch) xch = ok(pAllGather(d_h + 64 * (size_t)rank, d_h, 64, ncclChar, comm, s), "ncclAllGather(ipc handles)");
    if (xch) xch = cudaMemcpyAsync(all.data(), d_h, 64 * (size_t)world, cudaMemcpyDeviceToHost, s) == cudaSuccess;
    if (xch) xch = cudaStreamSynchronize(s) == cudaSuccess;
    if (d_h) cudaFree(d_h);
    if (!xch && ok_local) {
      ok_local = false;
      peer_why = "exchange of the IPC handles failed";
      cudaGetLastError();
    }
    int bad = agree_max(ok_local ? 0 : 1, d_int, s);
    if (bad == 0) {
      for (int r = 0; r < world && ok_local; ++r) {
        if (r == rank) {
          peer.tab.base[r] = peer.own;
          continue;
        }
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
          ok_local = false;
          peer_why = std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e);
          cudaGetLastError();
          break;
        }
        peer.tab.base[r] = p;
        peer.mapped[r] = true;
      }
      bad = agree_max(ok_local ? 0 : 1, d_int, s);  // also the barrier that orders every rank's zeroed epochs before the first push
    }
    if (bad != 0) {
      if (peer_why.empty()) peer_why = bad < 0 ? "agreement between the ranks failed" : "another rank could not set it up";
      // (the imports are closed, then an NCCL barrier, before any rank frees what another may still map)
      for (int r = 0; r < kMaxPeers; ++r)
        if (peer.mapped[r]) {
          cudaIpcCloseMemHandle(peer.tab.base[r]);
          peer.mapped[r] = false;
        }
      if (bad > 0) agree_max(0, d_int, s);
      peer_release(false, d_int, s);
      return false;
    }
    peer.on = peer.use = true;
    return true;
  }

  // `collective`: the context is being closed in step with the other ranks.  Nobody may free a buffer that
  // another rank still reads or maps: (1) this rank's stream is drained, (2) a bounded hand-shake over
  // the epochs themselves tells that every rank's is, (3) the imports are closed, (4) an NCCL barrier
  // (only if the hand-shake succeeded: the peers are alive and on their way to it), then the buffer is
  // freed.  A peer that never arrives costs the timeout and leaves the buffer to process exit.
  void peer_release(bool collective, int* d_int, cudaStream_t s) {
    bool free_own = true;
    if (collective && comm && peer.on) {
      int flag = 0;
      const unsigned long long budget = peer.timeout_ns < 10000000000ull ? peer.timeout_ns : 10000000000ull;
      free_own = cudaStreamSynchronize(s) == cudaSuccess &&
                 cudaMemcpyAsync(d_int, &flag, sizeof(int), cudaMemcpyHostToDevice, s) == cudaSuccess &&
                 peer_signal_wait(peer.tab, world, rank, kPeerClosing, 1, budget, d_int, s) == cudaSuccess &&
                 cudaMemcpyAsync(&flag, d_int, sizeof(int), cudaMemcpyDeviceToHost, s) == cudaSuccess &&
                 cudaStreamSynchronize(s) == cudaSuccess && flag == 0;
    }
    for (int r = 0; r < kMaxPeers; ++r)
      if (peer.mapped[r]) {
        cudaIpcCloseMemHandle(peer.tab.base[r]);
        peer.mapped[r] = false;
      }
    if (collective && comm && peer.on && free_own) free_own = agree_max(0, d_int, s) == 0;
    if (peer.own && free_own) cudaFree(peer.own);
    cudaGetLastError();
    peer.own = nullptr;
    peer.on = peer.use = false;
  }
  bool peer_active() const { return peer.on && peer.use; }
  template <typename real>
  real* peer_lam(int k = 0) const { return reinterpret_cast<real*>(static_cast<char*>(peer.own) + peer.lam_off + k * peer.region); }
  template <typename real>
  real* peer_mbuf(int k = 0) const { return reinterpret_cast<real*>(static_cast<char*>(peer.own) + peer.mbuf_off + k * peer.region); }

  // every rank's lam_all[k] <- concatenation of the ranks' slices (n_rhs slices pushed, ONE hand-shake)
  template <typename real>
  bool peer_allgather(const real* const* send_local, int n_rhs, int* err_flag, cudaStream_t s) {
    for (int k = 0; k < n_rhs; ++k) {
      const size_t off = peer.lam_off + k * peer.region + 3 * (size_t)first[rank] * sizeof(real);
      if (peer_push<real>(peer.tab, world, off, send_local[k], 3 * (size_t)count[rank], s) != cudaSuccess) return cuda_err("peer_push");
    }
    if (peer_signal_wait(peer.tab, world, rank, kPeerLambdaReady, ++peer.epoch[kPeerLambdaReady], peer.timeout_ns, err_flag, s) != cudaSuccess)
      return cuda_err("peer_signal_wait");
    return true;
  }
  // recv_local[k] <- this rank's rows of the sum over the ranks of their partial products k
  template <typename real>
  bool peer_reduce_scatter(real* const* recv_local, int n_rhs, int* err_flag, cudaStream_t s) {
    if (peer_signal_wait(peer.tab, world, rank, kPeerPartialReady, ++peer.epoch[kPeerPartialReady], peer.timeout_ns, err_flag, s) != cudaSuccess)
      return cuda_err("peer_signal_wait");
    for (int k = 0; k < n_rhs; ++k) {
      const size_t off = peer.mbuf_off + k * peer.region + 3 * (size_t)first[rank] * sizeof(real);
      if (peer_reduce<real>(peer.tab, world, off, 3 * (size_t)count[rank], recv_local[k], s) != cudaSuccess) return cuda_err("peer_reduce");
    }
    return true;
  }
  bool cuda_err(const char* what) {
    err = std::string(what) + ": " + cudaGetErrorString(cudaGetLastError());
    return false;
  }

  static void* open_lib(std::string* why) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h && why) *why = std::string("dlopen(libnccl.so.2): ") + dlerror();
    return h;
  }

  // ncclGetUniqueId without a communicator (rank 0 calls this, the host broadcasts the bytes)
  static bool unique_id(void* out128, std::string* why) {
    void* h = open_lib(why);
    if (!h) return false;
    auto f = (ncclResult_t(*)(ncclUniqueId*))dlsym(h, "ncclGetUniqueId");
    if (!f) {
      if (why) *why = "ncclGetUniqueId not found";
      return false;
    }
    ncclUniqueId id;
    ncclResult_t r = f(&id);
    if (r != ncclSuccess) {
      if (why) *why = "ncclGetUniqueId failed";
      return false;
    }
    memcpy(out128, &id, sizeof(id));
    return true;
  }

  bool ok(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return true;
    err = std::string(what) + ": " + (pGetErrorString ? pGetErrorString(r) : "NCCL error");
    return false;
  }

  bool init(const void* uid128, int rank_, int world_, const int* blobs_per_rank) {
    lib = open_lib(&err);
    if (!lib) return false;
#define RBL_SYM(member, name)                                   \
  member = (decltype(member))dlsym(lib, name);                  \
  if (!member) {                                                \
    err = std::string("NCCL symbol not found: ") + name;        \
    return false;                                               \
  }
    RBL_SYM(pCommInitRank, "ncclCommInitRank")
    RBL_SYM(pCommDestroy, "ncclCommDestroy")
    RBL_SYM(pAllGather, "ncclAllGather")
    RBL_SYM(pAllReduce, "ncclAllReduce")
    RBL_SYM(pReduceScatter, "ncclReduceScatter")
    RBL_SYM(pBroadcast, "ncclBroadcast")
    RBL_SYM(pReduce, "ncclReduce")
    RBL_SYM(pGroupStart, "ncclGroupStart")
    RBL_SYM(pGroupEnd, "ncclGroupEnd")
    RBL_SYM(pGetErrorString, "ncclGetErrorString")
#undef RBL_SYM
    rank = rank_;
    world = world_;
    count.assign(world, 0);
    first.assign(world, 0);
    n_all = 0;
    even = true;
    for (int r = 0; r < world; ++r) {
      count[r] = blobs_per_rank[r];
      first[r] = n_all;
      n_all += count[r];
      if (count[r] != count[0]) even = false;
    }
    ncclUniqueId id;
    memcpy(&id, uid128, sizeof(id));
    return ok(pCommInitRank(&comm, world, id, rank), "ncclCommInitRank");
  }

  template <typename real>
  static ncclDataType_t dtype() { return sizeof(real) == 8 ? ncclDouble : ncclFloat; }

  // recv_all[per * first[r] ...] <- rank r's `per * count[r]` reals, for every r
  template <typename real>
  bool allgatherv(const real* send_local, real* recv_all, int per, cudaStream_t s) {
    if (even) return ok(pAllGather(send_local, recv_all, (size_t)per * count[0], dtype<real>(), comm, s), "ncclAllGather");
    if (!ok(pGroupStart(), "ncclGroupStart")) return false;
    for (int r = 0; r < world; ++r) {
      real* dst = recv_all + (size_t)per * first[r];
      if (!ok(pBroadcast(r == rank ? (const void*)send_local : (const void*)dst, dst, (size_t)per * count[r], dtype<real>(), r, comm, s),
              "ncclBroadcast"))
        return false;
    }
    return ok(pGroupEnd(), "ncclGroupEnd");
  }

  // recv_local <- sum over ranks of send_all[per * first[rank] ...] (each rank keeps its own rows)
  template <typename real>
  bool reduce_scatterv(const real* send_all, real* recv_local, int per, cudaStream_t s) {
    if (even)
      return ok(pReduceScatter(send_all, recv_local, (size_t)per * count[0], dtype<real>(), ncclSum, comm, s), "ncclReduceScatter");
    if (!ok(pGroupStart(), "ncclGroupStart")) return false;
    for (int r = 0; r < world; ++r) {
      if (!ok(pReduce(send_all + (size_t)per * first[r], recv_local, (size_t)per * count[r], dtype<real>(), ncclSum, r, comm, s),
              "ncclReduce"))
        return false;
    }
    return ok(pGroupEnd(), "ncclGroupEnd");
  }

  template <typename real>
  bool allreduce_sum(real* buf, size_t n, cudaStream_t s) {
    return ok(pAllReduce(buf, buf, n, dtype<real>(), ncclSum, comm, s), "ncclAllReduce");
  }
  bool allreduce_max_int(int* buf, size_t n, cudaStream_t s) {
    return ok(pAllReduce(buf, buf, n, ncclInt, ncclMax, comm, s), "ncclAllReduce(max)");
  }
};

}  // namespace rbl
