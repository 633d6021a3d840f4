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

#include <cstring>
#include <string>
#include <vector>

namespace rbl {

struct Comm {
  void* lib = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  std::vector<long long> count;  // blobs held by each rank
  std::vector<long long> first;  // first global blob of each rank
  long long n_all = 0;
  bool even = true;
  std::string err;

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
    if (comm && pCommDestroy) pCommDestroy(comm);
    // the library handle is left open on purpose: torch may share it
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
