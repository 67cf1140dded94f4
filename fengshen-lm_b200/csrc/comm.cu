// fsb200 — communication helpers of the C ABI (fsb_comm_{init,reduce_scatter,all_gather,destroy}).
// The reference delegates the gradient reduce-scatter / parameter all-gather of ZeRO-1/2 to DeepSpeed, which issues them as
// NCCL collectives (fengshen/strategies/megatron_deepspeed.py:302-320 -> deepspeed.initialize). A host
// that is not PyTorch binds these entry points instead of torch.distributed; fsb200.engine.ZeroEngine(comm_backend="fsb")
// drives the same collectives through them. NCCL is resolved at run time (dlopen of libnccl.so.2 — the copy a hosting
// PyTorch process has already loaded is reused), so libfsb200.so carries no link-time dependency on it.
#include <dlfcn.h>
#include <string.h>

#include "host_common.h"

namespace fsb {

// the stable subset of nccl.h this file needs (declared here so that the build does not depend on an NCCL header path)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclSum = 0, kNcclFloat32 = 7, kNcclBfloat16 = 9 };

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};

static NcclApi* nccl() {
  static NcclApi api = [] {
    NcclApi a;
    memset(&a, 0, sizeof(a));
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // already in the process (PyTorch)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
#define FSB_SYM(field, name) *reinterpret_cast<void**>(&a.field) = dlsym(h, name)
    FSB_SYM(GetUniqueId, "ncclGetUniqueId");
    FSB_SYM(CommInitRank, "ncclCommInitRank");
    FSB_SYM(CommDestroy, "ncclCommDestroy");
    FSB_SYM(ReduceScatter, "ncclReduceScatter");
    FSB_SYM(AllGather, "ncclAllGather");
    FSB_SYM(AllReduce, "ncclAllReduce");
    FSB_SYM(GetErrorString, "ncclGetErrorString");
#undef FSB_SYM
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.ReduceScatter && a.AllGather && a.AllReduce;
    return a;
  }();
  return &api;
}

static int nccl_dtype(int dtype) { return dtype == FSB_F32 ? kNcclFloat32 : kNcclBfloat16; }

#define FSB_NCCL(call, what)                                                                        \
  do {                                                                                              \
    ncclResult_t r__ = (call);                                                                      \
    if (r__ != 0) {                                                                                 \
      set_error("%s: NCCL error %d (%s)", what, int(r__), a->GetErrorString ? a->GetErrorString(r__) : "?"); \
      return FSB_ERR_CUDA;                                                                          \
    }                                                                                               \
  } while (0)

}  // namespace fsb

using namespace fsb;

extern "C" int fsb_comm_unique_id(void* id128) {
  NcclApi* a = nccl();
  FSB_REQUIRE(a->ok, "comm: libnccl.so.2 could not be loaded");
  FSB_REQUIRE(id128 != nullptr, "comm_unique_id: null pointer");
  FSB_NCCL(a->GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)), "comm_unique_id");
  return FSB_OK;
}

extern "C" int fsb_comm_init(fsb_comm_t* comm, const void* id128, int world, int rank) {
  NcclApi* a = nccl();
  FSB_REQUIRE(a->ok, "comm: libnccl.so.2 could not be loaded");
  FSB_REQUIRE(comm && id128 && world > 0 && rank >= 0 && rank < world, "comm_init: bad arguments (world %d rank %d)", world, rank);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  FSB_NCCL(a->CommInitRank(&c, world, id, rank), "comm_init");
  *comm = reinterpret_cast<fsb_comm_t>(c);
  return FSB_OK;
}

extern "C" int fsb_comm_destroy(fsb_comm_t comm) {
  NcclApi* a = nccl();
  FSB_REQUIRE(a->ok && comm, "comm_destroy: no communicator");
  FSB_NCCL(a->CommDestroy(reinterpret_cast<ncclComm_t>(comm)), "comm_destroy");
  return FSB_OK;
}

extern "C" int fsb_comm_reduce_scatter(fsb_comm_t comm, const void* send, void* recv, int64_t recv_count, int dtype,
                                       fsb_stream_t st) {
  NcclApi* a = nccl();
  FSB_REQUIRE(a->ok && comm && send && recv && recv_count > 0, "comm_reduce_scatter: bad arguments");
  FSB_NCCL(a->ReduceScatter(send, recv, size_t(recv_count), nccl_dtype(dtype), kNcclSum, reinterpret_cast<ncclComm_t>(comm),
                            (cudaStream_t)st), "comm_reduce_scatter");
  return FSB_OK;
}

extern "C" int fsb_comm_all_gather(fsb_comm_t comm, const void* send, void* recv, int64_t send_count, int dtype,
                                   fsb_stream_t st) {
  NcclApi* a = nccl();
  FSB_REQUIRE(a->ok && comm && send && recv && send_count > 0, "comm_all_gather: bad arguments");
  FSB_NCCL(a->AllGather(send, recv, size_t(send_count), nccl_dtype(dtype), reinterpret_cast<ncclComm_t>(comm), (cudaStream_t)st),
           "comm_all_gather");
  return FSB_OK;
}

extern "C" int fsb_comm_all_reduce(fsb_comm_t comm, const void* send, void* recv, int64_t count, int dtype, fsb_stream_t st) {
  NcclApi* a = nccl();
  FSB_REQUIRE(a->ok && comm && send && recv && count > 0, "comm_all_reduce: bad arguments");
  FSB_NCCL(a->AllReduce(send, recv, size_t(count), nccl_dtype(dtype), kNcclSum, reinterpret_cast<ncclComm_t>(comm),
                        (cudaStream_t)st), "comm_all_reduce");
  return FSB_OK;
}
