// comm.cu — the path's one collective behind the C-ABI: an all-gather of the
// per-shard result rows (logits [B/G, classes], or per-cloud losses) over NCCL / NVLink.
//
// The eval forward is per-cloud, so there is no exchange step inside the path; the only
// communication is this gather, 10 KB per rank at the headline config — latency-bound. NCCL is
// resolved at run time from the process (dlopen of libnccl.so.2: the copy PyTorch already loaded,
// or the system one), so libsonet_b200.so keeps its single link dependency on libcudart and a
// single-GPU user never needs NCCL. The unique id is created by rank 0 and distributed by the
// host program with whatever transport it has (sonet_b200/dist.py uses the torch.distributed
// store); every call is asynchronous on the caller's stream and CUDA-graph capturable.
#include <dlfcn.h>

#include <mutex>

#include "common.cuh"

namespace {

struct NcclUniqueId {
  char internal[128];
};
typedef void* ncclComm_t;
typedef int ncclResult_t;
enum { kNcclInt8 = 0 };

struct NcclApi {
  ncclResult_t (*GetUniqueId)(NcclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // the copy already in the process
    if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) return;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(dlsym(h, "ncclGetVersion"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy;
  });
  return api;
}

struct Comm {
  ncclComm_t comm;
  int rank, world;
};

int nccl_fail(const char* what, ncclResult_t r) {
  NcclApi& a = nccl();
  sonet::set_error("%s: NCCL error %d (%s)", what, r, a.GetErrorString ? a.GetErrorString(r) : "?");
  return SONET_ERR_CUDA;
}

}  // namespace

extern "C" int sonet_comm_nccl_version(void) {
  NcclApi& a = nccl();
  int v = 0;
  if (!a.ok || a.GetVersion == nullptr || a.GetVersion(&v) != 0) return 0;
  return v;
}

extern "C" int sonet_comm_unique_id(void* id128) {
  using namespace sonet;
  SONET_REQUIRE(id128 != nullptr, "comm_unique_id: null pointer");
  NcclApi& a = nccl();
  if (!a.ok) SONET_FAIL(SONET_ERR_UNSUPPORTED, "comm: libnccl.so.2 is not loadable in this process");
  ncclResult_t r = a.GetUniqueId(static_cast<NcclUniqueId*>(id128));
  return r == 0 ? SONET_OK : nccl_fail("ncclGetUniqueId", r);
}

extern "C" int sonet_comm_init(const void* id128, int rank, int world, void** comm_out) {
  using namespace sonet;
  SONET_REQUIRE(id128 && comm_out && world >= 1 && rank >= 0 && rank < world, "comm_init: bad args");
  NcclApi& a = nccl();
  if (!a.ok) SONET_FAIL(SONET_ERR_UNSUPPORTED, "comm: libnccl.so.2 is not loadable in this process");
  NcclUniqueId id = *static_cast<const NcclUniqueId*>(id128);
  ncclComm_t c = nullptr;
  ncclResult_t r = a.CommInitRank(&c, world, id, rank);      // on the current CUDA device
  if (r != 0) return nccl_fail("ncclCommInitRank", r);
  *comm_out = new Comm{c, rank, world};
  return SONET_OK;
}

extern "C" int sonet_allgather(void* comm, const void* send, void* recv, long long bytes_per_rank,
                               sonet_stream_t stream) {
  using namespace sonet;
  SONET_REQUIRE(comm && send && recv && bytes_per_rank >= 0, "allgather: bad args");
  Comm* c = static_cast<Comm*>(comm);
  if (bytes_per_rank == 0) return SONET_OK;
  ncclResult_t r = nccl().AllGather(send, recv, static_cast<size_t>(bytes_per_rank), kNcclInt8, c->comm,
                                    as_stream(stream));
  return r == 0 ? SONET_OK : nccl_fail("ncclAllGather", r);
}

extern "C" int sonet_comm_destroy(void* comm) {
  if (comm == nullptr) return SONET_OK;
  Comm* c = static_cast<Comm*>(comm);
  ncclResult_t r = nccl().CommDestroy(c->comm);
  delete c;
  return r == 0 ? SONET_OK : nccl_fail("ncclCommDestroy", r);
}
