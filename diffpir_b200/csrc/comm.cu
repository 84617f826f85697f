// comm.cu - the ONE collective of the path, inside the library: broadcast of the packed weight arena over NCCL / NVLink
// (SURVEY.md section 8(b), 8(e)).  The reference is a single-process program (main_ddpir.py:214-236 loads the checkpoint on its
// only GPU); here one process per GPU runs an independent shard of the images, and the only data every rank needs from another is
// the model: rank `root` repacks the checkpoint once (dpir_load_weight) and ships the kernel-native arena - fp16 tensor-core
// layouts, tables - to its peers, which never touch the checkpoint.
//
// NCCL is bound at run time (dlopen) rather than at link time: a PyTorch process already carries its own libnccl.so.2, and a
// second copy of a different version in the same process is not something a library should cause.  The handle already loaded is
// used when there is one, else the system library.  No host-side fallback: without NCCL the calls fail with DPIR_E_STATE.
#include "common.cuh"

#include <dlfcn.h>
#include <cstring>
#include <mutex>

namespace {

// the five entry points used, with the types of nccl.h (2.x ABI) restated so that the build needs no NCCL headers
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm *NcclComm;
enum { kNcclSuccess = 0, kNcclUint8 = 1 };

struct NcclApi {
  void *handle = nullptr;
  int (*GetVersion)(int *) = nullptr;
  int (*GetUniqueId)(NcclUniqueId *) = nullptr;
  int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string why;            // why binding failed
};

NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // the copy this process already uses (PyTorch's), if any
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.why = std::string("libnccl.so.2 not found: ") + (dlerror() ? dlerror() : "?"); return; }
    auto sym = [&](const char *name) -> void * {
      void *p = dlsym(h, name);
      if (!p && api.why.empty()) api.why = std::string("libnccl: missing symbol ") + name;
      return p;
    };
    api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (api.why.empty()) api.handle = h;
  });
  return &api;
}

}  // namespace

struct dpir_comm {
  NcclComm comm = nullptr;
  int world = 0, rank = 0, device = 0;
  bool owned = true;          // false: adopted from the host program (dpir_comm_adopt), never destroyed here
};

#define DPIR_NCCL_API(api)                                                                      \
  NcclApi *api = nccl_api();                                                                    \
  if (!api->handle) {                                                                           \
    dpir::set_error("NCCL unavailable: %s", api->why.c_str());                                  \
    return DPIR_E_STATE;                                                                        \
  }

#define DPIR_NCCL_CHECK(api, expr)                                                              \
  do {                                                                                          \
    int _r = (expr);                                                                            \
    if (_r != kNcclSuccess) {                                                                   \
      dpir::set_error("%s failed: %s", #expr, api->GetErrorString(_r));                         \
      return DPIR_E_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

extern "C" {

int dpir_nccl_version(int32_t *version) {
  DPIR_REQUIRE(version, "dpir_nccl_version: null pointer");
  DPIR_NCCL_API(api);
  int v = 0;
  DPIR_NCCL_CHECK(api, api->GetVersion(&v));
  *version = v;
  return DPIR_OK;
}

int dpir_comm_unique_id(void *id, size_t cap) {
  DPIR_REQUIRE(id && cap >= DPIR_COMM_ID_BYTES, "dpir_comm_unique_id: need a %d-byte buffer", DPIR_COMM_ID_BYTES);
  DPIR_NCCL_API(api);
  NcclUniqueId u;
  DPIR_NCCL_CHECK(api, api->GetUniqueId(&u));
  memcpy(id, u.internal, sizeof(u.internal));
  return DPIR_OK;
}

int dpir_comm_create(const void *id, int32_t world, int32_t rank, dpir_comm **out) {
  DPIR_REQUIRE(id && out, "dpir_comm_create: null pointer");
  DPIR_REQUIRE(world >= 1 && rank >= 0 && rank < world, "dpir_comm_create: rank %d outside [0, %d)", rank, world);
  DPIR_NCCL_API(api);
  NcclUniqueId u;
  memcpy(u.internal, id, sizeof(u.internal));
  dpir_comm *c = new dpir_comm();
  c->world = world; c->rank = rank;
  if (cudaGetDevice(&c->device) != cudaSuccess) { delete c; dpir::set_error("dpir_comm_create: no CUDA device"); return DPIR_E_CUDA; }
  int r = api->CommInitRank(&c->comm, world, u, rank);       // collective: every rank of the job calls it with the same id
  if (r != kNcclSuccess) {
    dpir::set_error("ncclCommInitRank(world %d, rank %d) failed: %s", world, rank, api->GetErrorString(r));
    delete c;
    return DPIR_E_CUDA;
  }
  *out = c;
  return DPIR_OK;
}

int dpir_comm_adopt(void *nccl_comm, int32_t world, int32_t rank, dpir_comm **out) {
  DPIR_REQUIRE(nccl_comm && out, "dpir_comm_adopt: null pointer");
  DPIR_REQUIRE(world >= 1 && rank >= 0 && rank < world, "dpir_comm_adopt: rank %d outside [0, %d)", rank, world);
  DPIR_NCCL_API(api);
  (void)api;
  dpir_comm *c = new dpir_comm();
  c->comm = reinterpret_cast<NcclComm>(nccl_comm);
  c->world = world; c->rank = rank; c->owned = false;
  if (cudaGetDevice(&c->device) != cudaSuccess) { delete c; dpir::set_error("dpir_comm_adopt: no CUDA device"); return DPIR_E_CUDA; }
  *out = c;
  return DPIR_OK;
}

int dpir_comm_destroy(dpir_comm *c) {
  if (!c) return DPIR_OK;
  NcclApi *api = nccl_api();
  if (c->owned && api->handle && c->comm) api->CommDestroy(c->comm);
  delete c;
  return DPIR_OK;
}

int dpir_bcast_weights(dpir_engine *e, dpir_comm *c, int32_t root, void *stream) {
  DPIR_REQUIRE(e && c, "dpir_bcast_weights: null pointer");
  DPIR_REQUIRE(root >= 0 && root < c->world, "dpir_bcast_weights: root %d outside [0, %d)", root, c->world);
  DPIR_NCCL_API(api);
  if (c->rank == root) DPIR_REQUIRE(dpir_missing_weights(e) == 0, "dpir_bcast_weights: the root's checkpoint is incomplete (%d tensors missing)", dpir_missing_weights(e));
  int dev = -1;
  DPIR_CUDA_CHECK(cudaGetDevice(&dev));
  DPIR_REQUIRE(dev == c->device, "dpir_bcast_weights: communicator was created on device %d, current device is %d", c->device, dev);
  void *arena = nullptr;
  size_t bytes = 0;
  int rc = dpir_weight_arena(e, &arena, &bytes);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  // every rank built the same arena layout from the same dpir_unet_config, so the byte count agrees by construction
  DPIR_NCCL_CHECK(api, api->Broadcast(arena, arena, bytes, kNcclUint8, root, c->comm, st));
  if (c->rank != root) {
    // the arena's weights are in place once the broadcast has run on `stream`; tables derived from them (FiLM) are rebuilt lazily
    rc = dpir_mark_weights_loaded(e);
  }
  return rc;
}

}  // extern "C"
