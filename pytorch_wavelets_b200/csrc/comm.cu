// comm.cu -- the multi-GPU part of the C ABI (include/b200wave.h): one NCCL communicator per process x device behind an
// explicit handle, and the single collective of the path: the all-gather of a transform's output tensors along the
// batch dimension (planes are independent, so there is no exchange during compute).
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the PyTorch process the library normally lives in has it loaded
// already, and a build without NCCL installed still produces a working single-GPU library (b200w_comm_* then return
// B200W_ENOTIMPL).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200wave.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclFloat = 7 };

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok;
};

const NcclApi& nccl() {
  static NcclApi api = [] {
    NcclApi a;
    memset(&a, 0, sizeof(a));
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return a;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllGather && a.CommDestroy;
    return a;
  }();
  return api;
}

thread_local char g_comm_error[256] = "";

int fail(ncclResult_t r, const char* what) {
  const NcclApi& a = nccl();
  snprintf(g_comm_error, sizeof(g_comm_error), "%s: %s", what, (a.GetErrorString ? a.GetErrorString(r) : "NCCL error"));
  return B200W_ECUDA;
}

}  // namespace

struct b200w_comm {
  ncclComm_t comm;
  int rank, world, device;
};

extern "C" {

const char* b200w_comm_last_error(void) { return g_comm_error; }

int b200w_comm_unique_id(void* id128) {
  if (!id128) return B200W_EARG;
  const NcclApi& a = nccl();
  if (!a.ok) return B200W_ENOTIMPL;
  ncclUniqueId id;
  const ncclResult_t r = a.GetUniqueId(&id);
  if (r != 0) return fail(r, "ncclGetUniqueId");
  memcpy(id128, &id, sizeof(id));
  return B200W_OK;
}

int b200w_comm_init(b200w_comm** out, int rank, int world, const void* id128) {
  if (!out || !id128 || world < 1 || rank < 0 || rank >= world) return B200W_EARG;
  const NcclApi& a = nccl();
  if (!a.ok) return B200W_ENOTIMPL;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  b200w_comm* c = new b200w_comm;
  c->rank = rank; c->world = world; c->device = 0;
  (void)cudaGetDevice(&c->device);
  const ncclResult_t r = a.CommInitRank(&c->comm, world, id, rank);
  if (r != 0) { delete c; return fail(r, "ncclCommInitRank"); }
  *out = c;
  return B200W_OK;
}

int b200w_comm_destroy(b200w_comm* c) {
  if (!c) return B200W_OK;
  const NcclApi& a = nccl();
  ncclResult_t r = 0;
  if (a.ok) r = a.CommDestroy(c->comm);
  delete c;
  return r != 0 ? fail(r, "ncclCommDestroy") : B200W_OK;
}

int b200w_allgather(b200w_comm* c, const float* send, float* recv, long long count, void* stream) {
  if (!c || !send || !recv || count < 0) return B200W_EARG;
  if (count == 0) return B200W_OK;
  const NcclApi& a = nccl();
  if (!a.ok) return B200W_ENOTIMPL;
  const ncclResult_t r = a.AllGather(send, recv, (size_t)count, kNcclFloat, c->comm, (cudaStream_t)stream);
  return r != 0 ? fail(r, "ncclAllGather") : B200W_OK;
}

}  // extern "C"
