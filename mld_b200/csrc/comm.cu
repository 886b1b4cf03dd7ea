// The one collective of the path: an all-gather of the finished motions over NVLink.
// Every motion is independent, so ranks sample contiguous batch shards with replicated weights and nothing
// is exchanged per step; the finished joints are written by k_feats2joints straight into this rank's slot of
// the gathered buffer and ONE in-place ncclAllGather on a side stream closes the batch - the next batch's
// graph runs on the caller's stream meanwhile.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 - the copy PyTorch already loaded, else the system
// one), so libmldb200.so has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "engine.h"

namespace {

struct NcclUniqueId { char internal[128]; };     // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* NcclComm;
constexpr int kNcclFloat32 = 7;                  // ncclFloat32

typedef int (*fn_GetUniqueId)(NcclUniqueId*);
typedef int (*fn_CommInitRank)(NcclComm*, int, NcclUniqueId, int);
typedef int (*fn_CommDestroy)(NcclComm);
typedef int (*fn_AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t);
typedef const char* (*fn_GetErrorString)(int);

struct NcclApi {
  void* lib = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr;
  fn_CommInitRank CommInitRank = nullptr;
  fn_CommDestroy CommDestroy = nullptr;
  fn_AllGather AllGather = nullptr;
  fn_GetErrorString GetErrorString = nullptr;
  bool ok = false;
};

NcclApi& nccl() {
  static NcclApi api;
  if (api.lib) return api;
  api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy torch loaded, if any
  if (!api.lib) api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!api.lib) return api;
  api.GetUniqueId = (fn_GetUniqueId)dlsym(api.lib, "ncclGetUniqueId");
  api.CommInitRank = (fn_CommInitRank)dlsym(api.lib, "ncclCommInitRank");
  api.CommDestroy = (fn_CommDestroy)dlsym(api.lib, "ncclCommDestroy");
  api.AllGather = (fn_AllGather)dlsym(api.lib, "ncclAllGather");
  api.GetErrorString = (fn_GetErrorString)dlsym(api.lib, "ncclGetErrorString");
  api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.GetErrorString;
  return api;
}

int fail(int code, const char* what, int nccl_rc = 0) {
  char buf[256];
  if (nccl_rc && nccl().GetErrorString) snprintf(buf, sizeof buf, "%s: %s", what, nccl().GetErrorString(nccl_rc));
  else snprintf(buf, sizeof buf, "%s", what);
  mldb_set_err(buf);
  return code;
}

int ensure_streams(mldb_handle* h) {
  if (h->comm_stream) return MLDB_OK;
  if (cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_local_done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_gather_done[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_gather_done[1], cudaEventDisableTiming) != cudaSuccess)
    return fail(MLDB_ERR_CUDA, "comm stream / events");
  return MLDB_OK;
}

}  // namespace

extern "C" int mldb_comm_unique_id(void* out128) {
  if (!out128) return fail(MLDB_ERR_INVALID, "null argument");
  if (!nccl().ok) return fail(MLDB_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  NcclUniqueId id;
  const int rc = nccl().GetUniqueId(&id);
  if (rc) return fail(MLDB_ERR_CUDA, "ncclGetUniqueId", rc);
  memcpy(out128, &id, sizeof id);
  return MLDB_OK;
}

extern "C" int mldb_comm_init(mldb_handle* h, const void* unique_id128, int32_t nranks, int32_t rank) {
  if (!h || !unique_id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(MLDB_ERR_INVALID, "bad argument");
  if (h->nccl_comm) return fail(MLDB_ERR_STATE, "this handle already has a communicator");
  if (!nccl().ok) return fail(MLDB_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  NcclUniqueId id;
  memcpy(&id, unique_id128, sizeof id);
  NcclComm comm = nullptr;
  const int rc = nccl().CommInitRank(&comm, nranks, id, rank);
  int st = rc ? fail(MLDB_ERR_CUDA, "ncclCommInitRank", rc) : ensure_streams(h);
  if (prev >= 0 && prev != h->device) cudaSetDevice(prev);
  if (st != MLDB_OK) return st;
  h->nccl_comm = comm; h->comm_owned = true; h->comm_world = nranks; h->comm_rank = rank;
  return MLDB_OK;
}

extern "C" int mldb_comm_attach(mldb_handle* h, void* nccl_comm, int32_t nranks, int32_t rank) {
  if (!h || !nccl_comm || nranks < 1 || rank < 0 || rank >= nranks) return fail(MLDB_ERR_INVALID, "bad argument");
  if (h->nccl_comm) return fail(MLDB_ERR_STATE, "this handle already has a communicator");
  if (!nccl().ok) return fail(MLDB_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  int st = ensure_streams(h);
  if (st != MLDB_OK) return st;
  h->nccl_comm = nccl_comm; h->comm_owned = false; h->comm_world = nranks; h->comm_rank = rank;
  return MLDB_OK;
}

void mldb_comm_release(mldb_handle* h) {
  if (h->nccl_comm && h->comm_owned && nccl().ok) nccl().CommDestroy((NcclComm)h->nccl_comm);
  h->nccl_comm = nullptr;
  if (h->comm_stream) { cudaStreamDestroy(h->comm_stream); h->comm_stream = nullptr; }
  if (h->ev_local_done) { cudaEventDestroy(h->ev_local_done); h->ev_local_done = nullptr; }
  for (int i = 0; i < 2; ++i)
    if (h->ev_gather_done[i]) { cudaEventDestroy(h->ev_gather_done[i]); h->ev_gather_done[i] = nullptr; }
}

extern "C" int mldb_comm_info(const mldb_handle* h, int32_t* nranks, int32_t* rank) {
  if (!h || !nranks || !rank) return fail(MLDB_ERR_INVALID, "null argument");
  *nranks = h->nccl_comm ? h->comm_world : 1;
  *rank = h->nccl_comm ? h->comm_rank : 0;
  return MLDB_OK;
}

// local [count] floats of every rank -> global [nranks * count]; in place when
// local == global + rank * count.  Enqueued on `stream`.
extern "C" int mldb_allgather(mldb_handle* h, const float* local, float* global, int64_t count, void* stream) {
  if (!h || !local || !global || count <= 0) return fail(MLDB_ERR_INVALID, "bad argument");
  if (!h->nccl_comm) return fail(MLDB_ERR_STATE, "call mldb_comm_init / mldb_comm_attach first");
  const int rc = nccl().AllGather(local, global, (size_t)count, kNcclFloat32, (NcclComm)h->nccl_comm, (cudaStream_t)stream);
  if (rc) return fail(MLDB_ERR_CUDA, "ncclAllGather", rc);
  return MLDB_OK;
}

// The gather of a finished batch, off the caller's stream: the side stream waits for everything enqueued
// on `stream` so far (the sample that wrote this rank's slot of `global`), gathers in place and records
// the completion that mldb_gather_wait makes a stream wait for.  Gathers alternate between two completion
// events so that a caller that alternates two output buffers can let batch i's gather overlap batch i+1's
// sampling graph: mldb_gather_begin makes `stream` wait for the gather issued TWO calls ago (the last user
// of the buffer about to be overwritten).
int mldb_gather_begin(mldb_handle* h, cudaStream_t stream) {
  if (!h->nccl_comm) return fail(MLDB_ERR_STATE, "call mldb_comm_init / mldb_comm_attach first");
  if (h->gather_count >= 2 &&
      cudaStreamWaitEvent(stream, h->ev_gather_done[h->gather_count & 1], 0) != cudaSuccess)
    return fail(MLDB_ERR_CUDA, "gather: event");
  return MLDB_OK;
}
int mldb_gather_async(mldb_handle* h, float* global, int64_t count, cudaStream_t stream) {
  if (!h->nccl_comm) return fail(MLDB_ERR_STATE, "call mldb_comm_init / mldb_comm_attach first");
  if (cudaEventRecord(h->ev_local_done, stream) != cudaSuccess ||
      cudaStreamWaitEvent(h->comm_stream, h->ev_local_done, 0) != cudaSuccess)
    return fail(MLDB_ERR_CUDA, "gather: event");
  const int rc = nccl().AllGather(global + (int64_t)h->comm_rank * count, global, (size_t)count, kNcclFloat32,
                                  (NcclComm)h->nccl_comm, h->comm_stream);
  if (rc) return fail(MLDB_ERR_CUDA, "ncclAllGather", rc);
  if (cudaEventRecord(h->ev_gather_done[h->gather_count & 1], h->comm_stream) != cudaSuccess)
    return fail(MLDB_ERR_CUDA, "gather: event");
  h->gather_count++;
  return MLDB_OK;
}

extern "C" int mldb_gather_wait(mldb_handle* h, void* stream) {
  if (!h) return fail(MLDB_ERR_INVALID, "null handle");
  if (h->gather_count == 0) return MLDB_OK;
  if (cudaStreamWaitEvent((cudaStream_t)stream, h->ev_gather_done[(h->gather_count - 1) & 1], 0) != cudaSuccess)
    return fail(MLDB_ERR_CUDA, "gather wait");
  return MLDB_OK;
}
