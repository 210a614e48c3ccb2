// fa_nccl.cpp -- the thin communication C ABI of the data-parallel machine-translation run (include/fa_b200.h,
// "gradient all-reduce"): fa_b200_nccl_{unique_id, init, allreduce_f32, destroy}.
//
// The reference has no distributed backend; config 5 of BASELINE.json asks for one NCCL all-reduce of
// the flattened fp32 gradients per step over NVLink 5 / NVSwitch.  One process per GPU; the 128-byte ncclUniqueId is
// created by rank 0 and handed to the other ranks by the host side (fa_b200/mt/comm.py: a TCP rendezvous, no framework).
//
// libnccl is resolved with dlopen at first use, so libfa_b200.so itself keeps depending on libcudart only and loads on
// machines without NCCL (the attention path has no collective).  Inside a process that already holds a libnccl.so.2
// (e.g. PyTorch's bundled one) the same library is reused.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/fa_b200.h"

namespace {

// the handful of NCCL declarations used here (stable since NCCL 2.10; values from nccl.h)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { kNcclFloat32 = 7, kNcclSum = 0, kNcclAvg = 4 };

struct Api {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

Api g_api;
ncclComm_t g_comm = nullptr;
int g_rank = -1, g_world = 0;
std::mutex g_mu;
thread_local char g_nccl_err[256] = "";

int fail(const char* what, const char* detail) {
  snprintf(g_nccl_err, sizeof(g_nccl_err), "%s: %s", what, detail ? detail : "?");
  return FA_ERR_CUDA;
}

int load_api() {
  if (g_api.handle) return FA_OK;
  // Order: an explicit path, then a libnccl.so.2 the process already holds (a framework's bundled copy: two different
  // NCCL builds under one SONAME cannot coexist, and whoever comes second would be bound to the first), then the
  // system library.  RTLD_LOCAL: our use must not leak symbols into later loads.
  void* h = nullptr;
  if (const char* e = getenv("FA_B200_NCCL_LIB")) {
    if (*e) h = dlopen(e, RTLD_NOW | RTLD_LOCAL);
  }
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return fail("fa_b200_nccl: cannot load libnccl.so.2 (set FA_B200_NCCL_LIB)", dlerror());
#define FA_SYM(field, name)                                                       \
  g_api.field = reinterpret_cast<decltype(g_api.field)>(dlsym(h, name));          \
  if (!g_api.field) return fail("fa_b200_nccl: missing symbol", name)
  FA_SYM(GetUniqueId, "ncclGetUniqueId");
  FA_SYM(CommInitRank, "ncclCommInitRank");
  FA_SYM(AllReduce, "ncclAllReduce");
  FA_SYM(CommDestroy, "ncclCommDestroy");
  FA_SYM(GetErrorString, "ncclGetErrorString");
  FA_SYM(GetVersion, "ncclGetVersion");
#undef FA_SYM
  g_api.handle = h;
  return FA_OK;
}

int check(ncclResult_t r, const char* what) { return r == 0 ? FA_OK : fail(what, g_api.GetErrorString(r)); }

}  // namespace

extern "C" {

const char* fa_b200_nccl_last_error(void) { return g_nccl_err; }

int fa_b200_nccl_version(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (load_api() != FA_OK) return -1;
  int v = 0;
  return g_api.GetVersion(&v) == 0 ? v : -1;
}

int fa_b200_nccl_unique_id(void* id_out) {
  if (!id_out) return fail("fa_b200_nccl_unique_id", "null pointer");
  std::lock_guard<std::mutex> lk(g_mu);
  if (int rc = load_api()) return rc;
  ncclUniqueId id;
  if (int rc = check(g_api.GetUniqueId(&id), "ncclGetUniqueId")) return rc;
  memcpy(id_out, &id, FA_B200_NCCL_UNIQUE_ID_BYTES);
  return FA_OK;
}

int fa_b200_nccl_init(const void* id, int rank, int world) {
  if (!id || world <= 0 || rank < 0 || rank >= world) return fail("fa_b200_nccl_init", "bad argument");
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_comm) return fail("fa_b200_nccl_init", "communicator already initialised (call fa_b200_nccl_destroy first)");
  if (int rc = load_api()) return rc;
  ncclUniqueId uid;
  memcpy(&uid, id, FA_B200_NCCL_UNIQUE_ID_BYTES);
  if (int rc = check(g_api.CommInitRank(&g_comm, world, uid, rank), "ncclCommInitRank")) {
    g_comm = nullptr;
    return rc;
  }
  g_rank = rank;
  g_world = world;
  return FA_OK;
}

int fa_b200_nccl_world(void) { return g_world; }
int fa_b200_nccl_rank(void) { return g_rank; }

int fa_b200_nccl_allreduce_f32(float* buf, size_t count, int average, void* stream) {
  if (!g_comm) return fail("fa_b200_nccl_allreduce_f32", "communicator not initialised");
  if (!buf && count) return fail("fa_b200_nccl_allreduce_f32", "null buffer");
  if (count == 0) return FA_OK;
  return check(g_api.AllReduce(buf, buf, count, kNcclFloat32, average ? kNcclAvg : kNcclSum, g_comm,
                               static_cast<cudaStream_t>(stream)),
               "ncclAllReduce");
}

int fa_b200_nccl_destroy(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_comm) return FA_OK;
  ncclResult_t r = g_api.CommDestroy(g_comm);
  g_comm = nullptr;
  g_rank = -1;
  g_world = 0;
  return check(r, "ncclCommDestroy");
}

}  // extern "C"
