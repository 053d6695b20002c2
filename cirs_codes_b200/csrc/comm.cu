// Multi-GPU plumbing of the update: one process per GPU, environments sharded over ranks, parameters
// replicated; the reference has no multi-GPU path.  The PPO learn loop runs inside ONE C call per update
// (cirs_ppo_learn): minibatch kernels -> gradient all-reduce -> clip + Adam, all stream-ordered, no interpreter
// between them.  The collective is NCCL's (ring / NVLS over NVLink 5 + NVSwitch): the library is the one the
// process already has loaded (torch's bundled libnccl.so.2), resolved with dlopen at the first cirs_comm_* call so
// that libcirs_b200.so has no link-time dependency on it; single-GPU users never touch it.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"
#include "../../include/cirs_b200.h"

namespace {

// the handful of NCCL prototypes used (nccl.h 2.2x: stable ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0 };
enum { ncclInt32 = 2, ncclFloat32 = 7, ncclFloat64 = 8 };

struct Nccl {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
Nccl g_nccl;

bool nccl_load() {
  if (g_nccl.h) return true;
  void* h = dlopen("libnccl.so.2", RTLD_LAZY | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_LAZY | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_LAZY | RTLD_GLOBAL);
  if (!h) {
    cirs_set_error("cirs_comm: libnccl.so.2 not found (import torch first, or put NCCL on the library path)");
    return false;
  }
  Nccl n;
  n.h = h;
  *(void**)&n.GetUniqueId = dlsym(h, "ncclGetUniqueId");
  *(void**)&n.CommInitRank = dlsym(h, "ncclCommInitRank");
  *(void**)&n.CommDestroy = dlsym(h, "ncclCommDestroy");
  *(void**)&n.AllReduce = dlsym(h, "ncclAllReduce");
  *(void**)&n.GetErrorString = dlsym(h, "ncclGetErrorString");
  *(void**)&n.GroupStart = dlsym(h, "ncclGroupStart");
  *(void**)&n.GroupEnd = dlsym(h, "ncclGroupEnd");
  if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllReduce || !n.GetErrorString || !n.GroupStart ||
      !n.GroupEnd) {
    cirs_set_error("cirs_comm: libnccl.so.2 lacks an expected symbol");
    return false;
  }
  g_nccl = n;
  return true;
}

struct Comm {
  ncclComm_t nccl;
  int rank, world;
};

int nccl_fail(const char* what, ncclResult_t r) {
  char msg[256];
  snprintf(msg, sizeof(msg), "%s: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
  cirs_set_error(msg);
  return CIRS_ERR_CUDA;
}

}  // namespace

// stream-ordered in-place sum over ranks; dtype: 0 float32, 1 float64, 2 int32 (used by ppo.cu as well)
int cirs_comm_allreduce_impl(void* comm, void* buf, int64_t count, int dtype, cudaStream_t st) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (!c || !buf || count < 0 || dtype < 0 || dtype > 2) {
    cirs_set_error("cirs_comm_allreduce: bad argument");
    return CIRS_ERR_ARG;
  }
  if (count == 0 || c->world == 1) return CIRS_OK;
  const int dt = dtype == 0 ? ncclFloat32 : (dtype == 1 ? ncclFloat64 : ncclInt32);
  const bool prof = cirs_profile_begin(dtype == 0 ? "nccl_allreduce_f32" : (dtype == 1 ? "nccl_allreduce_f64" : "nccl_allreduce_i32"), st);
  ncclResult_t r = g_nccl.AllReduce(buf, buf, (size_t)count, dt, ncclSum, c->nccl, st);
  if (prof) cirs_profile_end(st);
  if (r != 0) return nccl_fail("ncclAllReduce", r);
  return CIRS_OK;
}

extern "C" int cirs_comm_unique_id(void* id128_h) {
  if (!id128_h) {
    cirs_set_error("cirs_comm_unique_id: null argument");
    return CIRS_ERR_ARG;
  }
  if (!nccl_load()) return CIRS_ERR_CUDA;
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != 0) return nccl_fail("ncclGetUniqueId", r);
  memcpy(id128_h, &id, sizeof(id));
  return CIRS_OK;
}

extern "C" int cirs_comm_create(const void* id128_h, int32_t rank, int32_t world, void** comm_out) {
  if (!id128_h || !comm_out || world < 1 || rank < 0 || rank >= world) {
    cirs_set_error("cirs_comm_create: bad argument");
    return CIRS_ERR_ARG;
  }
  if (!nccl_load()) return CIRS_ERR_CUDA;
  ncclUniqueId id;
  memcpy(&id, id128_h, sizeof(id));
  Comm* c = new Comm{nullptr, rank, world};
  ncclResult_t r = g_nccl.CommInitRank(&c->nccl, world, id, rank);
  if (r != 0) {
    delete c;
    return nccl_fail("ncclCommInitRank", r);
  }
  *comm_out = c;
  return CIRS_OK;
}

extern "C" int cirs_comm_destroy(void* comm) {
  Comm* c = reinterpret_cast<Comm*>(comm);
  if (!c) return CIRS_OK;
  if (c->nccl && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl);
  delete c;
  return CIRS_OK;
}

// all-reduces issued between begin and end travel as ONE fused NCCL operation (ncclGroupStart / ncclGroupEnd)
extern "C" int cirs_comm_group_begin(void* comm) {
  if (!comm) return CIRS_OK;
  ncclResult_t r = g_nccl.GroupStart();
  return r ? nccl_fail("ncclGroupStart", r) : CIRS_OK;
}
extern "C" int cirs_comm_group_end(void* comm) {
  if (!comm) return CIRS_OK;
  ncclResult_t r = g_nccl.GroupEnd();
  return r ? nccl_fail("ncclGroupEnd", r) : CIRS_OK;
}

extern "C" int cirs_comm_allreduce(void* comm, void* buf, int64_t count, int32_t dtype, void* stream) {
  return cirs_comm_allreduce_impl(comm, buf, count, dtype, (cudaStream_t)stream);
}
