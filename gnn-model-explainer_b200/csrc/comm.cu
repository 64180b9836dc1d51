// comm.cu -- the one exchange step of the multi-GPU path: explained nodes are independent units dealt across ranks
// (one process per GPU), and the only collective is ONE all-gather of the packed edge masks at the end, over NVLink/NVSwitch.
// The reference has no distributed code (its explain_nodes loop is sequential, explain.py:225-236); this is the C-ABI form of
// "every rank ends up with every mask".
//
// NCCL is resolved at run time (dlopen "libnccl.so.2": the copy torch already loaded when the host program is PyTorch, the
// system library otherwise), so libgnnx.so has no link-time NCCL dependency and still loads on a box without it; the
// communicator is bootstrapped from a 128-byte id that the caller transports (torch.distributed broadcast, MPI, a file ...).
#include <dlfcn.h>
#include <string.h>

#include "gnnx_internal.cuh"

namespace {

// the slice of nccl.h this file needs (NCCL's ABI for these entry points is stable across 2.x)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat = 7;

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
  api.AllGather = (decltype(api.AllGather))dlsym(api.lib, "ncclAllGather");
  api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))dlsym(api.lib, "ncclGetVersion");
  if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather) { dlclose(api.lib); api.lib = nullptr; return nullptr; }
  return &api;
}

#define GX_NCCL_CHECK(expr)                                                                                      \
  do {                                                                                                           \
    ncclResult_t _r = (expr);                                                                                    \
    if (_r != 0) {                                                                                               \
      gx_set_error("%s failed: %s", #expr, N->GetErrorString ? N->GetErrorString(_r) : "nccl error");            \
      return GX_ERR_CUDA;                                                                                        \
    }                                                                                                            \
  } while (0)

// scatter of the gathered per-rank slots into global item order: item p (size sz[p]) lives at gathered[src[p] ..) and goes to out[dst[p] ..)
__global__ void __launch_bounds__(256)
unshard_kernel(const float* __restrict__ gathered, int items, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
               const int32_t* __restrict__ sz, float* __restrict__ out) {
  for (int p = blockIdx.x; p < items; p += gridDim.x) {
    const float* s = gathered + src[p];
    float* d = out + dst[p];
    for (int e = threadIdx.x; e < sz[p]; e += blockDim.x) d[e] = s[e];
  }
}

}  // namespace

struct GxComm {
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
};

int gx_comm_impl_unique_id(char* id128) {
  NcclApi* N = nccl();
  if (!N) { gx_set_error("gx_comm_unique_id: libnccl.so.2 not found (%s)", dlerror() ? dlerror() : "dlopen failed"); return GX_ERR_UNSUPPORTED; }
  ncclUniqueId id;
  GX_NCCL_CHECK(N->GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return GX_OK;
}

int gx_comm_impl_init(GxComm** out, int world, int rank, const char* id128) {
  NcclApi* N = nccl();
  if (!N) { gx_set_error("gx_comm_init: libnccl.so.2 not found"); return GX_ERR_UNSUPPORTED; }
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  GxComm* c = new GxComm();
  c->world = world; c->rank = rank;
  ncclResult_t r = N->CommInitRank(&c->comm, world, id, rank);
  if (r != 0) { gx_set_error("ncclCommInitRank failed: %s", N->GetErrorString ? N->GetErrorString(r) : "nccl error"); delete c; return GX_ERR_CUDA; }
  *out = c;
  return GX_OK;
}

void gx_comm_impl_destroy(GxComm* c) {
  if (!c) return;
  NcclApi* N = nccl();
  if (N && c->comm) N->CommDestroy(c->comm);
  delete c;
}

int gx_comm_impl_world(const GxComm* c) { return c ? c->world : 1; }
int gx_comm_impl_rank(const GxComm* c) { return c ? c->rank : 0; }

int gx_comm_impl_allgather(GxComm* c, const float* send, float* recv, size_t slot_floats, cudaStream_t s) {
  NcclApi* N = nccl();
  if (!N || !c) { gx_set_error("gx_allgather_masks: no communicator (call gx_comm_init)"); return GX_ERR_INVALID; }
  GX_NCCL_CHECK(N->AllGather(send, recv, slot_floats, kNcclFloat, c->comm, s));
  return GX_OK;
}

cudaError_t gx_launch_unshard(const float* gathered, int items, const int64_t* src, const int64_t* dst, const int32_t* sz, float* out, cudaStream_t s) {
  const int grid = items < 148 * 8 ? items : 148 * 8;
  unshard_kernel<<<grid > 0 ? grid : 1, 256, 0, s>>>(gathered, items, src, dst, sz, out);
  return cudaGetLastError();
}
