// Data-parallel gradient exchange of the dg_b200 C ABI: a thin wrapper over NCCL (NVLink 5 / NVSwitch) so that a host
// that is not PyTorch can run the one collective of the train step (one sum all-reduce of the flat generator /
// discriminator gradient arenas per step).  The reference is single-GPU (train_srgan.py:15); this is the multi-GPU seam.
//
// NCCL is resolved at run time with dlopen("libnccl.so.2"): the library has no link-time dependency on it (single-GPU
// users and the CPU-side symbol tests load libdg_b200.so without NCCL present), and inside a PyTorch process the already
// loaded NCCL (same soname) is the one that gets used -- one NCCL per process.  Only stable entry points are bound
// (ncclGetUniqueId, ncclCommInitRank, ncclAllReduce, ncclCommDestroy, ncclGetErrorString, ncclGroupStart/End).
//
// Buckets: parameters, gradients and Adam moments of a network live in ONE flat fp32 arena (params.py), so a bucket is a
// contiguous [offset, offset + count) range of the gradient arena -- the bucket_pack / bucket_unpack copies a per-variable
// layout would need do not exist here; dg_comm_allreduce takes the range directly.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "dg_common.cuh"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32 = 7, ncclSum = 0 };

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return 0;
  const char* names[] = {getenv("DG_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !n[0]) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) DG_FAIL("dg_comm: cannot load NCCL (libnccl.so.2): %s", dlerror());
#define DG_SYM(field, name)                                                  \
  *(void**)(&g_nccl.field) = dlsym(h, name);                                  \
  if (!g_nccl.field) { dlclose(h); DG_FAIL("dg_comm: NCCL lacks %s", name); }
  DG_SYM(GetUniqueId, "ncclGetUniqueId")
  DG_SYM(CommInitRank, "ncclCommInitRank")
  DG_SYM(AllReduce, "ncclAllReduce")
  DG_SYM(CommDestroy, "ncclCommDestroy")
  DG_SYM(GetErrorString, "ncclGetErrorString")
  DG_SYM(GroupStart, "ncclGroupStart")
  DG_SYM(GroupEnd, "ncclGroupEnd")
#undef DG_SYM
  g_nccl.handle = h;
  return 0;
}

}  // namespace

struct dg_comm {
  ncclComm_t comm;
  int rank, world, device;
};

extern "C" int dg_comm_unique_id_bytes(void) { return (int)sizeof(ncclUniqueId); }

extern "C" int dg_comm_unique_id(void* id_out) {
  DG_REQUIRE(id_out, "dg_comm_unique_id: null argument");
  if (load_nccl()) return 1;
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != 0) DG_FAIL("dg_comm_unique_id: %s", g_nccl.GetErrorString(r));
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

extern "C" int dg_comm_init(dg_comm** out, const void* unique_id, int rank, int world, int device) {
  DG_REQUIRE(out && unique_id && world >= 1 && rank >= 0 && rank < world, "dg_comm_init: bad argument (rank %d of %d)", rank, world);
  if (load_nccl()) return 1;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) DG_FAIL("dg_comm_init: cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclComm_t c = nullptr;
  ncclResult_t r = g_nccl.CommInitRank(&c, world, id, rank);
  if (r != 0) DG_FAIL("dg_comm_init: ncclCommInitRank: %s", g_nccl.GetErrorString(r));
  dg_comm* h = new dg_comm();
  h->comm = c; h->rank = rank; h->world = world; h->device = device;
  *out = h;
  return 0;
}

// In-place sum all-reduce of buf[0..count) (fp32, device memory) over all ranks, enqueued on `stream` (capturable into a CUDA
// graph like any NCCL call); returns immediately.
extern "C" int dg_comm_allreduce(dg_comm* comm, float* buf, long long count, void* stream) {
  DG_REQUIRE(comm && comm->comm && buf && count > 0, "dg_comm_allreduce: bad argument");
  ncclResult_t r = g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat32, ncclSum, comm->comm, (cudaStream_t)stream);
  if (r != 0) DG_FAIL("dg_comm_allreduce: ncclAllReduce: %s", g_nccl.GetErrorString(r));
  return 0;
}

extern "C" int dg_comm_rank(dg_comm* comm) { return comm ? comm->rank : -1; }
extern "C" int dg_comm_world(dg_comm* comm) { return comm ? comm->world : 0; }

extern "C" void dg_comm_destroy(dg_comm* comm) {
  if (!comm) return;
  if (comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm->comm);
  delete comm;
}
