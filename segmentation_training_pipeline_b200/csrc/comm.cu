// Thin C-ABI wrappers over NCCL for the one exchange step of the path -- the SUM all-reduce of the flat fp32 gradient
// (replaces what keras.utils.multi_gpu_model does on the CPU in the reference, FAQ.md:108-112).
// NCCL is bound at RUN TIME (dlopen of libnccl.so.2: the copy torch already loaded, else the system one), so libstp.so keeps
// loading on machines without NCCL; the Python engine itself uses torch.distributed for bootstrap and collectives (ddp.py),
// these entry points serve non-PyTorch hosts of the C ABI.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace stp {
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;   // ncclSuccess == 0
enum { kNcclFloat32 = 7, kNcclSum = 0 };

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (api.h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.h, "ncclAllReduce");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) api.h = nullptr;
    }
  }
  return api.h ? &api : nullptr;
}

int nccl_fail(const char* what, ncclResult_t r) {
  NcclApi* a = nccl();
  set_error("%s: NCCL error %d (%s)", what, r, (a && a->GetErrorString) ? a->GetErrorString(r) : "?");
  return STP_E_CUDA;
}

}  // namespace
}  // namespace stp

using namespace stp;

struct stp_comm {
  ncclComm_t comm;
  int world, rank;
};

extern "C" int stp_comm_unique_id(void* out128) {
  STP_REQUIRE(out128, "comm_unique_id: null");
  NcclApi* a = nccl();
  if (!a) {
    set_error("comm_unique_id: libnccl.so.2 not found");
    return STP_E_UNSUPPORTED;
  }
  ncclUniqueId id;
  ncclResult_t r = a->GetUniqueId(&id);
  if (r != 0) return nccl_fail("comm_unique_id", r);
  memcpy(out128, &id, sizeof(id));
  return STP_OK;
}

extern "C" int stp_comm_init(int32_t world, int32_t rank, const void* id128, stp_comm** out) {
  STP_REQUIRE(id128 && out && world >= 1 && rank >= 0 && rank < world, "comm_init: bad args");
  NcclApi* a = nccl();
  if (!a) {
    set_error("comm_init: libnccl.so.2 not found");
    return STP_E_UNSUPPORTED;
  }
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  ncclResult_t r = a->CommInitRank(&c, world, id, rank);   // the CALLER has selected the device (cudaSetDevice)
  if (r != 0) return nccl_fail("comm_init", r);
  *out = new stp_comm{c, world, rank};
  return STP_OK;
}

extern "C" int stp_allreduce(stp_comm* comm, float* d_buf, int64_t count, stp_stream stream) {
  STP_REQUIRE(comm && d_buf && count >= 0, "allreduce: bad args");
  if (comm->world == 1 || count == 0) return STP_OK;
  ncclResult_t r = nccl()->AllReduce(d_buf, d_buf, (size_t)count, kNcclFloat32, kNcclSum, comm->comm, (cudaStream_t)stream);
  if (r != 0) return nccl_fail("allreduce", r);
  return STP_OK;
}

extern "C" int stp_comm_destroy(stp_comm* comm) {
  if (!comm) return STP_OK;
  NcclApi* a = nccl();
  if (a && comm->comm) a->CommDestroy(comm->comm);
  delete comm;
  return STP_OK;
}
