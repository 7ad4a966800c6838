// comm.cpp -- NCCL plumbing for the distributed path (one process per GPU): ghost-value halo exchange by
// grouped ncclSend/ncclRecv and small fp64 all-reduces for the Krylov dot products.
//
// Replaces the MPI traffic of the reference's hot path: the VecScatter ghost update
// inside KSP / system.update() and the MPI_Allreduce of VecMDot / VecNorm.  The matrix/rhs stash exchange
// of MatAssemblyBegin/End and the whole-vector all-gather of check_solution have no counterpart: every
// rank assembles all elements touching its owned rows and clamps locally.
//
// NCCL is resolved with dlopen at rdc_create_distributed time so that the single-GPU path has no
// dependency on it; when torch is loaded in the same process its bundled libnccl.so.2 is reused.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "rdc_internal.h"

namespace rdc {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* load_nccl(std::string& err) {
  static NcclApi api;
  if (api.lib) return &api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return nullptr; }
#define SYM(field, name)                                                       \
  *(void**)(&api.field) = dlsym(api.lib, name);                                \
  if (!api.field) { err = std::string("missing NCCL symbol ") + name; api.lib = nullptr; return nullptr; }
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  return &api;
}

#define RDC_NCCL(call)                                                                           \
  do {                                                                                           \
    ncclResult_t r_ = (call);                                                                    \
    if (r_ != ncclSuccess) {                                                                     \
      c->err = std::string(#call) + ": " + c->nccl->GetErrorString(r_);                          \
      return RDC_E_COMM;                                                                         \
    }                                                                                            \
  } while (0)

int comm_unique_id(void* out128, std::string& err) {
  NcclApi* api = load_nccl(err);
  if (!api) return RDC_E_COMM;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) { err = "ncclGetUniqueId failed"; return RDC_E_COMM; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return 0;
}

int comm_init(rdc_ctx* c, const void* uid, std::string& err) {
  c->nccl = load_nccl(err);
  if (!c->nccl) return RDC_E_COMM;
  ncclUniqueId id;
  memcpy(&id, uid, 128);
  ncclComm_t comm;
  ncclResult_t r = c->nccl->CommInitRank(&comm, c->S.nranks, id, c->S.rank);
  if (r != ncclSuccess) { err = std::string("ncclCommInitRank: ") + c->nccl->GetErrorString(r); return RDC_E_COMM; }
  c->comm = (void*)comm;
  return 0;
}

void comm_destroy(rdc_ctx* c) {
  if (c->comm && c->nccl) c->nccl->CommDestroy((ncclComm_t)c->comm);
  c->comm = nullptr;
}

int launch_pack(rdc_ctx* c, const double* x, int ncomp);  // solver.cu

// fills the ghost part of x (nv values per node) from the owning ranks
int halo_exchange(rdc_ctx* c, double* x) {
  if (c->S.nranks == 1) return 0;
  const int nv = c->nv;
  int rc = launch_pack(c, x, nv);
  if (rc) return rc;
  ncclComm_t comm = (ncclComm_t)c->comm;
  RDC_NCCL(c->nccl->GroupStart());
  for (size_t k = 0; k < c->S.nbr_rank.size(); k++) {
    const int q = c->S.nbr_rank[k];
    const size_t ns = (size_t)(c->S.send_ptr[k + 1] - c->S.send_ptr[k]) * nv;
    const size_t nr = (size_t)(c->S.recv_ptr[k + 1] - c->S.recv_ptr[k]) * nv;
    if (ns) RDC_NCCL(c->nccl->Send(c->d_sendbuf + (size_t)c->S.send_ptr[k] * nv, ns, ncclDouble, q, comm, c->stream));
    if (nr) RDC_NCCL(c->nccl->Recv(x + ((size_t)c->S.n_owned + c->S.recv_ptr[k]) * nv, nr, ncclDouble, q, comm, c->stream));
  }
  RDC_NCCL(c->nccl->GroupEnd());
  return 0;
}

int allreduce_sum(rdc_ctx* c, double* d_buf, int n) {
  if (c->S.nranks == 1 || n == 0) return 0;
  RDC_NCCL(c->nccl->AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
  return 0;
}

int allreduce_max(rdc_ctx* c, double* d_buf, int n) {
  if (c->S.nranks == 1 || n == 0) return 0;
  RDC_NCCL(c->nccl->AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclMax, (ncclComm_t)c->comm, c->stream));
  return 0;
}

}  // namespace rdc
