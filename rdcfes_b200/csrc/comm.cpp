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
#include <stdlib.h>

#include <algorithm>

#include "rdc_internal.h"

namespace rdc {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
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
  SYM(AllGather, "ncclAllGather");
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
  if (c->p2p) {
    P2P* P = c->p2p;
    for (int q = 0; q < (int)P->peer.size(); q++)
      if (q != c->S.rank && P->peer[q]) cudaIpcCloseMemHandle(P->peer[q]);
    cudaFree(P->d_peer); cudaFree(P->d_scratch);
    if (c->comm && c->nccl && P->arena) {  // nobody unmaps an arena that a peer may still be writing to
      cudaStreamSynchronize(c->stream);
    }
    cudaFree(P->arena);
    delete P;
    c->p2p = nullptr;
  }
  if (c->comm && c->nccl) c->nccl->CommDestroy((ncclComm_t)c->comm);
  c->comm = nullptr;
}

static double* p2p_alloc_raw(P2P* P) {
  if (P->slots_used >= P->nslots) return nullptr;
  return (double*)(P->arena + P->header_bytes + (size_t)(P->slots_used++) * P->slot_bytes);
}

// ---- peer-memory set-up ------------------------------------------------------------------------------------
// Collective over the NCCL communicator: agrees on the arena geometry (slot = the largest local vector of any
// rank), exchanges the cudaIpc handles of the arenas and the table that tells a sender where its segment starts
// in each receiver's ghost tail.
// Protocol discipline: the three collectives (all-gather of the tables, all-gather of the handles, min all-reduce of
// the outcome) are issued by EVERY rank in the same order whatever happens locally -- all fallible local work only
// lowers the rank's `ok` flag -- and the transport is chosen from the all-reduced minimum alone, so the ranks can
// neither hang in a collective that a peer skipped nor end up on different transports.  The only early returns are
// hard errors (RDC_E_NOMEM for the 4 KB exchange buffer, RDC_E_COMM when NCCL itself fails): rdc_create fails then.
// A peer mapping is used for stores and polling loads only (tag-in-word slots, p2p_dev.cuh), which is correct over
// any cudaIpc peer mapping (NVLink or PCIe); RDC_P2P=0 forces the NCCL transport.
int p2p_init(rdc_ctx* c, std::string& note) {
  const char* e = getenv("RDC_P2P");
  const int nr = c->S.nranks, me = c->S.rank;
  if (nr < 2 || nr > RDC_MAX_RANKS || (e && atoi(e) == 0)) return 0;
  ncclComm_t comm = (ncclComm_t)c->comm;
  P2P* P = new P2P();
  c->p2p = P;
  const int NSLOT = 8;
  // table contributed by every rank: {n_owned, vec_len, ghosts received from rank 0..nr-1 (offsets), spare}
  const int TW = nr + 4, HW = 64 + 8;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::vector<long long> mine((size_t)TW, 0), all((size_t)TW * nr, 0);
  mine[0] = c->S.n_owned;
  mine[1] = (long long)c->S.n_loc * c->nv;
  {
    std::vector<long long> off((size_t)nr, -1);
    for (size_t k = 0; k < c->S.nbr_rank.size(); k++) off[c->S.nbr_rank[k]] = c->S.recv_ptr[k];
    for (int q = 0; q < nr; q++) mine[2 + q] = off[q];
  }
  // one device buffer for all three collectives
  unsigned char* d_x = nullptr;
  const size_t xbytes = std::max(sizeof(long long) * TW, (size_t)HW) * (nr + 1);
  if (cudaMalloc(&d_x, xbytes) != cudaSuccess) { c->err = "p2p: cannot allocate the exchange buffer"; return RDC_E_NOMEM; }
  auto hard = [&](const char* why) { cudaFree(d_x); c->err = why; return RDC_E_COMM; };
  // (1) tables
  long long* d_tab = (long long*)d_x;
  cudaMemcpyAsync(d_tab, mine.data(), sizeof(long long) * TW, cudaMemcpyHostToDevice, c->stream);
  if (c->nccl->AllGather(d_tab, d_tab + TW, (size_t)TW * sizeof(long long), ncclChar, comm, c->stream) != ncclSuccess)
    return hard("p2p: ncclAllGather failed");
  cudaMemcpyAsync(all.data(), d_tab + TW, sizeof(long long) * TW * nr, cudaMemcpyDeviceToHost, c->stream);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return hard("p2p: stream error after the table exchange");
  // local work: geometry, arena, handle, peer tables -- failures only clear `ok`
  int ok = 1;
  std::string why;
  auto soft = [&](const char* w) { if (ok) why = w; ok = 0; };
  long long max_len = 0;
  for (int q = 0; q < nr; q++) max_len = std::max(max_len, all[(size_t)q * TW + 1]);
  P->slot_bytes = (((size_t)max_len * sizeof(double)) + 255) / 256 * 256;
  P->nslots = NSLOT;
  P->arena_bytes = P->header_bytes + P->slot_bytes * NSLOT;
  if ((size_t)c->S.n_ghost * c->nv * 16 > P->slot_bytes) soft("ghost staging area larger than a vector slot");
  P->dst_node_off.clear();
  for (size_t k = 0; k < c->S.nbr_rank.size(); k++) {   // where my segment starts in each neighbour's ghost tail
    const long long off = all[(size_t)c->S.nbr_rank[k] * TW + 2 + me];
    if (off < 0 && c->S.send_ptr[k + 1] > c->S.send_ptr[k]) soft("inconsistent neighbour tables");
    P->dst_node_off.push_back(off < 0 ? 0 : off);
  }
  if (cudaMalloc((void**)&P->arena, P->arena_bytes) != cudaSuccess) { cudaGetLastError(); P->arena = nullptr; soft("cudaMalloc of the arena failed"); }
  if (cudaMalloc((void**)&P->d_peer, sizeof(void*) * nr) != cudaSuccess) { cudaGetLastError(); P->d_peer = nullptr; soft("cudaMalloc failed"); }
  if (cudaMalloc((void**)&P->d_scratch, sizeof(double) * 8) != cudaSuccess) { cudaGetLastError(); P->d_scratch = nullptr; soft("cudaMalloc failed"); }
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof(h));
  if (P->arena) {
    cudaMemsetAsync(P->arena, 0, P->arena_bytes, c->stream);
    if (cudaIpcGetMemHandle(&h, P->arena) != cudaSuccess) { cudaGetLastError(); soft("cudaIpcGetMemHandle failed"); }
  }
  // (2) handles (+ the local flag, so that nobody maps an arena of a rank that is giving up)
  std::vector<unsigned char> hmine((size_t)HW, 0), hall((size_t)HW * nr, 0);
  memcpy(hmine.data(), &h, 64);
  hmine[64] = (unsigned char)ok;
  cudaMemcpyAsync(d_x, hmine.data(), HW, cudaMemcpyHostToDevice, c->stream);
  if (c->nccl->AllGather(d_x, d_x + HW, (size_t)HW, ncclChar, comm, c->stream) != ncclSuccess) return hard("p2p: ncclAllGather failed");
  cudaMemcpyAsync(hall.data(), d_x + HW, (size_t)HW * nr, cudaMemcpyDeviceToHost, c->stream);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return hard("p2p: stream error after the handle exchange");
  for (int q = 0; q < nr; q++)
    if (!hall[(size_t)q * HW + 64]) soft("a peer could not set up its arena");
  P->peer.assign((size_t)nr, nullptr);
  if (ok) {
    for (int q = 0; q < nr; q++) {
      if (q == me) { P->peer[q] = P->arena; continue; }
      cudaIpcMemHandle_t hq;
      memcpy(&hq, hall.data() + (size_t)q * HW, 64);
      if (cudaIpcOpenMemHandle(&P->peer[q], hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        P->peer[q] = nullptr;
        soft("cudaIpcOpenMemHandle of a peer arena failed");
        break;
      }
    }
  }
  // (3) every rank takes the same path: min over the ranks of the local outcome
  {
    long long v = ok;
    cudaMemcpyAsync(d_tab, &v, sizeof(long long), cudaMemcpyHostToDevice, c->stream);
    if (c->nccl->AllReduce(d_tab, d_tab, 1, ncclInt64, ncclMin, comm, c->stream) != ncclSuccess) return hard("p2p: ncclAllReduce failed");
    cudaMemcpyAsync(&v, d_tab, sizeof(long long), cudaMemcpyDeviceToHost, c->stream);
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return hard("p2p: stream error after the outcome exchange");
    if (ok && !v) why = "a peer failed to map the arenas";
    ok = (int)v;
  }
  cudaFree(d_x);
  if (!ok) {   // agreed by all ranks: NCCL send/recv and all-reduce stay in use
    for (int q = 0; q < nr; q++)
      if (q != me && P->peer[q]) { cudaIpcCloseMemHandle(P->peer[q]); P->peer[q] = nullptr; }
    cudaFree(P->arena); cudaFree(P->d_peer); cudaFree(P->d_scratch);
    P->arena = nullptr; P->d_peer = nullptr; P->d_scratch = nullptr;
    P->peer.clear();
    note = "p2p off (" + why + "); NCCL transport";
    return 0;
  }
  // two staging areas for the tagged ghost exchange (16 B per ghost value), one per parity: NSLOT >= 2 by construction
  for (int par = 0; par < 2; par++) P->stage_off[par] = (size_t)((unsigned char*)p2p_alloc_raw(P) - P->arena);
  cudaMemcpyAsync(P->d_peer, P->peer.data(), sizeof(void*) * nr, cudaMemcpyHostToDevice, c->stream);
  cudaMemsetAsync(P->d_scratch, 0, sizeof(double) * 8, c->stream);
  cudaStreamSynchronize(c->stream);
  P->on = true;
  return 0;
}

double* p2p_alloc(rdc_ctx* c, size_t n_doubles) {
  P2P* P = c->p2p;
  if (!P || !P->on || P->slots_used >= P->nslots || n_doubles * sizeof(double) > P->slot_bytes) return nullptr;
  return (double*)(P->arena + P->header_bytes + (size_t)(P->slots_used++) * P->slot_bytes);
}

bool p2p_owns(const rdc_ctx* c, const void* p) {
  const P2P* P = c->p2p;
  return P && P->on && (const unsigned char*)p >= P->arena && (const unsigned char*)p < P->arena + P->arena_bytes;
}

int launch_pack(rdc_ctx* c, const double* x, int ncomp);  // solver.cu

// fills the ghost part of x (nv values per node) from the owning ranks
int halo_exchange(rdc_ctx* c, double* x, bool check_done) {
  if (c->S.nranks == 1) return 0;
  if (p2p_on(c)) return p2p_launch_halo(c, x, check_done);   // any device vector: only the staging areas are peer-mapped
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

int allreduce_sum(rdc_ctx* c, double* d_buf, int n, bool check_done) {
  if (c->S.nranks == 1 || n == 0) return 0;
  if (c->p2p && c->p2p->on && n <= 8) return p2p_launch_allreduce(c, d_buf, n, check_done);
  RDC_NCCL(c->nccl->AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream));
  return 0;
}

int allreduce_max(rdc_ctx* c, double* d_buf, int n) {
  if (c->S.nranks == 1 || n == 0) return 0;
  RDC_NCCL(c->nccl->AllReduce(d_buf, d_buf, (size_t)n, ncclDouble, ncclMax, (ncclComm_t)c->comm, c->stream));
  return 0;
}

}  // namespace rdc
