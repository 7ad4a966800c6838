// api.cu -- the C ABI of include/rdc.h.  No exceptions cross the boundary; every failure leaves a message
// for rdc_last_error().  There is no CPU fallback: without a CUDA device rdc_create fails.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

#include <algorithm>
#include <new>

#include "rdc_internal.h"

namespace rdc {
int pairs_per_cta_for(int model, int etype);  // assemble.cu
}
using namespace rdc;

static std::string g_create_err;

struct OptName { const char* name; int rdc_options::*field; };
static const OptName kOptions[] = {
    {"spmv_tma", &rdc_options::spmv_tma}, {"spmv_minb", &rdc_options::spmv_minb},
    {"spmv_ctas_per_sm", &rdc_options::spmv_ctas_per_sm}, {"tma_ctas_per_sm", &rdc_options::tma_ctas_per_sm},
    {"tma_stages", &rdc_options::tma_stages}, {"sync_every", &rdc_options::sync_every},
    {"p2p_fused_ar", &rdc_options::p2p_fused_ar}, {"p2p_fused_halo", &rdc_options::p2p_fused_halo},
    {"node_order", &rdc_options::node_order}, {"bicg_persist", &rdc_options::bicg_persist}, {"persist_timing", &rdc_options::persist_timing}, {"l2_evict_first", &rdc_options::l2_evict_first}, {"vec_reverse", &rdc_options::vec_reverse}, {"trace", &rdc_options::trace}};

static void options_from_env(rdc_options& o) {
  for (const OptName& k : kOptions) {
    std::string env = "RDC_";
    for (const char* p = k.name; *p; p++) env += (char)toupper(*p);
    if (const char* e = getenv(env.c_str())) o.*(k.field) = atoi(e);
  }
}

extern "C" int rdc_set_option(rdc_ctx* c, const char* name, int value) {
  if (!c || !name) return RDC_E_ARG;
  for (const OptName& k : kOptions)
    if (!strcmp(k.name, name)) { c->opt.*(k.field) = value; return RDC_OK; }
  c->err = std::string("rdc_set_option: unknown option ") + name;
  return RDC_E_ARG;
}

extern "C" const char* rdc_version(void) { return "rdcfes_b200 0.1 (sm_100a)"; }

extern "C" int rdc_model_nvars(int model) {
  switch (model) {
    case RDC_ADPM: case RDC_RIPF: case RDC_HCC: case RDC_SOLID: return 3;
    case RDC_PIHNA: case RDC_PROTEAS: return 5;
  }
  return -1;
}
extern "C" int rdc_model_nparams(int model) {
  switch (model) {
    case RDC_ADPM: return ADPM_NPARAMS;
    case RDC_PIHNA: return PIHNA_NPARAMS;
    case RDC_RIPF: return RIPF_NPARAMS;
    case RDC_PROTEAS: return PROTEAS_NPARAMS;
    case RDC_HCC: return HCC_NPARAMS;
    case RDC_SOLID: return 0;   // materials and boundary conditions go through rdc_solid_set_*
  }
  return -1;
}

extern "C" const char* rdc_last_error(const rdc_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }
extern "C" void rdc_free(void* p) { free(p); }

template <class T>
static int upload(rdc_ctx* c, T** dst, const std::vector<T>& src) {
  const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T) + 64;  // slack: bulk copies round up to 16 B
  RDC_CUDA(cudaMalloc((void**)dst, bytes));
  if (!src.empty()) RDC_CUDA(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream));
  return 0;
}

static int create_impl(rdc_ctx** out, int model, int elem_type, int64_t N, int64_t E, const int32_t* conn, const double* xyz,
                       const int32_t* node_dof_base, int device, int rank, int nranks, int partitioner, const void* uid) {
  if (!out) return RDC_E_ARG;
  *out = nullptr;
  const int nv = rdc_model_nvars(model);
  if (nv < 0 || (elem_type != RDC_TET4 && elem_type != RDC_HEX8) || !conn || !xyz || nranks < 1 || rank < 0 || rank >= nranks) {
    g_create_err = "rdc_create: invalid argument";
    return RDC_E_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    g_create_err = "rdc_create: no CUDA device available (this library has no CPU path)";
    return RDC_E_NODEVICE;
  }
  if (device >= 0) {
    if (device >= ndev || cudaSetDevice(device) != cudaSuccess) { g_create_err = "rdc_create: cannot select the CUDA device"; return RDC_E_NODEVICE; }
  } else if (cudaGetDevice(&device) != cudaSuccess) {
    g_create_err = "rdc_create: cudaGetDevice failed";
    return RDC_E_NODEVICE;
  }
  rdc_ctx* c = new (std::nothrow) rdc_ctx();
  if (!c) return RDC_E_NOMEM;
  c->model = model; c->etype = elem_type; c->nen = elem_type == RDC_TET4 ? 4 : 8; c->nv = nv;
  c->nqp = elem_type == RDC_TET4 ? 5 : 8;
  c->device = device;
  options_from_env(c->opt);
  c->kmask = model_kmask(model);
  c->nkv = __builtin_popcount(c->kmask);
  if (!spmv_masks_ok()) { g_create_err = "internal: SpMV entry masks differ from the model definitions"; delete c; return RDC_E_STATE; }
  int rc = 0;
  auto fail = [&](int code) { g_create_err = c->err; rdc_destroy(c); return code; };
  try {
    rc = build_setup(c->S, elem_type, nv, N, E, conn, xyz, rank, nranks, partitioner, pairs_per_cta_for(model, elem_type), c->opt.node_order, c->err);
  } catch (const std::bad_alloc&) { c->err = "out of host memory in set-up"; rc = RDC_E_NOMEM; }
  if (rc) return fail(rc);
  HostSetup& S = c->S;
  c->D_glob = N * nv;
  if (c->D_glob > 0x7fffffff) { c->err = "more than 2^31 dofs are not supported by the 32-bit dof map"; return fail(RDC_E_ARG); }
  if (node_dof_base) {
    c->dof_base.assign(node_dof_base, node_dof_base + N);
    c->identity_dofs = true;
    for (int64_t n = 0; n < N; n++) {
      if (node_dof_base[n] < 0 || (int64_t)node_dof_base[n] + nv > c->D_glob) { c->err = "node_dof_base out of range"; return fail(RDC_E_MESH); }
      if (node_dof_base[n] != (int32_t)(n * nv)) c->identity_dofs = false;
    }
  }
  if (nranks > 1) c->identity_dofs = false;
  for (int32_t l = 0; l < c->S.n_owned && c->identity_dofs; l++)
    if (c->S.loc2glob[l] != l) c->identity_dofs = false;   // Morton numbering: user vectors go through the dof map

  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { c->err = "cudaStreamCreate failed"; return fail(RDC_E_CUDA); }
  cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1);
  for (int k = 0; k < 2; k++) { cudaEventCreate(&c->ev_asm[k]); cudaEventCreate(&c->ev_sol[k]); cudaEventCreate(&c->ev_clamp[k]); }
  if (upload_fe_tables()) { c->err = "cannot upload the reference-element tables"; return fail(RDC_E_CUDA); }

  if (nranks > 1) {  // communicator and peer-memory arena first: the exchanged vectors live inside the arena
    if (!uid) { c->err = "rdc_create_distributed needs the NCCL unique id"; return fail(RDC_E_ARG); }
    if ((rc = comm_init(c, uid, c->err))) return fail(rc);
    std::string p2p_note;
    if ((rc = p2p_init(c, p2p_note))) return fail(rc);
    if (getenv("RDC_VERBOSE")) fprintf(stderr, "[rdc rank %d] peer-memory exchange: %s %s\n", rank, (c->p2p && c->p2p->on) ? "on" : "off", p2p_note.c_str());
  }
  auto vec_alloc = [&](double** dst, size_t n) -> int {  // arena slot when available (ghost exchange over NVLink stores)
    double* p = p2p_alloc(c, n);
    if (p) { *dst = p; return 0; }
    RDC_CUDA(cudaMalloc(dst, n * sizeof(double)));
    return 0;
  };
  auto go = [&]() -> int {
    int r;
    if ((r = upload(c, &c->d_conn, S.conn))) return r;
    if ((r = upload(c, &c->d_n2e_ptr, S.n2e_ptr))) return r;
    if ((r = upload(c, &c->d_pair, S.pair_rec))) return r;
    if ((r = upload(c, &c->d_rowptr, S.rowptr))) return r;
    if ((r = upload(c, &c->d_col, S.col))) return r;
    if ((r = upload(c, &c->d_diag_blk, S.diag_blk))) return r;
    {  // 48-byte CTA descriptors {node0, nnode, pair0, npairs | blk0, nblk, contributor base, first task | ntask, 0, 0, 0}
      const int32_t nc = (int32_t)S.cta_node.size() - 1;
      std::vector<int32_t> desc((size_t)nc * 12, 0);
      for (int32_t k = 0; k < nc; k++) {
        const int32_t a = S.cta_node[k], b = S.cta_node[k + 1];
        int32_t* d = desc.data() + (size_t)k * 12;
        d[0] = a; d[1] = b - a; d[2] = S.n2e_ptr[a]; d[3] = S.n2e_ptr[b] - S.n2e_ptr[a];
        d[4] = S.rowptr[a]; d[5] = S.rowptr[b] - S.rowptr[a]; d[6] = S.cptr[S.rowptr[a]];
        d[7] = S.task_ptr[k]; d[8] = S.task_ptr[k + 1] - S.task_ptr[k];
      }
      if ((r = upload(c, &c->d_cta_node, desc))) return r;
    }
    if ((r = upload(c, &c->d_task, S.task))) return r;
    if ((r = upload(c, &c->d_clist, S.clist))) return r;
    c->ncta = (int)S.cta_node.size() - 1;
    c->nnzb = S.rowptr[S.n_owned];
    // padded coordinates
    std::vector<double> x4((size_t)S.n_loc * 4, 0.0);
    for (int32_t l = 0; l < S.n_loc; l++)
      for (int d = 0; d < 3; d++) x4[(size_t)l * 4 + d] = xyz[(size_t)S.loc2glob[l] * 3 + d];
    if ((r = upload(c, &c->d_xyz, x4))) return r;
    // dof map of the local dofs
    std::vector<int32_t> dm((size_t)S.n_loc * nv);
    for (int32_t l = 0; l < S.n_loc; l++) {
      const int32_t g = S.loc2glob[l];
      const int32_t base = c->dof_base.empty() ? g * nv : c->dof_base[g];
      for (int a = 0; a < nv; a++) dm[(size_t)l * nv + a] = base + a;
    }
    if ((r = upload(c, &c->d_dofmap, dm))) return r;
    {  // local nodes sorted by the global dof they map to: zero-copy transfers walk host memory in ascending order
      std::vector<int32_t> by((size_t)S.n_loc);
      for (int32_t l = 0; l < S.n_loc; l++) by[l] = l;
      std::sort(by.begin(), by.end(), [&](int32_t a, int32_t b) { return dm[(size_t)a * nv] < dm[(size_t)b * nv]; });
      if ((r = upload(c, &c->d_node_by_glob, by))) return r;
    }
    const size_t vb = (size_t)S.n_loc * nv * sizeof(double);
    if ((r = vec_alloc(&c->d_u, (size_t)S.n_loc * nv))) return r;
    RDC_CUDA(cudaMalloc(&c->d_uold, vb)); RDC_CUDA(cudaMalloc(&c->d_uolder, vb));
    RDC_CUDA(cudaMemsetAsync(c->d_u, 0, vb, c->stream)); RDC_CUDA(cudaMemsetAsync(c->d_uold, 0, vb, c->stream));
    RDC_CUDA(cudaMemsetAsync(c->d_uolder, 0, vb, c->stream));
    RDC_CUDA(cudaMalloc(&c->d_rhs, (size_t)S.n_owned * nv * sizeof(double)));
    RDC_CUDA(cudaMalloc(&c->d_dinv, (size_t)S.n_owned * nv * sizeof(double)));
    RDC_CUDA(cudaMalloc(&c->d_val, (size_t)c->nnzb * c->nkv * sizeof(double) + 64));
    RDC_CUDA(cudaMalloc(&c->d_stage, (size_t)c->D_glob * sizeof(double)));
    if (model == RDC_RIPF) {
      if ((r = vec_alloc(&c->d_td, (size_t)S.n_loc * 3))) return r;
      RDC_CUDA(cudaMemsetAsync(c->d_td, 0, (size_t)S.n_loc * 3 * sizeof(double), c->stream));
      RDC_CUDA(cudaMalloc(&c->d_prev, (size_t)S.n_owned * 3 * sizeof(double)));
    }
    if (nranks > 1) {
      if ((r = upload(c, &c->d_send_idx, S.send_idx))) return r;
      RDC_CUDA(cudaMalloc(&c->d_sendbuf, std::max<size_t>(S.send_idx.size(), 1) * nv * sizeof(double)));
    }
    if ((r = solver_init(c))) return r;
    RDC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  };
  if ((rc = go())) return fail(rc);
  // algorithmic byte counts, local to this rank
  {
    const int64_t Nl = S.n_loc, El = S.E_loc, nnzb = c->nnzb, No = S.n_owned;
    const int f_e = model == RDC_ADPM ? 3 : 0;
    const int f_n = model == RDC_RIPF ? 3 : (model == RDC_PROTEAS ? 2 : 0);
    // operator entries are counted in the format actually traversed: NKV stored planes per block, not v*v
    const int64_t nkv = c->nkv;
    c->st.bytes_assemble = 4LL * c->nen * El + 24LL * Nl + 8LL * nv * Nl + 8LL * f_e * El + 8LL * f_n * Nl + 8LL * nkv * nnzb + 8LL * nv * No;
    c->st.bytes_spmv = nnzb * (8LL * nkv + 4) + 4LL * (No + 1) + 16LL * nv * No;
    c->st.bytes_index = 4LL * (int64_t)S.pair_rec.size() + 4LL * (No + 1) * 2 + 48LL * ((int64_t)S.cta_node.size() - 1) + 4LL * (int64_t)S.task.size() +
                        2LL * (int64_t)S.clist.size();
    c->st.n_nodes_local = No; c->st.n_nodes_ghost = S.n_ghost; c->st.n_elems_local = El; c->st.nnzb_local = nnzb;
  }
  // the host copies of the big maps are no longer needed
  std::vector<int32_t>().swap(S.pair); std::vector<int32_t>().swap(S.cptr); std::vector<uint16_t>().swap(S.clist);
  std::vector<int32_t>().swap(S.task); std::vector<int32_t>().swap(S.pair_rec);
  std::vector<int32_t>().swap(S.conn);
  *out = c;
  return RDC_OK;
}

extern "C" int rdc_create(rdc_ctx** out, int model, int elem_type, int64_t n_nodes, int64_t n_elems, const int32_t* conn,
                          const double* xyz, const int32_t* node_dof_base, int device) {
  return create_impl(out, model, elem_type, n_nodes, n_elems, conn, xyz, node_dof_base, device, 0, 1, 0, nullptr);
}

extern "C" int rdc_create_distributed(rdc_ctx** out, int model, int elem_type, int64_t n_nodes, int64_t n_elems,
                                      const int32_t* conn, const double* xyz, const int32_t* node_dof_base, int device, int rank,
                                      int nranks, int partitioner, const void* nccl_unique_id) {
  return create_impl(out, model, elem_type, n_nodes, n_elems, conn, xyz, node_dof_base, device, rank, nranks, partitioner,
                     nccl_unique_id);
}

extern "C" int rdc_comm_unique_id(void* out128) {
  if (!out128) return RDC_E_ARG;
  return comm_unique_id(out128, g_create_err);
}

extern "C" void rdc_destroy(rdc_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  solver_free(c);
  region_free(c);
  solid_free(c);
  if (p2p_owns(c, c->d_u)) c->d_u = nullptr;
  if (p2p_owns(c, c->d_td)) c->d_td = nullptr;
  comm_destroy(c);
  cudaFree(c->d_conn); cudaFree(c->d_xyz); cudaFree(c->d_efield); cudaFree(c->d_n2e_ptr); cudaFree(c->d_pair);
  cudaFree(c->d_rowptr); cudaFree(c->d_col); cudaFree(c->d_diag_blk); cudaFree(c->d_cta_node); cudaFree(c->d_task);
  cudaFree(c->d_clist); cudaFree(c->d_dofmap); cudaFree(c->d_node_by_glob); cudaFree(c->d_val); cudaFree(c->d_rhs); cudaFree(c->d_dinv);
  cudaFree(c->d_u); cudaFree(c->d_uold); cudaFree(c->d_uolder); cudaFree(c->d_stage); cudaFree(c->d_td); cudaFree(c->d_rt);
  cudaFree(c->d_prev); cudaFree(c->d_aux); cudaFree(c->d_send_idx); cudaFree(c->d_sendbuf);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  for (int k = 0; k < 2; k++) {
    if (c->ev_asm[k]) cudaEventDestroy(c->ev_asm[k]);
    if (c->ev_sol[k]) cudaEventDestroy(c->ev_sol[k]);
    if (c->ev_clamp[k]) cudaEventDestroy(c->ev_clamp[k]);
  }
  if (c->stream && c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

#define CHECK_CTX(c)            \
  if (!(c)) return RDC_E_ARG;   \
  cudaSetDevice((c)->device)

extern "C" int64_t rdc_n_dofs(const rdc_ctx* c) { return c ? c->D_glob : -1; }

extern "C" int rdc_set_stream(rdc_ctx* c, void* s) {
  CHECK_CTX(c);
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)s;
  c->own_stream = false;
  return RDC_OK;
}

extern "C" int rdc_set_params(rdc_ctx* c, const double* p, int n) {
  CHECK_CTX(c);
  if (!p || n != rdc_model_nparams(c->model)) { c->err = "rdc_set_params: wrong parameter count for this model"; return RDC_E_ARG; }
  c->params.assign(p, p + n);
  c->have_params = true;
  return RDC_OK;
}

extern "C" int rdc_set_time(rdc_ctx* c, double t) {
  CHECK_CTX(c);
  c->time = t;
  return RDC_OK;
}

extern "C" int rdc_set_elem_field(rdc_ctx* c, int slot, const double* f, int ncomp) {
  CHECK_CTX(c);
  if (c->model != RDC_ADPM || slot != 0 || ncomp != 3 || !f) { c->err = "rdc_set_elem_field: only ADPM slot 0 with 3 components exists"; return RDC_E_ARG; }
  const HostSetup& S = c->S;
  std::vector<double> loc((size_t)S.E_loc * 3);
  for (int64_t le = 0; le < S.E_loc; le++)
    for (int d = 0; d < 3; d++) loc[(size_t)le * 3 + d] = f[(size_t)S.elem_glob[le] * 3 + d];
  if (!c->d_efield) RDC_CUDA(cudaMalloc(&c->d_efield, loc.size() * sizeof(double)));
  RDC_CUDA(cudaMemcpyAsync(c->d_efield, loc.data(), loc.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  return RDC_OK;
}

extern "C" int rdc_set_nodal_field(rdc_ctx* c, int slot, const double* f, int ncomp) {
  CHECK_CTX(c);
  const HostSetup& S = c->S;
  if (slot != 0 || !f || ncomp != 2 || (c->model != RDC_RIPF && c->model != RDC_PROTEAS)) {
    c->err = "rdc_set_nodal_field: slot 0 with 2 components exists for RIPF (RT broad/focus) and PROTEAS (AUX)";
    return RDC_E_ARG;
  }
  if (c->model == RDC_RIPF) {
    std::vector<double> loc((size_t)S.n_loc * 3, 0.0);  // {broad, focus, total}; total is written by rdc_clamp
    for (int32_t l = 0; l < S.n_loc; l++) {
      loc[(size_t)l * 3] = f[(size_t)S.loc2glob[l] * 2];
      loc[(size_t)l * 3 + 1] = f[(size_t)S.loc2glob[l] * 2 + 1];
    }
    if (!c->d_rt) RDC_CUDA(cudaMalloc(&c->d_rt, loc.size() * sizeof(double)));
    RDC_CUDA(cudaMemcpyAsync(c->d_rt, loc.data(), loc.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  } else {
    std::vector<double> loc((size_t)S.n_loc * 2);
    for (int32_t l = 0; l < S.n_loc; l++) {
      loc[(size_t)l * 2] = f[(size_t)S.loc2glob[l] * 2];
      loc[(size_t)l * 2 + 1] = f[(size_t)S.loc2glob[l] * 2 + 1];
    }
    if (!c->d_aux) RDC_CUDA(cudaMalloc(&c->d_aux, loc.size() * sizeof(double)));
    RDC_CUDA(cudaMemcpyAsync(c->d_aux, loc.data(), loc.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  return RDC_OK;
}

extern "C" int rdc_update_coords(rdc_ctx* c, const double* xyz) {
  CHECK_CTX(c);
  if (!xyz) return RDC_E_ARG;
  const HostSetup& S = c->S;
  std::vector<double> x4((size_t)S.n_loc * 4, 0.0);
  for (int32_t l = 0; l < S.n_loc; l++)
    for (int d = 0; d < 3; d++) x4[(size_t)l * 4 + d] = xyz[(size_t)S.loc2glob[l] * 3 + d];
  RDC_CUDA(cudaMemcpyAsync(c->d_xyz, x4.data(), x4.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  c->assembled = false;
  return RDC_OK;
}

// true when `p` is pinned / registered host memory that kernels can address directly (zero-copy over PCIe)
static bool host_is_mapped(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost && a.devicePointer != nullptr;
}

// host vector in global dof order -> local device vector (owned + ghosts)
static int to_device(rdc_ctx* c, const double* host, double* d_loc) {
  if (c->identity_dofs) {
    RDC_CUDA(cudaMemcpyAsync(d_loc, host, (size_t)c->D_glob * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return 0;
  }
  if (c->S.nranks > 1 && host_is_mapped(host)) {
    // a rank needs only its own n_loc*v entries: gather them straight out of the pinned host buffer instead of
    // staging the whole global vector (distributed runs move 1/nranks of the bytes)
    cudaPointerAttributes a;
    cudaPointerGetAttributes(&a, host);
    return launch_gather(c, (const double*)a.devicePointer, d_loc);
  }
  RDC_CUDA(cudaMemcpyAsync(c->d_stage, host, (size_t)c->D_glob * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  return launch_gather(c, c->d_stage, d_loc);
}
// local device vector (owned part) -> host vector in global dof order (all ranks get everything)
static int to_host(rdc_ctx* c, const double* d_loc, double* host) {
  if (c->identity_dofs) {
    RDC_CUDA(cudaMemcpyAsync(host, d_loc, (size_t)c->D_glob * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    RDC_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  }
  if (c->S.nranks > 1) RDC_CUDA(cudaMemsetAsync(c->d_stage, 0, (size_t)c->D_glob * sizeof(double), c->stream));
  int rc = launch_scatter(c, d_loc, c->d_stage);
  if (rc) return rc;
  if (c->S.nranks > 1 && (rc = allreduce_sum(c, c->d_stage, (int)c->D_glob))) return rc;
  RDC_CUDA(cudaMemcpyAsync(host, c->d_stage, (size_t)c->D_glob * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  return p2p_check_error(c);
}

extern "C" int rdc_set_solution(rdc_ctx* c, const double* u) {
  CHECK_CTX(c);
  if (!u) return RDC_E_ARG;
  int rc = to_device(c, u, c->d_u);
  if (rc) return rc;
  c->u_ghost_fresh = true;     // the upload fills the ghost entries as well
  if (c->model == RDC_RIPF) {  // ripf.C:50-51: prev_soln starts as the initial solution
    RDC_CUDA(cudaMemcpyAsync(c->d_prev, c->d_u, (size_t)c->S.n_owned * 3 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->ripf_primed = false;
  }
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  return RDC_OK;
}

extern "C" int rdc_get_solution(rdc_ctx* c, double* u) {
  CHECK_CTX(c);
  if (!u) return RDC_E_ARG;
  return to_host(c, c->d_u, u);
}

// Only the entries of the dofs OWNED by this rank are written (global dof indexing, the rest of u is left alone):
// what a distributed caller needs to refresh its own part of a PETSc-style vector without the all-gather.
extern "C" int rdc_get_solution_owned(rdc_ctx* c, double* u) {
  CHECK_CTX(c);
  if (!u) return RDC_E_ARG;
  if (c->S.nranks == 1) return to_host(c, c->d_u, u);
  if (host_is_mapped(u)) {
    cudaPointerAttributes a;
    cudaPointerGetAttributes(&a, u);
    int rc = launch_scatter(c, c->d_u, (double*)a.devicePointer);   // zero-copy stores into the pinned buffer
    if (rc) return rc;
    RDC_CUDA(cudaStreamSynchronize(c->stream));
    return RDC_OK;
  }
  const size_t n = (size_t)c->S.n_owned * c->nv;
  std::vector<double> tmp(n);
  RDC_CUDA(cudaMemcpyAsync(tmp.data(), c->d_u, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  for (int32_t l = 0; l < c->S.n_owned; l++) {
    const int32_t g = c->S.loc2glob[l];
    const int32_t base = c->dof_base.empty() ? g * c->nv : c->dof_base[g];
    for (int a2 = 0; a2 < c->nv; a2++) u[base + a2] = tmp[(size_t)l * c->nv + a2];
  }
  return RDC_OK;
}

extern "C" int rdc_get_rhs(rdc_ctx* c, double* rhs) {
  CHECK_CTX(c);
  if (!rhs) return RDC_E_ARG;
  if (!c->assembled) { c->err = "rdc_get_rhs: nothing assembled"; return RDC_E_STATE; }
  // d_rhs holds the owned rows only; to_host reads the owned part of a local vector
  return to_host(c, c->d_rhs, rhs);
}

extern "C" int rdc_get_old_solution(rdc_ctx* c, double* u) {
  CHECK_CTX(c);
  if (!u) return RDC_E_ARG;
  return to_host(c, c->d_uold, u);
}

extern "C" int rdc_rotate(rdc_ctx* c) {
  CHECK_CTX(c);
  int rc = refresh_u_ghosts(c);  // system.update(): ghosts of the current solution
  if (rc) return rc;
  std::swap(c->d_uolder, c->d_uold);  // older <- old
  RDC_CUDA(cudaMemcpyAsync(c->d_uold, c->d_u, (size_t)c->S.n_loc * c->nv * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return RDC_OK;
}

// Elapsed times of the phases are read from their event pairs lazily: non-blocking at the start of the next phase of
// the same kind (the events of the previous step completed long ago), blocking only in rdc_get_stats.
static void resolve_timings(rdc_ctx* c, bool wait) {
  auto one = [&](bool& pending, cudaEvent_t* ev, double& last, double& sum) {
    if (!pending) return;
    if (wait ? cudaEventSynchronize(ev[1]) != cudaSuccess : cudaEventQuery(ev[1]) != cudaSuccess) { cudaGetLastError(); return; }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess) { last = ms; sum += ms; }
    pending = false;
  };
  one(c->t_asm_pending, c->ev_asm, c->st.ms_assemble, c->st.sum_ms_assemble);
  one(c->t_sol_pending, c->ev_sol, c->st.ms_solve, c->st.sum_ms_solve);
  one(c->t_clamp_pending, c->ev_clamp, c->st.ms_clamp, c->st.sum_ms_clamp);
}

extern "C" int rdc_assemble(rdc_ctx* c, double time, double dt) {
  CHECK_CTX(c);
  if (c->model == RDC_SOLID) { c->err = "rdc_assemble: use rdc_solid_assemble / rdc_solid_newton on a solid context"; return RDC_E_STATE; }
  if (!c->have_params) { c->err = "rdc_assemble: rdc_set_params has not been called"; return RDC_E_STATE; }
  if (!(dt > 0.0)) { c->err = "rdc_assemble: dt must be positive"; return RDC_E_ARG; }
  c->time = time; c->dt = dt;
  // no host synchronisation here: the elapsed time is read when the statistics are asked for
  resolve_timings(c, false);
  cudaEventRecord(c->ev_asm[0], c->stream);
  int rc = launch_assemble(c);
  if (rc) return rc;
  cudaEventRecord(c->ev_asm[1], c->stream);
  c->t_asm_pending = true;
  c->assembled = true;
  return RDC_OK;
}

extern "C" int rdc_solve(rdc_ctx* c, int ksp, int pc, double rtol, int maxits, int restart, int* iterations, double* resnorm) {
  CHECK_CTX(c);
  if (!c->assembled) { c->err = "rdc_solve: the operator has not been assembled"; return RDC_E_STATE; }
  int its = 0;
  double res = 0;
  c->st.n_spmv = 0;
  resolve_timings(c, false);
  cudaEventRecord(c->ev_sol[0], c->stream);
  int rc = solver_solve(c, ksp, pc, rtol, maxits, restart, &its, &res);
  cudaEventRecord(c->ev_sol[1], c->stream);
  c->t_sol_pending = true;
  c->st.iterations = its;
  c->st.sum_iterations += its;
  c->st.sum_n_spmv += c->st.n_spmv;
  c->st.sum_ms_spmv += c->st.ms_spmv_total;
  c->st.n_solves++;
  c->st.resnorm = res;
  if (iterations) *iterations = its;
  if (resnorm) *resnorm = res;
  return rc;
}

extern "C" int rdc_clamp(rdc_ctx* c) {
  CHECK_CTX(c);
  if (c->model == RDC_SOLID) { c->err = "rdc_clamp: node positions are not clamped (no check_solution in solid_system.C)"; return RDC_E_STATE; }
  if (c->model == RDC_RIPF) {
    if (!c->have_params || !c->d_rt) { c->err = "rdc_clamp (RIPF): parameters and RT dose field are required"; return RDC_E_STATE; }
    if (!(c->dt > 0.0)) { c->err = "rdc_clamp (RIPF): time step unknown; call rdc_assemble/rdc_step or rdc_set_dt first"; return RDC_E_STATE; }
  }
  resolve_timings(c, false);
  cudaEventRecord(c->ev_clamp[0], c->stream);
  int rc = launch_clamp(c);
  if (rc) return rc;
  cudaEventRecord(c->ev_clamp[1], c->stream);
  c->t_clamp_pending = true;
  return RDC_OK;
}

extern "C" int rdc_set_dt(rdc_ctx* c, double dt) {
  CHECK_CTX(c);
  c->dt = dt;
  return RDC_OK;
}

extern "C" int rdc_step(rdc_ctx* c, double time, double dt, int ksp, int pc, double rtol, int maxits, int restart, int* iterations,
                        double* resnorm) {
  CHECK_CTX(c);
  int rc;
  if ((rc = rdc_rotate(c))) return rc;
  if ((rc = rdc_assemble(c, time, dt))) return rc;
  // BiCGStab runs as one cooperative launch that needs no host decision: the clamp is queued right behind it and the
  // host synchronises once per step, after both (RIPF's check_solution needs the host and keeps the plain order)
  if (ksp == RDC_KSP_BICGSTAB && c->model != RDC_RIPF) {
    resolve_timings(c, false);
    c->st.n_spmv = 0;
    cudaEventRecord(c->ev_sol[0], c->stream);
    rc = solver_persist_begin(c, pc, rtol, maxits);
    if (rc == 0) {
      cudaEventRecord(c->ev_sol[1], c->stream);
      c->t_sol_pending = true;
      int crc = rdc_clamp(c);
      int its = 0;
      double res = 0;
      rc = solver_persist_end(c, &its, &res);
      if (rc == RDC_E_DIVERGED && res == res) {
        // rho or omega vanished: continue with GMRES from the (clamped, still valid) iterate, then clamp again
        int its2 = 0;
        rc = rdc_solve(c, RDC_KSP_GMRES, pc, rtol, maxits, restart, &its2, &res);
        its += its2;
        if (!rc) crc = rdc_clamp(c);
      } else {
        c->st.sum_iterations += its;
        c->st.sum_n_spmv += c->st.n_spmv;
        c->st.sum_ms_spmv += c->st.ms_spmv_total;
        c->st.n_solves++;
      }
      c->st.iterations = its;
      c->st.resnorm = res;
      if (iterations) *iterations = its;
      if (resnorm) *resnorm = res;
      return rc ? rc : crc;
    }
    if (rc != 1) return rc;
  }
  if ((rc = rdc_solve(c, ksp, pc, rtol, maxits, restart, iterations, resnorm))) return rc;
  return rdc_clamp(c);
}

extern "C" int rdc_spmv(rdc_ctx* c, const double* x, double* y) {
  CHECK_CTX(c);
  if (!c->assembled || !x || !y) { c->err = "rdc_spmv: operator not assembled or null argument"; return RDC_E_STATE; }
  SolverWork* W = c->work;
  (void)W;
  double *dx = nullptr, *dy = nullptr;
  const size_t vb = (size_t)c->S.n_loc * c->nv * sizeof(double);
  RDC_CUDA(cudaMalloc(&dx, vb)); RDC_CUDA(cudaMalloc(&dy, vb));
  RDC_CUDA(cudaMemsetAsync(dy, 0, vb, c->stream));
  int rc = to_device(c, x, dx);
  if (!rc) rc = launch_spmv(c, dx, dy, nullptr);
  if (!rc) rc = to_host(c, dy, y);
  cudaFree(dx); cudaFree(dy);
  return rc;
}

extern "C" int rdc_bench_spmv(rdc_ctx* c, int reps, double* mean_ms) {
  CHECK_CTX(c);
  if (!c->assembled || reps < 1 || !mean_ms) return RDC_E_STATE;
  double* dy = nullptr;
  RDC_CUDA(cudaMalloc(&dy, (size_t)c->S.n_loc * c->nv * sizeof(double)));
  int rc = launch_spmv(c, c->d_u, dy, nullptr);  // warm-up
  cudaEventRecord(c->ev0, c->stream);
  for (int r = 0; r < reps && !rc; r++) rc = launch_spmv(c, c->d_u, dy, nullptr);
  cudaEventRecord(c->ev1, c->stream);
  cudaEventSynchronize(c->ev1);
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  *mean_ms = ms / reps;
  cudaFree(dy);
  return rc;
}

extern "C" int rdc_set_subdomains(rdc_ctx* c, const int32_t* region, int n_regions) {
  CHECK_CTX(c);
  return region_setup(c, region, n_regions);
}

extern "C" int rdc_region_volumes(rdc_ctx* c, int ncond, const struct rdc_range_cond* cond, double* vol) {
  CHECK_CTX(c);
  if (!cond || !vol) return RDC_E_ARG;
  int rc;
  if (!c->region && (rc = region_setup(c, nullptr, 1))) return rc;
  return region_volumes(c, ncond, cond, vol);
}

extern "C" int rdc_region_last_mean(rdc_ctx* c, int var, double* mean) {
  CHECK_CTX(c);
  if (!mean) return RDC_E_ARG;
  int rc;
  if (!c->region && (rc = region_setup(c, nullptr, 1))) return rc;
  return region_last_mean(c, var, mean);
}

extern "C" int rdc_bench_stream(rdc_ctx* c, int reps, int ctas_per_sm, double* mean_ms, int64_t* bytes) {
  CHECK_CTX(c);
  if (!c->assembled || reps < 1 || !mean_ms || ctas_per_sm < 1 || ctas_per_sm > 16) return RDC_E_STATE;
  int rc = launch_stream_probe(c, ctas_per_sm);  // warm-up
  cudaEventRecord(c->ev0, c->stream);
  for (int r = 0; r < reps && !rc; r++) rc = launch_stream_probe(c, ctas_per_sm);
  cudaEventRecord(c->ev1, c->stream);
  cudaEventSynchronize(c->ev1);
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  *mean_ms = ms / reps;
  if (bytes) *bytes = (int64_t)((size_t)c->nnzb * c->nkv / 2 * 16);
  return rc;
}

extern "C" int rdc_bench_dfma(rdc_ctx* c, double* tflops) {
  CHECK_CTX(c);
  if (!tflops) return RDC_E_ARG;
  const int iters = 20000;
  int rc = launch_dfma_probe(c, 1000);  // warm-up
  cudaEventRecord(c->ev0, c->stream);
  if (!rc) rc = launch_dfma_probe(c, iters);
  cudaEventRecord(c->ev1, c->stream);
  cudaEventSynchronize(c->ev1);
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  *tflops = 2.0 * 8.0 * iters * (148.0 * 8 * 256) / (ms * 1e-3) / 1e12;
  return rc;
}

extern "C" int rdc_bench_barrier(rdc_ctx* c, int reps, int ctas_per_sm, int mode, double* mean_us) {
  CHECK_CTX(c);
  if (reps < 1 || !mean_us || ctas_per_sm < 1 || ctas_per_sm > 8) return RDC_E_ARG;
  int rc = launch_barrier_probe(c, 10, ctas_per_sm, mode);  // warm-up
  cudaEventRecord(c->ev0, c->stream);
  if (!rc) rc = launch_barrier_probe(c, reps, ctas_per_sm, mode);
  cudaEventRecord(c->ev1, c->stream);
  cudaEventSynchronize(c->ev1);
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  *mean_us = 1e3 * ms / reps;
  return rc;
}

extern "C" int rdc_get_stats(rdc_ctx* c, struct rdc_stats* s) {
  if (!c || !s) return RDC_E_ARG;
  cudaSetDevice(c->device);
  resolve_timings(c, true);
  c->st.p2p_on = p2p_on(c) ? 1 : 0;
  c->st.p2p_fused = (p2p_on(c) && c->opt.p2p_fused_ar && c->opt.p2p_fused_halo) ? 1 : 0;
  *s = c->st;
  return RDC_OK;
}

// parity: expand the row-local block operator to scalar CSR in global dof numbering, sorted
extern "C" int rdc_download_csr(rdc_ctx* c, int64_t* n_rows, int64_t* nnz, int64_t** rows, int64_t** rowptr, int32_t** col,
                                double** val, double** rhs) {
  CHECK_CTX(c);
  if (!c->assembled) { c->err = "rdc_download_csr: nothing assembled"; return RDC_E_STATE; }
  const HostSetup& S = c->S;
  const int nv = c->nv, vv = nv * nv, nkv = c->nkv;
  const int32_t no = S.n_owned;
  int plane[25];  // entry (a,b) -> stored plane, -1 = structural zero (kept explicit in the reference's AIJ pattern)
  for (int ab = 0, s = 0; ab < vv; ab++) plane[ab] = (c->kmask >> ab & 1u) ? s++ : -1;
  std::vector<double> hval((size_t)c->nnzb * nkv), hrhs((size_t)no * nv);
  RDC_CUDA(cudaMemcpyAsync(hval.data(), c->d_val, hval.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  RDC_CUDA(cudaMemcpyAsync(hrhs.data(), c->d_rhs, hrhs.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  const int64_t nr = (int64_t)no * nv, nz = (int64_t)c->nnzb * vv;
  int64_t* o_rows = (int64_t*)malloc(sizeof(int64_t) * nr);
  int64_t* o_ptr = (int64_t*)malloc(sizeof(int64_t) * (nr + 1));
  int32_t* o_col = (int32_t*)malloc(sizeof(int32_t) * nz);
  double* o_val = (double*)malloc(sizeof(double) * nz);
  double* o_rhs = (double*)malloc(sizeof(double) * nr);
  if (!o_rows || !o_ptr || !o_col || !o_val || !o_rhs) { c->err = "out of host memory"; return RDC_E_NOMEM; }
  auto base_of = [&](int32_t l) { const int32_t g = S.loc2glob[l]; return c->dof_base.empty() ? g * nv : c->dof_base[g]; };
  // order the owned nodes by their dof base so that the rows come out sorted by global dof id
  std::vector<int32_t> order((size_t)no);
  for (int32_t i = 0; i < no; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return base_of(a) < base_of(b); });
  o_ptr[0] = 0;
  int64_t r = 0;
  std::vector<std::pair<int32_t, int32_t>> cols;  // (dof base of the column node, block slot)
  for (int32_t oi = 0; oi < no; oi++) {
    const int32_t i = order[oi];
    const int32_t r0 = S.rowptr[i], L = S.rowptr[i + 1] - r0;
    cols.clear();
    for (int32_t k = 0; k < L; k++) cols.emplace_back(base_of(S.col[r0 + k]), k);
    std::sort(cols.begin(), cols.end());
    for (int a = 0; a < nv; a++, r++) {
      o_rows[r] = base_of(i) + a;
      o_rhs[r] = hrhs[(size_t)i * nv + a];
      int64_t p = o_ptr[r];
      for (const auto& ck : cols)
        for (int b = 0; b < nv; b++, p++) {
          o_col[p] = ck.first + b;
          const int pl = plane[a * nv + b];
          o_val[p] = pl < 0 ? 0.0 : hval[(size_t)r0 * nkv + (size_t)pl * L + ck.second];
        }
      o_ptr[r + 1] = p;
    }
  }
  *n_rows = nr; *nnz = nz; *rows = o_rows; *rowptr = o_ptr; *col = o_col; *val = o_val; *rhs = o_rhs;
  return RDC_OK;
}

// Host-only probes of two pieces of set-up logic (no device needed; CPU tests check their invariants).
extern "C" int rdc_probe_spmv_tiles(int32_t n_rows, const int32_t* rowptr, int max_rows, int max_blocks, int32_t* n_tiles,
                                    int32_t** tiles) {
  if (!rowptr || !n_tiles || !tiles || n_rows < 0 || max_rows < 1 || max_blocks < 1) return RDC_E_ARG;
  std::vector<int32_t> t;
  const bool ok = cut_spmv_tiles(rowptr, n_rows, max_rows, max_blocks, t);
  *n_tiles = ok ? (int32_t)(t.size() / 4) : -1;
  *tiles = (int32_t*)malloc(sizeof(int32_t) * std::max<size_t>(t.size(), 1));
  if (!t.empty()) memcpy(*tiles, t.data(), sizeof(int32_t) * t.size());
  return RDC_OK;
}

extern "C" int rdc_probe_region_chunks(int64_t n_elems, const uint8_t* counted, const int32_t* region, int n_regions, int chunk,
                                       int64_t* n_counted, int32_t** perm, int32_t* n_chunks, int32_t** chunk_ptr,
                                       int32_t** rchunk_ptr) {
  if (!counted || !region || n_regions < 1 || chunk < 1) return RDC_E_ARG;
  std::vector<int32_t> p, cp, rp;
  bucket_regions(n_elems, counted, region, n_regions, chunk, p, cp, rp);
  auto dup = [](const std::vector<int32_t>& v) {
    int32_t* q = (int32_t*)malloc(sizeof(int32_t) * std::max<size_t>(v.size(), 1));
    if (!v.empty()) memcpy(q, v.data(), sizeof(int32_t) * v.size());
    return q;
  };
  *n_counted = (int64_t)p.size();
  *n_chunks = (int32_t)cp.size() - 1;
  *perm = dup(p); *chunk_ptr = dup(cp); *rchunk_ptr = dup(rp);
  return RDC_OK;
}

// Host-only probe of the partition / halo set-up (no device needed): lets the world_size-2 gloo tests check on
// CPU that what rank r sends is exactly what rank q expects to receive.  Buffers are malloc'ed; free with
// rdc_free.  send_glob / recv_glob hold GLOBAL node ids, grouped per neighbour (nbr_ptr offsets).
extern "C" int rdc_probe_partition(int elem_type, int nvars, int64_t n_nodes, int64_t n_elems, const int32_t* conn,
                                   const double* xyz, int rank, int nranks, int partitioner, int32_t* n_owned,
                                   int32_t* n_ghost, int64_t* n_elems_local, int32_t** owner, int32_t* n_nbr, int32_t** nbr_rank,
                                   int32_t** send_ptr, int32_t** send_glob, int32_t** recv_ptr, int32_t** recv_glob) {
  HostSetup S;
  std::string err;
  int rc;
  try {
    rc = build_setup(S, elem_type, nvars, n_nodes, n_elems, conn, xyz, rank, nranks, partitioner, 256, 1, err);
  } catch (const std::bad_alloc&) { rc = RDC_E_NOMEM; err = "out of host memory"; }
  if (rc) { g_create_err = err; return rc; }
  *n_owned = S.n_owned; *n_ghost = S.n_ghost; *n_elems_local = S.E_loc;
  auto dup = [](const std::vector<int32_t>& v) {
    int32_t* p = (int32_t*)malloc(sizeof(int32_t) * std::max<size_t>(v.size(), 1));
    if (!v.empty()) memcpy(p, v.data(), sizeof(int32_t) * v.size());
    return p;
  };
  std::vector<int32_t> own(S.owner_glob.begin(), S.owner_glob.end());
  if (own.empty()) own.assign((size_t)n_nodes, 0);
  *owner = dup(own);
  *n_nbr = (int32_t)S.nbr_rank.size();
  std::vector<int32_t> nb(S.nbr_rank.begin(), S.nbr_rank.end());
  *nbr_rank = dup(nb);
  *send_ptr = dup(S.send_ptr);
  std::vector<int32_t> sg(S.send_idx.size()), rg((size_t)S.n_ghost);
  for (size_t k = 0; k < S.send_idx.size(); k++) sg[k] = S.loc2glob[S.send_idx[k]];
  for (int32_t k = 0; k < S.n_ghost; k++) rg[k] = S.loc2glob[S.n_owned + k];
  *send_glob = dup(sg);
  *recv_ptr = dup(S.recv_ptr);
  *recv_glob = dup(rg);
  return RDC_OK;
}
