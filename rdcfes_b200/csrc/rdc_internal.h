// rdc_internal.h -- context layout shared by setup.cpp (host), assemble.cu, solver.cu and api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/rdc.h"

#define RDC_MAX_NEN 8
#define RDC_MAX_QP 8

namespace rdc {

// reference-element tables, computed on the host exactly like libMesh
// evaluates them (phi0 = 1 - xi - eta - zeta, left to right) and copied to __constant__ memory.
struct FeTable {
  int nen, nqp;
  double w[RDC_MAX_QP];
  double phi[RDC_MAX_NEN][RDC_MAX_QP];
  double dxi[RDC_MAX_NEN][RDC_MAX_QP], deta[RDC_MAX_NEN][RDC_MAX_QP], dzeta[RDC_MAX_NEN][RDC_MAX_QP];
};
void fe_table_fill(FeTable* T, int elem_type);

struct NcclApi;  // comm.cpp

// ---- NVLink peer-memory exchange (p2p.cu, comm.cpp) ----
#define RDC_MAX_RANKS 16
struct P2PHeader {                       // at offset 0 of every rank's arena; written by the peers
  unsigned long long ar_flag[2][RDC_MAX_RANKS];
  double ar_val[2][RDC_MAX_RANKS][8];
  // all-reduce fused into the producing kernel: one double = two 8-byte words {32 data bits | 32-bit tag}, so data
  // and flag arrive in the same (atomic) store and one NVLink crossing is the whole latency
  unsigned long long ll[2][RDC_MAX_RANKS][8][2];
  int error;
};
struct P2P {
  bool on = false;
  unsigned char* arena = nullptr;        // cudaMalloc'ed, exported with cudaIpcGetMemHandle
  size_t arena_bytes = 0, slot_bytes = 0, header_bytes = 8192;
  int nslots = 0, slots_used = 0;
  std::vector<void*> peer;               // [nranks] mapped arena base of every rank (own arena for self)
  void** d_peer = nullptr;
  std::vector<int64_t> dst_node_off;     // per neighbour: offset (nodes) of MY segment in ITS ghost ordering
  size_t stage_off[2] = {0, 0};          // arena offsets of the two ghost staging areas (16 B per value)
  double* d_scratch = nullptr;
  unsigned long long halo_seq = 0, ar_seq = 0;
};

// Host-side result of the one-time set-up (partition, numbering, pattern, assembly maps).
struct HostSetup {
  int nen = 0, nv = 0;
  int rank = 0, nranks = 1;
  int64_t N_glob = 0, E_glob = 0;
  int32_t n_owned = 0, n_ghost = 0, n_loc = 0;   // local node counts (owned first, then ghosts)
  int64_t E_loc = 0;
  std::vector<int32_t> loc2glob;                 // [n_loc]
  std::vector<int32_t> glob2loc;                 // [N_glob], -1 when not present on this rank
  std::vector<int64_t> elem_glob;                // [E_loc] global element id
  std::vector<int32_t> conn;                     // [E_loc*nen] local node ids
  std::vector<int32_t> n2e_ptr;                  // [n_owned+1] into pair[]
  std::vector<int32_t> pair;                     // [npairs] (local elem << 3) | local node index
  std::vector<int32_t> rowptr;                   // [n_owned+1] block rows
  std::vector<int32_t> col;                      // [nnzb] local node ids, sorted per row
  std::vector<int32_t> diag_blk;                 // [n_owned] block index of the diagonal block
  std::vector<int32_t> cta_node;                 // [ncta+1] node ranges of the assembly CTAs
  std::vector<int32_t> cptr;                     // [nnzb+1] into clist
  std::vector<uint16_t> clist;                   // j * pairs_per_cta + pair index inside the CTA
  // phase-2 work list of the assembly kernel: a block's contributor run is cut into pieces of at most `chunk` entries so
  // that the 24-contributor diagonal blocks do not hold up warps whose other lanes have ~6 (task = 2 x int32, see setup.cpp)
  std::vector<int32_t> task;                     // [2 * ntask]
  std::vector<int32_t> task_ptr;                 // [ncta+1]
  std::vector<int32_t> pair_rec;                 // [ncta * pairs_per_cta * (nen + 4)] padded pair records (setup.cpp)
  int pairs_per_cta = 0;
  // halo exchange (distributed): ghosts are ordered by owner rank, so each neighbour's ghosts are contiguous
  std::vector<int> nbr_rank;                     // neighbours
  std::vector<int32_t> send_ptr, send_idx;       // per neighbour: owned local node ids to send
  std::vector<int32_t> recv_ptr;                 // per neighbour: ghost range [recv_ptr[k], recv_ptr[k+1]) + n_owned
  std::vector<int32_t> owner_glob;               // [N_glob] owner rank of every node (empty when nranks == 1)
};

// partitioner: 0 = METIS k-way on the nodal graph, 1 = recursive coordinate bisection
int build_setup(HostSetup& S, int elem_type, int nv, int64_t N, int64_t E, const int32_t* conn, const double* xyz,
                int rank, int nranks, int partitioner, int pairs_per_cta, int node_order, std::string& err);

bool cut_spmv_tiles(const int32_t* rowptr, int32_t n_rows, int max_rows, int max_blocks, std::vector<int32_t>& tiles);
void bucket_regions(int64_t E_loc, const uint8_t* counted, const int32_t* region_of_local, int n_regions, int chunk,
                    std::vector<int32_t>& perm, std::vector<int32_t>& chunk_ptr, std::vector<int32_t>& rchunk_ptr);

}  // namespace rdc

struct SolverWork;  // solver.cu
struct RegionWork;  // reduce.cu
struct SolidWork;   // solid.cu

// Tuning switches of one context.  Defaults come from the environment (RDC_<NAME> upper case) at rdc_create and
// can be changed with rdc_set_option; the parity tests use them to run the alternative kernels side by side.
struct rdc_options {
  int spmv_tma = 1;            // 1: TMA bulk-copy staged SpMV ; 0: LDG SpMV
  int spmv_minb = 4;           // LDG SpMV: resident CTAs per SM the kernel is compiled for (4, 5, 6, 8)
  int spmv_ctas_per_sm = 0;    // LDG SpMV grid (0 = default)
  int tma_ctas_per_sm = 0;     // TMA SpMV grid (0 = default)
  int tma_stages = 0;          // TMA SpMV stages (0 = default 2 ; 3)
  int sync_every = 0;          // iterations queued ahead of the convergence flag (0 = default)
  int p2p_fused_ar = 1;        // all-reduce finished inside the producing kernel
  int p2p_fused_halo = 1;      // ghost exchange inside the BiCGStab vector kernels
  int node_order = 1;          // local numbering of the owned nodes: 1 Morton curve of the coordinates, 0 ascending global id (read at rdc_create)
  int bicg_persist = -1;       // BiCGStab as one cooperative launch (solver.cu k_bicgstab_persist): 1 always, 0 never (five launches per
                               // iteration), -1 automatic: when a CTA of the resident grid gets at most 40 operator tiles per SpMV
  int persist_timing = 1;      // the persistent solver times its phases (SpMV time of rdc_stats); 0 = no timer reads
  int l2_evict_first = 1;      // the SpMV's operator stream is fetched with the evict_first L2 priority (solver.cu bulk_g2s_hint)
  int vec_reverse = 1;         // BiCGStab s-update (elementwise, no reduction) walks its vectors back to front (L2 reuse of the SpMV's output tail)
  int trace = 0;               // print the device time of every operation of BiCGStab iteration 4
};

struct rdc_ctx {
  int model = 0, etype = 0, nen = 0, nv = 0, nqp = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = true;
  rdc::HostSetup S;
  std::vector<int32_t> dof_base;     // [N_glob] global dof id of var 0 (identity = nv*node when empty)
  bool identity_dofs = true;
  int64_t D_glob = 0;
  std::vector<double> params;
  bool have_params = false;
  double time = 0.0, dt = 0.0;

  // device: mesh + maps
  int32_t* d_conn = nullptr;
  double* d_xyz = nullptr;           // [n_loc*4] padded (x,y,z,0): one 32-byte sector per node
  double* d_efield = nullptr;        // [E_loc*3]
  int32_t *d_n2e_ptr = nullptr, *d_pair = nullptr, *d_rowptr = nullptr, *d_col = nullptr, *d_diag_blk = nullptr;
  int32_t *d_cta_node = nullptr, *d_task = nullptr;
  uint16_t* d_clist = nullptr;
  int32_t* d_dofmap = nullptr;       // [n_loc*nv] global dof id of each local dof (gather/scatter of user vectors)
  int32_t* d_node_by_glob = nullptr; // [n_loc] local node ids in ascending order of their global dof base (host-memory side of zero-copy transfers)
  int ncta = 0;
  int64_t nnzb = 0;
  // structurally non-zero entries (a,b) of the model's v x v node block (bit a*nv+b).  Only these NKV entry
  // planes are stored, assembled and streamed by SpMV; rdc_download_csr re-inserts the explicit zeros the
  // reference keeps in its AIJ pattern.
  unsigned kmask = 0;
  int nkv = 0;

  // device: operator and vectors.  Vectors are [n_loc*nv]: owned part first, ghost part after it.
  double *d_val = nullptr, *d_rhs = nullptr, *d_dinv = nullptr;
  double *d_u = nullptr, *d_uold = nullptr, *d_uolder = nullptr;
  double* d_stage = nullptr;         // [D_glob] staging for user vectors in global dof order
  bool assembled = false;
  bool u_ghost_fresh = false;        // the ghost part of d_u matches the owners' values (distributed runs)

  // model state
  double *d_td = nullptr, *d_rt = nullptr, *d_prev = nullptr;   // RIPF: TD [n_loc*3], RT [n_loc*3], prev [n_owned*3]
  double* d_aux = nullptr;                                      // PROTEAS AUX [n_loc*2]
  int ripf_rt_max = 0;
  bool ripf_primed = false;

  // solver
  SolverWork* work = nullptr;
  RegionWork* region = nullptr;      // save_solution reductions (rdc_set_subdomains)
  SolidWork* solid = nullptr;        // RDC_SOLID: reference configuration, materials, boundary conditions (solid.cu)

  // comm
  rdc::P2P* p2p = nullptr;
  int* work_state = nullptr;         // solver's device flags (state[0] = done) for the peer-memory kernels
  rdc::NcclApi* nccl = nullptr;
  void* comm = nullptr;              // ncclComm_t
  int32_t* d_send_idx = nullptr;
  double* d_sendbuf = nullptr;

  rdc_options opt;
  // stats
  rdc_stats st = {};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;          // scratch pair (rdc_bench_spmv)
  cudaEvent_t ev_asm[2] = {nullptr, nullptr}, ev_sol[2] = {nullptr, nullptr}, ev_clamp[2] = {nullptr, nullptr};
  bool t_asm_pending = false, t_sol_pending = false, t_clamp_pending = false;  // elapsed times are read lazily
  std::string err;
};

namespace rdc {
// assemble.cu
int launch_assemble(rdc_ctx* c);   // K, F and the point-Jacobi scaling dinv = 1/diag(K)
unsigned model_kmask(int model);
int upload_fe_tables();
// solver.cu
int solver_init(rdc_ctx* c);
void solver_free(rdc_ctx* c);
int solver_solve(rdc_ctx* c, int ksp, int pc, double rtol, int maxits, int restart, int* its, double* res);
int launch_spmv(rdc_ctx* c, const double* x, double* y, const double* rowscale, bool check_done = false);
int launch_clamp(rdc_ctx* c);
int refresh_u_ghosts(rdc_ctx* c);
int solver_persist_begin(rdc_ctx* c, int pc, double rtol, int maxits);   // 1: not applicable (use solver_solve)
int solver_persist_end(rdc_ctx* c, int* its, double* res);
int launch_stream_probe(rdc_ctx* c, int ctas_per_sm);
int launch_barrier_probe(rdc_ctx* c, int reps, int ctas_per_sm, int mode);
int launch_dfma_probe(rdc_ctx* c, int iters);
int spmv_masks_ok();                // solver.cu's entry masks agree with the model definitions
int launch_gather(rdc_ctx* c, const double* src_glob, double* dst_loc);     // dst_loc[l] = src[dofmap[l]]
int launch_scatter(rdc_ctx* c, const double* src_loc, double* dst_glob);    // owned part only
int halo_exchange(rdc_ctx* c, double* x, bool check_done = false);          // fills the ghost part of x
int allreduce_sum(rdc_ctx* c, double* d_buf, int n, bool check_done = false);
int allreduce_max(rdc_ctx* c, double* d_buf, int n);
// reduce.cu
int region_setup(rdc_ctx* c, const int32_t* region, int n_regions);
void region_free(rdc_ctx* c);
int region_volumes(rdc_ctx* c, int ncond, const rdc_range_cond* cond, double* vol);
int region_last_mean(rdc_ctx* c, int var, double* mean);
// solid.cu
void solid_free(rdc_ctx* c);
// comm.cpp
int comm_unique_id(void* out128, std::string& err);
int comm_init(rdc_ctx* c, const void* uid, std::string& err);
void comm_destroy(rdc_ctx* c);
int p2p_init(rdc_ctx* c, std::string& err);        // after comm_init: arena, IPC handles, peer tables
double* p2p_alloc(rdc_ctx* c, size_t n_doubles);   // vector slot inside the arena (nullptr: not available)
bool p2p_owns(const rdc_ctx* c, const void* p);
inline bool p2p_on(const rdc_ctx* c) { return c->p2p && c->p2p->on; }
// p2p.cu
int p2p_launch_halo(rdc_ctx* c, double* x, bool check_done);
int p2p_launch_allreduce(rdc_ctx* c, double* d_buf, int n, bool check_done);
int p2p_check_error(rdc_ctx* c);
struct HaloArgs;
void p2p_fill_halo_args(rdc_ctx* c, HaloArgs* A, int* max_blk, int* total_blk, unsigned long long* seq);
}  // namespace rdc

#define RDC_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess) {                                                             \
      c->err = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
      return RDC_E_CUDA;                                                                 \
    }                                                                                    \
  } while (0)
