// solid.cu -- the solid-mechanics Newton path on the device: neo-Hookean residual and
// tangent with growth and fibres, penalty boundary conditions, the Newton load step, stress post-processing.
//
// Replaces SolidSystem::element_time_derivative / side_time_derivative (solid_system.C:146-371) under
// [upstream] FEMSystem::assembly, SolidSystem::run_solver (:373-392, i.e. [upstream] NewtonSolver::solve with the options of
// :80-98) and SolidSystem::post_process (:394-538).  The unknowns are the CURRENT node positions (solid.C:27-30); they
// live in the context's solution vector, the undeformed positions ("SolidSystem::auxiliary", solid_system.C:26-48) in a
// second nodal array.  The mesh "moves" implicitly: every kernel reads the geometry from the solution vector.
//
// Assembly: the pair decomposition, staging and deterministic phase 2 of assemble.cu (asm_common.cuh) with a dense 3 x 3
// node block (9 entry planes).  One thread = one (node, element) pair = row i of the element tangent (solid_dev.cuh).
// The penalty terms touch only boundary rows: one thread per boundary node adds the contributions of its sides in a
// fixed order straight into the operator and refreshes the Jacobi scaling of that row -- no atomics, bit-reproducible.
// The linear systems J d = R go through the context's Krylov solvers (solver.cu) with d = 0 as the initial guess.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "asm_common.cuh"
#include "rdc_internal.h"
#include "solid_dev.cuh"

#define SOLID_MAX_MAT 16

namespace rdc {

__constant__ FeTable c_fe_solid[2];  // [0] TET4, [1] HEX8 (a second copy: constants are per translation unit)

struct SolidArgs {
  const double* xund;       // [n_loc*3]
  const int32_t* mat_of;    // [E_loc] or null
  const double* fibres;     // [E_loc*3] or null
  double pseudo_time;
  int use_symmetry;         // solver/assembly_use_symmetry (solid_system.C:180)
  double mats[SOLID_MAX_MAT][6];
};

template <int NEN, int PAIRS>
__global__ void __launch_bounds__(PAIRS, 1) k_solid_assemble(const AsmArgs A, const SolidArgs S) {
  constexpr int NV = 3;
  constexpr unsigned KMASK = 0x1FF;
  constexpr int NKV = 9;
  constexpr int TI = NEN == 4 ? 0 : 1;
  extern __shared__ double smem[];
  double* stageK = smem;                              // [NKV][NEN][PAIRS]
  double* stageF = smem + (size_t)NEN * NKV * PAIRS;  // [NV][PAIRS]
  __shared__ int s_rowptr[PAIRS + 1];
  __shared__ int s_n2e[PAIRS + 1];
  __shared__ int s_diag[PAIRS];
  __shared__ __align__(4) unsigned short s_clist_raw[PAIRS * NEN + 4];

  const int tid = threadIdx.x;
  AsmCta<NEN, PAIRS> cta;
  asm_prologue<NEN, PAIRS>(A, tid, cta, s_rowptr, s_n2e, s_diag, s_clist_raw);

  constexpr int REC4 = (NEN + 4) / 4;
  const int4* rec = reinterpret_cast<const int4*>(A.pair) + ((size_t)blockIdx.x * PAIRS + tid) * REC4;
  const int4 rec0 = __ldg(rec);
  if (rec0.x >= 0) {
    const int e = rec0.x >> 3, li = rec0.x & 7;
    int en[NEN];
    {
      const int4 c4 = __ldg(rec + 1);
      en[0] = c4.x; en[1] = c4.y; en[2] = c4.z; en[3] = c4.w;
      if constexpr (NEN == 8) {
        const int4 d4 = __ldg(rec + 2);
        en[4] = d4.x; en[5] = d4.y; en[6] = d4.z; en[7] = d4.w;
      }
    }
    double Xc[NEN][3], Xu[NEN][3];
#pragma unroll
    for (int l = 0; l < NEN; l++)
#pragma unroll
      for (int d = 0; d < 3; d++) {
        Xc[l][d] = A.u_old[(size_t)en[l] * 3 + d];     // the iterate = current positions
        Xu[l][d] = S.xund[(size_t)en[l] * 3 + d];
      }
    const int m = S.mat_of ? S.mat_of[e] : 0;
    double mat[6], eta[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 6; k++) mat[k] = S.mats[m][k];
    if (S.fibres) { eta[0] = S.fibres[(size_t)e * 3]; eta[1] = S.fibres[(size_t)e * 3 + 1]; eta[2] = S.fibres[(size_t)e * 3 + 2]; }
    double R[3] = {0.0, 0.0, 0.0};
    double* K = stageK + tid;
#pragma unroll
    for (int k = 0; k < NKV * NEN; k++) K[(size_t)k * PAIRS] = 0.0;
    solid_row<NEN>(c_fe_solid[TI], Xc, Xu, mat, S.pseudo_time, eta, li, R, K, PAIRS, S.use_symmetry);
#pragma unroll
    for (int a = 0; a < NV; a++) stageF[a * PAIRS + tid] = R[a];
  }
  asm_phase2<NV, KMASK, NEN, PAIRS>(A, tid, cta, stageK, stageF, s_rowptr, s_n2e, s_diag, s_clist_raw);
}

// ---- penalty boundary conditions ---------------------------------------------------------------------------------
struct BcArgs {
  int nrow;
  const int32_t* row_node;   // [nrow] owned local node
  const int32_t* row_ptr;    // [nrow+1] into ent_*
  const int32_t* ent_side;   // side record
  const int32_t* ent_pos;    // position of the row's node inside the side
  const int32_t* side_node;  // [nside*4] local node ids (-1 padded)
  const int32_t* side_bc;    // [nside]
  const double* bc_disp;     // [nbc*3]
  const double* u;
  const double* xund;
  const int32_t* rowptr;
  const int32_t* col;
  const int32_t* diag_blk;
  double* val;
  double* rhs;
  double* dinv;
  double pseudo_time, penalty;
};

template <int NS>
__global__ void __launch_bounds__(128) k_solid_bc(const BcArgs B) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B.nrow) return;
  const int r = B.row_node[t];
  const int r0 = B.rowptr[r], L = B.rowptr[r + 1] - r0;
  double* vrow = B.val + (size_t)r0 * 9;
  double Racc[3] = {0.0, 0.0, 0.0};
  for (int k = B.row_ptr[t]; k < B.row_ptr[t + 1]; k++) {
    const int s = B.ent_side[k], pos = B.ent_pos[k];
    int n[NS];
    double Xc[NS][3], Xu[NS][3];
#pragma unroll
    for (int j = 0; j < NS; j++) {
      n[j] = B.side_node[(size_t)s * 4 + j];
#pragma unroll
      for (int d = 0; d < 3; d++) { Xc[j][d] = B.u[(size_t)n[j] * 3 + d]; Xu[j][d] = B.xund[(size_t)n[j] * 3 + d]; }
    }
    double Kd[NS * 3];
#pragma unroll
    for (int j = 0; j < NS * 3; j++) Kd[j] = 0.0;
    solid_bc_row<NS>(Xc, Xu, B.bc_disp + (size_t)B.side_bc[s] * 3, B.pseudo_time, B.penalty, pos, Racc, Kd);
#pragma unroll
    for (int j = 0; j < NS; j++) {
      int lo = 0, hi = L - 1, kk = -1;   // the column of node n[j] in this (sorted) block row
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int cm = B.col[r0 + mid];
        if (cm == n[j]) { kk = mid; break; }
        if (cm < n[j]) lo = mid + 1; else hi = mid - 1;
      }
      if (kk < 0) continue;              // cannot happen: side nodes share an element with the row's node
#pragma unroll
      for (int d = 0; d < 3; d++) vrow[(size_t)(4 * d) * L + kk] += Kd[j * 3 + d];
    }
  }
#pragma unroll
  for (int d = 0; d < 3; d++) B.rhs[(size_t)r * 3 + d] += Racc[d];
  const int kd = B.diag_blk[r] - r0;
#pragma unroll
  for (int d = 0; d < 3; d++) B.dinv[(size_t)r * 3 + d] = 1.0 / vrow[(size_t)(4 * d) * L + kd];
}

// ---- small vector kernels of the Newton driver ---------------------------------------------------------------------
// deterministic sum of squares: fixed grid, fixed per-thread stride, block tree, then one block adds the partials
__global__ void __launch_bounds__(256) k_sumsq_partial(size_t n, const double* __restrict__ v, double* __restrict__ partial) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) acc += v[i] * v[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) k_sum_final(int nb, const double* __restrict__ partial, double* __restrict__ out) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) acc += partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0];
}
__global__ void __launch_bounds__(256) k_axpy(size_t n, double a, const double* __restrict__ x, double* __restrict__ y) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) y[i] += a * x[i];
}

template <int NEN>
__global__ void __launch_bounds__(128) k_solid_post(int64_t E, const int32_t* __restrict__ conn, const double* __restrict__ u,
                                                    const SolidArgs S, double* __restrict__ out /* [E*5] */) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= E) return;
  double Xc[NEN][3], Xu[NEN][3];
#pragma unroll
  for (int l = 0; l < NEN; l++) {
    const int nd = conn[e * NEN + l];
#pragma unroll
    for (int d = 0; d < 3; d++) { Xc[l][d] = u[(size_t)nd * 3 + d]; Xu[l][d] = S.xund[(size_t)nd * 3 + d]; }
  }
  const int m = S.mat_of ? S.mat_of[e] : 0;
  double mat[6], eta[3] = {0.0, 0.0, 0.0}, o[5];
#pragma unroll
  for (int k = 0; k < 6; k++) mat[k] = S.mats[m][k];
  if (S.fibres) { eta[0] = S.fibres[e * 3]; eta[1] = S.fibres[e * 3 + 1]; eta[2] = S.fibres[e * 3 + 2]; }
  solid_post_elem<NEN>(c_fe_solid[NEN == 4 ? 0 : 1], Xc, Xu, mat, S.pseudo_time, eta, o);
#pragma unroll
  for (int k = 0; k < 5; k++) out[e * 5 + k] = o[k];
}

}  // namespace rdc

using namespace rdc;

// ---- host state ---------------------------------------------------------------------------------------------------
struct SolidWork {
  double* d_xund = nullptr;
  int32_t* d_mat_of = nullptr;
  double* d_fibres = nullptr;
  double mats[SOLID_MAX_MAT][6];
  int nmat = 0;
  // boundary conditions
  int nrow = 0, nside = 0, ns = 3;
  int32_t *d_row_node = nullptr, *d_row_ptr = nullptr, *d_ent_side = nullptr, *d_ent_pos = nullptr, *d_side_node = nullptr, *d_side_bc = nullptr;
  double* d_bc_disp = nullptr;
  double penalty = 1.0e5;
  int use_symmetry = 0;
  // Newton scratch
  double *d_dx = nullptr, *d_partial = nullptr, *d_post = nullptr;
  bool tables = false;
};

namespace rdc {
void solid_free(rdc_ctx* c) {
  SolidWork* W = c->solid;
  if (!W) return;
  cudaFree(W->d_xund); cudaFree(W->d_mat_of); cudaFree(W->d_fibres); cudaFree(W->d_row_node); cudaFree(W->d_row_ptr);
  cudaFree(W->d_ent_side); cudaFree(W->d_ent_pos); cudaFree(W->d_side_node); cudaFree(W->d_side_bc); cudaFree(W->d_bc_disp);
  cudaFree(W->d_dx); cudaFree(W->d_partial); cudaFree(W->d_post);
  delete W;
  c->solid = nullptr;
}
}  // namespace rdc

#define CHECK_SOLID(c)                                                                     \
  if (!(c)) return RDC_E_ARG;                                                              \
  cudaSetDevice((c)->device);                                                              \
  if ((c)->model != RDC_SOLID) { (c)->err = "not a solid-mechanics context (RDC_SOLID)"; return RDC_E_STATE; }

static int solid_work(rdc_ctx* c, SolidWork** out) {
  if (!c->solid) {
    c->solid = new SolidWork();
    const double def[6] = {1.0e3, 0.3, 0.0, 0.0, 0.0, 0.0};   // solid.C:279-290 defaults
    memcpy(c->solid->mats[0], def, sizeof(def));
    c->solid->nmat = 1;
  }
  SolidWork* W = c->solid;
  if (!W->tables) {
    FeTable t[2];
    fe_table_fill(&t[0], RDC_TET4);
    fe_table_fill(&t[1], RDC_HEX8);
    RDC_CUDA(cudaMemcpyToSymbol(c_fe_solid, t, sizeof(t)));
    W->tables = true;
  }
  if (!W->d_dx) {
    RDC_CUDA(cudaMalloc(&W->d_dx, (size_t)c->S.n_loc * 3 * sizeof(double)));
    RDC_CUDA(cudaMalloc(&W->d_partial, 1024 * sizeof(double)));
  }
  *out = W;
  return 0;
}

template <class T>
static int put(rdc_ctx* c, T** dst, const std::vector<T>& src) {
  cudaFree(*dst);
  *dst = nullptr;
  RDC_CUDA(cudaMalloc((void**)dst, std::max<size_t>(src.size(), 1) * sizeof(T)));
  if (!src.empty()) RDC_CUDA(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// SolidSystem::save_initial_mesh (solid_system.C:26-48): the undeformed node positions, global node order
extern "C" int rdc_solid_set_reference(rdc_ctx* c, const double* xund) {
  CHECK_SOLID(c);
  if (!xund) return RDC_E_ARG;
  SolidWork* W;
  int rc = solid_work(c, &W);
  if (rc) return rc;
  const HostSetup& S = c->S;
  std::vector<double> loc((size_t)S.n_loc * 3);
  for (int32_t l = 0; l < S.n_loc; l++)
    for (int d = 0; d < 3; d++) loc[(size_t)l * 3 + d] = xund[(size_t)S.loc2glob[l] * 3 + d];
  return put(c, &W->d_xund, loc);
}

// es.parameters "material/<id>/Hyperelastic/..." (solid.C:276-291): {Young, Poisson, FibreStiffness, rate_0..2} per material,
// mat_of[e] = index of the element's subdomain in that table (NULL: material 0 everywhere)
extern "C" int rdc_solid_set_materials(rdc_ctx* c, int nmat, const double* mats, const int32_t* mat_of) {
  CHECK_SOLID(c);
  if (nmat < 1 || nmat > SOLID_MAX_MAT || !mats) { c->err = "rdc_solid_set_materials: 1..16 materials"; return RDC_E_ARG; }
  SolidWork* W;
  int rc = solid_work(c, &W);
  if (rc) return rc;
  memcpy(W->mats, mats, (size_t)nmat * 6 * sizeof(double));
  W->nmat = nmat;
  cudaFree(W->d_mat_of);
  W->d_mat_of = nullptr;
  if (mat_of) {
    const HostSetup& S = c->S;
    std::vector<int32_t> loc((size_t)S.E_loc);
    for (int64_t le = 0; le < S.E_loc; le++) {
      loc[le] = mat_of[S.elem_glob[le]];
      if (loc[le] < 0 || loc[le] >= nmat) { c->err = "rdc_solid_set_materials: material index out of range"; return RDC_E_ARG; }
    }
    if ((rc = put(c, &W->d_mat_of, loc))) return rc;
  }
  return RDC_OK;
}

// es.parameters "solver/assembly_use_symmetry" (solid.C:243-244, solid_system.C:180,248-262): blocks j >= i evaluated, the rest mirrored
extern "C" int rdc_solid_set_symmetry(rdc_ctx* c, int use_symmetry) {
  CHECK_SOLID(c);
  SolidWork* W;
  int rc = solid_work(c, &W);
  if (rc) return rc;
  W->use_symmetry = use_symmetry != 0;
  c->assembled = false;
  return RDC_OK;
}

// "SolidSystem::fibre" variables 0-2 (solid.C:303-337): reference fibre direction per element (NULL: none)
extern "C" int rdc_solid_set_fibres(rdc_ctx* c, const double* fibres) {
  CHECK_SOLID(c);
  SolidWork* W;
  int rc = solid_work(c, &W);
  if (rc) return rc;
  cudaFree(W->d_fibres);
  W->d_fibres = nullptr;
  if (!fibres) return RDC_OK;
  const HostSetup& S = c->S;
  std::vector<double> loc((size_t)S.E_loc * 3);
  for (int64_t le = 0; le < S.E_loc; le++)
    for (int d = 0; d < 3; d++) loc[(size_t)le * 3 + d] = fibres[(size_t)S.elem_glob[le] * 3 + d];
  return put(c, &W->d_fibres, loc);
}

// Host-only: which owned rows carry penalty terms and in which order.  conn = LOCAL connectivity (local node ids, owned nodes
// first), loc_of = global element -> local element or -1.  A side whose element is not local is skipped (every element that
// touches an owned node is local, so an owned boundary node sees all of its sides); an owned node's entries keep the input
// order of the sides -- the order the penalty terms are added in (fixed => bit-reproducible).
namespace rdc {
struct SolidBcLists {
  std::vector<int32_t> row_node, row_ptr, ent_side, ent_pos, side_node, side_bc;
};
int solid_build_bc_lists(int nen, int32_t n_owned, const std::vector<int32_t>& conn, const std::vector<int64_t>& loc_of, int nbc, int64_t nside,
                         const int64_t* side_elem, const int32_t* side_no, const int32_t* side_bc, SolidBcLists& L, std::string& err) {
  static const int tet[4][4] = {{0, 2, 1, -1}, {0, 1, 3, -1}, {1, 2, 3, -1}, {2, 0, 3, -1}};     // [upstream] side_nodes_map
  static const int hex[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
  const int ns = nen == 4 ? 3 : 4, nsides = nen == 4 ? 4 : 6;
  const int64_t E_glob = (int64_t)loc_of.size();
  L.side_node.assign((size_t)nside * 4, -1);
  L.side_bc.assign((size_t)nside, 0);
  std::vector<std::vector<std::pair<int32_t, int32_t>>> per_node((size_t)n_owned);   // owned node -> (side, position), input order
  for (int64_t k = 0; k < nside; k++) {
    if (side_elem[k] < 0 || side_elem[k] >= E_glob || side_no[k] < 0 || side_no[k] >= nsides || side_bc[k] < 0 || side_bc[k] >= nbc) {
      err = "rdc_solid_set_bcs: side out of range";
      return RDC_E_ARG;
    }
    const int64_t le = loc_of[side_elem[k]];
    L.side_bc[k] = side_bc[k];
    if (le < 0) continue;   // not on this rank
    for (int j = 0; j < ns; j++) {
      const int32_t nd = conn[(size_t)le * nen + (nen == 4 ? tet[side_no[k]][j] : hex[side_no[k]][j])];
      L.side_node[(size_t)k * 4 + j] = nd;
      if (nd < n_owned) per_node[nd].push_back({(int32_t)k, (int32_t)j});
    }
  }
  L.row_node.clear(); L.ent_side.clear(); L.ent_pos.clear();
  L.row_ptr.assign(1, 0);
  for (int32_t nd = 0; nd < n_owned; nd++) {
    if (per_node[nd].empty()) continue;
    L.row_node.push_back(nd);
    for (auto& sp : per_node[nd]) { L.ent_side.push_back(sp.first); L.ent_pos.push_back(sp.second); }
    L.row_ptr.push_back((int32_t)L.ent_side.size());
  }
  return 0;
}
}  // namespace rdc

// Boundary sides with their conditions: side k = side `side_no[k]` (libMesh side order) of element `side_elem[k]` carries
// boundary condition `side_bc[k]`, whose prescribed displacement is bc_disp[3*side_bc[k] ..] (NaN = component free);
// es.parameters "BCs", "BC/<id>/displacement", "BCs/displacement_penalty" + BoundaryInfo (solid_system.C:288-304).
extern "C" int rdc_solid_set_bcs(rdc_ctx* c, int nbc, const double* bc_disp, int64_t nside, const int64_t* side_elem, const int32_t* side_no,
                                 const int32_t* side_bc, double penalty) {
  CHECK_SOLID(c);
  if (nbc < 0 || nside < 0 || (nside > 0 && (!side_elem || !side_no || !side_bc || !bc_disp))) return RDC_E_ARG;
  SolidWork* W;
  int rc = solid_work(c, &W);
  if (rc) return rc;
  const HostSetup& S = c->S;
  const int nen = c->nen, ns = nen == 4 ? 3 : 4;
  // local connectivity back from the device (the host copy is dropped after rdc_create)
  std::vector<int32_t> conn((size_t)S.E_loc * nen);
  RDC_CUDA(cudaMemcpy(conn.data(), c->d_conn, conn.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
  std::vector<int64_t> loc_of(S.E_glob, -1);
  for (int64_t le = 0; le < S.E_loc; le++) loc_of[S.elem_glob[le]] = le;
  SolidBcLists L;
  if ((rc = solid_build_bc_lists(nen, S.n_owned, conn, loc_of, nbc, nside, side_elem, side_no, side_bc, L, c->err))) return rc;
  std::vector<double> disp(bc_disp, bc_disp + (size_t)nbc * 3);
  if ((rc = put(c, &W->d_row_node, L.row_node)) || (rc = put(c, &W->d_row_ptr, L.row_ptr)) || (rc = put(c, &W->d_ent_side, L.ent_side)) ||
      (rc = put(c, &W->d_ent_pos, L.ent_pos)) || (rc = put(c, &W->d_side_node, L.side_node)) || (rc = put(c, &W->d_side_bc, L.side_bc)) ||
      (rc = put(c, &W->d_bc_disp, disp)))
    return rc;
  W->nrow = (int)L.row_node.size(); W->nside = (int)nside; W->ns = ns; W->penalty = penalty;
  return RDC_OK;
}

// Host-only probe of the penalty-row lists of rank `rank` of an `nranks` job (no device needed; the CPU world-size-2 tests use
// it): rows as GLOBAL node ids, entries as (side index into the input, position of the row's node inside that side).  Arrays
// are malloc'ed by the library (rdc_free).
extern "C" int rdc_solid_probe_bc_rows(int elem_type, int64_t n_nodes, int64_t n_elems, const int32_t* conn, const double* xyz, int rank,
                                       int nranks, int partitioner, int64_t nside, const int64_t* side_elem, const int32_t* side_no,
                                       int32_t* n_rows, int32_t** row_node_glob, int32_t** row_ptr, int32_t** ent_side, int32_t** ent_pos) {
  if ((elem_type != RDC_TET4 && elem_type != RDC_HEX8) || !conn || !xyz || !n_rows || !row_node_glob || !row_ptr || !ent_side || !ent_pos)
    return RDC_E_ARG;
  HostSetup S;
  std::string err;
  int rc;
  try {
    rc = build_setup(S, elem_type, 3, n_nodes, n_elems, conn, xyz, rank, nranks, partitioner, 128, 1, err);
  } catch (const std::bad_alloc&) { rc = RDC_E_NOMEM; }
  if (rc) return rc;
  std::vector<int64_t> loc_of((size_t)n_elems, -1);
  for (int64_t le = 0; le < S.E_loc; le++) loc_of[S.elem_glob[le]] = le;
  std::vector<int32_t> bc((size_t)std::max<int64_t>(nside, 1), 0);
  SolidBcLists L;
  if ((rc = solid_build_bc_lists(S.nen, S.n_owned, S.conn, loc_of, 1, nside, side_elem, side_no, bc.data(), L, err))) return rc;
  for (auto& nd : L.row_node) nd = S.loc2glob[nd];
  auto dup = [](const std::vector<int32_t>& v) {
    int32_t* p = (int32_t*)malloc(sizeof(int32_t) * std::max<size_t>(v.size(), 1));
    if (!v.empty()) memcpy(p, v.data(), sizeof(int32_t) * v.size());
    return p;
  };
  *n_rows = (int32_t)L.row_node.size();
  *row_node_glob = dup(L.row_node); *row_ptr = dup(L.row_ptr); *ent_side = dup(L.ent_side); *ent_pos = dup(L.ent_pos);
  return RDC_OK;
}

static void fill_solid_args(const SolidWork* W, double pseudo_time, SolidArgs* S) {
  S->xund = W->d_xund; S->mat_of = W->d_mat_of; S->fibres = W->d_fibres; S->pseudo_time = pseudo_time;
  S->use_symmetry = W->use_symmetry;
  memcpy(S->mats, W->mats, sizeof(S->mats));
}

template <int NEN>
static int launch_solid_asm(rdc_ctx* c, const AsmArgs& A, const SolidArgs& S) {
  constexpr int PAIRS = 128;
  const size_t smem = ((size_t)NEN * 9 + 3) * PAIRS * sizeof(double);
  static bool attr_done_dev[64] = {};
  bool& attr_done = attr_done_dev[c->device & 63];
  if (!attr_done) {
    RDC_CUDA(cudaFuncSetAttribute(k_solid_assemble<NEN, PAIRS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  k_solid_assemble<NEN, PAIRS><<<c->ncta, PAIRS, smem, c->stream>>>(A, S);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

// Jacobian and residual at the current positions ([upstream] FEMSystem::assembly(true, true)): device resident
static int solid_assemble(rdc_ctx* c, SolidWork* W, double pseudo_time) {
  if (!W->d_xund) { c->err = "solid: rdc_solid_set_reference has not been called"; return RDC_E_STATE; }
  if (c->S.pairs_per_cta != 128) { c->err = "solid: the assembly kernel is built for 128 pairs per CTA"; return RDC_E_STATE; }
  int grc = refresh_u_ghosts(c);   // distributed: the positions of the ghost nodes (rows of owned nodes see ghost-layer elements)
  if (grc) return grc;
  AsmArgs A;
  A.conn = c->d_conn; A.xyz4 = c->d_xyz; A.u_old = c->d_u; A.efield = nullptr; A.aux0 = nullptr; A.aux1 = nullptr;
  A.n2e_ptr = c->d_n2e_ptr; A.pair = c->d_pair; A.rowptr = c->d_rowptr; A.cta_node = c->d_cta_node;
  A.task = reinterpret_cast<const int2*>(c->d_task); A.clist = c->d_clist; A.diag_blk = c->d_diag_blk;
  A.val = c->d_val; A.rhs = c->d_rhs; A.dinv = c->d_dinv;
  SolidArgs S;
  fill_solid_args(W, pseudo_time, &S);
  int rc = c->nen == 4 ? launch_solid_asm<4>(c, A, S) : launch_solid_asm<8>(c, A, S);
  if (rc) return rc;
  if (W->nrow > 0) {
    BcArgs B;
    B.nrow = W->nrow; B.row_node = W->d_row_node; B.row_ptr = W->d_row_ptr; B.ent_side = W->d_ent_side; B.ent_pos = W->d_ent_pos;
    B.side_node = W->d_side_node; B.side_bc = W->d_side_bc; B.bc_disp = W->d_bc_disp; B.u = c->d_u; B.xund = W->d_xund;
    B.rowptr = c->d_rowptr; B.col = c->d_col; B.diag_blk = c->d_diag_blk; B.val = c->d_val; B.rhs = c->d_rhs; B.dinv = c->d_dinv;
    B.pseudo_time = pseudo_time; B.penalty = W->penalty;
    const int grid = (W->nrow + 127) / 128;
    if (W->ns == 3) k_solid_bc<3><<<grid, 128, 0, c->stream>>>(B);
    else k_solid_bc<4><<<grid, 128, 0, c->stream>>>(B);
    c->st.kernel_launches++;
    RDC_CUDA(cudaGetLastError());
  }
  c->assembled = true;
  return 0;
}

extern "C" int rdc_solid_assemble(rdc_ctx* c, double pseudo_time) {
  CHECK_SOLID(c);
  SolidWork* W;
  int rc = solid_work(c, &W);
  if (rc) return rc;
  if ((rc = solid_assemble(c, W, pseudo_time))) return rc;
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  return RDC_OK;
}

// l2 norm of the owned part (first n entries) of a device vector, over all ranks: deterministic, and bit-identical on every
// rank (the ranks' sums are added in rank order), which the Newton driver's decisions rely on
static int dev_norm(rdc_ctx* c, SolidWork* W, const double* v, size_t n, double* out) {
  const int nb = 296;
  k_sumsq_partial<<<nb, 256, 0, c->stream>>>(n, v, W->d_partial);
  k_sum_final<<<1, 256, 0, c->stream>>>(nb, W->d_partial, W->d_partial + 512);
  c->st.kernel_launches += 2;
  int arc = allreduce_sum(c, W->d_partial + 512, 1);
  if (arc) return arc;
  double s = 0.0;
  RDC_CUDA(cudaMemcpyAsync(&s, W->d_partial + 512, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  *out = sqrt(s);
  return 0;
}

// One load step: SolidSystem::run_solver (solid_system.C:373-392) = [upstream] NewtonSolver::solve with the options of
// solid_system.C:80-98.  opts = {max_nonlinear_iterations, relative_step_tolerance, relative_residual_tolerance,
// absolute_residual_tolerance, require_reduction, max_linear_iterations, initial_linear_tolerance}; libMesh defaults for
// the rest (linear_tolerance_multiplier 1e-3, minimum_linear_tolerance 1e-12).  Every iteration: J and R at the iterate,
// linear tolerance = max(min(previous, 1e-3 ||R||, floor 1e-12), atol / ||R|| / 10), J d = R from d = 0 (ksp + point
// Jacobi, PETSc's relative test on the preconditioned residual), x -= d, then the convergence tests on the residual of
// the new iterate (absolute / relative residual always; relative step only after a finished linear solve).
// require_reduction backtracks by halving until the residual drops (libMesh's Brent refinement is not reproduced).
// info = {newton iterations, linear iterations, final residual, converged}.
extern "C" int rdc_solid_newton(rdc_ctx* c, double pseudo_time, const double* opts, int ksp, double* info) {
  CHECK_SOLID(c);
  if (!opts) return RDC_E_ARG;
  SolidWork* W;
  int rc = solid_work(c, &W);
  if (rc) return rc;
  const size_t D = (size_t)c->S.n_owned * 3;
  const int max_nl = (int)opts[0];
  const double rel_step = opts[1], rel_res = opts[2], abs_res = opts[3];
  const bool require_reduction = opts[4] != 0.0;
  const int max_lin = (int)opts[5];
  double lin_tol = opts[6];
  const double lin_mult = 1e-3, lin_min = 1e-12;
  double max_residual = 0.0, max_solution = 0.0, current_residual = 0.0, norm_delta = 0.0;
  int outer = 0, inner = 0, converged = 0;
  bool linear_finished = true;
  for (outer = 0; outer <= max_nl; outer++) {
    if ((rc = solid_assemble(c, W, pseudo_time))) return rc;
    double last_residual = current_residual;
    if ((rc = dev_norm(c, W, c->d_rhs, D, &current_residual))) return rc;
    if (current_residual != current_residual) { c->err = "solid: NaN residual"; rc = RDC_E_DIVERGED; break; }
    double nt = 0.0;
    if ((rc = dev_norm(c, W, c->d_u, D, &nt))) return rc;
    if (nt > max_solution) max_solution = nt;
    if (outer > 0) {
      if (require_reduction) {
        double step = 1.0;   // x currently = previous - d; walk back towards the previous iterate while the residual is not lower
        while (!(current_residual < last_residual) && step > 1e-6) {
          step *= 0.5;
          k_axpy<<<296, 256, 0, c->stream>>>(D, step, W->d_dx, c->d_u);
          c->st.kernel_launches++;
          c->u_ghost_fresh = false;
          if ((rc = solid_assemble(c, W, pseudo_time))) return rc;
          if ((rc = dev_norm(c, W, c->d_rhs, D, &current_residual))) return rc;
        }
        norm_delta *= step;
      }
      bool has = false;
      if (current_residual < abs_res) has = true;
      if (current_residual / max_residual < rel_res) has = true;
      if (linear_finished && max_solution != 0.0 && norm_delta / max_solution < rel_step) has = true;
      if (has) { converged = 1; break; }
    }
    if (outer == max_nl) break;
    if (current_residual == 0.0) { converged = 1; break; }
    if (current_residual > max_residual) max_residual = current_residual;
    if (current_residual * lin_mult < lin_tol) lin_tol = current_residual * lin_mult;
    if (lin_tol < lin_min) lin_tol = lin_min;
    if (lin_tol < abs_res / current_residual / 10.0) lin_tol = abs_res / current_residual / 10.0;
    // J d = R, d = 0: the Krylov solvers work on the context's solution vector, so it points at d for the solve
    RDC_CUDA(cudaMemsetAsync(W->d_dx, 0, (size_t)c->S.n_loc * 3 * sizeof(double), c->stream));
    std::swap(c->d_u, W->d_dx);
    int its = 0;
    double res = 0.0;
    c->st.n_spmv = 0;
    rc = solver_solve(c, ksp, RDC_PC_JACOBI, lin_tol, max_lin, 30, &its, &res);
    std::swap(c->d_u, W->d_dx);
    c->st.sum_iterations += its;
    c->st.n_solves++;
    if (rc) return rc;
    inner += its;
    linear_finished = its != max_lin;
    if ((rc = dev_norm(c, W, W->d_dx, D, &norm_delta))) return rc;
    k_axpy<<<296, 256, 0, c->stream>>>(D, -1.0, W->d_dx, c->d_u);   // newton_iterate.add(-1, linear_solution): owned dofs
    c->st.kernel_launches++;
    c->u_ghost_fresh = false;                                       // the ghost copies follow at the next assembly
    RDC_CUDA(cudaGetLastError());
  }
  c->u_ghost_fresh = false;
  if (info) { info[0] = outer; info[1] = inner; info[2] = current_residual; info[3] = converged; }
  return rc;
}

// SolidSystem::post_process (solid_system.C:394-538): per element mean normal stress, von Mises stress, current fibre
// vector (global element order; any of the outputs may be NULL)
extern "C" int rdc_solid_post_process(rdc_ctx* c, double pseudo_time, double* press, double* vm, double* fibre) {
  CHECK_SOLID(c);
  SolidWork* W;
  int rc = solid_work(c, &W);
  if (rc) return rc;
  if (!W->d_xund) { c->err = "solid: rdc_solid_set_reference has not been called"; return RDC_E_STATE; }
  const HostSetup& S = c->S;
  const int64_t E = S.E_loc;
  if (!W->d_post) RDC_CUDA(cudaMalloc(&W->d_post, (size_t)std::max<int64_t>(E, 1) * 5 * sizeof(double)));
  if ((rc = refresh_u_ghosts(c))) return rc;
  SolidArgs A;
  fill_solid_args(W, pseudo_time, &A);
  const int grid = (int)((E + 127) / 128);
  if (c->nen == 4) k_solid_post<4><<<grid, 128, 0, c->stream>>>(E, c->d_conn, c->d_u, A, W->d_post);
  else k_solid_post<8><<<grid, 128, 0, c->stream>>>(E, c->d_conn, c->d_u, A, W->d_post);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  std::vector<double> h((size_t)E * 5);
  RDC_CUDA(cudaMemcpyAsync(h.data(), W->d_post, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  RDC_CUDA(cudaStreamSynchronize(c->stream));
  std::vector<double> all;
  const double* src = h.data();
  std::vector<int64_t> glob_of;
  if (S.nranks > 1) {
    // every element is reported by exactly one rank -- the owner of its first node, on which it is always local -- and the
    // ranks' pieces are summed (zeros elsewhere): every rank receives the full arrays, like es.reinit() leaves them in libMesh
    std::vector<int32_t> first((size_t)E);
    {
      std::vector<int32_t> conn((size_t)E * c->nen);
      RDC_CUDA(cudaMemcpy(conn.data(), c->d_conn, conn.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
      for (int64_t le = 0; le < E; le++) first[le] = conn[(size_t)le * c->nen];
    }
    all.assign((size_t)S.E_glob * 5, 0.0);
    for (int64_t le = 0; le < E; le++)
      if (first[le] < S.n_owned)
        for (int k = 0; k < 5; k++) all[(size_t)S.elem_glob[le] * 5 + k] = h[(size_t)le * 5 + k];
    double* d_all = nullptr;
    RDC_CUDA(cudaMalloc(&d_all, all.size() * sizeof(double)));
    RDC_CUDA(cudaMemcpyAsync(d_all, all.data(), all.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    rc = allreduce_sum(c, d_all, (int)all.size());
    if (!rc && cudaMemcpyAsync(all.data(), d_all, all.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) rc = RDC_E_CUDA;
    if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = RDC_E_CUDA;
    cudaFree(d_all);
    if (rc) { if (c->err.empty()) c->err = "solid post-processing: exchange failed"; return rc; }
    src = all.data();
  }
  const int64_t n_out = S.nranks > 1 ? S.E_glob : E;
  for (int64_t k = 0; k < n_out; k++) {
    const int64_t g = S.nranks > 1 ? k : S.elem_glob[k];
    if (press) press[g] = src[(size_t)k * 5];
    if (vm) vm[g] = src[(size_t)k * 5 + 1];
    if (fibre) for (int d = 0; d < 3; d++) fibre[(size_t)g * 3 + d] = src[(size_t)k * 5 + 2 + d];
  }
  return RDC_OK;
}

// ---- host-only probes of the element arithmetic (no device needed; CPU tests hold them to the oracle) -------------------
static const FeTable& host_table(int elem_type) {
  static FeTable t[2];
  static bool done = false;
  if (!done) { fe_table_fill(&t[0], RDC_TET4); fe_table_fill(&t[1], RDC_HEX8); done = true; }
  return t[elem_type == RDC_TET4 ? 0 : 1];
}
// row `li` of one element: R[3], K[9*nen] (entry plane a*3+c, column node j at K[(a*3+c)*nen + j])
extern "C" int rdc_solid_probe_row(int elem_type, const double* Xc, const double* Xu, const double* mat6, double pseudo_time,
                                   const double* eta, int li, int use_symmetry, double* R, double* K) {
  if ((elem_type != RDC_TET4 && elem_type != RDC_HEX8) || !Xc || !Xu || !mat6 || !eta || !R || !K) return RDC_E_ARG;
  const int nen = elem_type == RDC_TET4 ? 4 : 8;
  for (int k = 0; k < 3; k++) R[k] = 0.0;
  for (int k = 0; k < 9 * nen; k++) K[k] = 0.0;
  if (nen == 4) solid_row<4>(host_table(elem_type), (const double (*)[3])Xc, (const double (*)[3])Xu, mat6, pseudo_time, eta, li, R, K, 1, use_symmetry);
  else solid_row<8>(host_table(elem_type), (const double (*)[3])Xc, (const double (*)[3])Xu, mat6, pseudo_time, eta, li, R, K, 1, use_symmetry);
  return RDC_OK;
}
// penalty row of node `i` of a side with ns nodes: R[3], Kd[ns*3]
extern "C" int rdc_solid_probe_bc_row(int ns, const double* Xc, const double* Xu, const double* disp, double pseudo_time, double penalty,
                                      int i, double* R, double* Kd) {
  if ((ns != 3 && ns != 4) || !Xc || !Xu || !disp || !R || !Kd) return RDC_E_ARG;
  for (int k = 0; k < 3; k++) R[k] = 0.0;
  for (int k = 0; k < ns * 3; k++) Kd[k] = 0.0;
  if (ns == 3) solid_bc_row<3>((const double (*)[3])Xc, (const double (*)[3])Xu, disp, pseudo_time, penalty, i, R, Kd);
  else solid_bc_row<4>((const double (*)[3])Xc, (const double (*)[3])Xu, disp, pseudo_time, penalty, i, R, Kd);
  return RDC_OK;
}
extern "C" int rdc_solid_probe_post(int elem_type, const double* Xc, const double* Xu, const double* mat6, double pseudo_time,
                                    const double* eta, double* out5) {
  if ((elem_type != RDC_TET4 && elem_type != RDC_HEX8) || !Xc || !Xu || !mat6 || !eta || !out5) return RDC_E_ARG;
  if (elem_type == RDC_TET4) solid_post_elem<4>(host_table(elem_type), (const double (*)[3])Xc, (const double (*)[3])Xu, mat6, pseudo_time, eta, out5);
  else solid_post_elem<8>(host_table(elem_type), (const double (*)[3])Xc, (const double (*)[3])Xu, mat6, pseudo_time, eta, out5);
  return RDC_OK;
}
