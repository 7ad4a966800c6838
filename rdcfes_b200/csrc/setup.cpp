// setup.cpp -- one-time host set-up: node partition, local numbering, block sparsity pattern and the
// integer maps the deterministic assembly needs.  Everything here is integer work and bit-exact.
//
// Replaces what libMesh does inside es.init() (adpm.C:43): METIS partition of the replicated mesh,
// DofMap distribution and sparsity preallocation.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "rdc_internal.h"

// METIS 5 as shipped in libmetis_static.a of the CUDA toolkit: idx_t = int64, real_t = float, no header.
extern "C" int METIS_SetDefaultOptions(int64_t* options);
extern "C" int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt,
                                   int64_t* vsize, int64_t* adjwgt, int64_t* nparts, float* tpwgts, float* ubvec,
                                   int64_t* options, int64_t* objval, int64_t* part);

namespace rdc {

void fe_table_fill(FeTable* T, int elem_type) {
  std::memset(T, 0, sizeof(*T));
  if (elem_type == RDC_TET4) {
    const double sixth = 1. / 6.;
    const double P[5][3] = {{.25, .25, .25}, {.5, sixth, sixth}, {sixth, .5, sixth}, {sixth, sixth, .5}, {sixth, sixth, sixth}};
    T->nen = 4; T->nqp = 5;
    T->w[0] = -2. / 15.;
    for (int q = 1; q < 5; q++) T->w[q] = .075;
    for (int q = 0; q < 5; q++) {
      const double z1 = P[q][0], z2 = P[q][1], z3 = P[q][2];
      const double z0 = 1. - z1 - z2 - z3;
      T->phi[0][q] = z0; T->phi[1][q] = z1; T->phi[2][q] = z2; T->phi[3][q] = z3;
      T->dxi[0][q] = -1.; T->deta[0][q] = -1.; T->dzeta[0][q] = -1.;
      T->dxi[1][q] = 1.; T->deta[2][q] = 1.; T->dzeta[3][q] = 1.;
    }
  } else {
    static const int i0[8] = {0, 1, 1, 0, 0, 1, 1, 0}, i1[8] = {0, 0, 1, 1, 0, 0, 1, 1}, i2[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    const double g = 5.7735026918962576450914878050196e-01;
    const double p1[2] = {-g, g};
    T->nen = 8; T->nqp = 8;
    int q = 0;
    for (int k = 0; k < 2; k++)
      for (int j = 0; j < 2; j++)
        for (int i = 0; i < 2; i++, q++) {
          const double xi = p1[i], eta = p1[j], zeta = p1[k];
          T->w[q] = 1.0;
          const double Lx[2] = {.5 * (1. - xi), .5 * (1. + xi)}, dL[2] = {-.5, .5};
          const double Ly[2] = {.5 * (1. - eta), .5 * (1. + eta)}, Lz[2] = {.5 * (1. - zeta), .5 * (1. + zeta)};
          for (int n = 0; n < 8; n++) {
            T->phi[n][q] = Lx[i0[n]] * Ly[i1[n]] * Lz[i2[n]];
            T->dxi[n][q] = dL[i0[n]] * Ly[i1[n]] * Lz[i2[n]];
            T->deta[n][q] = Lx[i0[n]] * dL[i1[n]] * Lz[i2[n]];
            T->dzeta[n][q] = Lx[i0[n]] * Ly[i1[n]] * dL[i2[n]];
          }
        }
  }
}

// ---- node partition ---------------------------------------------------------------------------
static void rcb(const double* xyz, std::vector<int32_t>& ids, int64_t lo, int64_t hi, int p_lo, int p_hi,
                std::vector<int32_t>& owner) {
  if (p_hi - p_lo == 1) {
    for (int64_t k = lo; k < hi; k++) owner[ids[k]] = p_lo;
    return;
  }
  double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
  for (int64_t k = lo; k < hi; k++)
    for (int d = 0; d < 3; d++) {
      mn[d] = std::min(mn[d], xyz[(int64_t)ids[k] * 3 + d]);
      mx[d] = std::max(mx[d], xyz[(int64_t)ids[k] * 3 + d]);
    }
  int ax = 0;
  for (int d = 1; d < 3; d++)
    if (mx[d] - mn[d] > mx[ax] - mn[ax]) ax = d;
  const int p_mid = (p_lo + p_hi) / 2;
  const int64_t mid = lo + (hi - lo) * (p_mid - p_lo) / (p_hi - p_lo);
  std::nth_element(ids.begin() + lo, ids.begin() + mid, ids.begin() + hi, [&](int32_t a, int32_t b) {
    const double xa = xyz[(int64_t)a * 3 + ax], xb = xyz[(int64_t)b * 3 + ax];
    return xa < xb || (xa == xb && a < b);
  });
  rcb(xyz, ids, lo, mid, p_lo, p_mid, owner);
  rcb(xyz, ids, mid, hi, p_mid, p_hi, owner);
}

static int partition_nodes(int64_t N, int64_t E, int nen, const int32_t* conn, const double* xyz, int nparts,
                           int partitioner, std::vector<int32_t>& owner, std::string& err) {
  owner.assign((size_t)N, 0);
  if (nparts == 1) return 0;
  if (partitioner == 1) {
    std::vector<int32_t> ids((size_t)N);
    std::iota(ids.begin(), ids.end(), 0);
    rcb(xyz, ids, 0, N, 0, nparts, owner);
    return 0;
  }
  // nodal graph without self loops for METIS
  std::vector<int64_t> deg((size_t)N + 1, 0);
  for (int64_t e = 0; e < E; e++)
    for (int i = 0; i < nen; i++) deg[conn[e * nen + i] + 1] += nen - 1;
  for (int64_t n = 0; n < N; n++) deg[n + 1] += deg[n];
  std::vector<int32_t> raw((size_t)deg[N]);
  {
    std::vector<int64_t> cur(deg.begin(), deg.end() - 1);
    for (int64_t e = 0; e < E; e++)
      for (int i = 0; i < nen; i++)
        for (int j = 0; j < nen; j++)
          if (i != j) raw[cur[conn[e * nen + i]]++] = conn[e * nen + j];
  }
  std::vector<int64_t> xadj((size_t)N + 1, 0);
  std::vector<int64_t> adj;
  adj.reserve((size_t)N * 16);
  for (int64_t n = 0; n < N; n++) {
    auto b = raw.begin() + deg[n], e = raw.begin() + deg[n + 1];
    std::sort(b, e);
    e = std::unique(b, e);
    for (auto it = b; it != e; ++it) adj.push_back(*it);
    xadj[n + 1] = (int64_t)adj.size();
  }
  raw.clear();
  raw.shrink_to_fit();
  int64_t nv = N, ncon = 1, np = nparts, objval = 0;
  int64_t options[40];
  METIS_SetDefaultOptions(options);
  options[8] = 12345;  // METIS_OPTION_SEED: same partition on every rank
  // Balance what the ranks actually do: a node's work in assembly (incident elements) and in SpMV (row length) grows with
  // its degree, and every rank waits for the slowest one in each of the three reductions of a Krylov iteration.  Vertex
  // weight = row length, load imbalance tolerance 0.5 % (METIS_OPTION_UFACTOR = 5, default 30): with unit weights and
  // the default tolerance the 8-way split of the 10 M-tet mesh left the busiest rank 2.5 % above the mean.
  std::vector<int64_t> vwgt((size_t)N);
  for (int64_t n = 0; n < N; n++) vwgt[n] = xadj[n + 1] - xadj[n] + 1;
  options[16] = 5;     // METIS_OPTION_UFACTOR
  std::vector<int64_t> part((size_t)N, 0);
  const int rc = METIS_PartGraphKway(&nv, &ncon, xadj.data(), adj.data(), vwgt.data(), nullptr, nullptr, &np, nullptr,
                                     nullptr, options, &objval, part.data());
  if (rc != 1) {
    err = "METIS_PartGraphKway failed";
    return RDC_E_MESH;
  }
  for (int64_t n = 0; n < N; n++) owner[n] = (int32_t)part[n];
  return 0;
}

// ---- the full set-up --------------------------------------------------------------------------
// 63-bit Morton key of a point quantised to 21 bits per axis inside the bounding box [lo, lo + 1/inv)
static inline uint64_t spread21(uint64_t v) {
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

int build_setup(HostSetup& S, int elem_type, int nv, int64_t N, int64_t E, const int32_t* conn, const double* xyz,
                int rank, int nranks, int partitioner, int pairs_per_cta, int node_order, std::string& err) {
  const int nen = elem_type == RDC_TET4 ? 4 : 8;
  S.nen = nen; S.nv = nv; S.rank = rank; S.nranks = nranks; S.N_glob = N; S.E_glob = E;
  S.pairs_per_cta = pairs_per_cta;
  if (N <= 0 || E <= 0 || N > (int64_t)0x7fffffff / 8) { err = "mesh size out of range"; return RDC_E_ARG; }
  if (E >= ((int64_t)1 << 28)) { err = "more than 2^28 elements per rank are not supported"; return RDC_E_ARG; }
  for (int64_t k = 0; k < E * nen; k++)
    if (conn[k] < 0 || conn[k] >= N) { err = "connectivity entry out of range"; return RDC_E_MESH; }

  std::vector<int32_t> owner;
  int rc = partition_nodes(N, E, nen, conn, xyz, nranks, partitioner, owner, err);
  if (rc) return rc;
  if (nranks > 1) S.owner_glob = owner;

  // local node set: owned, then ghosts ordered by (owner, global id).  Owned nodes are numbered along a Morton curve of
  // their coordinates (node_order 1, default): consecutive rows then belong to one small neighbourhood, so the 16 rows of
  // an SpMV tile and the ~5 nodes of an assembly CTA share most of the x / u entries they gather (L1 reuse) whatever
  // numbering the mesh generator chose; node_order 0 keeps ascending global ids.  User-facing dof ids are untouched
  // (d_dofmap translates), and rdc_download_csr sorts by global dof, so the bit-exact pattern comparison is unaffected.
  S.glob2loc.assign((size_t)N, -1);
  S.loc2glob.clear();
  for (int64_t g = 0; g < N; g++)
    if (owner[g] == rank) S.loc2glob.push_back((int32_t)g);
  S.n_owned = (int32_t)S.loc2glob.size();
  if (node_order == 1 && S.n_owned > 1) {
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int32_t g : S.loc2glob)
      for (int d = 0; d < 3; d++) { lo[d] = std::min(lo[d], xyz[(int64_t)g * 3 + d]); hi[d] = std::max(hi[d], xyz[(int64_t)g * 3 + d]); }
    double inv[3];
    for (int d = 0; d < 3; d++) inv[d] = hi[d] > lo[d] ? 2097151.0 / (hi[d] - lo[d]) : 0.0;
    std::vector<std::pair<uint64_t, int32_t>> key((size_t)S.n_owned);
#pragma omp parallel for schedule(static)
    for (int32_t k = 0; k < S.n_owned; k++) {
      const int32_t g = S.loc2glob[k];
      uint64_t m = 0;
      for (int d = 0; d < 3; d++) {
        double q = (xyz[(int64_t)g * 3 + d] - lo[d]) * inv[d];
        if (!(q >= 0.0)) q = 0.0;
        if (q > 2097151.0) q = 2097151.0;
        m |= spread21((uint64_t)q) << d;
      }
      key[k] = std::make_pair(m, g);
    }
    std::sort(key.begin(), key.end());   // ties (coincident quantised points) by global id: deterministic
    for (int32_t k = 0; k < S.n_owned; k++) S.loc2glob[k] = key[k].second;
  }
  for (int32_t k = 0; k < S.n_owned; k++) S.glob2loc[S.loc2glob[k]] = k;
  // local elements = elements with at least one owned node
  S.elem_glob.clear();
  std::vector<int32_t> ghosts;
  {
    std::vector<uint8_t> is_ghost((size_t)N, 0);
    for (int64_t e = 0; e < E; e++) {
      bool mine = false;
      for (int i = 0; i < nen; i++) mine |= (owner[conn[e * nen + i]] == rank);
      if (!mine) continue;
      S.elem_glob.push_back(e);
      for (int i = 0; i < nen; i++) {
        const int32_t g = conn[e * nen + i];
        if (owner[g] != rank && !is_ghost[g]) { is_ghost[g] = 1; ghosts.push_back(g); }
      }
    }
  }
  std::sort(ghosts.begin(), ghosts.end(), [&](int32_t a, int32_t b) {
    return owner[a] < owner[b] || (owner[a] == owner[b] && a < b);
  });
  for (int32_t g : ghosts) { S.glob2loc[g] = (int32_t)S.loc2glob.size(); S.loc2glob.push_back(g); }
  S.n_ghost = (int32_t)ghosts.size();
  S.n_loc = S.n_owned + S.n_ghost;
  S.E_loc = (int64_t)S.elem_glob.size();
  if (S.n_owned == 0) { err = "a rank owns no nodes"; return RDC_E_MESH; }

  S.conn.resize((size_t)S.E_loc * nen);
#pragma omp parallel for schedule(static)
  for (int64_t le = 0; le < S.E_loc; le++)
    for (int i = 0; i < nen; i++) S.conn[le * nen + i] = S.glob2loc[conn[S.elem_glob[le] * nen + i]];

  // node -> (element, local index) pairs of the owned nodes, ascending element id (deterministic sum order)
  const int32_t no = S.n_owned;
  S.n2e_ptr.assign((size_t)no + 1, 0);
  for (int64_t le = 0; le < S.E_loc; le++)
    for (int i = 0; i < nen; i++) {
      const int32_t l = S.conn[le * nen + i];
      if (l < no) S.n2e_ptr[l + 1]++;
    }
  int32_t max_inc = 0;
  for (int32_t n = 0; n < no; n++) {
    max_inc = std::max(max_inc, S.n2e_ptr[n + 1]);
    const int64_t s = (int64_t)S.n2e_ptr[n] + S.n2e_ptr[n + 1];
    if (s > 0x7fffffff) { err = "too many (node, element) pairs for 32-bit offsets"; return RDC_E_ARG; }
    S.n2e_ptr[n + 1] = (int32_t)s;
  }
  if (max_inc > pairs_per_cta) {
    err = "a node has " + std::to_string(max_inc) + " incident elements; at most " + std::to_string(pairs_per_cta) +
          " are supported";
    return RDC_E_MESH;
  }
  S.pair.resize((size_t)S.n2e_ptr[no]);
  {
    std::vector<int32_t> cur(S.n2e_ptr.begin(), S.n2e_ptr.end() - 1);
    for (int64_t le = 0; le < S.E_loc; le++)
      for (int i = 0; i < nen; i++) {
        const int32_t l = S.conn[le * nen + i];
        if (l < no) S.pair[cur[l]++] = (int32_t)((le << 3) | i);
      }
  }

  // block rows: sorted unique local node ids of all nodes sharing an element with the row node (+ itself)
  S.rowptr.assign((size_t)no + 1, 0);
  std::vector<int32_t> rowlen((size_t)no, 0);
#pragma omp parallel
  {
    std::vector<int32_t> tmp;
#pragma omp for schedule(static)
    for (int32_t n = 0; n < no; n++) {
      tmp.clear();
      tmp.push_back(n);
      for (int32_t p = S.n2e_ptr[n]; p < S.n2e_ptr[n + 1]; p++) {
        const int64_t le = S.pair[p] >> 3;
        for (int j = 0; j < nen; j++) tmp.push_back(S.conn[le * nen + j]);
      }
      std::sort(tmp.begin(), tmp.end());
      rowlen[n] = (int32_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
    }
  }
  for (int32_t n = 0; n < no; n++) {
    const int64_t s = (int64_t)S.rowptr[n] + rowlen[n];
    if (s > 0x7fffffff) { err = "too many blocks for 32-bit offsets"; return RDC_E_ARG; }
    S.rowptr[n + 1] = (int32_t)s;
  }
  const int64_t nnzb = S.rowptr[no];
  for (int32_t n = 0; n < no; n++)
    if (rowlen[n] > 255) { err = "a node has more than 255 neighbours"; return RDC_E_MESH; }
  S.col.resize((size_t)nnzb);
  S.diag_blk.resize((size_t)no);
#pragma omp parallel
  {
    std::vector<int32_t> tmp;
#pragma omp for schedule(static)
    for (int32_t n = 0; n < no; n++) {
      tmp.clear();
      tmp.push_back(n);
      for (int32_t p = S.n2e_ptr[n]; p < S.n2e_ptr[n + 1]; p++) {
        const int64_t le = S.pair[p] >> 3;
        for (int j = 0; j < nen; j++) tmp.push_back(S.conn[le * nen + j]);
      }
      std::sort(tmp.begin(), tmp.end());
      const int32_t len = (int32_t)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
      std::copy(tmp.begin(), tmp.begin() + len, S.col.begin() + S.rowptr[n]);
      S.diag_blk[n] = S.rowptr[n] + (int32_t)(std::lower_bound(tmp.begin(), tmp.begin() + len, n) - tmp.begin());
    }
  }

  // assembly CTAs: consecutive whole nodes, at most pairs_per_cta pairs each
  S.cta_node.clear();
  S.cta_node.push_back(0);
  {
    int32_t start_pairs = 0;
    for (int32_t n = 0; n < no; n++) {
      if (S.n2e_ptr[n + 1] - start_pairs > pairs_per_cta || n - S.cta_node.back() >= std::min(pairs_per_cta, 255)) {
        S.cta_node.push_back(n);
        start_pairs = S.n2e_ptr[n];
      }
    }
    S.cta_node.push_back(no);
  }
  const int32_t ncta = (int32_t)S.cta_node.size() - 1;

  // contributor lists: for every block, the (pair-in-CTA, j) entries that add into it, ascending pair order
  const int64_t ncontrib = (int64_t)S.pair.size() * nen;
  if (ncontrib > 0x7fffffff) { err = "too many contributions for 32-bit offsets"; return RDC_E_ARG; }
  S.cptr.assign((size_t)nnzb + 1, 0);
#pragma omp parallel for schedule(static)
  for (int32_t n = 0; n < no; n++) {
    const int32_t* rc0 = S.col.data() + S.rowptr[n];
    const int32_t len = S.rowptr[n + 1] - S.rowptr[n];
    for (int32_t p = S.n2e_ptr[n]; p < S.n2e_ptr[n + 1]; p++) {
      const int64_t le = S.pair[p] >> 3;
      for (int j = 0; j < nen; j++) {
        const int32_t k = (int32_t)(std::lower_bound(rc0, rc0 + len, S.conn[le * nen + j]) - rc0);
        S.cptr[(size_t)S.rowptr[n] + k + 1]++;
      }
    }
  }
  for (int64_t b = 0; b < nnzb; b++) S.cptr[b + 1] += S.cptr[b];
  S.clist.resize((size_t)ncontrib);
#pragma omp parallel for schedule(dynamic, 64)
  for (int32_t cta = 0; cta < ncta; cta++) {
    const int32_t pair0 = S.n2e_ptr[S.cta_node[cta]];
    for (int32_t n = S.cta_node[cta]; n < S.cta_node[cta + 1]; n++) {
      const int32_t* rc0 = S.col.data() + S.rowptr[n];
      const int32_t len = S.rowptr[n + 1] - S.rowptr[n];
      int32_t fill[64];
      std::vector<int32_t> fillv;
      int32_t* cur = fill;
      if (len > 64) { fillv.assign((size_t)len, 0); cur = fillv.data(); }
      for (int32_t k = 0; k < len; k++) cur[k] = S.cptr[(size_t)S.rowptr[n] + k];
      for (int32_t p = S.n2e_ptr[n]; p < S.n2e_ptr[n + 1]; p++) {
        const int64_t le = S.pair[p] >> 3;
        for (int j = 0; j < nen; j++) {
          const int32_t k = (int32_t)(std::lower_bound(rc0, rc0 + len, S.conn[le * nen + j]) - rc0);
          S.clist[(size_t)cur[k]++] = (uint16_t)(j * pairs_per_cta + (p - pair0));
        }
      }
    }
  }

  // phase-2 tasks.  Task = {x, y}: x = row inside the CTA | block inside the row << 8 | first contributor (relative to the
  // CTA's contributor base) << 16 ; y = number of contributors | piece index << 8 | pieces of this block << 16.  Pieces of
  // one block are consecutive tasks of ONE warp (the list is padded with empty tasks, y = 0, where a block would straddle
  // a multiple of 32): the kernel adds the partial sums of the pieces with shuffles, in piece order.
  {
    S.task_ptr.assign((size_t)ncta + 1, 0);
    std::vector<uint8_t> chunk_of((size_t)ncta, 8);
    auto layout = [&](int32_t cta, int ch, int32_t* out) -> int {   // emits (or only counts, out == nullptr) the padded list
      const int32_t cbase = S.cptr[S.rowptr[S.cta_node[cta]]];
      int nt = 0;
      for (int32_t n = S.cta_node[cta]; n < S.cta_node[cta + 1]; n++) {
        const int32_t r0 = S.rowptr[n], L = S.rowptr[n + 1] - r0;
        for (int32_t kk = 0; kk < L; kk++) {
          const int32_t c0 = S.cptr[(size_t)r0 + kk], cnt = S.cptr[(size_t)r0 + kk + 1] - c0;
          const int np = cnt > ch ? (cnt + ch - 1) / ch : 1;
          if (np > 32) return -1;
          if (np > 1 && (nt & 31) + np > 32)
            while (nt & 31) { if (out) { out[2 * nt] = 0; out[2 * nt + 1] = 0; } nt++; }
          for (int k = 0; k < np; k++, nt++) {
            if (!out) continue;
            const int start = c0 - cbase + k * ch, len = std::min(ch, cnt - k * ch);
            out[2 * nt] = (n - S.cta_node[cta]) | (kk << 8) | (start << 16);
            out[2 * nt + 1] = len | (k << 8) | (np << 16);
          }
        }
      }
      return nt;
    };
    bool bad = false;
#pragma omp parallel for schedule(static)
    for (int32_t cta = 0; cta < ncta; cta++) {
      int ch = 8, nt = layout(cta, 8, nullptr);
      while (nt < 0 && ch < 128) { ch *= 2; nt = layout(cta, ch, nullptr); }
      if (nt < 0) { bad = true; nt = 0; }
      chunk_of[cta] = (uint8_t)ch;
      S.task_ptr[(size_t)cta + 1] = nt;
    }
    if (bad) { err = "a block of the operator has more than 4096 contributing elements"; return RDC_E_MESH; }
    for (int32_t cta = 0; cta < ncta; cta++) {
      const int64_t t = (int64_t)S.task_ptr[cta] + S.task_ptr[(size_t)cta + 1];
      if (t > 0x3fffffff) { err = "too many assembly tasks for 32-bit offsets"; return RDC_E_ARG; }
      S.task_ptr[(size_t)cta + 1] = (int32_t)t;
    }
    S.task.assign((size_t)S.task_ptr[ncta] * 2, 0);
#pragma omp parallel for schedule(static)
    for (int32_t cta = 0; cta < ncta; cta++) layout(cta, chunk_of[cta], S.task.data() + (size_t)S.task_ptr[cta] * 2);
  }

  // per-CTA padded pair records {element << 3 | local index (-1 = padding), -, -, -, node ids of the element}: thread t of
  // CTA k finds its record at (k * pairs_per_cta + t), an address that depends on nothing but the block index, and the
  // node ids arrive in the same 32-byte sector -- the coordinate / solution gathers start one load level after the kernel
  // does (was: descriptor -> pair -> connectivity -> node data).  Two arrays (codes, node ids: 20 B per pair instead of 32)
  // measured 2 % slower.
  {
    const size_t rec = (size_t)nen + 4;   // ints per record (16-byte multiple): TET4 8, HEX8 12
    S.pair_rec.assign((size_t)ncta * pairs_per_cta * rec, -1);
#pragma omp parallel for schedule(static)
    for (int32_t cta = 0; cta < ncta; cta++) {
      const int32_t p0 = S.n2e_ptr[S.cta_node[cta]], p1 = S.n2e_ptr[S.cta_node[cta + 1]];
      for (int32_t p = p0; p < p1; p++) {
        int32_t* r = S.pair_rec.data() + ((size_t)cta * pairs_per_cta + (p - p0)) * rec;
        const int64_t le = S.pair[p] >> 3;
        r[0] = S.pair[p];
        for (int j = 0; j < nen; j++) r[4 + j] = S.conn[le * nen + j];
      }
    }
  }

  // halo lists
  S.nbr_rank.clear(); S.send_ptr.assign(1, 0); S.send_idx.clear(); S.recv_ptr.assign(1, 0);
  if (nranks > 1) {
    // receive side: ghosts grouped by owner
    std::vector<int32_t> recv_cnt((size_t)nranks, 0);
    for (int32_t l = no; l < S.n_loc; l++) recv_cnt[owner[S.loc2glob[l]]]++;
    // send side: my owned nodes that some other rank sees as ghost = owned nodes sharing an element with a
    // node owned by that rank.  Derived from the replicated mesh, so no communication is needed.
    std::vector<std::vector<int32_t>> send((size_t)nranks);
    {
      // one pass over the LOCAL elements (every element with a node of mine is local): an owned node goes to every
      // other rank that owns a node of the same element; duplicates are removed by the sort below
      int owners[RDC_MAX_NEN];
      for (int64_t le = 0; le < S.E_loc; le++) {
        const int64_t e = S.elem_glob[le];
        bool mixed = false;
        for (int i = 0; i < nen; i++) { owners[i] = owner[conn[e * nen + i]]; mixed |= owners[i] != rank; }
        if (!mixed) continue;
        for (int i = 0; i < nen; i++) {
          if (owners[i] != rank) continue;
          const int32_t g = conn[e * nen + i];
          for (int j = 0; j < nen; j++) {
            const int q = owners[j];
            if (q == rank) continue;
            bool seen = false;
            for (int k = 0; k < j; k++) seen |= owners[k] == q;
            if (!seen) send[q].push_back(g);
          }
        }
      }
      for (int q = 0; q < nranks; q++) {   // the receiver orders its ghosts by global id too
        std::sort(send[q].begin(), send[q].end());
        send[q].erase(std::unique(send[q].begin(), send[q].end()), send[q].end());
      }
    }
    for (int q = 0; q < nranks; q++) {
      if (q == rank) continue;
      if (recv_cnt[q] == 0 && send[q].empty()) continue;
      S.nbr_rank.push_back(q);
      for (int32_t g : send[q]) S.send_idx.push_back(S.glob2loc[g]);
      S.send_ptr.push_back((int32_t)S.send_idx.size());
      S.recv_ptr.push_back(S.recv_ptr.back() + recv_cnt[q]);
    }
  }
  return 0;
}

// ---- SpMV tiles (k_spmv_tma) --------------------------------------------------------------------
// Consecutive block rows, at most max_rows rows and max_blocks blocks per tile: {row0, nrows, first block, nblocks}.
// Returns false when a single row is longer than a tile (the caller keeps the LDG kernel).
bool cut_spmv_tiles(const int32_t* rowptr, int32_t n_rows, int max_rows, int max_blocks, std::vector<int32_t>& tiles) {
  tiles.clear();
  tiles.reserve(((size_t)n_rows / max_rows + 16) * 4);
  for (int32_t r = 0; r < n_rows;) {
    int32_t e = r;
    while (e < n_rows && e - r < max_rows && rowptr[e + 1] - rowptr[r] <= max_blocks) e++;
    if (e == r) { tiles.clear(); return false; }
    const int32_t t[4] = {r, e - r, rowptr[r], rowptr[e] - rowptr[r]};
    tiles.insert(tiles.end(), t, t + 4);
    r = e;
  }
  return true;
}

// ---- region buckets of the save_solution reductions (reduce.cu) -----------------------------------
// counted[le] != 0: this rank counts local element le.  Elements are bucketed by region, element order kept inside a
// region (stable), and every region's bucket is cut into chunks of at most `chunk` elements, so no chunk straddles
// two regions.  perm: bucketed local element ids; chunk_ptr: [n_chunks+1] into perm; rchunk_ptr: [n_regions+1] chunk
// range of every region.
void bucket_regions(int64_t E_loc, const uint8_t* counted, const int32_t* region_of_local, int n_regions, int chunk,
                    std::vector<int32_t>& perm, std::vector<int32_t>& chunk_ptr, std::vector<int32_t>& rchunk_ptr) {
  std::vector<int64_t> cnt((size_t)n_regions + 1, 0);
  for (int64_t le = 0; le < E_loc; le++)
    if (counted[le]) cnt[region_of_local[le] + 1]++;
  for (int r = 0; r < n_regions; r++) cnt[r + 1] += cnt[r];
  perm.assign((size_t)cnt[n_regions], 0);
  {
    std::vector<int64_t> cur(cnt.begin(), cnt.end() - 1);
    for (int64_t le = 0; le < E_loc; le++)
      if (counted[le]) perm[(size_t)cur[region_of_local[le]]++] = (int32_t)le;
  }
  chunk_ptr.assign(1, 0);
  rchunk_ptr.assign((size_t)n_regions + 1, 0);
  for (int r = 0; r < n_regions; r++) {
    for (int64_t a = cnt[r]; a < cnt[r + 1]; a += chunk) chunk_ptr.push_back((int32_t)std::min<int64_t>(a + chunk, cnt[r + 1]));
    rchunk_ptr[r + 1] = (int32_t)chunk_ptr.size() - 1;
  }
}

}  // namespace rdc
