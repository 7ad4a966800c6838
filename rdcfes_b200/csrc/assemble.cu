// assemble.cu -- deterministic, atomic-free FE assembly of K and F into the row-local block-CSR operator.
//
// Replaces the element loop of assemble_adpm / _pihna / _ripf / _proteas_model / _hcc
// (adpm.C:416-650 and siblings): dof_indices, DenseMatrix Ke/Fe, fe->reinit, the qp x i x j nest and
// add_matrix/add_vector (MatSetValues ADD_VALUES) are fused into one kernel.
//
// Work decomposition:
//   * one thread per (owned node, incident element) PAIR -- the row of Ke that belongs to that node.
//     Pairs are sorted by node, then element id; a CTA owns a run of whole nodes with <= PAIRS pairs.
//   * phase 1: each thread gathers its element (connectivity, padded coordinates, old solution, element
//     and nodal aux fields), evaluates geometry + the model's coefficient table at the quadrature points
//     and forms ITS ROW of Ke (nen blocks of v x v) and of Fe in registers; rows go to shared memory.
//   * phase 2: the staged contributions of every block of the CTA's rows are added in the fixed order of a
//     precomputed contributor list (ascending element id) and every value of K is written exactly once,
//     coalesced.  The work list is cut into TASKS of at most 8 (16) contributors -- a diagonal block has 24 and
//     more contributors, an off-diagonal one ~6, and a warp runs as long as its longest lane -- the pieces of a
//     split block leave partial sums in shared memory that the first piece adds in piece order.  No float
//     atomics; the summation order is fixed by the set-up, so two assemblies are bit-identical.
// Operator layout ("row-local SoA"): block row i has L_i blocks; only the NKV structurally non-zero entries
// (a,b) of the model's v x v node block are stored (KMASK); entry plane s = rank of bit a*v+b in KMASK:
//   val[rowptr[i]*NKV + s*L_i + k]   so that a half-warp reading one row is fully coalesced.
// The CTA that finishes the diagonal block of a row also writes the point-Jacobi scaling 1/K_ii.
#include <stdio.h>
#include <stdlib.h>

#include "models.cuh"
#include "rdc_internal.h"
#include "asm_common.cuh"

namespace rdc {

__constant__ FeTable c_fe[2];  // [0] TET4, [1] HEX8

int upload_fe_tables() {
  FeTable t[2];
  fe_table_fill(&t[0], RDC_TET4);
  fe_table_fill(&t[1], RDC_HEX8);
  return cudaMemcpyToSymbol(c_fe, t, sizeof(t)) == cudaSuccess ? 0 : -1;
}

// value at a quadrature point, summed exactly like the reference does (adpm.C:464-471): each product and
// each sum rounded separately, l = 0..nen-1 (bit-identical operands for the threshold decisions)
template <int NEN>
__device__ __forceinline__ double interp(const double* phi_q, const double* nodal) {
  double v = 0.0;
#pragma unroll
  for (int l = 0; l < NEN; l++) v = add_rn(v, mul_rn(phi_q[l], nodal[l]));
  return v;
}

template <class M>
__device__ __forceinline__ void load_aux(const AsmArgs& A, int node, double* out);
template <> __device__ __forceinline__ void load_aux<Adpm>(const AsmArgs&, int, double*) {}
template <> __device__ __forceinline__ void load_aux<Pihna>(const AsmArgs&, int, double*) {}
template <> __device__ __forceinline__ void load_aux<Hcc>(const AsmArgs&, int, double*) {}
template <> __device__ __forceinline__ void load_aux<Ripf>(const AsmArgs& A, int node, double* out) {
  out[0] = A.aux0[(size_t)node * 3 + 1];  // TD(cc), ripf.C:470
  out[1] = A.aux0[(size_t)node * 3 + 2];  // TD(fb), ripf.C:471
  out[2] = A.aux1[(size_t)node * 3 + 2];  // RT_total, ripf.C:477
}
template <> __device__ __forceinline__ void load_aux<Proteas>(const AsmArgs& A, int node, double* out) {
  out[0] = A.aux0[(size_t)node * 2 + 0];  // AUX variable 0 (proteas.C:472,481)
}

// gradients of the masked variables: g += dphi_l * U_l, l ascending, no contraction (adpm.C:469-470)
template <int NEN, int NVAR>
__device__ __forceinline__ void field_gradients(unsigned mask, const double (*dphi)[3], const double (*U)[NEN], double (*G)[3]) {
#pragma unroll
  for (int a = 0; a < NVAR; a++)
#pragma unroll
    for (int d = 0; d < 3; d++) {
      double g = 0.0;
      if (mask >> a & 1u) {
#pragma unroll
        for (int l = 0; l < NEN; l++) g = add_rn(g, mul_rn(dphi[l][d], U[a][l]));
      }
      G[a][d] = g;
    }
}

// [upstream] FEMap: inverse Jacobian entries from J = dx/dxi, same expression order as libMesh, no contraction
struct InvJac { double jac, xix, xiy, xiz, etax, etay, etaz, zex, zey, zez; };
__device__ __forceinline__ InvJac inv_jacobian(double dx_dxi, double dx_deta, double dx_dzeta, double dy_dxi, double dy_deta,
                                               double dy_dzeta, double dz_dxi, double dz_deta, double dz_dzeta) {
  InvJac r;
  const double c0 = sub_rn(mul_rn(dy_deta, dz_dzeta), mul_rn(dz_deta, dy_dzeta));
  const double c1 = sub_rn(mul_rn(dz_deta, dx_dzeta), mul_rn(dx_deta, dz_dzeta));
  const double c2 = sub_rn(mul_rn(dx_deta, dy_dzeta), mul_rn(dy_deta, dx_dzeta));
  r.jac = add_rn(add_rn(mul_rn(dx_dxi, c0), mul_rn(dy_dxi, c1)), mul_rn(dz_dxi, c2));
  const double inv = 1. / r.jac;
  r.xix = mul_rn(c0, inv); r.xiy = mul_rn(c1, inv); r.xiz = mul_rn(c2, inv);
  r.etax = mul_rn(sub_rn(mul_rn(dz_dxi, dy_dzeta), mul_rn(dy_dxi, dz_dzeta)), inv);
  r.etay = mul_rn(sub_rn(mul_rn(dx_dxi, dz_dzeta), mul_rn(dz_dxi, dx_dzeta)), inv);
  r.etaz = mul_rn(sub_rn(mul_rn(dy_dxi, dx_dzeta), mul_rn(dx_dxi, dy_dzeta)), inv);
  r.zex = mul_rn(sub_rn(mul_rn(dy_dxi, dz_deta), mul_rn(dz_dxi, dy_deta)), inv);
  r.zey = mul_rn(sub_rn(mul_rn(dz_dxi, dx_deta), mul_rn(dx_dxi, dz_deta)), inv);
  r.zez = mul_rn(sub_rn(mul_rn(dx_dxi, dy_deta), mul_rn(dy_dxi, dx_deta)), inv);
  return r;
}

template <class M, int NEN, int PAIRS, int MINB>
__global__ void __launch_bounds__(PAIRS, MINB) k_assemble(const AsmArgs A, const typename M::Params P) {
  constexpr int NV = M::NV;
  constexpr int VV = NV * NV;
  constexpr unsigned KMASK = M::CMASK | M::SMASK | M::TMASK;
  constexpr int NKV = popc(KMASK);
  constexpr int NSV = popc(M::SMASK) > 0 ? popc(M::SMASK) : 1;
  constexpr int NQP = NEN == 4 ? 5 : 8;
  constexpr int TI = NEN == 4 ? 0 : 1;
  constexpr int NAUX = M::N_NAUX;
  constexpr int NA_ = NAUX > 0 ? NAUX : 1;

  extern __shared__ double smem[];
  double* stageK = smem;                              // [NKV][NEN][PAIRS]
  double* stageF = smem + (size_t)NEN * NKV * PAIRS;  // [NV][PAIRS]
  __shared__ int s_rowptr[PAIRS + 1];
  __shared__ int s_n2e[PAIRS + 1];
  __shared__ int s_diag[PAIRS];
  __shared__ __align__(4) unsigned short s_clist_raw[PAIRS * NEN + 4];  // contributor codes j*PAIRS + pair (offset in a stage slot)

  const int tid = threadIdx.x;
  AsmCta<NEN, PAIRS> cta;
  asm_prologue<NEN, PAIRS>(A, tid, cta, s_rowptr, s_n2e, s_diag, s_clist_raw);

  // ------------------------------------------------------------------ phase 1: one pair per thread
  // my pair record: {element << 3 | local index, -, -, -, node ids} at an address known from the block index alone
  constexpr int REC4 = (NEN + 4) / 4;   // int4 words per record
  const int4* rec = reinterpret_cast<const int4*>(A.pair) + ((size_t)blockIdx.x * PAIRS + tid) * REC4;
  const int4 rec0 = __ldg(rec);
  if (rec0.x >= 0) {
    const int pk = rec0.x;
    const int e = pk >> 3, li = pk & 7;
    int en[NEN];
    {
      const int4 c4 = __ldg(rec + 1);
      en[0] = c4.x; en[1] = c4.y; en[2] = c4.z; en[3] = c4.w;
      if constexpr (NEN == 8) {
        const int4 d4 = __ldg(rec + 2);
        en[4] = d4.x; en[5] = d4.y; en[6] = d4.z; en[7] = d4.w;
      }
    }
    double X[NEN][3], U[NV][NEN], AX[NA_][NEN];
#pragma unroll
    for (int l = 0; l < NEN; l++) {
      const double2 xy = reinterpret_cast<const double2*>(A.xyz4)[2 * (size_t)en[l]];
      const double z = A.xyz4[4 * (size_t)en[l] + 2];
      X[l][0] = xy.x; X[l][1] = xy.y; X[l][2] = z;
#pragma unroll
      for (int a = 0; a < NV; a++) U[a][l] = A.u_old[(size_t)en[l] * NV + a];
      if (NAUX > 0) {
        double t[NA_];
        load_aux<M>(A, en[l], t);
#pragma unroll
        for (int m = 0; m < NAUX; m++) AX[m][l] = t[m];
      }
    }
    double ef[3] = {0.0, 0.0, 0.0};
    if (M::N_EFIELD == 3) { ef[0] = A.efield[(size_t)e * 3]; ef[1] = A.efield[(size_t)e * 3 + 1]; ef[2] = A.efield[(size_t)e * 3 + 2]; }

    double Facc[NV];
#pragma unroll
    for (int a = 0; a < NV; a++) Facc[a] = 0.0;
    const FeTable& T = c_fe[TI];

    if constexpr (NEN == 4) {
      // ---- TET4: affine map, J and dphi are element constants (FEMap, Appendix B-4)
      const InvJac ij = inv_jacobian(sub_rn(X[1][0], X[0][0]), sub_rn(X[2][0], X[0][0]), sub_rn(X[3][0], X[0][0]),
                                     sub_rn(X[1][1], X[0][1]), sub_rn(X[2][1], X[0][1]), sub_rn(X[3][1], X[0][1]),
                                     sub_rn(X[1][2], X[0][2]), sub_rn(X[2][2], X[0][2]), sub_rn(X[3][2], X[0][2]));
      double dphi[4][3];
      dphi[1][0] = ij.xix; dphi[1][1] = ij.xiy; dphi[1][2] = ij.xiz;
      dphi[2][0] = ij.etax; dphi[2][1] = ij.etay; dphi[2][2] = ij.etaz;
      dphi[3][0] = ij.zex; dphi[3][1] = ij.zey; dphi[3][2] = ij.zez;
#pragma unroll
      for (int d = 0; d < 3; d++) dphi[0][d] = sub_rn(sub_rn(-dphi[1][d], dphi[2][d]), dphi[3][d]);
      double Dg[M::NDIR], GG[4];
      {
        double G[NV][3], GA[NA_][3], dir[M::NDIR][3];
        field_gradients<4, NV>(M::GRADMASK, dphi, U, G);
        field_gradients<4, NA_>(NAUX > 0 ? M::AUXGRADMASK : 0u, dphi, AX, GA);
        M::directions(P, G, ef, GA, dir);
        double dNi[3] = {dphi[0][0], dphi[0][1], dphi[0][2]};
#pragma unroll
        for (int l = 1; l < 4; l++)
          if (li == l) { dNi[0] = dphi[l][0]; dNi[1] = dphi[l][1]; dNi[2] = dphi[l][2]; }
#pragma unroll
        for (int m = 0; m < M::NDIR; m++) Dg[m] = dot3(dir[m], dNi);
#pragma unroll
        for (int j = 0; j < 4; j++) GG[j] = dot3(dphi[j], dNi);
      }
      // The 5-point rule has phi = 1/4 at qp 0 and, at qp k >= 1, phi = 1/2 at ONE node (node k, node 0 for
      // k = 4) and 1/6 at the others, so  sum_q m_q phi_j(q) = [m_0/4 + (1/6) sum_{k>=1} m_k] + (1/3) m_{k(j)} :
      // a common base Kb plus one node-specific extra per block.  The extra of qp k goes straight to the
      // thread's own shared-memory slot of node k(j); the base (and the stiffness part) is added afterwards.
      // The qp loop is kept rolled: 5x less code and ~60 fewer live registers -> more resident warps.
      double Kb[NKV > 0 ? NKV : 1], Ss[NSV];
#pragma unroll
      for (int s = 0; s < NKV; s++) Kb[s] = 0.0;
#pragma unroll
      for (int s = 0; s < NSV; s++) Ss[s] = 0.0;
#pragma unroll 1
      for (int q = 0; q < NQP; q++) {
        double phi_q[4];
#pragma unroll
        for (int l = 0; l < 4; l++) phi_q[l] = T.phi[l][q];
        double Uq[NV], Aq[NA_];
#pragma unroll
        for (int a = 0; a < NV; a++) Uq[a] = interp<4>(phi_q, U[a]);
        Aq[0] = 0.0;
        if (NAUX > 0) {
          if (M::NV == 5 && NAUX == 1) Aq[0] = add_rn(0.0, mul_rn(phi_q[1], AX[0][1]));  // PROTEAS RTD, proteas.C:481
          else {
#pragma unroll
            for (int m = 0; m < NAUX; m++) Aq[m] = interp<4>(phi_q, AX[m]);
          }
        }
        double phi_i = phi_q[0];
#pragma unroll
        for (int l = 1; l < 4; l++)
          if (li == l) phi_i = phi_q[l];
        Coef<NV> k;
        M::coef(P, Uq, Aq, Dg, k);
        const double JxW = ij.jac * T.w[q];
        const double Wi = JxW * phi_i;
        const double wb = q == 0 ? 0.25 : (1.0 / 6.0);   // weight of this qp in the common base
        double* xslot = stageK + (size_t)(q & 3) * PAIRS + tid;  // node with phi = 1/2 at qp q>=1: q (4 -> node 0)
#pragma unroll
        for (int a = 0; a < NV; a++) Facc[a] += Wi * k.F0[a] + JxW * k.F1[a];
#pragma unroll
        for (int ab = 0; ab < VV; ab++) {
          if (!(KMASK >> ab & 1u)) continue;
          const int a = ab / NV, b = ab % NV;
          const int s = slot_of(KMASK, ab);
          double m = 0.0;
          if (M::CMASK >> ab & 1u) m = Wi * k.C[a][b];
          if (M::TMASK >> ab & 1u) m += JxW * k.T[a][b];
          Kb[s] += wb * m;
          if (q > 0) xslot[(size_t)s * NEN * PAIRS] = (1.0 / 3.0) * m;
          if (M::SMASK >> ab & 1u) Ss[slot_of(M::SMASK, ab)] += JxW * k.S[a][b];
        }
      }
      // complete my row: every node gets the base (+ stiffness part) on top of its own extra
#pragma unroll
      for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int ab = 0; ab < VV; ab++) {
          if (!(KMASK >> ab & 1u)) continue;
          const int s = slot_of(KMASK, ab);
          double v = Kb[s];
          if (M::SMASK >> ab & 1u) v += Ss[slot_of(M::SMASK, ab)] * GG[j];
          stageK[((size_t)s * NEN + j) * PAIRS + tid] += v;
        }
      }
    } else {
      // ---- HEX8: per-qp Jacobian; row accumulated directly
      double Kacc[NEN][NKV > 0 ? NKV : 1];
#pragma unroll
      for (int j = 0; j < NEN; j++)
#pragma unroll
        for (int s = 0; s < NKV; s++) Kacc[j][s] = 0.0;
#pragma unroll 1
      for (int q = 0; q < NQP; q++) {
        double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};  // J[c][r] = d x_c / d xi_r
#pragma unroll
        for (int n = 0; n < NEN; n++)
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            J[cc][0] = add_rn(J[cc][0], mul_rn(X[n][cc], T.dxi[n][q]));
            J[cc][1] = add_rn(J[cc][1], mul_rn(X[n][cc], T.deta[n][q]));
            J[cc][2] = add_rn(J[cc][2], mul_rn(X[n][cc], T.dzeta[n][q]));
          }
        const InvJac ij = inv_jacobian(J[0][0], J[0][1], J[0][2], J[1][0], J[1][1], J[1][2], J[2][0], J[2][1], J[2][2]);
        double dphi[NEN][3];
#pragma unroll
        for (int n = 0; n < NEN; n++) {
          dphi[n][0] = add_rn(add_rn(mul_rn(T.dxi[n][q], ij.xix), mul_rn(T.deta[n][q], ij.etax)), mul_rn(T.dzeta[n][q], ij.zex));
          dphi[n][1] = add_rn(add_rn(mul_rn(T.dxi[n][q], ij.xiy), mul_rn(T.deta[n][q], ij.etay)), mul_rn(T.dzeta[n][q], ij.zey));
          dphi[n][2] = add_rn(add_rn(mul_rn(T.dxi[n][q], ij.xiz), mul_rn(T.deta[n][q], ij.etaz)), mul_rn(T.dzeta[n][q], ij.zez));
        }
        double phi_q[NEN];
#pragma unroll
        for (int l = 0; l < NEN; l++) phi_q[l] = T.phi[l][q];
        double Uq[NV], Aq[NA_];
#pragma unroll
        for (int a = 0; a < NV; a++) Uq[a] = interp<NEN>(phi_q, U[a]);
        Aq[0] = 0.0;
        if (NAUX > 0) {
          if (M::NV == 5 && NAUX == 1) Aq[0] = add_rn(0.0, mul_rn(phi_q[1], AX[0][1]));
          else {
#pragma unroll
            for (int m = 0; m < NAUX; m++) Aq[m] = interp<NEN>(phi_q, AX[m]);
          }
        }
        double G[NV][3], GA[NA_][3], dir[M::NDIR][3];
        field_gradients<NEN, NV>(M::GRADMASK, dphi, U, G);
        field_gradients<NEN, NA_>(NAUX > 0 ? M::AUXGRADMASK : 0u, dphi, AX, GA);
        M::directions(P, G, ef, GA, dir);
        double dNi[3] = {dphi[0][0], dphi[0][1], dphi[0][2]};
        double phi_i = phi_q[0];
#pragma unroll
        for (int l = 1; l < NEN; l++)
          if (li == l) { dNi[0] = dphi[l][0]; dNi[1] = dphi[l][1]; dNi[2] = dphi[l][2]; phi_i = phi_q[l]; }
        double Dg[M::NDIR];
#pragma unroll
        for (int m = 0; m < M::NDIR; m++) Dg[m] = dot3(dir[m], dNi);
        Coef<NV> k;
        M::coef(P, Uq, Aq, Dg, k);
        const double JxW = ij.jac * T.w[q];
        const double Wi = JxW * phi_i;
#pragma unroll
        for (int a = 0; a < NV; a++) Facc[a] += Wi * k.F0[a] + JxW * k.F1[a];
#pragma unroll
        for (int ab = 0; ab < VV; ab++) {
          if (!(KMASK >> ab & 1u)) continue;
          const int a = ab / NV, b = ab % NV;
          const int s = slot_of(KMASK, ab);
          double mj = 0.0;
          if (M::CMASK >> ab & 1u) mj = Wi * k.C[a][b];
          if (M::TMASK >> ab & 1u) mj += JxW * k.T[a][b];
          const double sj = (M::SMASK >> ab & 1u) ? JxW * k.S[a][b] : 0.0;
#pragma unroll
          for (int j = 0; j < NEN; j++) {
            double acc = Kacc[j][s] + mj * phi_q[j];
            if (M::SMASK >> ab & 1u) acc += sj * dot3(dphi[j], dNi);
            Kacc[j][s] = acc;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < NEN; j++)
#pragma unroll
        for (int s = 0; s < NKV; s++) stageK[((size_t)s * NEN + j) * PAIRS + tid] = Kacc[j][s];
    }
#pragma unroll
    for (int a = 0; a < NV; a++) stageF[a * PAIRS + tid] = Facc[a];
  }
  asm_phase2<NV, KMASK, NEN, PAIRS>(A, tid, cta, stageK, stageF, s_rowptr, s_n2e, s_diag, s_clist_raw);
}

// ---------------------------------------------------------------------------------- host launchers
// utils.h:104,116,162: a law with cM <= 0 returns 0 -> pass cM = 0 and zero slopes to the device
static void fill_pulse(Pulse& o, const double* p) { o.cM = p[0] > 0.0 ? p[0] : 0.0; o.c0 = p[1]; o.c1 = p[2]; }
static void fill_sd(StepDecay& o, const double* p) {
  const bool on = p[0] > 0.0;
  o.cM = on ? p[0] : 0.0; o.c0 = p[1]; o.c1 = p[2]; o.slope = on ? p[0] / (p[2] - p[1]) : 0.0;
}
static void fill_tr(Trapezoid& o, const double* p) {
  const bool on = p[0] > 0.0;
  o.cM = on ? p[0] : 0.0; o.c0 = p[1]; o.c1 = p[2]; o.c2 = p[3]; o.c3 = p[4];
  o.up = on ? p[0] / (p[2] - p[1]) : 0.0; o.dn = on ? p[0] / (p[4] - p[3]) : 0.0;
}

template <class M, int NEN, int PAIRS, int MINB>
static int launch_t(rdc_ctx* c, const AsmArgs& A, const typename M::Params& P) {
  constexpr unsigned KMASK = M::CMASK | M::SMASK | M::TMASK;
  const size_t smem = ((size_t)NEN * popc(KMASK) + M::NV) * PAIRS * sizeof(double);
  static bool attr_done_dev[64] = {};   // kernel attributes are per device: a process may drive several GPUs
  bool& attr_done = attr_done_dev[c->device & 63];
  if (!attr_done) {
    RDC_CUDA(cudaFuncSetAttribute(k_assemble<M, NEN, PAIRS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RDC_CUDA(cudaFuncSetAttribute(k_assemble<M, NEN, PAIRS, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
    attr_done = true;
  }
  k_assemble<M, NEN, PAIRS, MINB><<<c->ncta, PAIRS, smem, c->stream>>>(A, P);
  c->st.kernel_launches++;
  RDC_CUDA(cudaGetLastError());
  return 0;
}

template <class M>
static int launch_m(rdc_ctx* c, const AsmArgs& A, const typename M::Params& P) {
  if (c->etype == RDC_TET4) {
    if constexpr (M::NV == 3) {  // 3-variable models: 2 CTAs of 256 pairs or 4 CTAs of 128 pairs per SM
      if (c->S.pairs_per_cta == 256) return launch_t<M, 4, 256, 2>(c, A, P);
      static int minb = -1;
      if (minb < 0) { const char* e = getenv("RDC_ASM_MINB"); minb = e ? atoi(e) : 4; }
      if (minb == 5) return launch_t<M, 4, 128, 5>(c, A, P);
      if (minb == 6) return launch_t<M, 4, 128, 6>(c, A, P);
      return launch_t<M, 4, 128, 4>(c, A, P);
    } else {
      return launch_t<M, 4, 128, 2>(c, A, P);  // 5-variable models: shared-memory stage of 128 pairs
    }
  }
  return launch_t<M, 8, 128, 1>(c, A, P);
}

// pairs per assembly CTA (= block size); bounded by the shared-memory stage of nen*nkv+v doubles per pair
int pairs_per_cta_for(int model, int etype) {
  const bool v5 = (model == RDC_PIHNA || model == RDC_PROTEAS);
  if (etype == RDC_TET4) {
    if (v5) return 128;
    const char* e = getenv("RDC_ASM_PAIRS");  // tuning knob: 128 (default, 4 CTAs/SM) or 256 pairs per CTA
    return (e && atoi(e) == 256) ? 256 : 128;
  }
  return 128;
}

int launch_assemble(rdc_ctx* c) {
  AsmArgs A;
  A.conn = c->d_conn; A.xyz4 = c->d_xyz; A.u_old = c->d_uold; A.efield = c->d_efield;
  A.aux0 = nullptr; A.aux1 = nullptr;
  A.n2e_ptr = c->d_n2e_ptr; A.pair = c->d_pair; A.rowptr = c->d_rowptr; A.cta_node = c->d_cta_node;
  A.task = reinterpret_cast<const int2*>(c->d_task); A.clist = c->d_clist; A.diag_blk = c->d_diag_blk; A.val = c->d_val; A.rhs = c->d_rhs; A.dinv = c->d_dinv;
  const double* p = c->params.data();
  const double h = c->dt / 2.0;
  switch (c->model) {
    case RDC_ADPM: {
      AdpmParams P;
      P.dt2 = h;
      fill_pulse(P.decay_PrP, p + ADPM_DECAY_PRP);
      {
        const double cm = p[ADPM_DECAY_PRP] * pow(c->time, p[ADPM_GAMMA]);  // adpm.C:369
        P.decay_PrP.cM = cm > 0.0 ? cm : 0.0;
      }
      fill_pulse(P.diffuse_A, p + ADPM_DIFFUSE_AB); fill_pulse(P.taxis1_A, p + ADPM_TAXIS1_AB);
      fill_pulse(P.taxis2_A, p + ADPM_TAXIS2_AB); fill_sd(P.produce_A, p + ADPM_PRODUCE_AB);
      fill_tr(P.transform_A, p + ADPM_TRANSFORM_AB); fill_pulse(P.decay_A, p + ADPM_DECAY_AB);
      fill_pulse(P.diffuse_T, p + ADPM_DIFFUSE_TAU); fill_pulse(P.taxis1_T, p + ADPM_TAXIS1_TAU);
      fill_pulse(P.taxis2_T, p + ADPM_TAXIS2_TAU); fill_sd(P.produce_T, p + ADPM_PRODUCE_TAU);
      fill_tr(P.transform_T, p + ADPM_TRANSFORM_TAU); fill_pulse(P.decay_T, p + ADPM_DECAY_TAU);
      P.omega_A = cos(p[ADPM_ANGLE_AB]); P.omega_T = cos(p[ADPM_ANGLE_TAU]);
      if (!c->d_efield) { c->err = "ADPM needs the tract vectors (rdc_set_elem_field slot 0)"; return RDC_E_STATE; }
      return launch_m<Adpm>(c, A, P);
    }
    case RDC_PIHNA: {
      PihnaParams P;
      P.dt2 = h;
      P.Lambda_k = p[PIHNA_LAMBDA_K]; P.Kappa_k = p[PIHNA_KAPPA_K]; P.Kappa_a = p[PIHNA_KAPPA_A]; P.ek = p[PIHNA_EK];
      P.nec_c = p[PIHNA_NECROSIS_C] / P.Kappa_k; P.nec_h = p[PIHNA_NECROSIS_H] / P.Kappa_k; P.nec_v = p[PIHNA_NECROSIS_V] / P.Kappa_k;
      P.dif_c = p[PIHNA_DIFFUSE_C]; P.tax_c = p[PIHNA_TAXIS_C]; P.dif_h = p[PIHNA_DIFFUSE_H]; P.tax_h = p[PIHNA_TAXIS_H];
      P.prod_c = p[PIHNA_PRODUCE_C]; P.c2h = p[PIHNA_SWITCH_C2H]; P.h2c = p[PIHNA_SWITCH_H2C]; P.h2n = p[PIHNA_SWITCH_H2N];
      P.dif_v = p[PIHNA_DIFFUSE_V]; P.tax_v = p[PIHNA_TAXIS_V]; P.prod_v = p[PIHNA_PRODUCE_V];
      P.sec_c = p[PIHNA_SECRETE_A_C]; P.sec_h = p[PIHNA_SECRETE_A_H]; P.upt_v = p[PIHNA_UPTAKE_A_V]; P.dec_a = p[PIHNA_DECAY_A];
      return launch_m<Pihna>(c, A, P);
    }
    case RDC_RIPF: {
      RipfParams P;
      P.dt2 = h;
      P.VF_s = p[RIPF_VF_STROMA]; P.VF_p = p[RIPF_VF_PARENCHYMA]; P.VF_e = p[RIPF_VF_EXPONENT]; P.VF_min = p[RIPF_VF_MIN_VACANT];
      P.phi_cc_B = p[RIPF_PHI_CC_B]; P.phi_cc_D = p[RIPF_PHI_CC_D]; P.phi_cc = p[RIPF_PHI_CC];
      P.phi_fb_B = p[RIPF_PHI_FB_B]; P.phi_fb_D = p[RIPF_PHI_FB_D]; P.phi_fb = p[RIPF_PHI_FB]; P.phi_tol = p[RIPF_PHI_TOL];
      P.kappa = p[RIPF_KAPPA]; P.kappa_RT_c = p[RIPF_KAPPA_RT_C]; P.delta = p[RIPF_DELTA];
      P.delta_RT_a = p[RIPF_DELTA_RT_A]; P.delta_RT_b = p[RIPF_DELTA_RT_B];
      P.lambda = p[RIPF_LAMBDA];
      P.lambda_RT_r = p[RIPF_LAMBDA_RT_R] != 0.0 ? p[RIPF_LAMBDA_RT_R] : (double)c->ripf_rt_max;  // ripf.C:398-399
      P.lambda_HU_r = p[RIPF_LAMBDA_HU_R]; P.omicro = p[RIPF_OMICRO];
      P.omicro_RT_r = p[RIPF_OMICRO_RT_R] != 0.0 ? p[RIPF_OMICRO_RT_R] : (double)c->ripf_rt_max;  // ripf.C:402-403
      P.omicro_fb_b = p[RIPF_OMICRO_FB_B]; P.omega = p[RIPF_OMEGA]; P.diffusion = p[RIPF_DIFFUSION];
      P.haptotaxis = p[RIPF_HAPTOTAXIS]; P.radiotaxis = p[RIPF_RADIOTAXIS];
      if (!c->d_rt || !c->ripf_primed) {
        c->err = "RIPF needs the RT dose field (rdc_set_nodal_field slot 0) and the pre-loop rdc_clamp (ripf.C:53)";
        return RDC_E_STATE;
      }
      A.aux0 = c->d_td; A.aux1 = c->d_rt;
      return launch_m<Ripf>(c, A, P);
    }
    case RDC_PROTEAS: {
      ProteasParams P;
      P.dt2 = h;
      P.T_max = p[PROTEAS_T_MAX]; P.RT_max = p[PROTEAS_RT_MAX]; P.rho_h = p[PROTEAS_RHO_H]; P.u_h = p[PROTEAS_U_H];
      P.delta_h = p[PROTEAS_DELTA_H]; P.a_RT_h = p[PROTEAS_A_RT_H]; P.b_RT_h = p[PROTEAS_B_RT_H]; P.nu_h = p[PROTEAS_NU_H];
      P.D_c = p[PROTEAS_D_C]; P.D_c_h = p[PROTEAS_D_C_H]; P.rho_c = p[PROTEAS_RHO_C]; P.u_c = p[PROTEAS_U_C];
      P.delta_c = p[PROTEAS_DELTA_C]; P.a_RT_c = p[PROTEAS_A_RT_C]; P.b_RT_c = p[PROTEAS_B_RT_C]; P.nu_c = p[PROTEAS_NU_C];
      P.psi_n = p[PROTEAS_PSI_N]; P.k_n = p[PROTEAS_K_N]; P.u_n = p[PROTEAS_U_N]; P.rho_v = p[PROTEAS_RHO_V]; P.nu_v = p[PROTEAS_NU_V];
      P.D_e = p[PROTEAS_D_E]; P.rho_e = p[PROTEAS_RHO_E]; P.u_e = p[PROTEAS_U_E]; P.xi_e = p[PROTEAS_XI_E];
      P.p_RT_e = p[PROTEAS_P_RT_E]; P.psi_e = p[PROTEAS_PSI_E];
      if (!c->d_aux) { c->err = "PROTEAS needs the AUX field (rdc_set_nodal_field slot 0)"; return RDC_E_STATE; }
      A.aux0 = c->d_aux;
      return launch_m<Proteas>(c, A, P);
    }
    case RDC_HCC: {
      HccParams P;
      P.dt2 = h;
      P.Lambda_k = p[HCC_LAMBDA_K]; P.Kappa_k = p[HCC_KAPPA_K]; P.ek = p[HCC_EK]; P.produce_l = p[HCC_PRODUCE_L];
      P.diffuse_c = p[HCC_DIFFUSE_C]; P.mechano_c = p[HCC_MECHANO_C]; P.produce_c = p[HCC_PRODUCE_C];
      P.nec_l = p[HCC_NECROSIS_L] / P.Kappa_k; P.nec_c = p[HCC_NECROSIS_C] / P.Kappa_k;
      return launch_m<Hcc>(c, A, P);
    }
  }
  c->err = "unknown model";
  return RDC_E_ARG;
}

unsigned model_kmask(int model) {
  switch (model) {
    case RDC_ADPM: return Adpm::CMASK | Adpm::SMASK | Adpm::TMASK;
    case RDC_PIHNA: return Pihna::CMASK | Pihna::SMASK | Pihna::TMASK;
    case RDC_RIPF: return Ripf::CMASK | Ripf::SMASK | Ripf::TMASK;
    case RDC_PROTEAS: return Proteas::CMASK | Proteas::SMASK | Proteas::TMASK;
    case RDC_HCC: return Hcc::CMASK | Hcc::SMASK | Hcc::TMASK;
    case RDC_SOLID: return 0x1FF;   // dense 3 x 3 node block (solid.cu)
  }
  return 0;
}

}  // namespace rdc
