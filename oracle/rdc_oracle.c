/* rdc_oracle.c -- CPU restatement of the rdcFEs per-time-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rdcfes_b200/ may import, link or execute this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the
 * checker and as the reported CPU baseline.
 *
 * PARITY STATUS: pinned to the reference's own sources for everything that is in-tree.  oracle/ref.py compiles
 * /root/reference/src/{adpm,pihna,ripf,proteas,coupled_hcc}.C UNCHANGED (g++, serial libMesh stand-in
 * oracle/ref_shim/) into oracle/_ref/; tests/test_ref_pin.py runs this restatement next to them (K, F, check_solution,
 * the RIPF state machine, save_solution, input()) and against the vectors they produced (tests/golden/ref_*.npz):
 * F and the clamped states agree bit for bit, K entries to <= 1e-12 pure relative.  What stays "unpinned upstream" is
 * only what is not under /root/reference: libMesh's FE tables / FEMap / dof numbering and PETSc's KSP (the reference
 * ships no tests or expected outputs); those are restated identically in the oracle and in
 * the stand-in and checked by independent known-answer tests (tests/test_oracle_kat.py: analytic P1/Q1 element
 * matrices, quadrature exactness, constant preservation, row sums, finite-difference Jacobians, sparse direct solve).
 *
 * What is restated (reference file:line):
 *   rate laws            utils.h:69-90,100-187
 *   assemble_adpm        adpm.C:324-652       (qp body 460-593)
 *   assemble_pihna       pihna.C:318-758      (coefficients 444-509, Fe 514-566, Ke 571-747)
 *   assemble_ripf        ripf.C:337-673       (coefficients 486-561, Fe 566-594, Ke 599-662)
 *   assemble_proteas     proteas.C:338-705    (coefficients 488-514, Fe 517-564, Ke 571-694)
 *   assemble_hcc         coupled_hcc.C:414-649
 *   check_solution       adpm.C:654-688, pihna.C:760-803, proteas.C:707-750, coupled_hcc.C:695-731,
 *                        ripf.C:675-775
 *   time loop body       adpm.C:60-84 (rotate, solve = zero+assemble+KSP, check_solution)
 * Third-party semantics that are NOT in /root/reference (libMesh @d3bda6c, PETSc @746207a; SURVEY.md
 * Appendix B) are restated from their published algorithms and marked [upstream]:
 *   TET4/HEX8 Lagrange shape functions, QGauss THIRD rules, FEMap (J, JxW, dphi), node-blocked dof
 *   numbering, full v x v nodal coupling sparsity, MatSetValues(ADD_VALUES) into AIJ,
 *   KSPGMRES(30) + left PCILU(0) / PCBJACOBI, rtol on the preconditioned residual, non-zero guess.
 *
 * Dof convention inside the oracle: dof(node, var) = nvars*node + var (node-blocked, Appendix B-5).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/rdc.h"

#define MAXNEN 8
#define MAXQP 8
#define MAXV 5
#define MAXND (MAXNEN * MAXV)

/* ------------------------------------------------------------------------------------------------
 * [upstream] reference element tables: libMesh FE<3,LAGRANGE> FIRST + QGauss(3, THIRD)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int nen, nqp;
  double w[MAXQP];
  double phi[MAXNEN][MAXQP];
  double dxi[MAXNEN][MAXQP], deta[MAXNEN][MAXQP], dzeta[MAXNEN][MAXQP];
} fe_table;

static void fe_table_tet4(fe_table* T) {
  /* QGauss TET THIRD, negative-weight 5-point rule */
  const double sixth = 1. / 6.;
  const double P[5][3] = {{.25, .25, .25}, {.5, sixth, sixth}, {sixth, .5, sixth}, {sixth, sixth, .5},
                          {sixth, sixth, sixth}};
  T->nen = 4;
  T->nqp = 5;
  T->w[0] = -2. / 15.;
  for (int q = 1; q < 5; q++) T->w[q] = .075;
  for (int q = 0; q < 5; q++) {
    const double z1 = P[q][0], z2 = P[q][1], z3 = P[q][2];
    const double z0 = 1. - z1 - z2 - z3; /* TET4: phi = (zeta0, zeta1, zeta2, zeta3) */
    T->phi[0][q] = z0; T->phi[1][q] = z1; T->phi[2][q] = z2; T->phi[3][q] = z3;
    T->dxi[0][q] = -1.; T->deta[0][q] = -1.; T->dzeta[0][q] = -1.;
    T->dxi[1][q] = 1.;  T->deta[1][q] = 0.;  T->dzeta[1][q] = 0.;
    T->dxi[2][q] = 0.;  T->deta[2][q] = 1.;  T->dzeta[2][q] = 0.;
    T->dxi[3][q] = 0.;  T->deta[3][q] = 0.;  T->dzeta[3][q] = 1.;
  }
}

static void fe_table_hex8(fe_table* T) {
  /* QGauss HEX THIRD = 2x2x2 Gauss-Legendre tensor product, x fastest (Appendix B-2);
   * HEX8 trilinear shapes with libMesh/Gmsh node order (Appendix B-3). */
  static const int i0[8] = {0, 1, 1, 0, 0, 1, 1, 0};
  static const int i1[8] = {0, 0, 1, 1, 0, 0, 1, 1};
  static const int i2[8] = {0, 0, 0, 0, 1, 1, 1, 1};
  const double g = 5.7735026918962576450914878050196e-01;
  const double p1[2] = {-g, g};
  T->nen = 8;
  T->nqp = 8;
  int q = 0;
  for (int k = 0; k < 2; k++)
    for (int j = 0; j < 2; j++)
      for (int i = 0; i < 2; i++, q++) {
        const double xi = p1[i], eta = p1[j], zeta = p1[k];
        T->w[q] = 1.0;
        const double Lx[2] = {.5 * (1. - xi), .5 * (1. + xi)}, dLx[2] = {-.5, .5};
        const double Ly[2] = {.5 * (1. - eta), .5 * (1. + eta)};
        const double Lz[2] = {.5 * (1. - zeta), .5 * (1. + zeta)};
        for (int n = 0; n < 8; n++) {
          T->phi[n][q] = Lx[i0[n]] * Ly[i1[n]] * Lz[i2[n]];
          T->dxi[n][q] = dLx[i0[n]] * Ly[i1[n]] * Lz[i2[n]];
          T->deta[n][q] = Lx[i0[n]] * dLx[i1[n]] * Lz[i2[n]];
          T->dzeta[n][q] = Lx[i0[n]] * Ly[i1[n]] * dLx[i2[n]];
        }
      }
}

static int fe_table_init(fe_table* T, int elem_type) {
  memset(T, 0, sizeof(*T));
  if (elem_type == RDC_TET4) { fe_table_tet4(T); return 0; }
  if (elem_type == RDC_HEX8) { fe_table_hex8(T); return 0; }
  return -1;
}

/* [upstream] FEMap::compute_single_point_map for a 3D element: J = dx/dxi, JxW = det J * w,
 * dphi = J^-T grad_xi phi (Appendix B-4).  Evaluated per quadrature point for every element type. */
static void fe_reinit(const fe_table* T, const double (*X)[3], double* JxW, double (*dphi)[MAXQP][3]) {
  for (int q = 0; q < T->nqp; q++) {
    double dx_dxi = 0, dx_deta = 0, dx_dzeta = 0, dy_dxi = 0, dy_deta = 0, dy_dzeta = 0, dz_dxi = 0,
           dz_deta = 0, dz_dzeta = 0;
    for (int n = 0; n < T->nen; n++) {
      dx_dxi += X[n][0] * T->dxi[n][q]; dx_deta += X[n][0] * T->deta[n][q]; dx_dzeta += X[n][0] * T->dzeta[n][q];
      dy_dxi += X[n][1] * T->dxi[n][q]; dy_deta += X[n][1] * T->deta[n][q]; dy_dzeta += X[n][1] * T->dzeta[n][q];
      dz_dxi += X[n][2] * T->dxi[n][q]; dz_deta += X[n][2] * T->deta[n][q]; dz_dzeta += X[n][2] * T->dzeta[n][q];
    }
    const double jac = dx_dxi * (dy_deta * dz_dzeta - dz_deta * dy_dzeta) +
                       dy_dxi * (dz_deta * dx_dzeta - dx_deta * dz_dzeta) +
                       dz_dxi * (dx_deta * dy_dzeta - dy_deta * dx_dzeta);
    JxW[q] = jac * T->w[q];
    const double inv = 1. / jac;
    const double dxidx = (dy_deta * dz_dzeta - dz_deta * dy_dzeta) * inv;
    const double dxidy = (dz_deta * dx_dzeta - dx_deta * dz_dzeta) * inv;
    const double dxidz = (dx_deta * dy_dzeta - dy_deta * dx_dzeta) * inv;
    const double detadx = (dz_dxi * dy_dzeta - dy_dxi * dz_dzeta) * inv;
    const double detady = (dx_dxi * dz_dzeta - dz_dxi * dx_dzeta) * inv;
    const double detadz = (dy_dxi * dx_dzeta - dx_dxi * dy_dzeta) * inv;
    const double dzetadx = (dy_dxi * dz_deta - dz_dxi * dy_deta) * inv;
    const double dzetady = (dz_dxi * dx_deta - dx_dxi * dz_deta) * inv;
    const double dzetadz = (dx_dxi * dy_deta - dy_dxi * dx_deta) * inv;
    for (int n = 0; n < T->nen; n++) {
      dphi[n][q][0] = T->dxi[n][q] * dxidx + T->deta[n][q] * detadx + T->dzeta[n][q] * dzetadx;
      dphi[n][q][1] = T->dxi[n][q] * dxidy + T->deta[n][q] * detady + T->dzeta[n][q] * dzetady;
      dphi[n][q][2] = T->dxi[n][q] * dxidz + T->deta[n][q] * detadz + T->dzeta[n][q] * dzetadz;
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * rate laws, utils.h:69-90,100-187
 * ---------------------------------------------------------------------------------------------- */
static inline double sq(double v) { return v * v; }                       /* utils.h:69 pow2 */
static inline double heaviside(double x) { return x > 0 ? 1. : 0.; }       /* utils.h:84 */
static inline double lbound(double L, double X) { return X < L ? L : X; }  /* utils.h:86 */
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline double norm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

static double law_Pi(double C, const double* p) { /* utils.h:100-110 rectangular pulse */
  if (0.0 >= p[0]) return 0.0;
  if (C < p[1]) return 0.0;
  if (C < p[2]) return p[0];
  return 0.0;
}
static double law_SD(double C, const double* p) { /* utils.h:112-122 step decay */
  if (0.0 >= p[0]) return 0.0;
  if (C < p[1]) return p[0];
  if (C < p[2]) return p[0] * (p[2] - C) / (p[2] - p[1]);
  return 0.0;
}
static double law_dSD(double C, const double* p) { /* utils.h:123-133 */
  if (0.0 >= p[0]) return 0.0;
  if (C < p[1]) return 0.0;
  if (C < p[2]) return -p[0] / (p[2] - p[1]);
  return 0.0;
}
static double law_Tr(double C, const double* p) { /* utils.h:158-172 trapezoid */
  if (0.0 >= p[0]) return 0.0;
  if (C < p[1]) return 0.0;
  if (C < p[2]) return p[0] * (C - p[1]) / (p[2] - p[1]);
  if (C < p[3]) return p[0];
  if (C < p[4]) return p[0] * (p[4] - C) / (p[4] - p[3]);
  return 0.0;
}
static double law_dTr(double C, const double* p) { /* utils.h:173-187 */
  if (0.0 >= p[0]) return 0.0;
  if (C < p[1]) return 0.0;
  if (C < p[2]) return p[0] / (p[2] - p[1]);
  if (C < p[3]) return 0.0;
  if (C < p[4]) return -p[0] / (p[4] - p[3]);
  return 0.0;
}

/* ------------------------------------------------------------------------------------------------
 * element kernels.  Dense Ke/Fe use the reference's var-major sub-block layout (adpm.C:432-449):
 * row = a*nen + i, col = b*nen + j.  U[a][l] = old solution of variable a at local node l.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const fe_table* T;
  int nen, nqp, nv, nd;
  const double* JxW;
  double (*dphi)[MAXQP][3];
  double U[MAXV][MAXNEN];
  double DT_2, time;
  const double* p;      /* flat parameter vector (include/rdc.h) */
  const double* efield; /* ADPM: tract vector [3] of this element */
  double aux[6][MAXNEN];/* RIPF: TD cc, TD fb, RT_total at the element nodes; PROTEAS: AUX var0 */
  int ripf_rt_total_max;
} elem_ctx;

#define KE(a, b, i, j) Ke[((a) * nen + (i)) * nd + (b) * nen + (j)]
#define FE(a, i) Fe[(a) * nen + (i)]
#define N_(i) (T->phi[i][qp])
#define DN(i) (c->dphi[i][qp])

static void interp(const elem_ctx* c, int qp, int a, double* val, double* grad) {
  /* adpm.C:464-471: value and (optionally) gradient of variable a at quadrature point qp */
  double v = 0.0, g[3] = {0.0, 0.0, 0.0};
  for (int l = 0; l < c->nen; l++) {
    v += c->T->phi[l][qp] * c->U[a][l];
    if (grad)
      for (int d = 0; d < 3; d++) g[d] += c->dphi[l][qp][d] * c->U[a][l];
  }
  *val = v;
  if (grad) { grad[0] = g[0]; grad[1] = g[1]; grad[2] = g[2]; }
}

/* ---- ADPM, adpm.C:460-593 ---- */
static void elem_adpm(const elem_ctx* c, double* Ke, double* Fe) {
  const fe_table* T = c->T;
  const int nen = c->nen, nd = c->nd;
  const double DT_2 = c->DT_2;
  const double* p = c->p;
  /* adpm.C:368-371: decay/PrP scaled by pow(time, gamma) */
  const double decay_PrP[3] = {p[ADPM_DECAY_PRP] * pow(c->time, p[ADPM_GAMMA]), p[ADPM_DECAY_PRP + 1],
                               p[ADPM_DECAY_PRP + 2]};
  const double *diffuse_A = p + ADPM_DIFFUSE_AB, *taxis1_A = p + ADPM_TAXIS1_AB, *taxis2_A = p + ADPM_TAXIS2_AB,
               *produce_A = p + ADPM_PRODUCE_AB, *transform_A = p + ADPM_TRANSFORM_AB, *decay_A = p + ADPM_DECAY_AB,
               *diffuse_T = p + ADPM_DIFFUSE_TAU, *taxis1_T = p + ADPM_TAXIS1_TAU, *taxis2_T = p + ADPM_TAXIS2_TAU,
               *produce_T = p + ADPM_PRODUCE_TAU, *transform_T = p + ADPM_TRANSFORM_TAU, *decay_T = p + ADPM_DECAY_TAU;
  const double omega_A = cos(p[ADPM_ANGLE_AB]), omega_T = cos(p[ADPM_ANGLE_TAU]); /* adpm.C:413-414 */
  const double* tr = c->efield;                                                 /* adpm.C:453-458 */

  for (int qp = 0; qp < c->nqp; qp++) {
    double P, A, Tu, gA[3], gT[3];
    interp(c, qp, 0, &P, NULL);
    interp(c, qp, 1, &A, gA);
    interp(c, qp, 2, &Tu, gT);
    /* adpm.C:473-492 tract alignment */
    double tA[3] = {0, 0, 0}, tT[3] = {0, 0, 0};
    const double nA = norm3(gA), nT = norm3(gT);
    if (nA) {
      const double u[3] = {gA[0] / nA, gA[1] / nA, gA[2] / nA};
      const double d = dot3(u, tr);
      if (d > +omega_A) { tA[0] = tr[0]; tA[1] = tr[1]; tA[2] = tr[2]; }
      else if (d < -omega_A) { tA[0] = -tr[0]; tA[1] = -tr[1]; tA[2] = -tr[2]; }
    }
    if (nT) {
      const double u[3] = {gT[0] / nT, gT[1] / nT, gT[2] / nT};
      const double d = dot3(u, tr);
      if (d > +omega_T) { tT[0] = tr[0]; tT[1] = tr[1]; tT[2] = tr[2]; }
      else if (d < -omega_T) { tT[0] = -tr[0]; tT[1] = -tr[1]; tT[2] = -tr[2]; }
    }
    const double TrA = law_Tr(A, transform_A), TrT = law_Tr(Tu, transform_T);
    const double dTrA = law_dTr(A, transform_A), dTrT = law_dTr(Tu, transform_T);
    const double P0 = law_Pi(P, decay_PrP);
    const double SA = law_SD(A, produce_A), dSA = law_dSD(A, produce_A), PA = law_Pi(A, decay_A);
    const double DA = law_Pi(A, diffuse_A), X1A = law_Pi(A, taxis1_A), X2A = law_Pi(Tu, taxis2_A);
    const double ST = law_SD(Tu, produce_T), dST = law_dSD(Tu, produce_T), PT = law_Pi(Tu, decay_T);
    const double DT = law_Pi(Tu, diffuse_T), X1T = law_Pi(Tu, taxis1_T), X2T = law_Pi(A, taxis2_T);
    const double W = c->JxW[qp];

    for (int i = 0; i < nen; i++) {
      const double Ni = N_(i);
      const double* dNi = DN(i);
      /* adpm.C:497-504 */
      FE(0, i) += W * (P * Ni + DT_2 * (-TrA * P * Ni - TrT * P * Ni - P0 * P * Ni));
      /* adpm.C:506-517 */
      FE(1, i) += W * (A * Ni + DT_2 * (SA * A * Ni + TrA * P * Ni - PA * A * Ni - DA * dot3(gA, dNi) -
                                        X1A * A * dot3(tA, dNi) + X2A * A * dot3(tT, dNi)));
      /* adpm.C:519-530 */
      FE(2, i) += W * (Tu * Ni + DT_2 * (ST * Tu * Ni + TrT * P * Ni - PT * Tu * Ni - DT * dot3(gT, dNi) -
                                         X1T * Tu * dot3(tT, dNi) + X2T * Tu * dot3(tA, dNi)));
      for (int j = 0; j < nen; j++) {
        const double Nj = N_(j);
        const double* dNj = DN(j);
        const double NN = Nj * Ni;
        KE(0, 0, i, j) += W * (NN - DT_2 * (-TrA * NN - TrT * NN - P0 * NN)); /* adpm.C:535-542 */
        KE(0, 1, i, j) += W * (-DT_2 * (-dTrA * P * NN));                      /* adpm.C:543-547 */
        KE(0, 2, i, j) += W * (-DT_2 * (-dTrT * P * NN));                      /* adpm.C:548-552 */
        KE(1, 0, i, j) += W * (-DT_2 * (+TrA * NN));                           /* adpm.C:554-558 */
        KE(1, 1, i, j) += W * (NN - DT_2 * (SA * NN + dSA * A * NN + dTrA * P * NN - PA * NN -
                                            DA * dot3(dNj, dNi) - X1A * Nj * dot3(tA, dNi) +
                                            X2A * Nj * dot3(tT, dNi)));        /* adpm.C:559-571 */
        KE(2, 0, i, j) += W * (-DT_2 * (+TrT * NN));                           /* adpm.C:573-577 */
        KE(2, 2, i, j) += W * (NN - DT_2 * (ST * NN + dST * Tu * NN + dTrT * P * NN - PT * NN -
                                            DT * dot3(dNj, dNi) - X1T * Nj * dot3(tT, dNi) +
                                            X2T * Nj * dot3(tA, dNi)));        /* adpm.C:578-590 */
      }
    }
  }
}

/* ---- PIHNA, pihna.C:427-750 ---- */
static void elem_pihna(const elem_ctx* c, double* Ke, double* Fe) {
  const fe_table* T = c->T;
  const int nen = c->nen, nd = c->nd;
  const double DT_2 = c->DT_2;
  const double* p = c->p;
  const double Lambda_k = p[PIHNA_LAMBDA_K], Kappa_k = p[PIHNA_KAPPA_K], Kappa_a = p[PIHNA_KAPPA_A], ek = p[PIHNA_EK];
  const double nec_c = p[PIHNA_NECROSIS_C] / Kappa_k, nec_h = p[PIHNA_NECROSIS_H] / Kappa_k,
               nec_v = p[PIHNA_NECROSIS_V] / Kappa_k; /* pihna.C:364-366 */
  const double dif_c_ = p[PIHNA_DIFFUSE_C], tax_c_ = p[PIHNA_TAXIS_C], dif_h_ = p[PIHNA_DIFFUSE_H],
               tax_h_ = p[PIHNA_TAXIS_H], prod_c = p[PIHNA_PRODUCE_C], c2h = p[PIHNA_SWITCH_C2H],
               h2c = p[PIHNA_SWITCH_H2C], h2n = p[PIHNA_SWITCH_H2N], dif_v_ = p[PIHNA_DIFFUSE_V],
               tax_v_ = p[PIHNA_TAXIS_V], prod_v = p[PIHNA_PRODUCE_V], sec_c = p[PIHNA_SECRETE_A_C],
               sec_h = p[PIHNA_SECRETE_A_H], upt_v = p[PIHNA_UPTAKE_A_V], dec_a = p[PIHNA_DECAY_A];

  for (int qp = 0; qp < c->nqp; qp++) {
    double n, cc, h, v, a, gc[3], gh[3], gv[3], ga[3];
    interp(c, qp, 0, &n, NULL);
    interp(c, qp, 1, &cc, gc);
    interp(c, qp, 2, &h, gh);
    interp(c, qp, 3, &v, gv);
    interp(c, qp, 4, &a, ga);
    /* pihna.C:444-472 volume filling */
    double Tau, dTn, dTc, dTh, dTv;
    {
      const double Te_ = (n + cc + h + v) / Kappa_k;
      if (Te_ <= 0.0) { Tau = 1.0; dTn = dTc = dTh = dTv = 0.0; }
      else if (Te_ >= 1.0) { Tau = 0.0; dTn = dTc = dTh = dTv = 0.0; }
      else { Tau = pow(1.0 - Te_, ek); dTn = dTc = dTh = dTv = (-ek / Kappa_k) * pow(1.0 - Te_, ek - 1.0); }
    }
    /* pihna.C:474-499 vascular fraction (0/0 -> NaN propagates like the reference) */
    double Ve, dVc, dVh, dVv;
    {
      const double Ve_ = v / (cc + h + v);
      if (Ve_ <= 0.0) { Ve = 0.0; dVc = dVh = dVv = 0.0; }
      else if (Ve_ >= 1.0) { Ve = 1.0; dVc = dVh = dVv = 0.0; }
      else { Ve = Ve_; dVc = dVh = -Ve_ / (cc + h + v); dVv = (1.0 - Ve_) / (cc + h + v); }
    }
    const double Ua = a / (a + Kappa_a), dUa = 1.0 / (a + Kappa_a) - Ua / (a + Kappa_a); /* pihna.C:501-502 */
    /* pihna.C:504-509 */
    const double dif_c = cc > Lambda_k ? dif_c_ : 0.0, tax_c = cc > Lambda_k ? tax_c_ : 0.0;
    const double dif_h = h > Lambda_k ? dif_h_ : 0.0, tax_h = h > Lambda_k ? tax_h_ : 0.0;
    const double dif_v = v > Lambda_k ? dif_v_ : 0.0, tax_v = v > Lambda_k ? tax_v_ : 0.0;
    const double W = c->JxW[qp];

    for (int i = 0; i < nen; i++) {
      const double Ni = N_(i);
      const double* dNi = DN(i);
      const double Gc = dot3(gc, dNi), Gh = dot3(gh, dNi), Gv = dot3(gv, dNi), Ga = dot3(ga, dNi);
      /* pihna.C:514-522 */
      FE(0, i) += W * (n * Ni + DT_2 * (nec_c * cc * n * Ni + nec_h * h * n * Ni + nec_v * v * n * Ni +
                                        h2n * (1.0 - Ve) * h * Ni));
      /* pihna.C:524-534 */
      FE(1, i) += W * (cc * Ni + DT_2 * (prod_c * Tau * cc * Ni - c2h * (1.0 - Ve) * cc * Ni + h2c * Ve * h * Ni -
                                         nec_c * cc * n * Ni - dif_c * Tau * Gc - tax_c * Tau * cc * Gv));
      /* pihna.C:536-546 */
      FE(2, i) += W * (h * Ni + DT_2 * (c2h * (1.0 - Ve) * cc * Ni - h2c * Ve * h * Ni - nec_h * h * n * Ni -
                                        dif_h * Tau * Gh - tax_h * Tau * h * Gv - h2n * (1.0 - Ve) * h * Ni));
      /* pihna.C:548-556 */
      FE(3, i) += W * (v * Ni + DT_2 * (prod_v * Tau * Ua * v * Ni - nec_v * v * n * Ni - dif_v * Tau * Gv -
                                        tax_v * Tau * v * Ga));
      /* pihna.C:558-566 */
      FE(4, i) += W * (a * Ni + DT_2 * (sec_c * cc * Ni + sec_h * h * Ni - upt_v * v * a * Ni - dec_a * a * Ni));

      for (int j = 0; j < nen; j++) {
        const double Nj = N_(j);
        const double NN = Nj * Ni;
        const double DD = dot3(DN(j), dNi);
        /* row n, pihna.C:571-597 */
        KE(0, 0, i, j) += W * (NN - DT_2 * (nec_c * cc * NN + nec_h * h * NN + nec_v * v * NN));
        KE(0, 1, i, j) += W * (-DT_2 * (nec_c * Nj * n * Ni + h2n * (-dVc) * Nj * h * Ni));
        KE(0, 2, i, j) += W * (-DT_2 * (nec_h * Nj * n * Ni + h2n * (-dVh) * Nj * h * Ni + h2n * (1.0 - Ve) * NN));
        KE(0, 3, i, j) += W * (-DT_2 * (nec_v * Nj * n * Ni + h2n * (-dVv) * Nj * h * Ni));
        /* row c, pihna.C:599-641 */
        KE(1, 0, i, j) += W * (-DT_2 * (prod_c * dTn * Nj * cc * Ni - nec_c * cc * NN - dif_c * dTn * Nj * Gc -
                                        tax_c * dTn * Nj * cc * Gv));
        KE(1, 1, i, j) += W * (NN - DT_2 * (prod_c * Tau * NN + prod_c * dTc * Nj * cc * Ni - c2h * (1.0 - Ve) * NN -
                                            c2h * (-dVc) * Nj * cc * Ni + h2c * dVc * Nj * h * Ni - nec_c * Nj * n * Ni -
                                            dif_c * dTc * Nj * Gc - dif_c * Tau * DD - tax_c * dTc * Nj * cc * Gv -
                                            tax_c * Tau * Nj * Gv));
        KE(1, 2, i, j) += W * (-DT_2 * (prod_c * dTh * Nj * cc * Ni - c2h * (-dVh) * Nj * cc * Ni + h2c * dVh * Nj * h * Ni +
                                        h2c * Ve * NN - dif_c * dTh * Nj * Gc - tax_c * dTh * Nj * cc * Gv));
        KE(1, 3, i, j) += W * (-DT_2 * (prod_c * dTv * Nj * cc * Ni - c2h * (-dVv) * Nj * cc * Ni + h2c * dVv * Nj * h * Ni -
                                        dif_c * dTv * Nj * Gc - tax_c * dTv * Nj * cc * Gv - tax_c * Tau * cc * DD));
        /* row h, pihna.C:643-684 */
        KE(2, 0, i, j) += W * (-DT_2 * (-nec_h * h * NN - dif_h * dTn * Nj * Gh - tax_h * dTn * Nj * h * Gv));
        KE(2, 1, i, j) += W * (-DT_2 * (c2h * (1.0 - Ve) * NN + c2h * (-dVc) * Nj * cc * Ni - h2c * dVc * Nj * h * Ni -
                                        dif_h * dTc * Nj * Gh - tax_h * dTc * Nj * h * Gv - h2n * (-dVc) * Nj * h * Ni));
        KE(2, 2, i, j) += W * (NN - DT_2 * (c2h * (-dVh) * Nj * cc * Ni - h2c * dVh * Nj * h * Ni - h2c * Ve * NN -
                                            nec_h * Nj * n * Ni - dif_h * dTh * Nj * Gh - dif_h * Tau * DD -
                                            tax_h * dTh * Nj * h * Gv - tax_h * Tau * Nj * Gv -
                                            h2n * (-dVh) * Nj * h * Ni - h2n * (1.0 - Ve) * NN));
        KE(2, 3, i, j) += W * (-DT_2 * (c2h * (-dVv) * Nj * cc * Ni - h2c * dVv * Nj * h * Ni - dif_h * dTv * Nj * Gh -
                                        tax_h * dTv * Nj * h * Gv - tax_h * Tau * h * DD - h2n * (-dVv) * Nj * h * Ni));
        /* row v, pihna.C:686-724 */
        KE(3, 0, i, j) += W * (-DT_2 * (prod_v * dTn * Nj * Ua * v * Ni - nec_v * v * NN - dif_v * dTn * Nj * Gv -
                                        tax_v * dTn * Nj * v * Ga));
        KE(3, 1, i, j) += W * (-DT_2 * (prod_v * dTc * Nj * Ua * v * Ni - dif_v * dTc * Nj * Gv - tax_v * dTc * Nj * v * Ga));
        KE(3, 2, i, j) += W * (-DT_2 * (prod_v * dTh * Nj * Ua * v * Ni - dif_v * dTh * Nj * Gv - tax_v * dTh * Nj * v * Ga));
        KE(3, 3, i, j) += W * (NN - DT_2 * (prod_v * dTv * Nj * Ua * v * Ni - nec_v * Nj * n * Ni - dif_v * dTv * Nj * Gv -
                                            dif_v * Tau * DD - tax_v * dTv * Nj * v * Ga - tax_v * Tau * Nj * Ga));
        KE(3, 4, i, j) += W * (-DT_2 * (prod_v * Tau * dUa * Nj * v * Ni - tax_v * Tau * v * DD));
        /* row a, pihna.C:726-747 */
        KE(4, 1, i, j) += W * (-DT_2 * (sec_c * NN));
        KE(4, 2, i, j) += W * (-DT_2 * (sec_h * NN));
        KE(4, 3, i, j) += W * (-DT_2 * (-upt_v * Nj * a * Ni));
        KE(4, 4, i, j) += W * (NN - DT_2 * (-upt_v * v * NN - dec_a * NN));
      }
    }
  }
}

/* ---- RIPF, ripf.C:449-665.  aux[0] = TD(cc), aux[1] = TD(fb), aux[2] = RT_total at the nodes ---- */
static void elem_ripf(const elem_ctx* c, double* Ke, double* Fe) {
  const fe_table* T = c->T;
  const int nen = c->nen, nd = c->nd;
  const double DT_2 = c->DT_2;
  const double* p = c->p;
  const double VF_s = p[RIPF_VF_STROMA], VF_p = p[RIPF_VF_PARENCHYMA], VF_e = p[RIPF_VF_EXPONENT],
               VF_min = p[RIPF_VF_MIN_VACANT];
  const double phi_cc_B = p[RIPF_PHI_CC_B], phi_cc_D = p[RIPF_PHI_CC_D], phi_cc = p[RIPF_PHI_CC],
               phi_fb_B = p[RIPF_PHI_FB_B], phi_fb_D = p[RIPF_PHI_FB_D], phi_fb = p[RIPF_PHI_FB],
               phi_tol = p[RIPF_PHI_TOL];
  const double kappa = p[RIPF_KAPPA], kappa_RT_c = p[RIPF_KAPPA_RT_C], delta = p[RIPF_DELTA],
               delta_RT_a = p[RIPF_DELTA_RT_A], delta_RT_b = p[RIPF_DELTA_RT_B];
  const double lambda = p[RIPF_LAMBDA];
  /* ripf.C:398-403: zero falls back to the INT parameter RT_dose/total/max */
  const double lambda_RT_r = p[RIPF_LAMBDA_RT_R] ? p[RIPF_LAMBDA_RT_R] : (double)c->ripf_rt_total_max;
  const double lambda_HU_r = p[RIPF_LAMBDA_HU_R], omicro = p[RIPF_OMICRO];
  const double omicro_RT_r = p[RIPF_OMICRO_RT_R] ? p[RIPF_OMICRO_RT_R] : (double)c->ripf_rt_total_max;
  const double omicro_fb_b = p[RIPF_OMICRO_FB_B], omega = p[RIPF_OMEGA], diffusion = p[RIPF_DIFFUSION],
               haptotaxis = p[RIPF_HAPTOTAXIS], radiotaxis = p[RIPF_RADIOTAXIS];

  for (int qp = 0; qp < c->nqp; qp++) {
    double HU, cc, fb, gHU[3], gfb[3];
    interp(c, qp, 0, &HU, gHU);
    interp(c, qp, 1, &cc, NULL);
    interp(c, qp, 2, &fb, gfb);
    double cc_dt = 0.0, fb_dt = 0.0; /* ripf.C:467-472 */
    for (int l = 0; l < nen; l++) {
      cc_dt += T->phi[l][qp] * c->aux[0][l];
      fb_dt += T->phi[l][qp] * c->aux[1][l];
    }
    double RT = 0.0, gRT[3] = {0, 0, 0}; /* ripf.C:473-484 */
    for (int l = 0; l < nen; l++) {
      RT += T->phi[l][qp] * c->aux[2][l];
      for (int d = 0; d < 3; d++) gRT[d] += c->dphi[l][qp][d] * c->aux[2][l];
    }
    {
      const double l2 = norm3(gRT);
      if (l2) { gRT[0] /= l2; gRT[1] /= l2; gRT[2] /= l2; }
      else { gRT[0] = gRT[1] = gRT[2] = 0.0; }
    }
    /* ripf.C:486-489 */
    const double kappa_RT = kappa * exp(-kappa_RT_c * RT);
    const double delta_RT = delta * (1.0 - exp(-delta_RT_a * RT - delta_RT_b * sq(RT)));
    const double lambda_RT = lambda * (RT / lambda_RT_r);
    const double omicro_RT = omicro * lbound(0.0, 4.0 * ((RT / omicro_RT_r) - sq(RT / omicro_RT_r)));
    /* ripf.C:491-496 */
    double eps_cc = 0.0, eps_fb = 0.0;
    if (cc_dt > phi_tol) eps_cc = phi_cc_B; else if (cc_dt < -phi_tol) eps_cc = phi_cc_D;
    if (fb_dt > phi_tol) eps_fb = phi_fb_B; else if (fb_dt < -phi_tol) eps_fb = phi_fb_D;
    /* ripf.C:498-514 */
    const double VF_total = VF_s + VF_p + (cc + fb);
    double Tau = 0.0, dTcc = 0.0, dTfb = 0.0;
    if (VF_total < 1.0) {
      Tau = pow(1.0 - VF_total, VF_e);
      dTcc = dTfb = -VF_e * pow(1.0 - VF_total, VF_e - 1.0);
      if (Tau < VF_min) { Tau = 0.0; dTcc = dTfb = 0.0; }
    }
    /* ripf.C:516-523 */
    double Koppa = 0.0, dKoppa = 0.0;
    if (cc < 0.0) {}
    else if (cc < 1.0) { Koppa = 4.0 * (cc - cc * cc); dKoppa = 4.0 - 8.0 * cc; }
    /* ripf.C:525-561 */
    double Lom = 0.0, dLom_HU = 0.0, dLom_cc = 0.0, dLom_fb = 0.0;
    double Ome = 0.0, dOme_HU = 0.0, dOme_cc = 0.0, dOme_fb = 0.0;
    if (fb < 0.0) {}
    else if (fb < 1.0) {
      if (HU > lambda_HU_r && HU < 0.0) {
        Lom = (1.0 - sq(fb)) * (HU / lambda_HU_r);
        dLom_HU = (1.0 - sq(fb)) / lambda_HU_r;
        dLom_cc = 0.0;
        dLom_fb = -(2.0 * fb) * (HU / lambda_HU_r);
      } else if (HU < lambda_HU_r) {
        Lom = (1.0 - sq(fb));
        dLom_HU = 0.0; dLom_cc = 0.0;
        dLom_fb = -(2.0 * fb);
      }
      if (fb <= omicro_fb_b) { Ome = 4.0 * (omicro_fb_b - sq(omicro_fb_b)); dOme_HU = dOme_cc = dOme_fb = 0.0; }
      else { Ome = 4.0 * (fb - sq(fb)); dOme_HU = dOme_cc = 0.0; dOme_fb = 4.0 - 8.0 * fb; }
    }
    const double W = c->JxW[qp];

    for (int i = 0; i < nen; i++) {
      const double Ni = N_(i);
      const double* dNi = DN(i);
      const double Gfb = dot3(gfb, dNi), GHU = dot3(gHU, dNi), GRT = dot3(gRT, dNi);
      /* ripf.C:566-574 */
      FE(0, i) += W * (HU * Ni + DT_2 * (eps_cc * cc * Ni + eps_fb * fb * Ni + phi_cc * cc_dt * Ni + phi_fb * fb_dt * Ni));
      /* ripf.C:576-582 */
      FE(1, i) += W * (cc * Ni + DT_2 * (kappa_RT * Tau * Koppa * Ni - delta_RT * cc * Ni));
      /* ripf.C:584-594 */
      FE(2, i) += W * (fb * Ni + DT_2 * (lambda_RT * Tau * Lom * Ni + omicro_RT * Tau * Ome * Ni - omega * fb * Ni -
                                         diffusion * Tau * Gfb - haptotaxis * Tau * (GHU * fb) - radiotaxis * Tau * (GRT * fb)));
      for (int j = 0; j < nen; j++) {
        const double Nj = N_(j);
        const double NN = Nj * Ni;
        const double DD = dot3(DN(j), dNi);
        KE(0, 0, i, j) += W * (NN);                                   /* ripf.C:599-603 */
        KE(0, 1, i, j) += W * (-DT_2 * (eps_cc * NN));                /* ripf.C:604-608 */
        KE(0, 2, i, j) += W * (-DT_2 * (eps_fb * NN));                /* ripf.C:609-613 */
        KE(1, 1, i, j) += W * (NN - DT_2 * (kappa_RT * dTcc * Koppa * NN + kappa_RT * Tau * dKoppa * NN - delta_RT * NN)); /* :615-622 */
        KE(1, 2, i, j) += W * (-DT_2 * (kappa_RT * dTfb * Koppa * NN)); /* ripf.C:623-627 */
        KE(2, 0, i, j) += W * (-DT_2 * (lambda_RT * Tau * dLom_HU * NN + omicro_RT * Tau * dOme_HU * NN -
                                        haptotaxis * Tau * (DD * fb)));  /* ripf.C:629-635 */
        KE(2, 1, i, j) += W * (-DT_2 * (lambda_RT * dTcc * Lom * NN + lambda_RT * Tau * dLom_cc * NN +
                                        omicro_RT * dTcc * Ome * NN + omicro_RT * Tau * dOme_cc * NN -
                                        diffusion * dTcc * Nj * Gfb - haptotaxis * dTcc * Nj * (GHU * fb) -
                                        radiotaxis * dTcc * Nj * (GRT * fb))); /* ripf.C:636-646 */
        KE(2, 2, i, j) += W * (NN - DT_2 * (lambda_RT * dTfb * Lom * NN + lambda_RT * Tau * dLom_fb * NN +
                                            omicro_RT * dTfb * Ome * NN + omicro_RT * Tau * dOme_fb * NN - omega * NN -
                                            diffusion * dTfb * Nj * Gfb - diffusion * Tau * DD -
                                            haptotaxis * dTfb * Nj * (GHU * fb) - haptotaxis * Tau * (GHU * Nj) -
                                            radiotaxis * dTfb * Nj * (GRT * fb) - radiotaxis * Tau * (GRT * Nj))); /* :647-662 */
      }
    }
  }
}

/* ---- PROTEAS, proteas.C:454-698.  aux[0][l] = AUX variable 0 at local node l (Appendix C-4) ---- */
static void elem_proteas(const elem_ctx* c, double* Ke, double* Fe) {
  const fe_table* T = c->T;
  const int nen = c->nen, nd = c->nd;
  const double DT_2 = c->DT_2;
  const double* p = c->p;
  const double T_max = p[PROTEAS_T_MAX], RT_max = p[PROTEAS_RT_MAX];
  const double rho_h = p[PROTEAS_RHO_H], u_h = p[PROTEAS_U_H], delta_h = p[PROTEAS_DELTA_H], a_RT_h = p[PROTEAS_A_RT_H],
               b_RT_h = p[PROTEAS_B_RT_H], nu_h = p[PROTEAS_NU_H];
  const double D_c = p[PROTEAS_D_C], D_c_h = p[PROTEAS_D_C_H], rho_c = p[PROTEAS_RHO_C], u_c = p[PROTEAS_U_C],
               delta_c = p[PROTEAS_DELTA_C], a_RT_c = p[PROTEAS_A_RT_C], b_RT_c = p[PROTEAS_B_RT_C], nu_c = p[PROTEAS_NU_C];
  const double psi_n = p[PROTEAS_PSI_N], k_n = p[PROTEAS_K_N], u_n = p[PROTEAS_U_N];
  const double rho_v = p[PROTEAS_RHO_V], nu_v = p[PROTEAS_NU_V];
  const double D_e = p[PROTEAS_D_E], rho_e = p[PROTEAS_RHO_E], u_e = p[PROTEAS_U_E], xi_e = p[PROTEAS_XI_E],
               p_RT_e = p[PROTEAS_P_RT_E], psi_e = p[PROTEAS_PSI_E];

  for (int qp = 0; qp < c->nqp; qp++) {
    double hos, tum, nec, vsc, oed, ghos[3], gtum[3], goed[3];
    interp(c, qp, 0, &hos, ghos);
    interp(c, qp, 1, &tum, gtum);
    interp(c, qp, 2, &nec, NULL);
    interp(c, qp, 3, &vsc, NULL);
    interp(c, qp, 4, &oed, goed);
    /* proteas.C:470-486: only ONE shape function each; RTD uses variable-0 dof of local node 1.
     * (HU, GRAD_HU, GRAD_RTD are computed by the reference but never used afterwards.) */
    const double RTD = T->phi[1][qp] * c->aux[0][1];
    /* proteas.C:488-491 */
    const double Tt = hos + tum + nec + vsc;
    double Kappa = 1.0 - Tt / T_max;
    Kappa = fmin(fmax(Kappa, 0.0), 1.0);
    const double dKappa = -1.0 / T_max;
    /* proteas.C:493-514 */
    const double host_prol = rho_h * Kappa * heaviside(vsc - u_h);
    const double dhost_prol = rho_h * dKappa * heaviside(vsc - u_h);
    const double host_RT_death = delta_h * (1.0 - exp(-a_RT_h * RTD - b_RT_h * sq(RTD)));
    const double host_nec = nu_h * nec;
    const double tumour_prol = rho_c * Kappa * heaviside(vsc - u_c);
    const double dtumour_prol = rho_c * dKappa * heaviside(vsc - u_c);
    const double tumour_RT_death = delta_c * (1.0 - exp(-a_RT_c * RTD - b_RT_c * sq(RTD)));
    const double tumour_nec = nu_c * nec;
    const double nec_prol = nu_h * hos + nu_c * tum + nu_v * vsc;
    const double nec_clearance = psi_n * (1.0 - tanh(k_n * vsc - u_n));
    const double dnec_clearance_dv = psi_n * -k_n / (cosh(k_n * vsc - u_n) * cosh(k_n * vsc - u_n));
    const double vsc_prol = rho_v * Kappa * tum;
    const double dvsc_prol = rho_v * dKappa * tum;
    const double vsc_nec = nu_v * nec;
    const double oed_prol = rho_e * tum * (1.0 - tum);
    const double doed_prol_dc = rho_e * (1.0 - 2.0 * tum);
    const double oed_RT = xi_e * pow(RTD / RT_max, p_RT_e);
    const double oed_clearance = psi_e * (1.0 - heaviside(vsc - u_e));
    const double W = c->JxW[qp];

    for (int i = 0; i < nen; i++) {
      const double Ni = N_(i);
      const double* dNi = DN(i);
      const double Gt = dot3(gtum, dNi), Gh = dot3(ghos, dNi), Ge = dot3(goed, dNi);
      FE(0, i) += W * (hos * Ni + DT_2 * (+host_prol * hos * (1.0 - hos) * Ni - host_RT_death * hos * Ni - host_nec * hos * Ni)); /* :520-527 */
      FE(1, i) += W * (tum * Ni + DT_2 * (-D_c * Kappa * Gt - D_c_h * Kappa * (Gh * tum) + tumour_prol * tum * Ni -
                                          tumour_RT_death * tum * Ni - tumour_nec * tum * Ni)); /* :529-538 */
      FE(2, i) += W * (nec * Ni + DT_2 * (+nec_prol * nec * Ni - nec_clearance * nec * Ni)); /* :540-546 */
      FE(3, i) += W * (vsc * Ni + DT_2 * (+vsc_prol * vsc * Ni - vsc_nec * vsc * Ni));        /* :548-554 */
      FE(4, i) += W * (oed * Ni + DT_2 * (-D_e * Ge + oed_prol * oed * Ni - oed_RT * oed * Ni - oed_clearance * oed * Ni)); /* :556-564 */
      for (int j = 0; j < nen; j++) {
        const double Nj = N_(j);
        const double NN = Nj * Ni;
        const double DD = dot3(DN(j), dNi);
        /* host, proteas.C:571-595 */
        KE(0, 0, i, j) += W * (NN - DT_2 * (+dhost_prol * hos * (1.0 - hos) * NN + host_prol * (1.0 - 2.0 * hos) * NN -
                                            host_RT_death * NN - host_nec * NN));
        KE(0, 1, i, j) += W * (-DT_2 * (+dhost_prol * hos * (1.0 - hos) * NN));
        KE(0, 2, i, j) += W * (-DT_2 * (+dhost_prol * hos * (1.0 - hos) * NN - nu_h * Nj * hos * Ni));
        KE(0, 3, i, j) += W * (-DT_2 * (+dhost_prol * hos * (1.0 - hos) * NN));
        /* tumour, proteas.C:597-630 */
        KE(1, 0, i, j) += W * (-DT_2 * (-D_c * dKappa * Nj * Gt - D_c_h * dKappa * Nj * (Gh * tum) - D_c_h * Kappa * (DD * tum) +
                                        dtumour_prol * Nj * tum * Ni));
        KE(1, 1, i, j) += W * (NN - DT_2 * (-D_c * dKappa * Nj * Gt - D_c * Kappa * DD + dtumour_prol * Nj * tum * Ni +
                                            tumour_prol * NN - tumour_RT_death * NN - tumour_nec * NN));
        KE(1, 2, i, j) += W * (-DT_2 * (-D_c * dKappa * Nj * Gt - D_c_h * dKappa * Nj * (Gh * tum) + dtumour_prol * Nj * tum * Ni -
                                        nu_c * Nj * tum * Ni));
        KE(1, 3, i, j) += W * (-DT_2 * (-D_c * dKappa * Nj * Gt - D_c_h * dKappa * Nj * (Gh * tum) + dtumour_prol * Nj * tum * Ni));
        /* necrotic, proteas.C:632-654 */
        KE(2, 0, i, j) += W * (-DT_2 * (+nu_h * Nj * nec * Ni));
        KE(2, 1, i, j) += W * (-DT_2 * (+nu_c * Nj * nec * Ni));
        KE(2, 2, i, j) += W * (NN - DT_2 * (+nec_prol * NN - nec_clearance * NN));
        KE(2, 3, i, j) += W * (-DT_2 * (+nu_v * Nj * nec * Ni - dnec_clearance_dv * Nj * nec * Ni));
        /* vascular, proteas.C:656-679 */
        KE(3, 0, i, j) += W * (-DT_2 * (+dvsc_prol * Nj * vsc * Ni));
        KE(3, 1, i, j) += W * (-DT_2 * (+dvsc_prol * Nj * vsc * Ni));
        KE(3, 2, i, j) += W * (-DT_2 * (+dvsc_prol * Nj * vsc * Ni - nu_v * Nj * vsc * Ni));
        KE(3, 3, i, j) += W * (NN - DT_2 * (+dvsc_prol * Nj * vsc * Ni + vsc_prol * NN - vsc_nec * NN));
        /* oedema, proteas.C:681-694 */
        KE(4, 1, i, j) += W * (-DT_2 * (+doed_prol_dc * Nj * oed * Ni));
        KE(4, 4, i, j) += W * (NN - DT_2 * (-D_e * DD + oed_prol * NN - oed_RT * NN - oed_clearance * NN));
      }
    }
  }
}

/* ---- HCC, coupled_hcc.C:496-640 (quirks of Appendix C-3 reproduced) ---- */
static void elem_hcc(const elem_ctx* c, double* Ke, double* Fe) {
  const fe_table* T = c->T;
  const int nen = c->nen, nd = c->nd;
  const double DT_2 = c->DT_2;
  const double* p = c->p;
  const double Lambda_k = p[HCC_LAMBDA_K], Kappa_k = p[HCC_KAPPA_K], ek = p[HCC_EK], produce_l = p[HCC_PRODUCE_L];
  const double diffuse_c_ = p[HCC_DIFFUSE_C], mechano_c_ = p[HCC_MECHANO_C], produce_c = p[HCC_PRODUCE_C];
  const double nec_l = p[HCC_NECROSIS_L] / Kappa_k, nec_c = p[HCC_NECROSIS_C] / Kappa_k; /* coupled_hcc.C:459-460 */

  for (int qp = 0; qp < c->nqp; qp++) {
    double l_, c_, n_, gc[3];
    interp(c, qp, 0, &l_, NULL);
    interp(c, qp, 1, &c_, gc);
    interp(c, qp, 2, &n_, NULL);
    const double gsig[3] = {0.0, 0.0, 0.0}; /* coupled_hcc.C:508 */
    double Tau, dTl, dTc, dTn;              /* coupled_hcc.C:510-532 */
    {
      const double Te_ = (l_ + c_ + n_) / Kappa_k;
      if (Te_ <= 0.0) { Tau = 1.0; dTl = dTc = dTn = 0.0; }
      else if (Te_ >= 1.0) { Tau = 0.0; dTl = dTc = dTn = 0.0; }
      else { Tau = pow(1.0 - Te_, ek); dTl = dTc = dTn = (-ek / Kappa_k) * pow(1.0 - Te_, ek - 1.0); }
    }
    const double diffuse_c = c_ > Lambda_k ? diffuse_c_ : 0.0, mechano_c = c_ > Lambda_k ? mechano_c_ : 0.0; /* :534-535 */
    const double W = c->JxW[qp];

    for (int i = 0; i < nen; i++) {
      const double Ni = N_(i);
      const double* dNi = DN(i);
      const double Gc = dot3(gc, dNi), Gs = dot3(gsig, dNi);
      FE(0, i) += W * (l_ * Ni + DT_2 * (produce_l * Tau * l_ * Ni - nec_l * l_ * n_ * Ni)); /* :540-546 */
      FE(1, i) += W * (c_ * Ni + DT_2 * (produce_c * Tau * c_ * Ni - nec_c * c_ * n_ * Ni - diffuse_c * Tau * Gc -
                                         mechano_c * Tau * c_ * Gs));                         /* :548-556 */
      FE(2, i) += W * (n_ * Ni + DT_2 * (nec_l * l_ * n_ * Ni + nec_c * c_ * n_ * Ni));       /* :558-564 */
      for (int j = 0; j < nen; j++) {
        const double Nj = N_(j);
        const double NN = Nj * Ni;
        const double DD = dot3(DN(j), dNi);
        KE(0, 0, i, j) += W * (NN - DT_2 * (produce_l * Tau * NN + produce_l * dTl * Nj * l_ * Ni - nec_l * Nj * n_ * Ni)); /* :569-576 */
        KE(0, 1, i, j) += W * (NN - DT_2 * (produce_l * dTc * Nj * l_ * Ni));                 /* :577-582 (capacity on off-diagonal) */
        KE(0, 2, i, j) += W * (NN - DT_2 * (produce_l * dTn * Nj * l_ * Ni - nec_l * l_ * NN)); /* :583-589 */
        KE(1, 0, i, j) += W * (NN - DT_2 * (produce_c * dTl * Nj * c_ * Ni - diffuse_c * dTl * Nj * Gc -
                                            mechano_c * dTl * Nj * c_ * Gs));                 /* :591-598 */
        KE(1, 1, i, j) += W * (NN - DT_2 * (produce_c * Tau * NN + produce_c * dTc * Nj * c_ * Ni - nec_c * Nj * n_ * Ni -
                                            diffuse_c * dTc * Nj * Gc - diffuse_c * Tau * DD -
                                            mechano_c * dTc * Nj * c_ * Gs - mechano_c * Tau * Nj * Gs)); /* :599-610 */
        KE(1, 1, i, j) += W * (NN - DT_2 * (produce_c * dTn * Nj * c_ * Ni - nec_c * c_ * NN - diffuse_c * dTn * Nj * Gc -
                                            mechano_c * dTn * Nj * c_ * Gs));                 /* :611-619 (second write to [1][1]) */
        KE(2, 0, i, j) += W * (-DT_2 * (nec_l * Nj * n_ * Ni));                               /* :621-625 */
        KE(2, 1, i, j) += W * (-DT_2 * (nec_c * Nj * n_ * Ni));                               /* :626-630 */
        KE(2, 2, i, j) += W * (NN - DT_2 * (nec_l * l_ * NN + nec_c * c_ * NN));              /* :631-637 */
      }
    }
  }
}

int orc_model_nvars(int model) { return (model == RDC_PIHNA || model == RDC_PROTEAS) ? 5 : 3; }

/* ------------------------------------------------------------------------------------------------
 * [upstream] sparsity: (node graph + I) (x) dense v x v, scalar CSR with sorted columns (App. B-6)
 * ---------------------------------------------------------------------------------------------- */
static int cmp_i32(const void* a, const void* b) {
  const int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

/* node-level adjacency: nptr[N+1], nadj (sorted, includes self).  Returns 0 / -1. */
int orc_node_graph(int64_t N, int64_t E, int nen, const int32_t* conn, int64_t** nptr_out, int32_t** nadj_out) {
  int64_t* cnt = (int64_t*)calloc((size_t)N + 1, sizeof(int64_t));
  if (!cnt) return -1;
  for (int64_t e = 0; e < E; e++)
    for (int i = 0; i < nen; i++) cnt[conn[e * nen + i] + 1] += nen;
  for (int64_t n = 0; n < N; n++) cnt[n + 1] += cnt[n];
  int32_t* raw = (int32_t*)malloc(sizeof(int32_t) * (size_t)(cnt[N] > 0 ? cnt[N] : 1));
  int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N + 1));
  if (!raw || !fill) return -1;
  memcpy(fill, cnt, sizeof(int64_t) * (size_t)(N + 1));
  for (int64_t e = 0; e < E; e++)
    for (int i = 0; i < nen; i++) {
      const int32_t n = conn[e * nen + i];
      for (int j = 0; j < nen; j++) raw[fill[n]++] = conn[e * nen + j];
    }
  int64_t* nptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(N + 1));
  nptr[0] = 0;
  /* sort + unique in place per node, then compact */
  for (int64_t n = 0; n < N; n++) {
    int32_t* a = raw + cnt[n];
    const int64_t m = cnt[n + 1] - cnt[n];
    int64_t u = 0;
    if (m > 0) {
      qsort(a, (size_t)m, sizeof(int32_t), cmp_i32);
      u = 1;
      for (int64_t k = 1; k < m; k++)
        if (a[k] != a[u - 1]) a[u++] = a[k];
    } else { /* isolated node keeps its diagonal */
      u = 0;
    }
    fill[n] = u;
    nptr[n + 1] = nptr[n] + (u > 0 ? u : 1);
  }
  int32_t* nadj = (int32_t*)malloc(sizeof(int32_t) * (size_t)(nptr[N] > 0 ? nptr[N] : 1));
  for (int64_t n = 0; n < N; n++) {
    if (fill[n] == 0) nadj[nptr[n]] = (int32_t)n;
    else memcpy(nadj + nptr[n], raw + cnt[n], sizeof(int32_t) * (size_t)fill[n]);
  }
  free(raw); free(fill); free(cnt);
  *nptr_out = nptr;
  *nadj_out = nadj;
  return 0;
}

/* scalar CSR pattern in dof numbering nv*node+var.  rowptr [D+1] int64, col [nnz] int32. */
int orc_build_pattern(int64_t N, int64_t E, int nen, int nv, const int32_t* conn, int64_t* nnz_out, int64_t** rowptr_out,
                      int32_t** col_out) {
  int64_t* nptr; int32_t* nadj;
  if (orc_node_graph(N, E, nen, conn, &nptr, &nadj)) return -1;
  const int64_t D = N * nv;
  int64_t* rowptr = (int64_t*)malloc(sizeof(int64_t) * (size_t)(D + 1));
  rowptr[0] = 0;
  for (int64_t n = 0; n < N; n++)
    for (int a = 0; a < nv; a++) rowptr[n * nv + a + 1] = rowptr[n * nv + a] + (nptr[n + 1] - nptr[n]) * nv;
  int32_t* col = (int32_t*)malloc(sizeof(int32_t) * (size_t)rowptr[D]);
  for (int64_t n = 0; n < N; n++)
    for (int a = 0; a < nv; a++) {
      int64_t k = rowptr[n * nv + a];
      for (int64_t q = nptr[n]; q < nptr[n + 1]; q++)
        for (int b = 0; b < nv; b++) col[k++] = nadj[q] * nv + b;
    }
  free(nptr); free(nadj);
  *nnz_out = rowptr[D];
  *rowptr_out = rowptr;
  *col_out = col;
  return 0;
}

void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------------------------------------
 * global assembly: zero K,F; loop elements; MatSetValues/VecSetValues(ADD_VALUES)  (adpm.C:416-650)
 *
 * nodal_aux: RIPF [N*6] = {TD_HU, TD_cc, TD_fb, RT_broad, RT_focus, RT_total}; PROTEAS [N*2] = AUX.
 * Threading (CPU baseline only): rows are split into contiguous chunks, each thread adds only the
 * rows it owns, visiting every element (so the per-entry summation order equals the serial one and
 * the result is bit-identical for any thread count).
 * ---------------------------------------------------------------------------------------------- */
static int64_t csr_find(const int64_t* rowptr, const int32_t* col, int64_t row, int32_t c) {
  int64_t lo = rowptr[row], hi = rowptr[row + 1] - 1;
  while (lo <= hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (col[mid] == c) return mid;
    if (col[mid] < c) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

int orc_assemble(int model, int elem_type, int64_t N, int64_t E, const int32_t* conn, const double* xyz,
                 const double* u_old, const double* params, const double* elem_field, const double* nodal_aux,
                 int ripf_rt_total_max, double time, double dt, const int64_t* rowptr, const int32_t* col,
                 double* val, double* rhs, int nthreads) {
  fe_table T;
  if (fe_table_init(&T, elem_type)) return -1;
  const int nen = T.nen, nv = orc_model_nvars(model), nd = nen * nv;
  const int64_t D = N * nv;
  memset(val, 0, sizeof(double) * (size_t)rowptr[D]);
  memset(rhs, 0, sizeof(double) * (size_t)D);
  if (nthreads < 1) nthreads = 1;
  int err = 0;
#pragma omp parallel num_threads(nthreads)
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    const int tid = 0, nt = 1;
#endif
    const int64_t n_lo = N * tid / nt, n_hi = N * (tid + 1) / nt; /* owned node range */
    double Ke[MAXND * MAXND], Fe[MAXND], JxW[MAXQP];
    double dphi[MAXNEN][MAXQP][3];
    double X[MAXNEN][3];
    elem_ctx c;
    memset(&c, 0, sizeof(c));
    c.T = &T; c.nen = nen; c.nqp = T.nqp; c.nv = nv; c.nd = nd;
    c.JxW = JxW; c.dphi = dphi; c.DT_2 = dt / 2.0; c.time = time; c.p = params;
    c.ripf_rt_total_max = ripf_rt_total_max;
    for (int64_t e = 0; e < E; e++) {
      const int32_t* en = conn + e * nen;
      int mine = 0;
      for (int i = 0; i < nen; i++) mine |= (en[i] >= n_lo && en[i] < n_hi);
      if (!mine) continue;
      for (int i = 0; i < nen; i++) {
        for (int d = 0; d < 3; d++) X[i][d] = xyz[(int64_t)en[i] * 3 + d];
        for (int a = 0; a < nv; a++) c.U[a][i] = u_old[(int64_t)en[i] * nv + a];
      }
      if (model == RDC_ADPM) c.efield = elem_field + e * 3;
      if (model == RDC_RIPF)
        for (int i = 0; i < nen; i++) {
          c.aux[0][i] = nodal_aux[(int64_t)en[i] * 6 + 1];
          c.aux[1][i] = nodal_aux[(int64_t)en[i] * 6 + 2];
          c.aux[2][i] = nodal_aux[(int64_t)en[i] * 6 + 5];
        }
      if (model == RDC_PROTEAS)
        for (int i = 0; i < nen; i++) c.aux[0][i] = nodal_aux[(int64_t)en[i] * 2 + 0];
      fe_reinit(&T, (const double(*)[3])X, JxW, dphi);
      memset(Ke, 0, sizeof(double) * (size_t)(nd * nd));
      memset(Fe, 0, sizeof(double) * (size_t)nd);
      switch (model) {
        case RDC_ADPM: elem_adpm(&c, Ke, Fe); break;
        case RDC_PIHNA: elem_pihna(&c, Ke, Fe); break;
        case RDC_RIPF: elem_ripf(&c, Ke, Fe); break;
        case RDC_PROTEAS: elem_proteas(&c, Ke, Fe); break;
        case RDC_HCC: elem_hcc(&c, Ke, Fe); break;
        default: err = -1;
      }
      /* add_matrix / add_vector with dof_indices = all dofs var-major (adpm.C:418-423,648-649) */
      for (int a = 0; a < nv; a++)
        for (int i = 0; i < nen; i++) {
          if (en[i] < n_lo || en[i] >= n_hi) continue;
          const int64_t row = (int64_t)en[i] * nv + a;
          rhs[row] += Fe[a * nen + i];
          for (int b = 0; b < nv; b++)
            for (int j = 0; j < nen; j++) {
              const int64_t k = csr_find(rowptr, col, row, en[j] * nv + b);
              if (k < 0) { err = -2; continue; }
              val[k] += Ke[(a * nen + i) * nd + b * nen + j];
            }
        }
    }
  }
  return err;
}

/* single element (for the known-answer tests): returns dense Ke [nd*nd], Fe [nd], JxW, dphi */
int orc_element(int model, int elem_type, const double* X /*[nen*3]*/, const double* U /*[nen*nv] node-major*/,
                const double* params, const double* efield, const double* aux /*[nen*6] or [nen*2]*/,
                int ripf_rt_total_max, double time, double dt, double* Ke, double* Fe, double* JxW_out,
                double* dphi_out /*[nen*nqp*3]*/) {
  fe_table T;
  if (fe_table_init(&T, elem_type)) return -1;
  const int nen = T.nen, nv = orc_model_nvars(model), nd = nen * nv;
  double JxW[MAXQP];
  double dphi[MAXNEN][MAXQP][3];
  double Xl[MAXNEN][3];
  elem_ctx c;
  memset(&c, 0, sizeof(c));
  c.T = &T; c.nen = nen; c.nqp = T.nqp; c.nv = nv; c.nd = nd;
  c.JxW = JxW; c.dphi = dphi; c.DT_2 = dt / 2.0; c.time = time; c.p = params; c.efield = efield;
  c.ripf_rt_total_max = ripf_rt_total_max;
  for (int i = 0; i < nen; i++) {
    for (int d = 0; d < 3; d++) Xl[i][d] = X[i * 3 + d];
    for (int a = 0; a < nv; a++) c.U[a][i] = U[i * nv + a];
    if (model == RDC_RIPF && aux) { c.aux[0][i] = aux[i * 6 + 1]; c.aux[1][i] = aux[i * 6 + 2]; c.aux[2][i] = aux[i * 6 + 5]; }
    if (model == RDC_PROTEAS && aux) c.aux[0][i] = aux[i * 2 + 0];
  }
  fe_reinit(&T, (const double(*)[3])Xl, JxW, dphi);
  memset(Ke, 0, sizeof(double) * (size_t)(nd * nd));
  memset(Fe, 0, sizeof(double) * (size_t)nd);
  switch (model) {
    case RDC_ADPM: elem_adpm(&c, Ke, Fe); break;
    case RDC_PIHNA: elem_pihna(&c, Ke, Fe); break;
    case RDC_RIPF: elem_ripf(&c, Ke, Fe); break;
    case RDC_PROTEAS: elem_proteas(&c, Ke, Fe); break;
    case RDC_HCC: elem_hcc(&c, Ke, Fe); break;
    default: return -1;
  }
  if (JxW_out) memcpy(JxW_out, JxW, sizeof(double) * (size_t)T.nqp);
  if (dphi_out)
    for (int i = 0; i < nen; i++)
      for (int q = 0; q < T.nqp; q++)
        for (int d = 0; d < 3; d++) dphi_out[(i * T.nqp + q) * 3 + d] = dphi[i][q][d];
  return 0;
}

/* reference-element tables for the quadrature KATs */
int orc_fe_tables(int elem_type, int* nen, int* nqp, double* w, double* phi /*[nen*nqp]*/) {
  fe_table T;
  if (fe_table_init(&T, elem_type)) return -1;
  *nen = T.nen; *nqp = T.nqp;
  for (int q = 0; q < T.nqp; q++) w[q] = T.w[q];
  for (int i = 0; i < T.nen; i++)
    for (int q = 0; q < T.nqp; q++) phi[i * T.nqp + q] = T.phi[i][q];
  return 0;
}


/* ------------------------------------------------------------------------------------------------
 * save_solution: the per-region post-step reductions
 *   adpm.C:690-829 (per parcellation ID: concentration of the LAST element of the region -- the reference
 *   assigns, it does not accumulate (:780-783) -- and the volume of the elements whose nodes are all in range),
 *   pihna.C:842-976 (four thresholded volumes, one region), ripf.C:777-864 (two volumes, two conditions per node).
 * A condition is  lo <= (sum_a w[a] * u[a]) / div <= hi  at a node, the sum taken over the non-zero weights in
 * ascending variable order like the reference's expressions (c+h, (n+c+h+v)/Kappa_k, HU, cc >= min ...).
 * ---------------------------------------------------------------------------------------------- */
/* [upstream] Tet4::volume(): triple product of the edge vectors / 6; other types: sum of JxW */
static double elem_volume(const fe_table* T, int elem_type, const double (*X)[3]) {
  if (elem_type == RDC_TET4) {
    const double a[3] = {X[3][0] - X[0][0], X[3][1] - X[0][1], X[3][2] - X[0][2]};
    const double b[3] = {X[1][0] - X[0][0], X[1][1] - X[0][1], X[1][2] - X[0][2]};
    const double c[3] = {X[2][0] - X[0][0], X[2][1] - X[0][1], X[2][2] - X[0][2]};
    return (a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0])) / 6.;
  }
  double JxW[MAXQP], dphi[MAXNEN][MAXQP][3], v = 0.0;
  fe_reinit(T, X, JxW, dphi);
  for (int q = 0; q < T->nqp; q++) v += JxW[q];
  return v;
}

/* cond: ncond records of {w[5], div, lo, hi} (8 doubles).  vol[region] += Volume for the elements whose every node
 * satisfies every condition; serial element loop in element order, exactly like the reference. */
int orc_region_volumes(int elem_type, int nv, int64_t N, int64_t E, const int32_t* conn, const double* xyz, const double* u,
                       const int32_t* region, int n_regions, int ncond, const double* cond, double* vol) {
  fe_table T;
  if (fe_table_init(&T, elem_type)) return -1;
  (void)N;
  for (int r = 0; r < n_regions; r++) vol[r] = 0.0;
  for (int64_t e = 0; e < E; e++) {
    double X[MAXNEN][3];
    for (int l = 0; l < T.nen; l++)
      for (int d = 0; d < 3; d++) X[l][d] = xyz[(int64_t)conn[e * T.nen + l] * 3 + d];
    int consider = 1;
    for (int l = 0; l < T.nen && consider; l++) {
      const double* un = u + (int64_t)conn[e * T.nen + l] * nv;
      for (int k = 0; k < ncond && consider; k++) {
        const double* c = cond + 8 * k;
        double s = 0.0;
        int first = 1;
        for (int a = 0; a < nv; a++)
          if (c[a] != 0.0) { s = first ? c[a] * un[a] : s + c[a] * un[a]; first = 0; }
        s /= c[5];
        if (!(s >= c[6] && s <= c[7])) consider = 0;
      }
    }
    if (consider) vol[region ? region[e] : 0] += elem_volume(&T, elem_type, X);
  }
  return 0;
}

/* mean[region] = (sum_qp JxW sum_l phi_l u_l[var]) / Volume of the LAST element of that region (adpm.C:763-783) */
int orc_region_last_mean(int elem_type, int nv, int64_t N, int64_t E, const int32_t* conn, const double* xyz, const double* u,
                         const int32_t* region, int n_regions, int var, double* mean) {
  fe_table T;
  if (fe_table_init(&T, elem_type)) return -1;
  (void)N;
  for (int r = 0; r < n_regions; r++) mean[r] = 0.0;
  for (int64_t e = 0; e < E; e++) {
    double X[MAXNEN][3], JxW[MAXQP], dphi[MAXNEN][MAXQP][3];
    for (int l = 0; l < T.nen; l++)
      for (int d = 0; d < 3; d++) X[l][d] = xyz[(int64_t)conn[e * T.nen + l] * 3 + d];
    fe_reinit(&T, X, JxW, dphi);
    double avg = 0.0;
    for (int q = 0; q < T.nqp; q++) {
      double val = 0.0;
      for (int l = 0; l < T.nen; l++) val += T.phi[l][q] * u[(int64_t)conn[e * T.nen + l] * nv + var];
      avg += JxW[q] * val;
    }
    mean[region ? region[e] : 0] = avg / elem_volume(&T, elem_type, X);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------------
 * check_solution
 * ---------------------------------------------------------------------------------------------- */
/* adpm.C:665-682, pihna.C:772-797, proteas.C:719-744, coupled_hcc.C:707-724 */
void orc_clamp_nonneg(int64_t D, double* u) {
  for (int64_t k = 0; k < D; k++)
    if (u[k] < 0.0) u[k] = 0.0;
}

/* ripf.C:675-775.  u [N*3] in/out (clamped), prev [N*3] in/out (receives the UNCLAMPED u, :770),
 * aux [N*6] = {TD[3], RT_broad, RT_focus, RT_total}: TD and RT_total are rewritten.
 * Returns int(RT_total_max) (:772) or a negative number when RT_total_max <= 0 (:773 libmesh_error). */
int orc_ripf_check(int64_t N, double* u, double* prev, double* aux, const double* params, double time, double dt) {
  const double DT_R = 1.0 / dt;
  const double HU_min = params[RIPF_HU_MIN], HU_max = params[RIPF_HU_MAX];
  const double bf = params[RIPF_RT_BROAD_FRAC], ff = params[RIPF_RT_FOCUS_FRAC], tf = bf + ff;
  const int day = (int)floor(time);
  double RT_total_max = -1.0;
  for (int64_t n = 0; n < N; n++) {
    const double s0 = u[n * 3], s1 = u[n * 3 + 1], s2 = u[n * 3 + 2];
    double HU = s0, cc = s1, fb = s2;
    if (HU < HU_min) HU = HU_min; else if (HU > HU_max) HU = HU_max;
    if (cc < 0.0) cc = 0.0;
    if (fb < 0.0) fb = 0.0;
    aux[n * 6 + 0] = (HU - prev[n * 3 + 0]) * DT_R;
    aux[n * 6 + 1] = (cc - prev[n * 3 + 1]) * DT_R;
    aux[n * 6 + 2] = (fb - prev[n * 3 + 2]) * DT_R;
    const double RT_broad = aux[n * 6 + 3], RT_focus = aux[n * 6 + 4];
    double RT_total;
    if (day < bf) RT_total = RT_broad / bf * (day + 1);
    else if (day < tf) RT_total = RT_focus / ff * ((day + 1) - bf) + RT_broad;
    else RT_total = RT_broad + RT_focus;
    aux[n * 6 + 5] = RT_total;
    if (RT_total > RT_total_max) RT_total_max = RT_total;
    prev[n * 3] = s0; prev[n * 3 + 1] = s1; prev[n * 3 + 2] = s2;
    u[n * 3] = HU; u[n * 3 + 1] = cc; u[n * 3 + 2] = fb;
  }
  if (RT_total_max <= 0.0) return -1;
  return (int)RT_total_max;
}

/* ------------------------------------------------------------------------------------------------
 * [upstream] linear solve: PETSc KSPGMRES(restart) with LEFT preconditioning, classical Gram-Schmidt
 * (no refinement), convergence on the preconditioned residual ||B r|| <= max(rtol*||B b||, 1e-50),
 * non-zero initial guess.  pc: 0 = ILU(0) per block (nblocks contiguous row blocks = PCBJACOBI with
 * one block per MPI rank; nblocks = 1 is serial PCILU), 1 = point Jacobi, 2 = none.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  int pc, nblocks;
  int64_t D;
  const int64_t* rowptr;
  const int32_t* col;
  const double* val;
  double* lu;       /* ILU(0) factors on the pattern of the block-diagonal part */
  int64_t* diag;    /* position of the diagonal in each row */
  int64_t* blk_lo;  /* [nblocks+1] */
  double* dinv;
  int nthreads;
} pc_t;

static void spmv(const pc_t* P, const double* x, double* y) {
  const int64_t D = P->D;
#pragma omp parallel for num_threads(P->nthreads) schedule(static)
  for (int64_t i = 0; i < D; i++) {
    double s = 0.0;
    for (int64_t k = P->rowptr[i]; k < P->rowptr[i + 1]; k++) s += P->val[k] * x[P->col[k]];
    y[i] = s;
  }
}

static int pc_setup(pc_t* P) {
  const int64_t D = P->D;
  P->diag = (int64_t*)malloc(sizeof(int64_t) * (size_t)D);
  for (int64_t i = 0; i < D; i++) {
    P->diag[i] = csr_find(P->rowptr, P->col, i, (int32_t)i);
    if (P->diag[i] < 0) return -1;
  }
  if (P->pc == 1) {
    P->dinv = (double*)malloc(sizeof(double) * (size_t)D);
    for (int64_t i = 0; i < D; i++) P->dinv[i] = 1.0 / P->val[P->diag[i]];
    return 0;
  }
  if (P->pc != 0) return 0;
  P->blk_lo = (int64_t*)malloc(sizeof(int64_t) * (size_t)(P->nblocks + 1));
  for (int b = 0; b <= P->nblocks; b++) P->blk_lo[b] = D * b / P->nblocks;
  P->lu = (double*)malloc(sizeof(double) * (size_t)P->rowptr[D]);
  memcpy(P->lu, P->val, sizeof(double) * (size_t)P->rowptr[D]);
  int bad = 0;
  /* IKJ ILU(0) restricted to each diagonal block (entries outside the block are ignored) */
#pragma omp parallel for num_threads(P->nthreads) schedule(static, 1)
  for (int b = 0; b < P->nblocks; b++) {
    const int64_t lo = P->blk_lo[b], hi = P->blk_lo[b + 1];
    for (int64_t i = lo; i < hi; i++) {
      for (int64_t kk = P->rowptr[i]; kk < P->diag[i]; kk++) {
        const int64_t k = P->col[kk];
        if (k < lo) continue;
        const double piv = P->lu[P->diag[k]];
        if (piv == 0.0) { bad = 1; continue; }
        const double f = P->lu[kk] / piv;
        P->lu[kk] = f;
        /* row_i[j] -= f * row_k[j] for j > k on the pattern of row i */
        int64_t pj = P->diag[k] + 1;
        for (int64_t jj = kk + 1; jj < P->rowptr[i + 1]; jj++) {
          const int32_t j = P->col[jj];
          if (j >= hi) break;
          while (pj < P->rowptr[k + 1] && P->col[pj] < j) pj++;
          if (pj < P->rowptr[k + 1] && P->col[pj] == j) P->lu[jj] -= f * P->lu[pj];
        }
      }
    }
  }
  return bad ? -2 : 0;
}

static void pc_apply(const pc_t* P, const double* r, double* z) {
  const int64_t D = P->D;
  if (P->pc == 2) { memcpy(z, r, sizeof(double) * (size_t)D); return; }
  if (P->pc == 1) {
#pragma omp parallel for num_threads(P->nthreads) schedule(static)
    for (int64_t i = 0; i < D; i++) z[i] = P->dinv[i] * r[i];
    return;
  }
#pragma omp parallel for num_threads(P->nthreads) schedule(static, 1)
  for (int b = 0; b < P->nblocks; b++) {
    const int64_t lo = P->blk_lo[b], hi = P->blk_lo[b + 1];
    for (int64_t i = lo; i < hi; i++) { /* L y = r (unit lower) */
      double s = r[i];
      for (int64_t kk = P->rowptr[i]; kk < P->diag[i]; kk++) {
        const int64_t k = P->col[kk];
        if (k >= lo) s -= P->lu[kk] * z[k];
      }
      z[i] = s;
    }
    for (int64_t i = hi - 1; i >= lo; i--) { /* U z = y */
      double s = z[i];
      for (int64_t kk = P->diag[i] + 1; kk < P->rowptr[i + 1]; kk++) {
        const int64_t k = P->col[kk];
        if (k < hi) s -= P->lu[kk] * z[k];
      }
      z[i] = s / P->lu[P->diag[i]];
    }
  }
}

static void pc_free(pc_t* P) { free(P->lu); free(P->diag); free(P->blk_lo); free(P->dinv); }

static double vdot(int64_t D, const double* a, const double* b, int nt) {
  double s = 0.0;
#pragma omp parallel for num_threads(nt) reduction(+ : s) schedule(static)
  for (int64_t i = 0; i < D; i++) s += a[i] * b[i];
  return s;
}

/* returns 0 converged, 1 maxits reached, <0 error.  x: in = initial guess, out = solution */
int orc_gmres(int64_t D, const int64_t* rowptr, const int32_t* col, const double* val, const double* b, double* x,
              int pc, int nblocks, int restart, double rtol, int maxits, int nthreads, int* its_out, double* res_out,
              double* res0_out) {
  pc_t P;
  memset(&P, 0, sizeof(P));
  P.pc = pc; P.nblocks = nblocks < 1 ? 1 : nblocks; P.D = D; P.rowptr = rowptr; P.col = col; P.val = val;
  P.nthreads = nthreads < 1 ? 1 : nthreads;
  int rc = pc_setup(&P);
  if (rc) { pc_free(&P); return -10 + rc; }
  const int m = restart;
  double* V = (double*)malloc(sizeof(double) * (size_t)D * (size_t)(m + 1));
  double* w = (double*)malloc(sizeof(double) * (size_t)D);
  double* t = (double*)malloc(sizeof(double) * (size_t)D);
  double* H = (double*)calloc((size_t)(m + 1) * (size_t)m, sizeof(double));
  double* cs = (double*)calloc((size_t)m, sizeof(double));
  double* sn = (double*)calloc((size_t)m, sizeof(double));
  double* g = (double*)calloc((size_t)m + 1, sizeof(double));
  double* y = (double*)calloc((size_t)m, sizeof(double));
  const int nt = P.nthreads;
  /* reference norm: ||B b|| */
  pc_apply(&P, b, w);
  const double bnorm = sqrt(vdot(D, w, w, nt));
  const double target = fmax(rtol * bnorm, 1e-50);
  int its = 0, status = 1;
  double res = 0.0;
  for (;;) {
    spmv(&P, x, t);
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t i = 0; i < D; i++) t[i] = b[i] - t[i];
    pc_apply(&P, t, V); /* v0 = B (b - A x) */
    double beta = sqrt(vdot(D, V, V, nt));
    res = beta;
    if (beta <= target) { status = 0; break; }
    if (its >= maxits) { status = 1; break; }
#pragma omp parallel for num_threads(nt) schedule(static)
    for (int64_t i = 0; i < D; i++) V[i] /= beta;
    memset(g, 0, sizeof(double) * (size_t)(m + 1));
    g[0] = beta;
    int j = 0, done = 0;
    for (; j < m && its < maxits; j++) {
      double* vj = V + (size_t)D * (size_t)j;
      double* vn = V + (size_t)D * (size_t)(j + 1);
      spmv(&P, vj, t);
      pc_apply(&P, t, vn);
      /* classical Gram-Schmidt: all dots against the unmodified vn, then one update */
      for (int k = 0; k <= j; k++) H[k * m + j] = vdot(D, V + (size_t)D * (size_t)k, vn, nt);
      for (int k = 0; k <= j; k++) {
        const double h = H[k * m + j];
        const double* vk = V + (size_t)D * (size_t)k;
#pragma omp parallel for num_threads(nt) schedule(static)
        for (int64_t i = 0; i < D; i++) vn[i] -= h * vk[i];
      }
      const double hn = sqrt(vdot(D, vn, vn, nt));
      H[(j + 1) * m + j] = hn;
      if (hn != 0.0) {
#pragma omp parallel for num_threads(nt) schedule(static)
        for (int64_t i = 0; i < D; i++) vn[i] /= hn;
      }
      for (int k = 0; k < j; k++) { /* previous rotations */
        const double a = H[k * m + j], c2 = H[(k + 1) * m + j];
        H[k * m + j] = cs[k] * a + sn[k] * c2;
        H[(k + 1) * m + j] = -sn[k] * a + cs[k] * c2;
      }
      {
        const double a = H[j * m + j], c2 = H[(j + 1) * m + j];
        const double r = hypot(a, c2);
        cs[j] = r == 0.0 ? 1.0 : a / r;
        sn[j] = r == 0.0 ? 0.0 : c2 / r;
        H[j * m + j] = r;
        H[(j + 1) * m + j] = 0.0;
        g[j + 1] = -sn[j] * g[j];
        g[j] = cs[j] * g[j];
      }
      its++;
      res = fabs(g[j + 1]);
      if (res <= target || hn == 0.0) { j++; done = 1; break; }
    }
    /* back substitution and update x += V y */
    for (int k = j - 1; k >= 0; k--) {
      double s = g[k];
      for (int l = k + 1; l < j; l++) s -= H[k * m + l] * y[l];
      y[k] = s / H[k * m + k];
    }
    for (int k = 0; k < j; k++) {
      const double yk = y[k];
      const double* vk = V + (size_t)D * (size_t)k;
#pragma omp parallel for num_threads(nt) schedule(static)
      for (int64_t i = 0; i < D; i++) x[i] += yk * vk[i];
    }
    if (done) { status = 0; break; }
    if (its >= maxits) { status = 1; break; }
  }
  if (its_out) *its_out = its;
  if (res_out) *res_out = res;
  if (res0_out) *res0_out = bnorm;
  free(V); free(w); free(t); free(H); free(cs); free(sn); free(g); free(y);
  pc_free(&P);
  return status;
}

/* y = A x (parity helper) */
void orc_spmv(int64_t D, const int64_t* rowptr, const int32_t* col, const double* val, const double* x, double* y,
              int nthreads) {
  pc_t P;
  memset(&P, 0, sizeof(P));
  P.D = D; P.rowptr = rowptr; P.col = col; P.val = val; P.nthreads = nthreads < 1 ? 1 : nthreads;
  spmv(&P, x, y);
}

/* ------------------------------------------------------------------------------------------------
 * one time step, adpm.C:63-76: (time already advanced by the caller) rotate -> zero+assemble -> KSP ->
 * check_solution.  u: in = current solution, out = new clamped solution; u_old receives the rotated
 * copy.  state (RIPF only): prev [N*3] and aux [N*6]; rt_max in/out.
 * ---------------------------------------------------------------------------------------------- */
int orc_step(int model, int elem_type, int64_t N, int64_t E, const int32_t* conn, const double* xyz, double* u,
             double* u_old, const double* params, const double* elem_field, double* nodal_aux, double* ripf_prev,
             int* ripf_rt_total_max, double time, double dt, const int64_t* rowptr, const int32_t* col, double* val,
             double* rhs, int pc, int nblocks, int restart, double rtol, int maxits, int nthreads, int* its,
             double* res, double* t_assemble, double* t_solve) {
  const int nv = orc_model_nvars(model);
  const int64_t D = N * nv;
  memcpy(u_old, u, sizeof(double) * (size_t)D);
#ifdef _OPENMP
  double t0 = omp_get_wtime();
#endif
  int rc = orc_assemble(model, elem_type, N, E, conn, xyz, u_old, params, elem_field, nodal_aux,
                        ripf_rt_total_max ? *ripf_rt_total_max : 0, time, dt, rowptr, col, val, rhs, nthreads);
  if (rc) return rc;
#ifdef _OPENMP
  double t1 = omp_get_wtime();
#endif
  rc = orc_gmres(D, rowptr, col, val, rhs, u, pc, nblocks, restart, rtol, maxits, nthreads, its, res, NULL);
#ifdef _OPENMP
  double t2 = omp_get_wtime();
  if (t_assemble) *t_assemble = t1 - t0;
  if (t_solve) *t_solve = t2 - t1;
#endif
  if (rc < 0) return rc;
  if (model == RDC_RIPF) {
    const int m = orc_ripf_check(N, u, ripf_prev, nodal_aux, params, time, dt);
    if (m < 0) return -20;
    *ripf_rt_total_max = m;
  } else {
    orc_clamp_nonneg(D, u);
  }
  return rc;
}

/* the solid-mechanics Newton path uses the FE tables, pattern builder and GMRES above */
#include "solid_oracle.c"
