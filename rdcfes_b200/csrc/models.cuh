// models.cuh -- per-quadrature-point coefficient tables of the five RDC models (device code).
//
// Every reference callback has the separable form
//   Fe_a(i)    += JxW [ (U_a + dt/2 f_a) phi_i + dt/2 (q_a . grad phi_i) ]
//   Ke_ab(i,j) += JxW [ (cap_ab - dt/2 r_ab) phi_j phi_i - dt/2 s_ab (grad phi_j . grad phi_i)
//                       - dt/2 phi_j (t_ab . grad phi_i) ]
// with f, q, r, s, t evaluated from the OLD state at the quadrature point.  A model here supplies, for
// one quadrature point and one test function i, the scalars
//   F0[a]   = U_a + dt/2 f_a                 (multiplies JxW phi_i)
//   F1[a]   = dt/2 (q_a . grad phi_i)        (multiplies JxW)
//   C[a][b] = cap_ab - dt/2 r_ab             (multiplies JxW phi_j phi_i)
//   S[a][b] = -dt/2 s_ab                     (multiplies JxW grad phi_j . grad phi_i)
//   T[a][b] = -dt/2 (t_ab . grad phi_i)      (multiplies JxW phi_j)
// Only entries whose bit is set in CMASK / SMASK / TMASK are produced and consumed (bit a*NV+b), so the
// structurally empty blocks cost nothing and are written as explicit zeros.
// The vectors q and t are combinations of a few per-model DIRECTION vectors (field gradients, tract
// vectors); the caller hands in their dot products with grad phi_i (Dg[]).
//
// Reference lines: ADPM adpm.C:460-593; PIHNA pihna.C:427-750; RIPF ripf.C:449-665;
// PROTEAS proteas.C:454-698; HCC coupled_hcc.C:496-640; rate laws utils.h:69-187.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace rdc {

// bit helpers for the block masks
__host__ __device__ constexpr unsigned bit(int nv, int a, int b) { return 1u << (a * nv + b); }

// Rounding discipline.  The discrete decisions of the models (Pi_/SD_/Tr_ branches, taxis alignment,
// capacity switches, epsilon switches) depend on the interpolated state, on field gradients and on the
// element geometry.  Those quantities are computed with mul_rn/add_rn (no FMA contraction) in exactly the
// operation order of the reference build (x86-64 baseline has no FMA), so every comparison sees
// bit-identical operands.  Everything downstream of the decisions (coefficient products, quadrature
// accumulation) may be contracted by the compiler: it changes values by O(1e-16) relative, never a branch.
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
// a0*b0 + a1*b1 + a2*b2, left to right, each product and sum rounded (TypeVector::operator* / norm_sq)
__device__ __forceinline__ double dot3_rn(const double* a, const double* b) {
  return add_rn(add_rn(mul_rn(a[0], b[0]), mul_rn(a[1], b[1])), mul_rn(a[2], b[2]));
}

// ------------------------------------------------------------------------------------ rate laws
// utils.h:100-110.  "cM <= 0 disables the law" (Appendix C-7) is applied on the host: the parameter structs
// carry cM = 0 and zero slopes for disabled laws, so the device code needs no cM test.
__device__ __forceinline__ double law_pulse(double C, double cM, double c0, double c1) {
  return (C >= c0 && C < c1) ? cM : 0.0;
}
struct StepDecay { double cM, c0, c1, slope; };  // slope = cM/(c1-c0) precomputed on the host
__device__ __forceinline__ void law_stepdecay(double C, const StepDecay& p, double& v, double& dv) {  // utils.h:112-133
  v = 0.0; dv = 0.0;
  if (C < p.c0) v = p.cM;
  else if (C < p.c1) { v = (p.c1 - C) * p.slope; dv = -p.slope; }
}
struct Trapezoid { double cM, c0, c1, c2, c3, up, dn; };  // up = cM/(c1-c0), dn = cM/(c3-c2)
__device__ __forceinline__ void law_trapezoid(double C, const Trapezoid& p, double& v, double& dv) {  // utils.h:158-187
  v = 0.0; dv = 0.0;
  if (C < p.c0) {}
  else if (C < p.c1) { v = (C - p.c0) * p.up; dv = p.up; }
  else if (C < p.c2) { v = p.cM; }
  else if (C < p.c3) { v = (p.c3 - C) * p.dn; dv = -p.dn; }
}
struct Pulse { double cM, c0, c1; };

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// x^e and x^(e-1) for the capacity laws Tau = (1 - T/kappa)^e (pihna.C:444-472, coupled_hcc.C:510-532).
// The shipped configurations use e = 3 (run/PIHNA/input.dat:26, run/Coupled/HCC/input.dat:30); small integer exponents
// are formed by multiplication (within 2 ulp of pow, against a 1e-12 tolerance), anything else by pow -- the branch
// is uniform over the grid.  The generic pow costs more than the rest of the quadrature point.
__device__ __forceinline__ void pow_pair(double x, double e, double& xe, double& xe1) {
  if (e == 1.0) { xe1 = 1.0; xe = x; }
  else if (e == 2.0) { xe1 = x; xe = x * x; }
  else if (e == 3.0) { xe1 = x * x; xe = xe1 * x; }
  else if (e == 4.0) { const double x2 = x * x; xe1 = x2 * x; xe = x2 * x2; }
  else { xe = pow(x, e); xe1 = pow(x, e - 1.0); }
}

template <int NV>
struct Coef {
  double F0[NV], F1[NV];
  double C[NV][NV], S[NV][NV], T[NV][NV];
};

// =========================================================================================== ADPM
struct AdpmParams {
  double dt2;
  Pulse decay_PrP;  // cM already multiplied by pow(time, gamma) on the host (adpm.C:369)
  Pulse diffuse_A, taxis1_A, taxis2_A, decay_A, diffuse_T, taxis1_T, taxis2_T, decay_T;
  StepDecay produce_A, produce_T;
  Trapezoid transform_A, transform_T;
  double omega_A, omega_T;  // cos(angle), adpm.C:413-414
};

struct Adpm {
  static constexpr int NV = 3;
  static constexpr int NDIR = 4;           // gA, gT, tract_A, tract_T
  static constexpr unsigned GRADMASK = 0b110;  // gradients of A_b and Tau
  static constexpr unsigned CMASK = bit(3, 0, 0) | bit(3, 0, 1) | bit(3, 0, 2) | bit(3, 1, 0) | bit(3, 1, 1) | bit(3, 2, 0) | bit(3, 2, 2);
  static constexpr unsigned SMASK = bit(3, 1, 1) | bit(3, 2, 2);
  static constexpr unsigned TMASK = bit(3, 1, 1) | bit(3, 2, 2);
  static constexpr int N_EFIELD = 3;       // tract vector per element
  static constexpr int N_NAUX = 0;
  static constexpr unsigned AUXGRADMASK = 0;
  typedef AdpmParams Params;

  // direction vectors from the field gradients G[var][3], element field ef[3] (adpm.C:473-492)
  __device__ static __forceinline__ void directions(const Params& p, const double (*G)[3], const double* ef,
                                                    const double (*GA)[3], double (*dir)[3]) {
    (void)GA;
    for (int d = 0; d < 3; d++) { dir[0][d] = G[1][d]; dir[1][d] = G[2][d]; dir[2][d] = 0.0; dir[3][d] = 0.0; }
    align(G[1], ef, p.omega_A, dir[2]);
    align(G[2], ef, p.omega_T, dir[3]);
  }
  // adpm.C:473-492: tract direction +ef / -ef / 0 from d = (g/|g|) . ef against +-omega.  The reference's expression
  // (unit vector by three divisions, ordered dot product) costs a square root and three divisions; d is first
  // estimated as (g . ef) / |g| from one reciprocal square root, and only when that estimate is within 1e-11 of a
  // threshold (relative to |ef| + omega: thousands of times the rounding error of either expression) is the
  // reference's own sequence evaluated, so the decision is always the one the reference takes.
  __device__ static __forceinline__ void align(const double* g, const double* ef, double omega, double* out) {
    out[0] = out[1] = out[2] = 0.0;
    const double gg = dot3_rn(g, g);
    if (gg == 0.0) return;                       // |g| != 0 test of the reference (sqrt(x) == 0 only for x == 0)
    const double est = dot3(g, ef) * rsqrt(gg);
    const double margin = 1e-11 * (fabs(ef[0]) + fabs(ef[1]) + fabs(ef[2]) + fabs(omega));
    double d = est;
    if (!(fabs(fabs(est) - fabs(omega)) > margin)) {   // too close to call (or NaN): the reference's sequence
      const double n = sqrt(gg);
      const double u[3] = {g[0] / n, g[1] / n, g[2] / n};
      d = dot3_rn(u, ef);
    }
    if (d > +omega) { out[0] = ef[0]; out[1] = ef[1]; out[2] = ef[2]; }
    else if (d < -omega) { out[0] = -ef[0]; out[1] = -ef[1]; out[2] = -ef[2]; }
  }

  // U[var] at the qp, A[] nodal aux at the qp (unused), Dg[dir] = dir . grad phi_i
  __device__ static __forceinline__ void coef(const Params& p, const double* U, const double* A, const double* Dg,
                                              Coef<3>& k) {
    (void)A;
    const double P = U[0], Ab = U[1], Ta = U[2];
    double TrA, dTrA, TrT, dTrT, SA, dSA, ST, dST;
    law_trapezoid(Ab, p.transform_A, TrA, dTrA);
    law_trapezoid(Ta, p.transform_T, TrT, dTrT);
    law_stepdecay(Ab, p.produce_A, SA, dSA);
    law_stepdecay(Ta, p.produce_T, ST, dST);
    const double P0 = law_pulse(P, p.decay_PrP.cM, p.decay_PrP.c0, p.decay_PrP.c1);
    const double PA = law_pulse(Ab, p.decay_A.cM, p.decay_A.c0, p.decay_A.c1);
    const double DA = law_pulse(Ab, p.diffuse_A.cM, p.diffuse_A.c0, p.diffuse_A.c1);
    const double X1A = law_pulse(Ab, p.taxis1_A.cM, p.taxis1_A.c0, p.taxis1_A.c1);
    const double X2A = law_pulse(Ta, p.taxis2_A.cM, p.taxis2_A.c0, p.taxis2_A.c1);
    const double PT = law_pulse(Ta, p.decay_T.cM, p.decay_T.c0, p.decay_T.c1);
    const double DT = law_pulse(Ta, p.diffuse_T.cM, p.diffuse_T.c0, p.diffuse_T.c1);
    const double X1T = law_pulse(Ta, p.taxis1_T.cM, p.taxis1_T.c0, p.taxis1_T.c1);
    const double X2T = law_pulse(Ab, p.taxis2_T.cM, p.taxis2_T.c0, p.taxis2_T.c1);
    const double h = p.dt2;
    const double loss = TrA + TrT + P0;
    // load vector, adpm.C:497-530
    k.F0[0] = P - h * (loss * P);
    k.F0[1] = Ab + h * (SA * Ab + TrA * P - PA * Ab);
    k.F0[2] = Ta + h * (ST * Ta + TrT * P - PT * Ta);
    k.F1[0] = 0.0;
    k.F1[1] = h * (-DA * Dg[0] - X1A * Ab * Dg[2] + X2A * Ab * Dg[3]);
    k.F1[2] = h * (-DT * Dg[1] - X1T * Ta * Dg[3] + X2T * Ta * Dg[2]);
    // matrix, adpm.C:535-590
    k.C[0][0] = 1.0 + h * loss;
    k.C[0][1] = h * (dTrA * P);
    k.C[0][2] = h * (dTrT * P);
    k.C[1][0] = -h * TrA;
    k.C[1][1] = 1.0 - h * (SA + dSA * Ab + dTrA * P - PA);
    k.C[2][0] = -h * TrT;
    k.C[2][2] = 1.0 - h * (ST + dST * Ta + dTrT * P - PT);
    k.S[1][1] = h * DA;
    k.S[2][2] = h * DT;
    k.T[1][1] = h * (X1A * Dg[2] - X2A * Dg[3]);
    k.T[2][2] = h * (X1T * Dg[3] - X2T * Dg[2]);
  }
};

// ========================================================================================== PIHNA
struct PihnaParams {
  double dt2;
  double Lambda_k, Kappa_k, Kappa_a, ek;
  double nec_c, nec_h, nec_v;  // already / Kappa_k (pihna.C:364-366)
  double dif_c, tax_c, dif_h, tax_h, prod_c, c2h, h2c, h2n, dif_v, tax_v, prod_v, sec_c, sec_h, upt_v, dec_a;
};

struct Pihna {
  static constexpr int NV = 5;
  static constexpr int NDIR = 4;  // grad c, grad h, grad v, grad a
  static constexpr unsigned GRADMASK = 0b11110;
  static constexpr unsigned CMASK =
      bit(5, 0, 0) | bit(5, 0, 1) | bit(5, 0, 2) | bit(5, 0, 3) | bit(5, 1, 0) | bit(5, 1, 1) | bit(5, 1, 2) | bit(5, 1, 3) |
      bit(5, 2, 0) | bit(5, 2, 1) | bit(5, 2, 2) | bit(5, 2, 3) | bit(5, 3, 0) | bit(5, 3, 1) | bit(5, 3, 2) | bit(5, 3, 3) |
      bit(5, 3, 4) | bit(5, 4, 1) | bit(5, 4, 2) | bit(5, 4, 3) | bit(5, 4, 4);
  static constexpr unsigned SMASK = bit(5, 1, 1) | bit(5, 1, 3) | bit(5, 2, 2) | bit(5, 2, 3) | bit(5, 3, 3) | bit(5, 3, 4);
  static constexpr unsigned TMASK = bit(5, 1, 0) | bit(5, 1, 1) | bit(5, 1, 2) | bit(5, 1, 3) | bit(5, 2, 0) | bit(5, 2, 1) |
                                    bit(5, 2, 2) | bit(5, 2, 3) | bit(5, 3, 0) | bit(5, 3, 1) | bit(5, 3, 2) | bit(5, 3, 3);
  static constexpr int N_EFIELD = 0;
  static constexpr int N_NAUX = 0;
  static constexpr unsigned AUXGRADMASK = 0;
  typedef PihnaParams Params;

  __device__ static __forceinline__ void directions(const Params&, const double (*G)[3], const double*,
                                                    const double (*GA)[3], double (*dir)[3]) {
    (void)GA;
    for (int d = 0; d < 3; d++) { dir[0][d] = G[1][d]; dir[1][d] = G[2][d]; dir[2][d] = G[3][d]; dir[3][d] = G[4][d]; }
  }

  __device__ static __forceinline__ void coef(const Params& p, const double* U, const double* A, const double* Dg,
                                              Coef<5>& k) {
    (void)A;
    const double n = U[0], c = U[1], hh = U[2], v = U[3], a = U[4];
    const double Gc = Dg[0], Gh = Dg[1], Gv = Dg[2], Ga = Dg[3];
    // pihna.C:444-472
    double Tau, dT;
    {
      const double Te = (n + c + hh + v) / p.Kappa_k;
      if (Te <= 0.0) { Tau = 1.0; dT = 0.0; }
      else if (Te >= 1.0) { Tau = 0.0; dT = 0.0; }
      else { double xe1; pow_pair(1.0 - Te, p.ek, Tau, xe1); dT = (-p.ek / p.Kappa_k) * xe1; }
    }
    // pihna.C:474-499 (0/0 = NaN fails both comparisons like on the host)
    double Ve, dVc, dVv;
    {
      const double chv = c + hh + v;
      const double Ve_ = v / chv;
      if (Ve_ <= 0.0) { Ve = 0.0; dVc = 0.0; dVv = 0.0; }
      else if (Ve_ >= 1.0) { Ve = 1.0; dVc = 0.0; dVv = 0.0; }
      else { const double rc = 1.0 / chv; Ve = Ve_; dVc = -Ve_ * rc; dVv = (1.0 - Ve_) * rc; }   // one reciprocal for both derivatives
    }
    const double dVh = dVc;
    const double ra = 1.0 / (a + p.Kappa_a);
    const double Ua = a * ra, dUa = ra - Ua * ra;  // pihna.C:501-502 (a/(a+K), 1/(a+K) - Ua/(a+K)) from one reciprocal
    const double dif_c = c > p.Lambda_k ? p.dif_c : 0.0, tax_c = c > p.Lambda_k ? p.tax_c : 0.0;  // pihna.C:504-509
    const double dif_h = hh > p.Lambda_k ? p.dif_h : 0.0, tax_h = hh > p.Lambda_k ? p.tax_h : 0.0;
    const double dif_v = v > p.Lambda_k ? p.dif_v : 0.0, tax_v = v > p.Lambda_k ? p.tax_v : 0.0;
    const double h = p.dt2;
    const double oneV = 1.0 - Ve;
    // load vector, pihna.C:514-566
    k.F0[0] = n + h * (p.nec_c * c * n + p.nec_h * hh * n + p.nec_v * v * n + p.h2n * oneV * hh);
    k.F0[1] = c + h * (p.prod_c * Tau * c - p.c2h * oneV * c + p.h2c * Ve * hh - p.nec_c * c * n);
    k.F0[2] = hh + h * (p.c2h * oneV * c - p.h2c * Ve * hh - p.nec_h * hh * n - p.h2n * oneV * hh);
    k.F0[3] = v + h * (p.prod_v * Tau * Ua * v - p.nec_v * v * n);
    k.F0[4] = a + h * (p.sec_c * c + p.sec_h * hh - p.upt_v * v * a - p.dec_a * a);
    k.F1[0] = 0.0;
    k.F1[1] = h * (-dif_c * Tau * Gc - tax_c * Tau * c * Gv);
    k.F1[2] = h * (-dif_h * Tau * Gh - tax_h * Tau * hh * Gv);
    k.F1[3] = h * (-dif_v * Tau * Gv - tax_v * Tau * v * Ga);
    k.F1[4] = 0.0;
    // row n, pihna.C:571-597
    k.C[0][0] = 1.0 - h * (p.nec_c * c + p.nec_h * hh + p.nec_v * v);
    k.C[0][1] = -h * (p.nec_c * n - p.h2n * dVc * hh);
    k.C[0][2] = -h * (p.nec_h * n - p.h2n * dVh * hh + p.h2n * oneV);
    k.C[0][3] = -h * (p.nec_v * n - p.h2n * dVv * hh);
    // row c, pihna.C:599-641.  All Tau derivatives are equal (dT), so the flux derivative is shared.
    const double fc = -dif_c * Gc - tax_c * c * Gv;  // d(q_c . grad phi_i)/dTau
    k.C[1][0] = -h * (p.prod_c * dT * c - p.nec_c * c);
    k.C[1][1] = 1.0 - h * (p.prod_c * Tau + p.prod_c * dT * c - p.c2h * oneV + p.c2h * dVc * c + p.h2c * dVc * hh - p.nec_c * n);
    k.C[1][2] = -h * (p.prod_c * dT * c + p.c2h * dVh * c + p.h2c * dVh * hh + p.h2c * Ve);
    k.C[1][3] = -h * (p.prod_c * dT * c + p.c2h * dVv * c + p.h2c * dVv * hh);
    k.T[1][0] = -h * (dT * fc);
    k.T[1][1] = -h * (dT * fc - tax_c * Tau * Gv);
    k.T[1][2] = -h * (dT * fc);
    k.T[1][3] = -h * (dT * fc);
    k.S[1][1] = h * (dif_c * Tau);
    k.S[1][3] = h * (tax_c * Tau * c);
    // row h, pihna.C:643-684
    const double fh = -dif_h * Gh - tax_h * hh * Gv;
    k.C[2][0] = -h * (-p.nec_h * hh);
    k.C[2][1] = -h * (p.c2h * oneV - p.c2h * dVc * c - p.h2c * dVc * hh + p.h2n * dVc * hh);
    k.C[2][2] = 1.0 - h * (-p.c2h * dVh * c - p.h2c * dVh * hh - p.h2c * Ve - p.nec_h * n + p.h2n * dVh * hh - p.h2n * oneV);
    k.C[2][3] = -h * (-p.c2h * dVv * c - p.h2c * dVv * hh + p.h2n * dVv * hh);
    k.T[2][0] = -h * (dT * fh);
    k.T[2][1] = -h * (dT * fh);
    k.T[2][2] = -h * (dT * fh - tax_h * Tau * Gv);
    k.T[2][3] = -h * (dT * fh);
    k.S[2][2] = h * (dif_h * Tau);
    k.S[2][3] = h * (tax_h * Tau * hh);
    // row v, pihna.C:686-724
    const double fv = -dif_v * Gv - tax_v * v * Ga;
    const double pv = p.prod_v * dT * Ua * v;
    k.C[3][0] = -h * (pv - p.nec_v * v);
    k.C[3][1] = -h * pv;
    k.C[3][2] = -h * pv;
    k.C[3][3] = 1.0 - h * (pv - p.nec_v * n);
    k.C[3][4] = -h * (p.prod_v * Tau * dUa * v);
    k.T[3][0] = -h * (dT * fv);
    k.T[3][1] = -h * (dT * fv);
    k.T[3][2] = -h * (dT * fv);
    k.T[3][3] = -h * (dT * fv - tax_v * Tau * Ga);
    k.S[3][3] = h * (dif_v * Tau);
    k.S[3][4] = h * (tax_v * Tau * v);
    // row a, pihna.C:726-747
    k.C[4][1] = -h * p.sec_c;
    k.C[4][2] = -h * p.sec_h;
    k.C[4][3] = h * (p.upt_v * a);
    k.C[4][4] = 1.0 + h * (p.upt_v * v + p.dec_a);
  }
};

// =========================================================================================== RIPF
struct RipfParams {
  double dt2;
  double VF_s, VF_p, VF_e, VF_min;
  double phi_cc_B, phi_cc_D, phi_cc, phi_fb_B, phi_fb_D, phi_fb, phi_tol;
  double kappa, kappa_RT_c, delta, delta_RT_a, delta_RT_b;
  double lambda, lambda_RT_r, lambda_HU_r, omicro, omicro_RT_r, omicro_fb_b, omega, diffusion, haptotaxis, radiotaxis;
};

struct Ripf {
  static constexpr int NV = 3;
  static constexpr int NDIR = 3;  // grad fb, grad HU, unit grad RT_total
  static constexpr unsigned GRADMASK = 0b101;
  static constexpr unsigned CMASK = bit(3, 0, 0) | bit(3, 0, 1) | bit(3, 0, 2) | bit(3, 1, 1) | bit(3, 1, 2) | bit(3, 2, 0) | bit(3, 2, 1) | bit(3, 2, 2);
  static constexpr unsigned SMASK = bit(3, 2, 0) | bit(3, 2, 2);
  static constexpr unsigned TMASK = bit(3, 2, 1) | bit(3, 2, 2);
  static constexpr int N_EFIELD = 0;
  static constexpr int N_NAUX = 3;  // TD(cc), TD(fb), RT_total ; gradient needed for aux 2
  static constexpr unsigned AUXGRADMASK = 0b100;
  typedef RipfParams Params;

  __device__ static __forceinline__ void directions(const Params&, const double (*G)[3], const double*,
                                                    const double (*GA)[3], double (*dir)[3]) {
    for (int d = 0; d < 3; d++) { dir[0][d] = G[2][d]; dir[1][d] = G[0][d]; }
    // ripf.C:481-484 normalised dose gradient
    const double l2 = sqrt(dot3_rn(GA[2], GA[2]));
    if (l2 != 0.0) { dir[2][0] = GA[2][0] / l2; dir[2][1] = GA[2][1] / l2; dir[2][2] = GA[2][2] / l2; }
    else { dir[2][0] = dir[2][1] = dir[2][2] = 0.0; }
  }

  __device__ static __forceinline__ void coef(const Params& p, const double* U, const double* A, const double* Dg,
                                              Coef<3>& k) {
    const double HU = U[0], cc = U[1], fb = U[2];
    const double cc_dt = A[0], fb_dt = A[1], RT = A[2];
    const double Gfb = Dg[0], GHU = Dg[1], GRT = Dg[2];
    // ripf.C:486-489
    const double kappa_RT = p.kappa * exp(-p.kappa_RT_c * RT);
    const double delta_RT = p.delta * (1.0 - exp(-p.delta_RT_a * RT - p.delta_RT_b * (RT * RT)));
    const double lambda_RT = p.lambda * (RT / p.lambda_RT_r);
    const double rr = RT / p.omicro_RT_r;
    const double bump = 4.0 * (rr - rr * rr);
    const double omicro_RT = p.omicro * (bump < 0.0 ? 0.0 : bump);
    // ripf.C:491-496
    double eps_cc = 0.0, eps_fb = 0.0;
    if (cc_dt > p.phi_tol) eps_cc = p.phi_cc_B; else if (cc_dt < -p.phi_tol) eps_cc = p.phi_cc_D;
    if (fb_dt > p.phi_tol) eps_fb = p.phi_fb_B; else if (fb_dt < -p.phi_tol) eps_fb = p.phi_fb_D;
    // ripf.C:498-514
    const double VF_total = p.VF_s + p.VF_p + (cc + fb);
    double Tau = 0.0, dT = 0.0;
    if (VF_total < 1.0) {
      Tau = pow(1.0 - VF_total, p.VF_e);   // feeds the comparison below: kept as the reference computes it
      dT = -p.VF_e * pow(1.0 - VF_total, p.VF_e - 1.0);
      if (Tau < p.VF_min) { Tau = 0.0; dT = 0.0; }
    }
    // ripf.C:516-523
    double Koppa = 0.0, dKoppa = 0.0;
    if (cc >= 0.0 && cc < 1.0) { Koppa = 4.0 * (cc - cc * cc); dKoppa = 4.0 - 8.0 * cc; }
    // ripf.C:525-561
    double Lom = 0.0, dLom_HU = 0.0, dLom_fb = 0.0, Ome = 0.0, dOme_fb = 0.0;
    if (fb >= 0.0 && fb < 1.0) {
      if (HU > p.lambda_HU_r && HU < 0.0) {
        Lom = (1.0 - fb * fb) * (HU / p.lambda_HU_r);
        dLom_HU = (1.0 - fb * fb) / p.lambda_HU_r;
        dLom_fb = -(2.0 * fb) * (HU / p.lambda_HU_r);
      } else if (HU < p.lambda_HU_r) {
        Lom = 1.0 - fb * fb;
        dLom_fb = -(2.0 * fb);
      }
      if (fb <= p.omicro_fb_b) { Ome = 4.0 * (p.omicro_fb_b - p.omicro_fb_b * p.omicro_fb_b); }
      else { Ome = 4.0 * (fb - fb * fb); dOme_fb = 4.0 - 8.0 * fb; }
    }
    const double h = p.dt2;
    // flux of fb: q = -D Tau grad fb - hapto Tau fb grad HU - radio Tau fb gradRT  (ripf.C:590-592)
    const double flux = p.diffusion * Gfb + p.haptotaxis * (GHU * fb) + p.radiotaxis * (GRT * fb);
    k.F0[0] = HU + h * (eps_cc * cc + eps_fb * fb + p.phi_cc * cc_dt + p.phi_fb * fb_dt);  // ripf.C:566-574
    k.F0[1] = cc + h * (kappa_RT * Tau * Koppa - delta_RT * cc);                            // ripf.C:576-582
    k.F0[2] = fb + h * (lambda_RT * Tau * Lom + omicro_RT * Tau * Ome - p.omega * fb);      // ripf.C:584-594
    k.F1[0] = 0.0;
    k.F1[1] = 0.0;
    k.F1[2] = -h * (Tau * flux);
    k.C[0][0] = 1.0;                                                                        // ripf.C:599-603
    k.C[0][1] = -h * eps_cc;
    k.C[0][2] = -h * eps_fb;
    k.C[1][1] = 1.0 - h * (kappa_RT * dT * Koppa + kappa_RT * Tau * dKoppa - delta_RT);     // ripf.C:615-622
    k.C[1][2] = -h * (kappa_RT * dT * Koppa);
    k.C[2][0] = -h * (lambda_RT * Tau * dLom_HU);                                           // ripf.C:629-635 (dOme_HU = 0)
    k.S[2][0] = h * (p.haptotaxis * Tau * fb);
    k.C[2][1] = -h * (lambda_RT * dT * Lom + omicro_RT * dT * Ome);                         // ripf.C:636-646 (d./dcc = 0)
    k.T[2][1] = h * (dT * flux);
    k.C[2][2] = 1.0 - h * (lambda_RT * dT * Lom + lambda_RT * Tau * dLom_fb + omicro_RT * dT * Ome +
                           omicro_RT * Tau * dOme_fb - p.omega);                            // ripf.C:647-662
    k.S[2][2] = h * (p.diffusion * Tau);
    k.T[2][2] = h * (dT * flux + p.haptotaxis * Tau * GHU + p.radiotaxis * Tau * GRT);
  }
};

// ======================================================================================== PROTEAS
struct ProteasParams {
  double dt2;
  double T_max, RT_max, rho_h, u_h, delta_h, a_RT_h, b_RT_h, nu_h;
  double D_c, D_c_h, rho_c, u_c, delta_c, a_RT_c, b_RT_c, nu_c;
  double psi_n, k_n, u_n, rho_v, nu_v, D_e, rho_e, u_e, xi_e, p_RT_e, psi_e;
};

struct Proteas {
  static constexpr int NV = 5;
  static constexpr int NDIR = 3;  // grad tum, grad hos, grad oed
  static constexpr unsigned GRADMASK = 0b10011;
  static constexpr unsigned CMASK =
      bit(5, 0, 0) | bit(5, 0, 1) | bit(5, 0, 2) | bit(5, 0, 3) | bit(5, 1, 0) | bit(5, 1, 1) | bit(5, 1, 2) | bit(5, 1, 3) |
      bit(5, 2, 0) | bit(5, 2, 1) | bit(5, 2, 2) | bit(5, 2, 3) | bit(5, 3, 0) | bit(5, 3, 1) | bit(5, 3, 2) | bit(5, 3, 3) |
      bit(5, 4, 1) | bit(5, 4, 4);
  static constexpr unsigned SMASK = bit(5, 1, 0) | bit(5, 1, 1) | bit(5, 4, 4);
  static constexpr unsigned TMASK = bit(5, 1, 0) | bit(5, 1, 1) | bit(5, 1, 2) | bit(5, 1, 3);
  static constexpr int N_EFIELD = 0;
  static constexpr int N_NAUX = 1;  // RTD = phi_1(qp) * AUX(var 0, local node 1)  (proteas.C:481, Appendix C-4)
  static constexpr unsigned AUXGRADMASK = 0;
  typedef ProteasParams Params;

  __device__ static __forceinline__ void directions(const Params&, const double (*G)[3], const double*,
                                                    const double (*GA)[3], double (*dir)[3]) {
    (void)GA;
    for (int d = 0; d < 3; d++) { dir[0][d] = G[1][d]; dir[1][d] = G[0][d]; dir[2][d] = G[4][d]; }
  }

  __device__ static __forceinline__ void coef(const Params& p, const double* U, const double* A, const double* Dg,
                                              Coef<5>& k) {
    const double hos = U[0], tum = U[1], nec = U[2], vsc = U[3], oed = U[4];
    const double RTD = A[0];
    const double Gt = Dg[0], Gh = Dg[1], Ge = Dg[2];
    // proteas.C:488-491
    const double Tt = hos + tum + nec + vsc;
    double Kappa = 1.0 - Tt / p.T_max;
    Kappa = fmin(fmax(Kappa, 0.0), 1.0);
    const double dK = -1.0 / p.T_max;
    // proteas.C:493-514
    const double Hh = (vsc - p.u_h) > 0.0 ? 1.0 : 0.0, Hc = (vsc - p.u_c) > 0.0 ? 1.0 : 0.0, He = (vsc - p.u_e) > 0.0 ? 1.0 : 0.0;
    const double host_prol = p.rho_h * Kappa * Hh, dhost_prol = p.rho_h * dK * Hh;
    const double host_RTd = p.delta_h * (1.0 - exp(-p.a_RT_h * RTD - p.b_RT_h * (RTD * RTD)));
    const double host_nec = p.nu_h * nec;
    const double tum_prol = p.rho_c * Kappa * Hc, dtum_prol = p.rho_c * dK * Hc;
    const double tum_RTd = p.delta_c * (1.0 - exp(-p.a_RT_c * RTD - p.b_RT_c * (RTD * RTD)));
    const double tum_nec = p.nu_c * nec;
    const double nec_prol = p.nu_h * hos + p.nu_c * tum + p.nu_v * vsc;
    const double arg = p.k_n * vsc - p.u_n;
    const double nec_clear = p.psi_n * (1.0 - tanh(arg));
    const double ch = cosh(arg);
    const double dnec_clear_dv = p.psi_n * -p.k_n / (ch * ch);
    const double vsc_prol = p.rho_v * Kappa * tum, dvsc_prol = p.rho_v * dK * tum;
    const double vsc_nec = p.nu_v * nec;
    const double oed_prol = p.rho_e * tum * (1.0 - tum), doed_prol_dc = p.rho_e * (1.0 - 2.0 * tum);
    const double oed_RT = p.xi_e * pow(RTD / p.RT_max, p.p_RT_e);
    const double oed_clear = p.psi_e * (1.0 - He);
    const double h = p.dt2;
    const double logi = hos * (1.0 - hos);
    // load vector, proteas.C:520-564
    k.F0[0] = hos + h * (host_prol * logi - host_RTd * hos - host_nec * hos);
    k.F0[1] = tum + h * (tum_prol * tum - tum_RTd * tum - tum_nec * tum);
    k.F0[2] = nec + h * (nec_prol * nec - nec_clear * nec);
    k.F0[3] = vsc + h * (vsc_prol * vsc - vsc_nec * vsc);
    k.F0[4] = oed + h * (oed_prol * oed - oed_RT * oed - oed_clear * oed);
    k.F1[0] = 0.0;
    k.F1[1] = h * (-p.D_c * Kappa * Gt - p.D_c_h * Kappa * (Gh * tum));
    k.F1[2] = 0.0;
    k.F1[3] = 0.0;
    k.F1[4] = h * (-p.D_e * Ge);
    // host row, proteas.C:571-595
    k.C[0][0] = 1.0 - h * (dhost_prol * logi + host_prol * (1.0 - 2.0 * hos) - host_RTd - host_nec);
    k.C[0][1] = -h * (dhost_prol * logi);
    k.C[0][2] = -h * (dhost_prol * logi - p.nu_h * hos);
    k.C[0][3] = -h * (dhost_prol * logi);
    // tumour row, proteas.C:597-630 (block [1][1] has no D_c_h terms: reference inconsistency kept)
    const double fl_full = -p.D_c * dK * Gt - p.D_c_h * dK * (Gh * tum);
    k.C[1][0] = -h * (dtum_prol * tum);
    k.T[1][0] = -h * fl_full;
    k.S[1][0] = h * (p.D_c_h * Kappa * tum);
    k.C[1][1] = 1.0 - h * (dtum_prol * tum + tum_prol - tum_RTd - tum_nec);
    k.T[1][1] = -h * (-p.D_c * dK * Gt);
    k.S[1][1] = h * (p.D_c * Kappa);
    k.C[1][2] = -h * (dtum_prol * tum - p.nu_c * tum);
    k.T[1][2] = -h * fl_full;
    k.C[1][3] = -h * (dtum_prol * tum);
    k.T[1][3] = -h * fl_full;
    // necrotic row, proteas.C:632-654
    k.C[2][0] = -h * (p.nu_h * nec);
    k.C[2][1] = -h * (p.nu_c * nec);
    k.C[2][2] = 1.0 - h * (nec_prol - nec_clear);
    k.C[2][3] = -h * (p.nu_v * nec - dnec_clear_dv * nec);
    // vascular row, proteas.C:656-679 ([3][1] omits rho_v*Kappa*vsc: reference inconsistency kept)
    k.C[3][0] = -h * (dvsc_prol * vsc);
    k.C[3][1] = -h * (dvsc_prol * vsc);
    k.C[3][2] = -h * (dvsc_prol * vsc - p.nu_v * vsc);
    k.C[3][3] = 1.0 - h * (dvsc_prol * vsc + vsc_prol - vsc_nec);
    // oedema row, proteas.C:681-694
    k.C[4][1] = -h * (doed_prol_dc * oed);
    k.C[4][4] = 1.0 - h * (oed_prol - oed_RT - oed_clear);
    k.S[4][4] = h * p.D_e;
  }
};

// ============================================================================================ HCC
struct HccParams {
  double dt2;
  double Lambda_k, Kappa_k, ek, produce_l, diffuse_c, mechano_c, produce_c, nec_l, nec_c;  // nec_* already / Kappa_k
};

struct Hcc {
  static constexpr int NV = 3;
  static constexpr int NDIR = 1;  // grad c   (GRAD_sigma == 0, coupled_hcc.C:508: mechano terms vanish)
  static constexpr unsigned GRADMASK = 0b010;
  static constexpr unsigned CMASK = bit(3, 0, 0) | bit(3, 0, 1) | bit(3, 0, 2) | bit(3, 1, 0) | bit(3, 1, 1) | bit(3, 2, 0) | bit(3, 2, 1) | bit(3, 2, 2);
  static constexpr unsigned SMASK = bit(3, 1, 1);
  static constexpr unsigned TMASK = bit(3, 1, 0) | bit(3, 1, 1);
  static constexpr int N_EFIELD = 0;
  static constexpr int N_NAUX = 0;
  static constexpr unsigned AUXGRADMASK = 0;
  typedef HccParams Params;

  __device__ static __forceinline__ void directions(const Params&, const double (*G)[3], const double*,
                                                    const double (*GA)[3], double (*dir)[3]) {
    (void)GA;
    for (int d = 0; d < 3; d++) dir[0][d] = G[1][d];
  }

  __device__ static __forceinline__ void coef(const Params& p, const double* U, const double* A, const double* Dg,
                                              Coef<3>& k) {
    (void)A;
    const double l = U[0], c = U[1], n = U[2];
    const double Gc = Dg[0];
    double Tau, dT;  // coupled_hcc.C:510-532
    {
      const double Te = (l + c + n) / p.Kappa_k;
      if (Te <= 0.0) { Tau = 1.0; dT = 0.0; }
      else if (Te >= 1.0) { Tau = 0.0; dT = 0.0; }
      else { double xe1; pow_pair(1.0 - Te, p.ek, Tau, xe1); dT = (-p.ek / p.Kappa_k) * xe1; }
    }
    const double dif = c > p.Lambda_k ? p.diffuse_c : 0.0;  // coupled_hcc.C:534
    const double h = p.dt2;
    k.F0[0] = l + h * (p.produce_l * Tau * l - p.nec_l * l * n);  // :540-546
    k.F0[1] = c + h * (p.produce_c * Tau * c - p.nec_c * c * n);  // :548-556
    k.F0[2] = n + h * (p.nec_l * l * n + p.nec_c * c * n);        // :558-564
    k.F1[0] = 0.0;
    k.F1[1] = h * (-dif * Tau * Gc);
    k.F1[2] = 0.0;
    // capacity sits on [0][1], [0][2], [1][0] too, and the d/dn block of row c lands on [1][1] again
    // (coupled_hcc.C:577-619): reproduced, not fixed.
    k.C[0][0] = 1.0 - h * (p.produce_l * Tau + p.produce_l * dT * l - p.nec_l * n);
    k.C[0][1] = 1.0 - h * (p.produce_l * dT * l);
    k.C[0][2] = 1.0 - h * (p.produce_l * dT * l - p.nec_l * l);
    k.C[1][0] = 1.0 - h * (p.produce_c * dT * c);
    k.T[1][0] = h * (dif * dT * Gc);
    k.C[1][1] = (1.0 - h * (p.produce_c * Tau + p.produce_c * dT * c - p.nec_c * n)) +
                (1.0 - h * (p.produce_c * dT * c - p.nec_c * c));
    k.T[1][1] = h * (dif * dT * Gc) + h * (dif * dT * Gc);
    k.S[1][1] = h * (dif * Tau);
    k.C[2][0] = -h * (p.nec_l * n);
    k.C[2][1] = -h * (p.nec_c * n);
    k.C[2][2] = 1.0 - h * (p.nec_l * l + p.nec_c * c);
  }
};

}  // namespace rdc
