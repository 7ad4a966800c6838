// solid_dev.cuh -- element-level arithmetic of the solid-mechanics path, written once for
// the device kernels of solid.cu and for the host-only probes the CPU tests call (rdc_solid_probe_*).
//
// Reference: SolidSystem::element_time_derivative (solid_system.C:146-271), side_time_derivative (:273-371),
// post_process (:394-538); material hyperelastic.h:31-99 + hyperlastic_inline.h:1-189.
//
// The reference forms the spatial tangent by a 3^8 push-forward of dS/dC (hyperlastic_inline.h:141-157) and multiplies it
// with two Voigt B matrices per node pair.  With its energy derivatives (dW/dI2 = d2W/dI1^2 = d2W/dI4^2 = 0) that chain has
// the closed form used here.  Let grad_X = dX/dx, F = grad_X^-1, Fp = diag(lambda) (growth), Q = F Fp F^-1 = F Fp grad_X,
// Je = det F / det Fp, b = Je dW/dJe, a = Je (dW/dJe + Je d2W/dJe2).  Then
//   sigma    = [ mu F F^T + b Q Q^T - K_f (F A)(F A)^T ] / det F
//   c_ijkl   = [ a (Q Q^T)_ij delta_kl - b (Q_ik Q_jl + Q_il Q_jk) ] / det F
//   K_ij[a][c] = JxW { delta_ac (g . sigma h) + [ a (QQ^T g)_a h_c - b ( Q_ac (g . Q h) + (Q h)_a (Q^T g)_c ) ] / det F }
// with g = grad phi_i, h = grad phi_j on the CURRENT configuration.  tests/test_solid_host.py holds every function below
// to the CPU oracle (which follows the reference term by term and is pinned to the compiled reference sources).
#pragma once
#include <math.h>

#include "rdc_internal.h"

namespace rdc {

#ifdef __CUDACC__
#define RDC_HD __host__ __device__ __forceinline__
#else
#define RDC_HD inline
#endif

// [upstream] FEMap for a 3-D element at quadrature point q: physical gradients and JxW on the configuration X
template <int NEN>
RDC_HD void solid_geometry(const FeTable& T, int q, const double (*X)[3], double (*dphi)[3], double& JxW) {
  double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // J[c][r] = d x_c / d xi_r
#pragma unroll
  for (int n = 0; n < NEN; n++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      J[c][0] += X[n][c] * T.dxi[n][q];
      J[c][1] += X[n][c] * T.deta[n][q];
      J[c][2] += X[n][c] * T.dzeta[n][q];
    }
  const double c0 = J[1][1] * J[2][2] - J[2][1] * J[1][2];
  const double c1 = J[2][1] * J[0][2] - J[0][1] * J[2][2];
  const double c2 = J[0][1] * J[1][2] - J[1][1] * J[0][2];
  const double jac = J[0][0] * c0 + J[1][0] * c1 + J[2][0] * c2;
  const double inv = 1.0 / jac;
  const double xix = c0 * inv, xiy = c1 * inv, xiz = c2 * inv;
  const double etax = (J[2][0] * J[1][2] - J[1][0] * J[2][2]) * inv;
  const double etay = (J[0][0] * J[2][2] - J[2][0] * J[0][2]) * inv;
  const double etaz = (J[1][0] * J[0][2] - J[0][0] * J[1][2]) * inv;
  const double zex = (J[1][0] * J[2][1] - J[2][0] * J[1][1]) * inv;
  const double zey = (J[2][0] * J[0][1] - J[0][0] * J[2][1]) * inv;
  const double zez = (J[0][0] * J[1][1] - J[1][0] * J[0][1]) * inv;
#pragma unroll
  for (int n = 0; n < NEN; n++) {
    dphi[n][0] = T.dxi[n][q] * xix + T.deta[n][q] * etax + T.dzeta[n][q] * zex;
    dphi[n][1] = T.dxi[n][q] * xiy + T.deta[n][q] * etay + T.dzeta[n][q] * zey;
    dphi[n][2] = T.dxi[n][q] * xiz + T.deta[n][q] * etaz + T.dzeta[n][q] * zez;
  }
  JxW = jac * T.w[q];
}

struct SolidPoint {       // the constitutive state at one quadrature point
  double F[3][3], Q[3][3], P[3][3], sig[3][3];
  double a, b, invJ;
};

// mat = {Young, Poisson, FibreStiffness, rate_0, rate_1, rate_2}; hyperelastic.h:31-56, hyperlastic_inline.h:17-101
template <int NEN>
RDC_HD void solid_point(const double (*Xu)[3], const double (*dphi)[3], const double* mat, double pseudo_time, const double* eta,
                        SolidPoint& S) {
  double G[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // grad_X[d][j] = d X_d / d x_j  (solid_system.C:216-226)
#pragma unroll
  for (int l = 0; l < NEN; l++)
#pragma unroll
    for (int d = 0; d < 3; d++)
#pragma unroll
      for (int j = 0; j < 3; j++) G[d][j] += Xu[l][d] * dphi[l][j];
  const double detG = G[0][0] * (G[1][1] * G[2][2] - G[1][2] * G[2][1]) - G[0][1] * (G[1][0] * G[2][2] - G[1][2] * G[2][0]) +
                      G[0][2] * (G[1][0] * G[2][1] - G[1][1] * G[2][0]);
  const double ig = 1.0 / detG;
  double (*F)[3] = S.F;                                  // F = grad_X^-1
  F[0][0] = (G[1][1] * G[2][2] - G[1][2] * G[2][1]) * ig; F[0][1] = -(G[0][1] * G[2][2] - G[0][2] * G[2][1]) * ig;
  F[0][2] = (G[0][1] * G[1][2] - G[0][2] * G[1][1]) * ig; F[1][0] = -(G[1][0] * G[2][2] - G[1][2] * G[2][0]) * ig;
  F[1][1] = (G[0][0] * G[2][2] - G[0][2] * G[2][0]) * ig; F[1][2] = -(G[0][0] * G[1][2] - G[0][2] * G[1][0]) * ig;
  F[2][0] = (G[1][0] * G[2][1] - G[1][1] * G[2][0]) * ig; F[2][1] = -(G[0][0] * G[2][1] - G[0][1] * G[2][0]) * ig;
  F[2][2] = (G[0][0] * G[1][1] - G[0][1] * G[1][0]) * ig;
  const double Jdet = F[0][0] * (F[1][1] * F[2][2] - F[1][2] * F[2][1]) - F[0][1] * (F[1][0] * F[2][2] - F[1][2] * F[2][0]) +
                      F[0][2] * (F[1][0] * F[2][1] - F[1][1] * F[2][0]);
  double lam[3];
#pragma unroll
  for (int d = 0; d < 3; d++) lam[d] = 1.0 + pseudo_time * mat[3 + d];        // solid_system.C:229-231
  const double Je = Jdet / (lam[0] * lam[1] * lam[2]);                          // det(F Fp^-1)
  const double Young = mat[0], Poisson = mat[1], Kf = mat[2];
  const double mu = 0.5 * Young / (1.0 + Poisson);
  const double lambda = Young * Poisson / ((1.0 + Poisson) * (1.0 - 2.0 * Poisson));
  const double dW = (-mu / Je) + (lambda / 2.0 * Je - lambda / 2.0 / Je);       // hyperlastic_inline.h:41
  const double d2W = (mu / Je / Je) + (lambda / 2.0 + lambda / 2.0 / Je / Je);  // :46
  S.b = Je * dW;
  S.a = Je * (dW + Je * d2W);
  S.invJ = 1.0 / Jdet;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) S.Q[i][j] = F[i][0] * lam[0] * G[0][j] + F[i][1] * lam[1] * G[1][j] + F[i][2] * lam[2] * G[2][j];
  double FA[3] = {0, 0, 0};
  if (Kf > 0.0) {                                                               // hyperelastic.h:53: A = f.unit()
    const double l = sqrt(eta[0] * eta[0] + eta[1] * eta[1] + eta[2] * eta[2]);
#pragma unroll
    for (int i = 0; i < 3; i++) FA[i] = (F[i][0] * eta[0] + F[i][1] * eta[1] + F[i][2] * eta[2]) / l;
  }
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = i; j < 3; j++) {
      const double p = S.Q[i][0] * S.Q[j][0] + S.Q[i][1] * S.Q[j][1] + S.Q[i][2] * S.Q[j][2];
      const double ff = F[i][0] * F[j][0] + F[i][1] * F[j][1] + F[i][2] * F[j][2];
      const double s = (mu * ff + S.b * p - Kf * FA[i] * FA[j]) * S.invJ;       // 2 dW/dI4 = -FibreStiffness (:43)
      S.P[i][j] = p; S.P[j][i] = p;
      S.sig[i][j] = s; S.sig[j][i] = s;
    }
}

// Row `li` of the element residual and tangent (element term only), accumulated over the quadrature points:
//   R[a] and K[(a*3 + c) * NEN + j] * kstride  (entry plane a*3+c, column node j); the caller zeroes K and R.
// use_symmetry = es.parameters "solver/assembly_use_symmetry" (solid_system.C:180,248-262): the reference then evaluates
// the blocks j >= i only and mirrors them, K_ji = (K_ij)^T; with growth or fibres the exact tangent is not symmetric, so this
// is a different matrix.  Row li therefore takes its blocks j < li as the transpose of block (j, li).
template <int NEN>
RDC_HD void solid_row(const FeTable& T, const double (*Xc)[3], const double (*Xu)[3], const double* mat, double pseudo_time,
                      const double* eta, int li, double* R, double* K, int kstride, int use_symmetry = 0) {
  // TET4: the map is affine, so grad phi, F and with them sigma and the tangent are the same at all five points of the
  // rule -- one evaluation carries the whole weight sum(w_q) = 1/6 (w_0 = -2/15 times -5/4).  HEX8: per point.
  constexpr int NQP = NEN == 4 ? 1 : 8;
#pragma unroll 1
  for (int q = 0; q < NQP; q++) {
    double dphi[NEN][3], JxW;
    solid_geometry<NEN>(T, q, Xc, dphi, JxW);
    if (NEN == 4) JxW *= -1.25;
    SolidPoint S;
    solid_point<NEN>(Xu, dphi, mat, pseudo_time, eta, S);
    double g[3] = {dphi[0][0], dphi[0][1], dphi[0][2]};
#pragma unroll
    for (int l = 1; l < NEN; l++)
      if (li == l) { g[0] = dphi[l][0]; g[1] = dphi[l][1]; g[2] = dphi[l][2]; }
    double sg[3], Pg[3], Qtg[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      sg[a] = S.sig[a][0] * g[0] + S.sig[a][1] * g[1] + S.sig[a][2] * g[2];
      Pg[a] = S.P[a][0] * g[0] + S.P[a][1] * g[1] + S.P[a][2] * g[2];
      Qtg[a] = S.Q[0][a] * g[0] + S.Q[1][a] * g[1] + S.Q[2][a] * g[2];
      R[a] += JxW * sg[a];                                                      // hyperelastic.h:58-74, solid_system.C:243-246
    }
    const double wa = JxW * S.a * S.invJ, wb = JxW * S.b * S.invJ;
#pragma unroll
    for (int j = 0; j < NEN; j++) {
      const double* h = dphi[j];
      double Qh[3];
#pragma unroll
      for (int a = 0; a < 3; a++) Qh[a] = S.Q[a][0] * h[0] + S.Q[a][1] * h[1] + S.Q[a][2] * h[2];
      const double gQh = g[0] * Qh[0] + g[1] * Qh[1] + g[2] * Qh[2];
      const double geo = JxW * (sg[0] * h[0] + sg[1] * h[1] + sg[2] * h[2]);    // G_NN, hyperelastic.h:83-85
      if (use_symmetry && j < li) {
        // mirrored block: K_{li,j}[a][c] = K_{j,li}[c][a], i.e. the formula with the roles of g and h exchanged, transposed
        double Ph[3], Qg[3], Qth[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
          Ph[a] = S.P[a][0] * h[0] + S.P[a][1] * h[1] + S.P[a][2] * h[2];
          Qg[a] = S.Q[a][0] * g[0] + S.Q[a][1] * g[1] + S.Q[a][2] * g[2];
          Qth[a] = S.Q[0][a] * h[0] + S.Q[1][a] * h[1] + S.Q[2][a] * h[2];
        }
        const double hQg = h[0] * Qg[0] + h[1] * Qg[1] + h[2] * Qg[2];
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
          for (int c = 0; c < 3; c++) {
            double v = wa * Ph[c] * g[a] - wb * (S.Q[c][a] * hQg + Qg[c] * Qth[a]);
            if (a == c) v += geo;
            K[(size_t)((a * 3 + c) * NEN + j) * kstride] += v;
          }
        continue;
      }
#pragma unroll
      for (int a = 0; a < 3; a++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
          double v = wa * Pg[a] * h[c] - wb * (S.Q[a][c] * gQh + Qh[a] * Qtg[c]);
          if (a == c) v += geo;
          K[(size_t)((a * 3 + c) * NEN + j) * kstride] += v;
        }
    }
  }
}

// [upstream] QGauss(2, THIRD) on a TRI3 / QUAD4 side: shape values and reference derivatives of the side's own nodes
template <int NS>
RDC_HD void solid_side_rule(int q, double* N, double* dxi, double* deta, double& w) {
  if (NS == 3) {
    const double px = q == 0 ? 1.0 / 3.0 : (q == 3 ? 0.6 : 0.2);
    const double py = q == 0 ? 1.0 / 3.0 : (q == 1 ? 0.6 : 0.2);
    w = q == 0 ? -27.0 / 96.0 : 25.0 / 96.0;
    N[0] = 1.0 - px - py; N[1] = px; N[2] = py;
    dxi[0] = -1.0; dxi[1] = 1.0; dxi[2] = 0.0; deta[0] = -1.0; deta[1] = 0.0; deta[2] = 1.0;
  } else {
    const double g = 5.7735026918962576450914878050196e-01;
    const double xi = (q & 1) ? g : -g, eta = (q & 2) ? g : -g;
    const double Lx[2] = {0.5 * (1.0 - xi), 0.5 * (1.0 + xi)}, Ly[2] = {0.5 * (1.0 - eta), 0.5 * (1.0 + eta)};
    w = 1.0;
    // node order of a QUAD4: (0,0) (1,0) (1,1) (0,1)
    N[0] = Lx[0] * Ly[0]; N[1] = Lx[1] * Ly[0]; N[2] = Lx[1] * Ly[1]; N[NS - 1] = Lx[0] * Ly[1];
    dxi[0] = -0.5 * Ly[0]; dxi[1] = 0.5 * Ly[0]; dxi[2] = 0.5 * Ly[1]; dxi[NS - 1] = -0.5 * Ly[1];
    deta[0] = Lx[0] * -0.5; deta[1] = Lx[1] * -0.5; deta[2] = Lx[1] * 0.5; deta[NS - 1] = Lx[0] * 0.5;
  }
}

// Penalty term of ONE boundary condition on ONE side for the row of the side's node `i` (solid_system.C:273-371):
//   R[d] += JxW N_i diff_d penalty ;  Kd[j*3 + d] += JxW N_i N_j penalty  (diagonal entries (d,d) of block (i, j) only).
// Xc/Xu: current / undeformed positions of the side's NS nodes; disp: prescribed displacement, NaN = component left free.
template <int NS>
RDC_HD void solid_bc_row(const double (*Xc)[3], const double (*Xu)[3], const double* disp, double pseudo_time, double penalty, int i,
                         double* R, double* Kd) {
  const double ratio = pseudo_time * 1.000001;   // solid_system.C:285-286
#pragma unroll 1
  for (int q = 0; q < 4; q++) {
    double N[NS], dxi[NS], deta[NS], w;
    solid_side_rule<NS>(q, N, dxi, deta, w);
    double x[3] = {0, 0, 0}, X0[3] = {0, 0, 0}, a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
#pragma unroll
    for (int n = 0; n < NS; n++)
#pragma unroll
      for (int d = 0; d < 3; d++) {
        x[d] += Xc[n][d] * N[n]; X0[d] += N[n] * Xu[n][d];
        a[d] += Xc[n][d] * dxi[n]; b[d] += Xc[n][d] * deta[n];
      }
    const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
    const double JxW = sqrt(cx * cx + cy * cy + cz * cz) * w;   // [upstream] FEMap::compute_face_map
    double Ni = N[0];
#pragma unroll
    for (int n = 1; n < NS; n++)
      if (i == n) Ni = N[n];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const double diff = x[d] - X0[d] - disp[d] * ratio;       // :336-337 (NaN displacement -> NaN -> skipped)
      if (diff != diff) continue;
      R[d] += JxW * Ni * diff * penalty;
#pragma unroll
      for (int j = 0; j < NS; j++) Kd[j * 3 + d] += JxW * Ni * N[j] * penalty;
    }
  }
}

// post_process of one element (solid_system.C:394-538): mean Cauchy stress over the quadrature points, its mean normal
// stress and von Mises stress, and the mean of F eta.  The reference takes the eigenvalues of the mean stress (eig3.C)
// and forms (e0+e1+e2)/3 and sqrt(e0^2+e1^2+e2^2-e0e1-e0e2-e1e2): both are invariants, evaluated here from the tensor
// itself (trace / 3, sqrt(3 J2)), using the upper triangle like solid_system.C:509-511.
template <int NEN>
RDC_HD void solid_post_elem(const FeTable& T, const double (*Xc)[3], const double (*Xu)[3], const double* mat, double pseudo_time,
                            const double* eta, double* out /* {p, vm, f0, f1, f2} */) {
  constexpr int NQP = NEN == 4 ? 1 : 8;   // TET4: constant stress, the mean over the rule's points is the value itself
  double s00 = 0, s11 = 0, s22 = 0, s01 = 0, s12 = 0, s02 = 0, f[3] = {0, 0, 0};
#pragma unroll 1
  for (int q = 0; q < NQP; q++) {
    double dphi[NEN][3], JxW;
    solid_geometry<NEN>(T, q, Xc, dphi, JxW);
    SolidPoint S;
    solid_point<NEN>(Xu, dphi, mat, pseudo_time, eta, S);
    s00 += S.sig[0][0]; s11 += S.sig[1][1]; s22 += S.sig[2][2]; s01 += S.sig[0][1]; s12 += S.sig[1][2]; s02 += S.sig[0][2];
#pragma unroll
    for (int i = 0; i < 3; i++) f[i] += S.F[i][0] * eta[0] + S.F[i][1] * eta[1] + S.F[i][2] * eta[2];
  }
  const double inv = 1.0 / (double)NQP;
  s00 *= inv; s11 *= inv; s22 *= inv; s01 *= inv; s12 *= inv; s02 *= inv;
  out[0] = (s00 + s11 + s22) / 3.0;
  const double d0 = s00 - s11, d1 = s11 - s22, d2 = s00 - s22;
  out[1] = sqrt(0.5 * (d0 * d0 + d1 * d1 + d2 * d2) + 3.0 * (s01 * s01 + s12 * s12 + s02 * s02));
  out[2] = f[0] * inv; out[3] = f[1] * inv; out[4] = f[2] * inv;
}

}  // namespace rdc
