/* solid_oracle.c -- CPU restatement of the reference's solid-mechanics Newton path.
 *
 * TEST INFRASTRUCTURE (oracle/).  Included at the end of rdc_oracle.c (one translation unit: it uses the FE tables,
 * the sparsity builder and the GMRES of that file).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
 * may call it; the product path never does.
 *
 * Parity status: PINNED for the in-tree arithmetic -- tests/test_solid_pin.py holds every function below to the
 * reference's own src/solid_system.C + hyperelastic.h + hyperlastic_inline.h + eig3.C compiled unchanged
 * (oracle/ref_shim/ref_solid.cpp -> oracle/_ref/libref_solid.so).  Restated [upstream] pieces: FE tables / FEMap
 * (rdc_oracle.c), the side quadrature rules and side map (QGauss 2-D THIRD, FEMap::compute_face_map), and libMesh's
 * NewtonSolver::solve driver (the reference only configures it, solid_system.C:80-98).
 *
 * Reference lines:
 *   material            hyperelastic.h:31-56 (initialize), hyperlastic_inline.h:17-189 (calculate_stress)
 *   B matrix, residual  hyperlastic_inline.h:1-15, hyperelastic.h:58-74
 *   tangent             hyperelastic.h:76-99
 *   element loop        solid_system.C:146-271
 *   penalty side term   solid_system.C:273-371
 *   post-processing     solid_system.C:394-538, eig3.C (JAMA tred2/tql2)
 *   load stepping       solid.C:81-108, solid_system.C:373-392 (run_solver), :101-120 (update = move the mesh)
 *
 * Unknowns are the CURRENT node positions (solid.C:27-30); the mesh is moved to the iterate before every residual
 * evaluation ([upstream] FEMContext::elem_position_set), so dphi and JxW live on the current configuration and
 * grad_X = dX/dx, F = (grad_X)^-1.  Dof convention: dof(node, d) = 3*node + d.
 */

/* ---- 3x3 helpers in libMesh's TypeTensor order ------------------------------------------------------------------ */
static double m3_det(const double a[3][3]) {
  return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
         a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}
static void m3_inv(const double a[3][3], double r[3][3]) { /* [upstream] TypeTensor::inverse: adjugate / det */
  const double d = m3_det(a);
  r[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) / d; r[0][1] = -(a[0][1] * a[2][2] - a[0][2] * a[2][1]) / d;
  r[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) / d; r[1][0] = -(a[1][0] * a[2][2] - a[1][2] * a[2][0]) / d;
  r[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) / d; r[1][2] = -(a[0][0] * a[1][2] - a[0][2] * a[1][0]) / d;
  r[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) / d; r[2][1] = -(a[0][0] * a[2][1] - a[0][1] * a[2][0]) / d;
  r[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) / d;
}
static void m3_mul(const double a[3][3], const double b[3][3], double r[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += a[i][k] * b[k][j];
      r[i][j] = s;
    }
}

typedef struct {
  double F[3][3];       /* deformation gradient (hyperelastic.h:44) */
  double sigma[3][3];   /* Cauchy stress */
  double tangent[6][6]; /* spatial tangent in Voigt form {00,11,22,01,12,02} */
} solid_mat;

/* hyperelastic.h:31-56 + hyperlastic_inline.h:17-189.  gradX[d][j] = d X_d / d x_j. */
static void solid_material(const double gradX[3][3], const double lambda_g[3], const double eta[3], double Young, double Poisson,
                           double FibreStiffness, int want_tangent, solid_mat* M) {
  double Fp[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, FpINV[3][3], Fe[3][3], A[3] = {0, 0, 0};
  m3_inv(gradX, M->F);                                  /* F = dX_dy.inverse() */
  for (int l = 0; l < 3; l++) Fp[l][l] = lambda_g[l];
  m3_inv(Fp, FpINV);
  m3_mul(M->F, FpINV, Fe);                              /* Fe = F * Fp.inverse() */
  if (FibreStiffness > 0.0) {                           /* A = f.unit() */
    const double l = sqrt(eta[0] * eta[0] + eta[1] * eta[1] + eta[2] * eta[2]);
    for (int d = 0; d < 3; d++) A[d] = eta[d] / l;
  }
  const double (*F)[3] = M->F;
  /* hyperlastic_inline.h:20-50 */
  const double mu = 0.5 * Young / (1.0 + Poisson);
  const double lambda = Young * Poisson / ((1.0 + Poisson) * (1.0 - 2.0 * Poisson));
  const double koppa = FibreStiffness / 2.0;
  double FeT[3][3], Ce[3][3], CeINV[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) FeT[i][j] = Fe[j][i];
  m3_mul(FeT, Fe, Ce);
  m3_inv(Ce, CeINV);
  const double Je = m3_det(Fe);
  const double J_recip = 1.0 / m3_det(F);
  const double dWdI1 = (mu / 2.0), dWdI2 = 0.0;
  const double dWdJe = (-mu / Je) + (lambda / 2.0 * Je - lambda / 2.0 / Je);
  const double dWdI4 = (-koppa);
  const double d2WdI1dI1 = 0.0, d2WdI2dI2 = 0.0, d2WdI4dI4 = 0.0;
  const double d2WdJedJe = (mu / Je / Je) + (lambda / 2.0 + lambda / 2.0 / Je / Je);
  const double I1 = Ce[0][0] + Ce[1][1] + Ce[2][2];
  double dI1dCe[3][3], dI2dCe[3][3], dJedCe[3][3], dI4dCe[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      const double dl = i == j ? 1.0 : 0.0;
      dI1dCe[i][j] = dl;
      dI2dCe[i][j] = dl * I1 - Ce[i][j];
      dJedCe[i][j] = 0.5 * Je * CeINV[i][j];
      dI4dCe[i][j] = A[i] * A[j];
    }
  double S2pk[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      S2pk[i][j] = 2.0 * dWdI1 * dI1dCe[i][j] + 2.0 * dWdI2 * dI2dCe[i][j] + 2.0 * dWdJe * dJedCe[i][j] + 2.0 * dWdI4 * dI4dCe[i][j];
  for (int i = 0; i < 3; i++)                            /* :89-101 push forward with F (not Fe), / det F */
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int I = 0; I < 3; I++)
        for (int J = 0; J < 3; J++) s += F[i][I] * F[j][J] * S2pk[I][J];
      M->sigma[i][j] = s * J_recip;
    }
  if (!want_tangent) return;
  /* :107-139: dSdCe, dCedC, dSdC */
  static const double h = 0.5;
  double dSdC[3][3][3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) {
          double acc = 0.0;
          for (int m = 0; m < 3; m++)
            for (int n = 0; n < 3; n++) {
              const double dij = i == j, dmn = m == n, dim = i == m, djn = j == n, din = i == n, djm = j == m;
              const double d2I2 = dij * dmn - h * dim * djn - h * din * djm;
              const double d2Je = 0.25 * Je * CeINV[i][j] * CeINV[m][n] - 0.25 * Je * CeINV[i][m] * CeINV[j][n] -
                                  0.25 * Je * CeINV[i][n] * CeINV[j][m];
              const double dSdCe = 4.0 * dWdI2 * d2I2 + 4.0 * dWdJe * d2Je + 4.0 * d2WdI1dI1 * dI1dCe[i][j] * dI1dCe[m][n] +
                                   4.0 * d2WdI2dI2 * dI2dCe[i][j] * dI2dCe[m][n] + 4.0 * d2WdJedJe * dJedCe[i][j] * dJedCe[m][n] +
                                   4.0 * d2WdI4dI4 * dI4dCe[i][j] * dI4dCe[m][n];
              const double dCedC = 0.5 * FpINV[k][m] * FpINV[n][l] + 0.5 * FpINV[l][m] * FpINV[k][n];
              acc += dSdCe * dCedC;
            }
          dSdC[i][j][k][l] = acc;
        }
  /* :141-157 tsm_ijkl = F_iI F_jJ F_kK F_lL dSdC_IJKL / det F.  The reference's 3^8 loop is evaluated here as four
   * successive single-index contractions (same products, different summation tree: 1e-16 relative). */
  double t1[3][3][3][3], t2[3][3][3][3];
  for (int i = 0; i < 3; i++) for (int J = 0; J < 3; J++) for (int K = 0; K < 3; K++) for (int L = 0; L < 3; L++) {
    double s = 0; for (int I = 0; I < 3; I++) s += F[i][I] * dSdC[I][J][K][L]; t1[i][J][K][L] = s; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int K = 0; K < 3; K++) for (int L = 0; L < 3; L++) {
    double s = 0; for (int J = 0; J < 3; J++) s += F[j][J] * t1[i][J][K][L]; t2[i][j][K][L] = s; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int L = 0; L < 3; L++) {
    double s = 0; for (int K = 0; K < 3; K++) s += F[k][K] * t2[i][j][K][L]; t1[i][j][k][L] = s; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) {
    double s = 0; for (int L = 0; L < 3; L++) s += F[l][L] * t1[i][j][k][L]; t2[i][j][k][l] = s * J_recip; }
  static const int vm[6][2] = {{0, 0}, {1, 1}, {2, 2}, {0, 1}, {1, 2}, {0, 2}};   /* :161-196 */
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) M->tangent[a][b] = t2[vm[a][0]][vm[a][1]][vm[b][0]][vm[b][1]];
}

/* hyperlastic_inline.h:1-15 */
static void solid_B(const double g[3], double B[3][6]) {
  memset(B, 0, 18 * sizeof(double));
  B[0][0] = g[0]; B[1][1] = g[1]; B[2][2] = g[2];
  B[0][3] = g[1]; B[1][3] = g[0]; B[1][4] = g[2]; B[2][4] = g[1]; B[0][5] = g[2]; B[2][5] = g[0];
}

/* material constants of one subdomain: {Young, Poisson, FibreStiffness, rate_0, rate_1, rate_2} (solid_system.C:182-189) */
#define SOLID_NMAT 6

/* solid_system.C:146-271.  Re [3*nen] and Ke [(3*nen)^2] are variable-major (entry (a*nen+i, b*nen+j)) and are
 * ADDED to (the side terms follow into the same arrays). */
static void solid_element(const fe_table* T, const double (*Xc)[3], const double (*Xu)[3], const double* mat, double pseudo_time,
                          const double* eta, int want_jac, int use_symmetry, double* Re, double* Ke) {
  const int nen = T->nen, nd = 3 * nen;
  double JxW[MAXQP], dphi[MAXNEN][MAXQP][3];
  fe_reinit(T, Xc, JxW, dphi);
  for (int qp = 0; qp < T->nqp; qp++) {
    double gradX[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int d = 0; d < 3; d++)
      for (int l = 0; l < nen; l++)
        for (int j = 0; j < 3; j++) gradX[d][j] += Xu[l][d] * dphi[l][qp][j];   /* add_scaled(dphi, XYZ_undefo) */
    double lam[3];
    for (int d = 0; d < 3; d++) lam[d] = 1.0 + pseudo_time * mat[3 + d];
    solid_mat M;
    solid_material(gradX, lam, eta, mat[0], mat[1], mat[2], want_jac, &M);
    const double SV[6] = {M.sigma[0][0], M.sigma[1][1], M.sigma[2][2], M.sigma[0][1], M.sigma[1][2], M.sigma[0][2]};
    for (int i = 0; i < nen; i++) {
      double BL[3][6];
      solid_B(dphi[i][qp], BL);
      for (int ii = 0; ii < 3; ii++) {               /* get_residual: R = B_L * SV, then scale(JxW) */
        double r = 0;
        for (int v = 0; v < 6; v++) r += BL[ii][v] * SV[v];
        Re[ii * nen + i] += r * JxW[qp];
      }
      if (!want_jac) continue;
      double BLT[3][6];                               /* B_L.right_multiply(tangent) */
      for (int a = 0; a < 3; a++)
        for (int w = 0; w < 6; w++) {
          double s = 0;
          for (int v = 0; v < 6; v++) s += BL[a][v] * M.tangent[v][w];
          BLT[a][w] = s;
        }
      for (int j = use_symmetry ? i : 0; j < nen; j++) {
        double BK[3][6], D[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        /* G_NN = (dphi_i * sigma) * dphi_j */
        double rs[3];
        for (int c = 0; c < 3; c++) rs[c] = dphi[i][qp][0] * M.sigma[0][c] + dphi[i][qp][1] * M.sigma[1][c] + dphi[i][qp][2] * M.sigma[2][c];
        const double G_NN = rs[0] * dphi[j][qp][0] + rs[1] * dphi[j][qp][1] + rs[2] * dphi[j][qp][2];
        for (int n = 0; n < 3; n++) D[n][n] += G_NN;
        solid_B(dphi[j][qp], BK);
        for (int a = 0; a < 3; a++)                   /* right_multiply_transpose(B_K); D += B_L */
          for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int w = 0; w < 6; w++) s += BLT[a][w] * BK[c][w];
            D[a][c] += s;
          }
        for (int ii = 0; ii < 3; ii++)
          for (int jj = 0; jj < 3; jj++) {
            Ke[(size_t)(ii * nen + i) * nd + jj * nen + j] += D[ii][jj] * JxW[qp];
            if (use_symmetry && i != j) Ke[(size_t)(jj * nen + j) * nd + ii * nen + i] += D[ii][jj] * JxW[qp];
          }
      }
    }
  }
}

/* [upstream] Tet4/Hex8::side_nodes_map */
static const int solid_side_tet[4][4] = {{0, 2, 1, 0}, {0, 1, 3, 0}, {1, 2, 3, 0}, {2, 0, 3, 0}};
static const int solid_side_hex[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};

/* [upstream] QGauss(2, THIRD) on the side + FEMap::compute_face_map; shape values of the side's nodes */
static int solid_side_rule(int nen, double w[4], double N[4][4] /*[node][qp]*/, double dxi[4][4], double deta[4][4]) {
  if (nen == 4) {
    static const double P[4][2] = {{1. / 3., 1. / 3.}, {.2, .6}, {.2, .2}, {.6, .2}};
    static const double W[4] = {-27. / 96., 25. / 96., 25. / 96., 25. / 96.};
    for (int q = 0; q < 4; q++) {
      w[q] = W[q];
      N[0][q] = 1. - P[q][0] - P[q][1]; N[1][q] = P[q][0]; N[2][q] = P[q][1];
      dxi[0][q] = -1.; dxi[1][q] = 1.; dxi[2][q] = 0.; deta[0][q] = -1.; deta[1][q] = 0.; deta[2][q] = 1.;
    }
    return 3;
  }
  const double g = 5.7735026918962576450914878050196e-01, p1[2] = {-g, g}, dL[2] = {-.5, .5};
  static const int i0[4] = {0, 1, 1, 0}, i1[4] = {0, 0, 1, 1};
  int q = 0;
  for (int j = 0; j < 2; j++)
    for (int i = 0; i < 2; i++, q++) {
      const double Lx[2] = {.5 * (1. - p1[i]), .5 * (1. + p1[i])}, Ly[2] = {.5 * (1. - p1[j]), .5 * (1. + p1[j])};
      w[q] = 1.0;
      for (int n = 0; n < 4; n++) { N[n][q] = Lx[i0[n]] * Ly[i1[n]]; dxi[n][q] = dL[i0[n]] * Ly[i1[n]]; deta[n][q] = Lx[i0[n]] * dL[i1[n]]; }
    }
  return 4;
}

/* solid_system.C:273-371 for ONE boundary condition on ONE side.  disp[3] = BC/<id>/displacement (NaN = free). */
static void solid_side(int nen, int side, const double (*Xc)[3], const double (*Xu)[3], const double* disp, double pseudo_time,
                       double penalty, int want_jac, double* Re, double* Ke) {
  const int nd = 3 * nen;
  const int* sn = nen == 4 ? solid_side_tet[side] : solid_side_hex[side];
  double w[4], N[4][4], dxi[4][4], deta[4][4];
  const int ns = solid_side_rule(nen, w, N, dxi, deta);
  const double ratio = pseudo_time * 1.000001;                       /* :285-286 */
  double dv[3];
  for (int d = 0; d < 3; d++) dv[d] = disp[d] * ratio;               /* :299-301 */
  for (int qp = 0; qp < 4; qp++) {
    double x[3] = {0, 0, 0}, X0[3] = {0, 0, 0}, a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
    for (int n = 0; n < ns; n++)
      for (int d = 0; d < 3; d++) {
        x[d] += Xc[sn[n]][d] * N[n][qp];
        X0[d] += N[n][qp] * Xu[sn[n]][d];                            /* :326-333 */
        a[d] += Xc[sn[n]][d] * dxi[n][qp];
        b[d] += Xc[sn[n]][d] * deta[n][qp];
      }
    const double cx = a[1] * b[2] - a[2] * b[1], cy = -a[0] * b[2] + a[2] * b[0], cz = a[0] * b[1] - a[1] * b[0];
    const double JxW = sqrt(cx * cx + cy * cy + cz * cz) * w[qp];
    double diff[3];
    for (int d = 0; d < 3; d++) diff[d] = x[d] - X0[d] - dv[d];      /* :336-337 */
    for (int i = 0; i < ns; i++) {
      for (int di = 0; di < 3; di++) {
        if (isnan(diff[di])) continue;
        Re[di * nen + sn[i]] += JxW * N[i][qp] * diff[di] * penalty; /* :344-347 */
      }
      if (!want_jac) continue;
      for (int j = 0; j < ns; j++)
        for (int dj = 0; dj < 3; dj++) {
          if (isnan(diff[dj])) continue;
          Ke[(size_t)(dj * nen + sn[i]) * nd + dj * nen + sn[j]] += JxW * N[i][qp] * N[j][qp] * penalty;   /* :357-360 */
        }
    }
  }
}

/* problem description shared by the entry points below */
typedef struct {
  int elem_type;
  int64_t N, E;
  const int32_t* conn;
  const double* xund;        /* [N*3] undeformed positions ("SolidSystem::auxiliary") */
  const int32_t* mat_of;     /* [E] material index or NULL (all 0) */
  const double* mats;        /* [nmat*6] */
  const double* fibres;      /* [E*3] or NULL */
  int64_t nside;
  const int64_t* side_elem;  /* [nside] */
  const int32_t* side_no;    /* [nside] local side number (libMesh side order) */
  const int32_t* side_bc;    /* [nside] index into bc_disp */
  const double* bc_disp;     /* [nbc*3], NaN = unconstrained component */
  double penalty;
  int use_symmetry;
} solid_problem;

static void solid_gather(const solid_problem* P, int64_t e, int nen, const double* x, double (*Xc)[3], double (*Xu)[3]) {
  for (int l = 0; l < nen; l++) {
    const int64_t n = P->conn[e * nen + l];
    for (int d = 0; d < 3; d++) { Xc[l][d] = x[3 * n + d]; Xu[l][d] = P->xund[3 * n + d]; }
  }
}

int orc_solid_element(int elem_type, const double* Xcur, const double* Xund, const double* mat, double pseudo_time, const double* eta,
                      int want_jac, int use_symmetry, double* Re, double* Ke) {
  fe_table T;
  if (fe_table_init(&T, elem_type)) return -1;
  const int nd = 3 * T.nen;
  memset(Re, 0, nd * sizeof(double));
  if (want_jac) memset(Ke, 0, (size_t)nd * nd * sizeof(double));
  solid_element(&T, (const double (*)[3])Xcur, (const double (*)[3])Xund, mat, pseudo_time, eta, want_jac, use_symmetry, Re, Ke);
  return 0;
}
int orc_solid_side(int elem_type, int side, const double* Xcur, const double* Xund, const double* disp, double pseudo_time, double penalty,
                   int want_jac, double* Re, double* Ke) {
  const int nen = elem_type == RDC_TET4 ? 4 : 8, nd = 3 * nen;
  memset(Re, 0, nd * sizeof(double));
  if (want_jac) memset(Ke, 0, (size_t)nd * nd * sizeof(double));
  solid_side(nen, side, (const double (*)[3])Xcur, (const double (*)[3])Xund, disp, pseudo_time, penalty, want_jac, Re, Ke);
  return 0;
}

/* [upstream] FEMSystem::assembly: elements in ascending order, element term + side terms, then add_vector/add_matrix.
 * rowptr/col = orc_build_pattern(N, E, nen, 3, ...).  val may be NULL (residual only). */
static int solid_assemble(const solid_problem* P, const double* x, double pseudo_time, const int64_t* rowptr, const int32_t* col,
                          double* val, double* rhs) {
  fe_table T;
  if (fe_table_init(&T, P->elem_type)) return -1;
  const int nen = T.nen, nd = 3 * nen;
  const int64_t D = 3 * P->N;
  memset(rhs, 0, D * sizeof(double));
  if (val) memset(val, 0, rowptr[D] * sizeof(double));
  /* sides grouped by element (input order kept inside an element) */
  int64_t* sptr = (int64_t*)calloc(P->E + 2, sizeof(int64_t));
  int64_t* sidx = (int64_t*)malloc((P->nside > 0 ? P->nside : 1) * sizeof(int64_t));
  if (!sptr || !sidx) { free(sptr); free(sidx); return -2; }
  for (int64_t s = 0; s < P->nside; s++) sptr[P->side_elem[s] + 2]++;
  for (int64_t e = 0; e < P->E; e++) sptr[e + 2] += sptr[e + 1];
  for (int64_t s = 0; s < P->nside; s++) sidx[sptr[P->side_elem[s] + 1]++] = s;
  double* Ke = (double*)malloc((size_t)nd * nd * sizeof(double));
  static const double zero3[3] = {0, 0, 0};
  int rc = 0;
  for (int64_t e = 0; e < P->E && !rc; e++) {
    double Xc[MAXNEN][3], Xu[MAXNEN][3], Re[3 * MAXNEN];
    solid_gather(P, e, nen, x, Xc, Xu);
    memset(Re, 0, sizeof(Re));
    if (val) memset(Ke, 0, (size_t)nd * nd * sizeof(double));
    const double* mat = P->mats + SOLID_NMAT * (P->mat_of ? P->mat_of[e] : 0);
    solid_element(&T, Xc, Xu, mat, pseudo_time, P->fibres ? P->fibres + 3 * e : zero3, val != NULL, P->use_symmetry, Re, Ke);
    for (int64_t k = sptr[e]; k < sptr[e + 1]; k++) {
      const int64_t s = sidx[k];
      solid_side(nen, P->side_no[s], Xc, Xu, P->bc_disp + 3 * P->side_bc[s], pseudo_time, P->penalty, val != NULL, Re, Ke);
    }
    for (int a = 0; a < 3 && !rc; a++)
      for (int i = 0; i < nen && !rc; i++) {
        const int64_t row = 3 * (int64_t)P->conn[e * nen + i] + a;
        rhs[row] += Re[a * nen + i];
        if (!val) continue;
        for (int b = 0; b < 3; b++)
          for (int j = 0; j < nen; j++) {
            const int64_t p = csr_find(rowptr, col, row, (int32_t)(3 * P->conn[e * nen + j] + b));
            if (p < 0) { rc = -3; break; }
            val[p] += Ke[(size_t)(a * nen + i) * nd + b * nen + j];
          }
      }
  }
  free(Ke); free(sptr); free(sidx);
  return rc;
}

static void solid_fill(solid_problem* P, int elem_type, int64_t N, int64_t E, const int32_t* conn, const double* xund, const int32_t* mat_of,
                       const double* mats, const double* fibres, int64_t nside, const int64_t* side_elem, const int32_t* side_no,
                       const int32_t* side_bc, const double* bc_disp, double penalty, int use_symmetry) {
  P->elem_type = elem_type; P->N = N; P->E = E; P->conn = conn; P->xund = xund; P->mat_of = mat_of; P->mats = mats; P->fibres = fibres;
  P->nside = nside; P->side_elem = side_elem; P->side_no = side_no; P->side_bc = side_bc; P->bc_disp = bc_disp; P->penalty = penalty;
  P->use_symmetry = use_symmetry;
}

int orc_solid_assemble(int elem_type, int64_t N, int64_t E, const int32_t* conn, const double* xcur, const double* xund,
                       const int32_t* mat_of, const double* mats, const double* fibres, double pseudo_time, int64_t nside,
                       const int64_t* side_elem, const int32_t* side_no, const int32_t* side_bc, const double* bc_disp, double penalty,
                       int use_symmetry, const int64_t* rowptr, const int32_t* col, double* val, double* rhs) {
  solid_problem P;
  solid_fill(&P, elem_type, N, E, conn, xund, mat_of, mats, fibres, nside, side_elem, side_no, side_bc, bc_disp, penalty, use_symmetry);
  return solid_assemble(&P, xcur, pseudo_time, rowptr, col, val, rhs);
}

static double solid_norm(int64_t D, const double* v) {
  double s = 0;
  for (int64_t i = 0; i < D; i++) s += v[i] * v[i];
  return sqrt(s);
}

/* One load step = SolidSystem::run_solver (solid_system.C:373-392) = [upstream] NewtonSolver::solve with the options of
 * solid_system.C:80-98.  opts = {max_nonlinear_iterations, relative_step_tolerance, relative_residual_tolerance,
 * absolute_residual_tolerance, require_reduction, max_linear_iterations, initial_linear_tolerance}; libMesh defaults for the
 * rest: linear_tolerance_multiplier 1e-3, minimum_linear_tolerance 1e-12, absolute_step_tolerance 0.
 * Restated: residual + Jacobian at the iterate, linear tolerance = max(min(previous, ||R|| * 1e-3, floor 1e-12),
 * atol/||R||/10), J d = R from d = 0 (GMRES(30) + ILU(0), PETSc's preconditioned relative test), x -= d, residual of the
 * full step, convergence tests (absolute / relative residual always, relative step only after a finished linear solve).
 * require_reduction = true backtracks by halving until the residual drops (libMesh refines the step with Brent's method
 * afterwards: not reproduced; no shipped input uses it).  x: in = start positions, out = converged positions.
 * info = {newton iterations, total linear iterations, final residual, converged flag}. */
int orc_solid_newton(int elem_type, int64_t N, int64_t E, const int32_t* conn, double* x, const double* xund, const int32_t* mat_of,
                     const double* mats, const double* fibres, double pseudo_time, int64_t nside, const int64_t* side_elem,
                     const int32_t* side_no, const int32_t* side_bc, const double* bc_disp, double penalty, int use_symmetry,
                     const double* opts, int pc, int nthreads, double* info) {
  solid_problem P;
  solid_fill(&P, elem_type, N, E, conn, xund, mat_of, mats, fibres, nside, side_elem, side_no, side_bc, bc_disp, penalty, use_symmetry);
  const int nen = elem_type == RDC_TET4 ? 4 : 8;
  const int64_t D = 3 * N;
  int64_t nnz = 0, *rowptr = NULL;
  int32_t* col = NULL;
  if (orc_build_pattern(N, E, nen, 3, conn, &nnz, &rowptr, &col)) return -1;
  double* val = (double*)malloc(nnz * sizeof(double));
  double* rhs = (double*)malloc(D * sizeof(double));
  double* dx = (double*)malloc(D * sizeof(double));
  const int max_nl = (int)opts[0];
  const double rel_step = opts[1], rel_res = opts[2], abs_res = opts[3];
  const int require_reduction = opts[4] != 0.0;
  const int max_lin = (int)opts[5];
  double lin_tol = opts[6];
  const double lin_mult = 1e-3, lin_min = 1e-12;
  double max_residual = 0.0, max_solution = 0.0;
  int outer = 0, inner = 0, converged = 0, rc = 0;
  double current_residual = 0.0;
  for (outer = 0; outer < max_nl; outer++) {
    if ((rc = solid_assemble(&P, x, pseudo_time, rowptr, col, val, rhs))) break;
    current_residual = solid_norm(D, rhs);
    if (current_residual != current_residual) { rc = -7; break; }
    if (current_residual == 0.0) { converged = 1; break; }   /* [upstream] "max_residual_norm == 0" guard: nothing to solve */
    if (current_residual > max_residual) max_residual = current_residual;
    const double norm_total = solid_norm(D, x);
    if (norm_total > max_solution) max_solution = norm_total;
    if (current_residual * lin_mult < lin_tol) lin_tol = current_residual * lin_mult;
    if (lin_tol < lin_min) lin_tol = lin_min;
    if (lin_tol < abs_res / current_residual / 10.0) lin_tol = abs_res / current_residual / 10.0;
    memset(dx, 0, D * sizeof(double));
    int its = 0;
    double res = 0, res0 = 0;
    const int grc = orc_gmres(D, rowptr, col, val, rhs, dx, pc, 1, 30, lin_tol, max_lin, nthreads, &its, &res, &res0);
    if (grc < 0) { rc = grc; break; }
    inner += its;
    const int linear_finished = its != max_lin;
    double norm_delta = solid_norm(D, dx);
    const double last_residual = current_residual;
    for (int64_t i = 0; i < D; i++) x[i] -= dx[i];
    if ((rc = solid_assemble(&P, x, pseudo_time, rowptr, col, NULL, rhs))) break;
    current_residual = solid_norm(D, rhs);
    double steplength = 1.0;
    if (require_reduction) {
      while (!(current_residual < last_residual) && steplength > 1e-6) {
        steplength *= 0.5;
        for (int64_t i = 0; i < D; i++) x[i] += steplength * dx[i];
        if ((rc = solid_assemble(&P, x, pseudo_time, rowptr, col, NULL, rhs))) break;
        current_residual = solid_norm(D, rhs);
      }
      if (rc) break;
      norm_delta *= steplength;
    }
    const double nt = solid_norm(D, x);
    if (nt > max_solution) max_solution = nt;
    int has = 0;
    if (current_residual < abs_res) has = 1;
    if (current_residual / max_residual < rel_res) has = 1;
    if (linear_finished && max_solution != 0.0 && norm_delta / max_solution < rel_step) has = 1;
    if (has) { converged = 1; outer++; break; }
  }
  if (info) { info[0] = outer; info[1] = inner; info[2] = current_residual; info[3] = converged; }
  free(val); free(rhs); free(dx); orc_free(rowptr); orc_free(col);
  return rc;
}

/* ---- eig3.C (public-domain JAMA tred2 + tql2, n = 3): eigenvalues ascending ----------------------------------------- */
static double solid_hypot2(double x, double y) { return sqrt(x * x + y * y); }
static void solid_tred2(double V[3][3], double d[3], double e[3]) {
  const int n = 3;
  for (int j = 0; j < n; j++) d[j] = V[n - 1][j];
  for (int i = n - 1; i > 0; i--) {
    double scale = 0.0, h = 0.0;
    for (int k = 0; k < i; k++) scale = scale + fabs(d[k]);
    if (scale == 0.0) {
      e[i] = d[i - 1];
      for (int j = 0; j < i; j++) { d[j] = V[i - 1][j]; V[i][j] = 0.0; V[j][i] = 0.0; }
    } else {
      for (int k = 0; k < i; k++) { d[k] /= scale; h += d[k] * d[k]; }
      double f = d[i - 1];
      double g = sqrt(h);
      if (f > 0) g = -g;
      e[i] = scale * g;
      h = h - f * g;
      d[i - 1] = f - g;
      for (int j = 0; j < i; j++) e[j] = 0.0;
      for (int j = 0; j < i; j++) {
        f = d[j];
        V[j][i] = f;
        g = e[j] + V[j][j] * f;
        for (int k = j + 1; k <= i - 1; k++) { g += V[k][j] * d[k]; e[k] += V[k][j] * f; }
        e[j] = g;
      }
      f = 0.0;
      for (int j = 0; j < i; j++) { e[j] /= h; f += e[j] * d[j]; }
      const double hh = f / (h + h);
      for (int j = 0; j < i; j++) e[j] -= hh * d[j];
      for (int j = 0; j < i; j++) {
        f = d[j];
        g = e[j];
        for (int k = j; k <= i - 1; k++) V[k][j] -= (f * e[k] + g * d[k]);
        d[j] = V[i - 1][j];
        V[i][j] = 0.0;
      }
    }
    d[i] = h;
  }
  for (int i = 0; i < n - 1; i++) {
    V[n - 1][i] = V[i][i];
    V[i][i] = 1.0;
    const double h = d[i + 1];
    if (h != 0.0) {
      for (int k = 0; k <= i; k++) d[k] = V[k][i + 1] / h;
      for (int j = 0; j <= i; j++) {
        double g = 0.0;
        for (int k = 0; k <= i; k++) g += V[k][i + 1] * V[k][j];
        for (int k = 0; k <= i; k++) V[k][j] -= g * d[k];
      }
    }
    for (int k = 0; k <= i; k++) V[k][i + 1] = 0.0;
  }
  for (int j = 0; j < n; j++) { d[j] = V[n - 1][j]; V[n - 1][j] = 0.0; }
  V[n - 1][n - 1] = 1.0;
  e[0] = 0.0;
}
static void solid_tql2(double V[3][3], double d[3], double e[3]) {
  const int n = 3;
  for (int i = 1; i < n; i++) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  double f = 0.0, tst1 = 0.0;
  const double eps = pow(2.0, -52.0);
  for (int l = 0; l < n; l++) {
    const double t = fabs(d[l]) + fabs(e[l]);
    tst1 = tst1 > t ? tst1 : t;
    int m = l;
    while (m < n) {
      if (fabs(e[m]) <= eps * tst1) break;
      m++;
    }
    if (m > l) {
      do {
        double g = d[l];
        double p = (d[l + 1] - g) / (2.0 * e[l]);
        double r = solid_hypot2(p, 1.0);
        if (p < 0) r = -r;
        d[l] = e[l] / (p + r);
        d[l + 1] = e[l] * (p + r);
        const double dl1 = d[l + 1];
        double h = g - d[l];
        for (int i = l + 2; i < n; i++) d[i] -= h;
        f = f + h;
        p = d[m];
        double c = 1.0, c2 = c, c3 = c;
        const double el1 = e[l + 1];
        double s = 0.0, s2 = 0.0;
        for (int i = m - 1; i >= l; i--) {
          c3 = c2; c2 = c; s2 = s;
          g = c * e[i];
          h = c * p;
          r = solid_hypot2(p, e[i]);
          e[i + 1] = s * r;
          s = e[i] / r;
          c = p / r;
          p = c * d[i] - s * g;
          d[i + 1] = h + s * (c * g + s * d[i]);
          for (int k = 0; k < n; k++) {
            h = V[k][i + 1];
            V[k][i + 1] = s * V[k][i] + c * h;
            V[k][i] = c * V[k][i] - s * h;
          }
        }
        p = -s * s2 * c3 * el1 * e[l] / dl1;
        e[l] = s * p;
        d[l] = c * p;
      } while (fabs(e[l]) > eps * tst1);
    }
    d[l] = d[l] + f;
    e[l] = 0.0;
  }
  for (int i = 0; i < n - 1; i++) {
    int k = i;
    double p = d[i];
    for (int j = i + 1; j < n; j++)
      if (d[j] < p) { k = j; p = d[j]; }
    if (k != i) {
      d[k] = d[i];
      d[i] = p;
      for (int j = 0; j < n; j++) { p = V[j][i]; V[j][i] = V[j][k]; V[j][k] = p; }
    }
  }
}
void orc_eig3(const double* A /*[9]*/, double* V /*[9]*/, double* d /*[3]*/) {
  double VV[3][3], e[3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) VV[i][j] = A[3 * i + j];
  solid_tred2(VV, d, e);
  solid_tql2(VV, d, e);
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) V[3 * i + j] = VV[i][j];
}

/* SolidSystem::post_process, solid_system.C:394-538 */
int orc_solid_post(int elem_type, int64_t N, int64_t E, const int32_t* conn, const double* xcur, const double* xund,
                   const int32_t* mat_of, const double* mats, const double* fibres, double pseudo_time, double* press, double* vm,
                   double* fibre_cur) {
  fe_table T;
  if (fe_table_init(&T, elem_type)) return -1;
  const int nen = T.nen;
  solid_problem P;
  solid_fill(&P, elem_type, N, E, conn, xund, mat_of, mats, fibres, 0, NULL, NULL, NULL, NULL, 0.0, 0);
  static const double zero3[3] = {0, 0, 0};
  for (int64_t e = 0; e < E; e++) {
    double Xc[MAXNEN][3], Xu[MAXNEN][3], JxW[MAXQP], dphi[MAXNEN][MAXQP][3];
    solid_gather(&P, e, nen, xcur, Xc, Xu);
    fe_reinit(&T, Xc, JxW, dphi);
    const double* mat = mats + SOLID_NMAT * (mat_of ? mat_of[e] : 0);
    const double* eta = fibres ? fibres + 3 * e : zero3;
    double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, fv[3] = {0, 0, 0};
    for (int qp = 0; qp < T.nqp; qp++) {
      double gradX[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, lam[3];
      for (int d = 0; d < 3; d++)
        for (int l = 0; l < nen; l++)
          for (int j = 0; j < 3; j++) gradX[d][j] += Xu[l][d] * dphi[l][qp][j];
      for (int d = 0; d < 3; d++) lam[d] = 1.0 + pseudo_time * mat[3 + d];
      solid_mat M;
      solid_material(gradX, lam, eta, mat[0], mat[1], mat[2], 0, &M);
      for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) S[i][j] += M.sigma[i][j];
        fv[i] += M.F[i][0] * eta[0] + M.F[i][1] * eta[1] + M.F[i][2] * eta[2];   /* F * eta (the raw eta, :502) */
      }
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) S[i][j] /= (double)T.nqp;
    const double Sc[9] = {S[0][0], S[0][1], S[0][2], S[0][1], S[1][1], S[1][2], S[0][2], S[1][2], S[2][2]};   /* :509-511 */
    double V[9], ev[3];
    orc_eig3(Sc, V, ev);
    press[e] = (ev[0] + ev[1] + ev[2]) / 3.0;
    vm[e] = sqrt(ev[0] * ev[0] + ev[1] * ev[1] + ev[2] * ev[2] - ev[0] * ev[1] - ev[0] * ev[2] - ev[1] * ev[2]);
    for (int d = 0; d < 3; d++) fibre_cur[3 * e + d] = fv[d] / (double)T.nqp;
  }
  return 0;
}
