// solid_adapter_check.cpp -- adapter/rdc_libmesh_adapter.h (RdcSolidAdapter) run next to the reference's OWN SolidSystem.
//
// TEST INFRASTRUCTURE.  One program holds both sides: the reference's src/solid_system.C (+ hyperelastic.h, eig3.C) compiled
// unchanged against the serial libMesh stand-in (through oracle/ref_shim/ref_solid.cpp) and the device path reached the way
// a maintainer would reach it, from libMesh objects through the adapter.  It compares, in place,
//   (A) the Jacobian and the residual of rdc_solid_assemble with the reference's element_time_derivative +
//       side_time_derivative assembled like FEMSystem::assembly does,
//   (B) RdcSolidAdapter::post_process with SolidSystem::post_process,
// and (C) runs load steps with RdcSolidAdapter::run_solver, writing the final positions for the Python side (oracle Newton).
// Built where /root/reference exists; the binary travels to the GPU box.
//   usage: solid_adapter_check <case file> <out file>
#include "../../oracle/ref_shim/ref_solid.cpp"
#include "../../adapter/rdc_libmesh_adapter.h"

#include <cmath>
#include <fstream>

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  std::ifstream in(argv[1]);
  int et, nmat, nbc, nload;
  long N, E, nside;
  double penalty, pseudo_time, loading_step;
  in >> et >> N >> E;
  std::vector<double> xyz(3 * N), xpert(3 * N), mats, fibres(3 * E), disp;
  std::vector<int32_t> conn((size_t)E * et), mat_of(E);
  for (auto& v : xyz) in >> v;
  for (auto& v : xpert) in >> v;
  for (auto& v : conn) in >> v;
  for (auto& v : mat_of) in >> v;
  in >> nmat;
  mats.resize(6 * nmat);
  for (auto& v : mats) in >> v;
  for (auto& v : fibres) in >> v;
  in >> nbc;
  std::vector<int> bc_id(nbc);
  disp.resize(3 * nbc);
  for (int b = 0; b < nbc; b++) {
    in >> bc_id[b];
    for (int d = 0; d < 3; d++) { std::string tok; in >> tok; disp[3 * b + d] = tok == "nan" ? NAN : atof(tok.c_str()); }
  }
  in >> nside;
  std::vector<long> se(nside);
  std::vector<int> sn(nside), sb(nside);
  for (long k = 0; k < nside; k++) in >> se[k] >> sn[k] >> sb[k];
  double opts[7];
  in >> penalty >> pseudo_time >> nload >> loading_step;
  for (double& o : opts) in >> o;
  if (!in) { fprintf(stderr, "bad case file\n"); return 2; }

  // ---- the reference's systems (solid.C:27-66) on this mesh, parameters as solid.C:input() stores them
  SolidCtx* c = (SolidCtx*)ref_solid_create(et, N, E, conn.data(), xyz.data(), mat_of.data());
  if (!c) return 3;
  const char* keys[6] = {"Young", "Poisson", "FibreStiffness", "VolumetricStretchRatio/rate_0", "VolumetricStretchRatio/rate_1",
                         "VolumetricStretchRatio/rate_2"};
  for (int m = 0; m < nmat; m++)
    for (int k = 0; k < 6; k++) ref_solid_set_real(c, ("material/" + std::to_string(m) + "/Hyperelastic/" + keys[k]).c_str(), mats[6 * m + k]);
  std::string bcs = " ";
  for (int b = 0; b < nbc; b++) {
    bcs += std::to_string(bc_id[b]) + " ";
    ref_solid_set_point(c, ("BC/" + std::to_string(bc_id[b]) + "/displacement").c_str(), disp[3 * b], disp[3 * b + 1], disp[3 * b + 2]);
  }
  ref_solid_set_string(c, "BCs", bcs.c_str());
  ref_solid_set_real(c, "BCs/displacement_penalty", penalty);
  for (long k = 0; k < nside; k++) ref_solid_add_side(c, se[k], sn[k], bc_id[sb[k]]);
  ref_solid_set_fibres(c, fibres.data());
  c->es.parameters.set<int>("solver/nonlinear/max_nonlinear_iterations") = (int)opts[0];
  c->es.parameters.set<Real>("solver/nonlinear/relative_step_tolerance") = opts[1];
  c->es.parameters.set<Real>("solver/nonlinear/relative_residual_tolerance") = opts[2];
  c->es.parameters.set<Real>("solver/nonlinear/absolute_residual_tolerance") = opts[3];
  c->es.parameters.set<bool>("solver/nonlinear/require_reduction") = opts[4] != 0.0;
  c->es.parameters.set<int>("solver/linear/max_linear_iterations") = (int)opts[5];
  c->es.parameters.set<Real>("solver/linear/initial_linear_tolerance") = opts[6];

  // ---- (A) one state away from equilibrium: device Jacobian / residual vs the reference's own
  ref_solid_set_real(c, "pseudo_time", pseudo_time);
  ref_solid_set_positions(c, xpert.data());
  rdcfes::RdcSolidAdapter ad(c->es, *c->model);
  ad.hand_over();
  if (rdc_solid_assemble(ad.ctx(), pseudo_time)) { fprintf(stderr, "%s\n", rdc_last_error(ad.ctx())); return 4; }
  int64_t n_rows = 0, nnz = 0, *rows = nullptr, *rowptr = nullptr;
  int32_t* col = nullptr;
  double *val = nullptr, *rhs = nullptr;
  if (rdc_download_csr(ad.ctx(), &n_rows, &nnz, &rows, &rowptr, &col, &val, &rhs)) return 4;
  ref_solid_assemble(c, 1);
  const int64_t rnnz = ref_solid_nnz(c);
  std::vector<int64_t> rrp(3 * N + 1);
  std::vector<int32_t> rcol(rnnz);
  std::vector<double> rval(rnnz), rrhs(3 * N);
  ref_solid_get_csr(c, rrp.data(), rcol.data(), rval.data(), rrhs.data());
  double kmax = 0, kerr = 0, rmax = 0, rerr = 0;
  bool same_pattern = rnnz == nnz && n_rows == 3 * N;
  for (int64_t i = 0; same_pattern && i <= n_rows; i++) same_pattern = rowptr[i] == rrp[i];
  for (int64_t p = 0; same_pattern && p < nnz; p++) {
    same_pattern = col[p] == rcol[p];
    kmax = std::max(kmax, std::fabs(rval[p]));
    kerr = std::max(kerr, std::fabs(val[p] - rval[p]));
  }
  for (int64_t i = 0; i < 3 * N && same_pattern; i++) { rmax = std::max(rmax, std::fabs(rrhs[i])); rerr = std::max(rerr, std::fabs(rhs[i] - rrhs[i])); }
  // ---- (B) post-processing through the adapter vs SolidSystem::post_process
  std::vector<double> p_ref(E), vm_ref(E), f_ref(3 * E), p_dev(E), vm_dev(E), f_dev(3 * E);
  ad.post_process();
  {
    libMesh::System& ps = c->es.get_system("SolidSystem::pressure");
    libMesh::System& vs = c->es.get_system("SolidSystem::von_mises");
    libMesh::System& fs = c->es.get_system("SolidSystem::fibre");
    for (long e = 0; e < E; e++) {
      p_dev[e] = (*ps.solution)(e); vm_dev[e] = (*vs.solution)(e);
      for (int d = 0; d < 3; d++) f_dev[3 * e + d] = (*fs.solution)(e * 6 + 3 + d);
    }
  }
  ref_solid_post_process(c, p_ref.data(), vm_ref.data(), f_ref.data());
  double smax = 0, perr = 0, ferr = 0;
  for (long e = 0; e < E; e++) {
    smax = std::max(smax, std::fabs(p_ref[e]) + vm_ref[e]);
    perr = std::max(perr, std::max(std::fabs(p_dev[e] - p_ref[e]), std::fabs(vm_dev[e] - vm_ref[e])));
    for (int d = 0; d < 3; d++) ferr = std::max(ferr, std::fabs(f_dev[3 * e + d] - f_ref[3 * e + d]));
  }
  // ---- (C) load steps from the undeformed state through RdcSolidAdapter::run_solver
  ref_solid_set_positions(c, xyz.data());
  ad.push_positions();
  double t = 0.0;
  int all_converged = 1;
  std::vector<double> x(3 * N);
  for (int l = 1; l <= nload; l++) {
    t += loading_step;
    ref_solid_set_real(c, "pseudo_time", t);
    all_converged &= ad.run_solver(RDC_KSP_GMRES) ? 1 : 0;
    for (long i = 0; i < 3 * N; i++) x[i] = (*c->model->solution)(i);
    ref_solid_set_positions(c, x.data());   // what SolidSystem::update -> mesh_position_set does in libMesh
  }
  std::ofstream out(argv[2]);
  out.precision(17);
  out << (same_pattern ? 1 : 0) << ' ' << kerr / kmax << ' ' << rerr / rmax << ' ' << perr / smax << ' ' << ferr << ' ' << all_converged << '\n';
  for (double v : x) out << v << '\n';
  printf("pattern %d  dJ/max|J| %.3e  dR/max|R| %.3e  post %.3e  fibre %.3e  converged %d\n", (int)same_pattern, kerr / kmax, rerr / rmax,
         perr / smax, ferr, all_converged);
  rdc_free(rows); rdc_free(rowptr); rdc_free(col); rdc_free(val); rdc_free(rhs);
  return 0;
}
