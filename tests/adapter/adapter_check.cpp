// adapter_check.cpp -- runs adapter/rdc_libmesh_adapter.h (the reference-side glue of INTEGRATION.md) the way a patched
// adpm.C / pihna.C / ripf.C / proteas.C / coupled_hcc.C would: EquationSystems + TransientLinearImplicitSystem, the
// assemble callback forwarding to RdcAdapter::assemble(), RdcLinearSolver installed on model.linear_solver, and the
// time-loop body of adpm.C:60-84.  libMesh itself is not installable here, so the program is built against the serial
// stand-in oracle/ref_shim/libmesh (same class and member names; the reference's own model files compile against it
// unchanged): what is verified is that the adapter compiles against libMesh's interface and drives the C ABI correctly.
// TEST CODE.  Input: a text file written by tests/test_gpu_adapter.py; output: the solution after the steps.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <memory>

#include "../../adapter/rdc_libmesh_adapter.h"

using namespace libMesh;

static std::unique_ptr<rdcfes::RdcAdapter> rdc;
static void assemble_cb(EquationSystems&, const std::string&) { rdc->assemble(); }   // body of assemble_<model>

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: adapter_check <input> <output>\n"); return 2; }
  std::ifstream in(argv[1]);
  int model, nen, nsteps, nparams, ncomp_elem, ncomp_nodal;
  long N, E;
  double dt;
  std::string ksp;
  in >> model >> nen >> N >> E >> dt >> nsteps >> ksp >> nparams >> ncomp_elem >> ncomp_nodal;
  LibMeshInit init;
  Mesh mesh(init.comm(), 3);
  EquationSystems es(mesh);
  static const char* sysname[] = {"ADPM", "PIHNA", "RIPF", "PROTEAS_model", "HCC"};
  static const std::vector<std::vector<const char*>> vars = {
      {"PrP", "A_b", "Tau"}, {"n", "c", "h", "v", "a"}, {"HU", "cc", "fb"}, {"hos", "tum", "nec", "vsc", "oed"}, {"l", "c", "n"}};
  for (int k = 0; k < nparams; k++) {
    std::string key, kind;
    double v;
    in >> key >> kind >> v;
    if (kind == "int") es.parameters.set<int>(key) = (int)v;
    else es.parameters.set<Real>(key) = v;
  }
  es.parameters.set<Real>("time_step") = dt;
  es.parameters.set<Real>("time") = 0.0;
  es.parameters.set<std::string>("rdc/ksp") = ksp;
  std::vector<double> xyz(3 * N), u0, ef, nf;
  std::vector<int32_t> conn((size_t)E * nen);
  for (auto& x : xyz) in >> x;
  for (auto& c : conn) in >> c;
  TransientLinearImplicitSystem& modelsys = es.add_system<TransientLinearImplicitSystem>(sysname[model]);
  for (const char* v : vars[model]) modelsys.add_variable(v, FIRST, LAGRANGE);
  modelsys.attach_assemble_function(assemble_cb);
  System* aux = nullptr;
  if (model == RDC_ADPM) {
    aux = &es.add_system<ExplicitSystem>("Tracts");
    for (const char* v : {"TractX", "TractY", "TractZ"}) aux->add_variable(v, CONSTANT, MONOMIAL);
  } else if (model == RDC_RIPF) {
    aux = &es.add_system<ExplicitSystem>("RT");
    for (const char* v : {"RT_dose/broad", "RT_dose/focus", "RT_dose/total"}) aux->add_variable(v, FIRST, LAGRANGE);
  } else if (model == RDC_PROTEAS) {
    aux = &es.add_system<ExplicitSystem>("AUX");
    for (const char* v : {"HU", "RTD"}) aux->add_variable(v, FIRST, LAGRANGE);
  }
  mesh.build(nen == 4 ? TET4 : HEX8, N, E, conn.data(), xyz.data(), nullptr);
  es.init();
  const int nv = (int)vars[model].size();
  u0.resize((size_t)N * nv);
  for (auto& x : u0) in >> x;
  for (long i = 0; i < N * nv; i++) modelsys.solution->set(i, u0[i]);
  modelsys.solution->close();
  modelsys.update();
  if (ncomp_elem) {
    ef.resize((size_t)E * ncomp_elem);
    for (auto& x : ef) in >> x;
    for (size_t i = 0; i < ef.size(); i++) aux->solution->set(i, ef[i]);
  }
  if (ncomp_nodal) {
    nf.resize((size_t)N * ncomp_nodal);
    for (auto& x : nf) in >> x;
    const unsigned navars = aux->n_vars();
    for (long n = 0; n < N; n++)
      for (int v = 0; v < ncomp_nodal; v++) aux->solution->set(n * navars + v, nf[n * ncomp_nodal + v]);
  }
  if (aux) { aux->solution->close(); aux->update(); }
  try {
    // ---- the lines a patched driver adds after es.init() (adapter header, bottom)
    rdc.reset(new rdcfes::RdcAdapter(es, sysname[model], model));
    rdc->hand_over();
    if (model == RDC_ADPM) rdc->set_elem_field(*aux, 3);
    if (model == RDC_RIPF || model == RDC_PROTEAS) rdc->set_nodal_field(*aux);
    modelsys.linear_solver.reset(new rdcfes::RdcLinearSolver(init.comm(), *rdc, rdcfes::rdc_ksp_from_parameters(es)));
    if (model == RDC_RIPF) rdc->check_solution();                        // ripf.C:53
    // ---- the time loop of adpm.C:60-84
    for (int t = 1; t <= nsteps; t++) {
      es.parameters.set<Real>("time") += es.parameters.get<Real>("time_step");
      modelsys.time = es.parameters.get<Real>("time");
      rdc->rotate();                                                     // instead of adpm.C:71-72
      modelsys.solve();                                                  // -> assemble_cb -> RdcLinearSolver::solve
      rdc->check_solution();                                             // instead of check_solution(es)
    }
    rdc->pull_solution();                                                // output step (adpm.C:79-83)
  } catch (const std::exception& e) {
    fprintf(stderr, "adapter_check: %s\n", e.what());
    return 1;
  }
  FILE* f = fopen(argv[2], "w");
  for (long i = 0; i < N * nv; i++) fprintf(f, "%.17g\n", (*modelsys.solution)(i));
  fclose(f);
  printf("adapter_check ok: %d steps, last solve %u iterations\n", nsteps, modelsys.n_linear_iterations());
  rdc.reset();
  return 0;
}
