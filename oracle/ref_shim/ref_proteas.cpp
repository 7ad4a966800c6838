// ref_proteas.cpp -- the reference's src/proteas.C, unchanged, behind extern "C" entry points.  TEST INFRASTRUCTURE.
#include <cstdio>
#include REF_SOURCE
PerfLog plog("rdcFEs");
void eigen_decomposition(double[3][3], double[3][3], double[3]) {}

#define REF_PREFIX(name) ref_proteas_##name
static const char* ref_main_system() { return "PROTEAS_model"; }
static void ref_setup_systems(EquationSystems& es) {   // proteas.C:27-41
  TransientLinearImplicitSystem& model = es.add_system<TransientLinearImplicitSystem>("PROTEAS_model");
  for (const char* v : {"hos", "tum", "nec", "vsc", "oed"}) model.add_variable(v, FIRST, LAGRANGE);
  model.attach_init_function(initial_proteas_model);
  model.attach_assemble_function(assemble_proteas_model);
  ExplicitSystem& aux = es.add_system<ExplicitSystem>("AUX");
  for (const char* v : {"HU", "RTD"}) aux.add_variable(v, FIRST, LAGRANGE);
  aux.attach_init_function(initial_aux_data);
}
struct RefCtx;
static void ref_call_assemble(EquationSystems& es) { assemble_proteas_model(es, "PROTEAS_model"); }
static void ref_call_input(const char* file, EquationSystems& es) { input(file, es); }
#include "ref_api.inc"
static void ref_call_check(RefCtx& c) { check_solution(c.es); }
static int ref_call_save(RefCtx&, const char*) { return -1; }   // proteas.C has no save_solution
