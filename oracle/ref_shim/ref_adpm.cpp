// ref_adpm.cpp -- compiles the reference's src/adpm.C unchanged (from where it lies) against the serial libMesh
// stand-in and exports its callbacks.  TEST INFRASTRUCTURE (oracle/); built into oracle/_ref/libref_adpm.so.
#include <cstdio>
#include REF_SOURCE   // "<reference>/src/adpm.C", given by oracle/build_ref.py
PerfLog plog("rdcFEs");                               // defined in main.C in the reference
void eigen_decomposition(double[3][3], double[3][3], double[3]) {}   // eig3.C: solid mechanics only, never called here

#define REF_PREFIX(name) ref_adpm_##name
static const char* ref_main_system() { return "ADPM"; }
static void ref_setup_systems(EquationSystems& es) {   // the system layout of adpm.C:23-37
  TransientLinearImplicitSystem& model = es.add_system<TransientLinearImplicitSystem>("ADPM");
  for (const char* v : {"PrP", "A_b", "Tau"}) model.add_variable(v, FIRST, LAGRANGE);
  model.attach_assemble_function(assemble_adpm);
  model.attach_init_function(initial_adpm);
  ExplicitSystem& tracts = es.add_system<ExplicitSystem>("Tracts");
  for (const char* v : {"TractX", "TractY", "TractZ"}) tracts.add_variable(v, CONSTANT, MONOMIAL);
  tracts.attach_init_function(initial_tracts);
}
struct RefCtx;
static void ref_call_assemble(EquationSystems& es) { assemble_adpm(es, "ADPM"); }
static void ref_call_input(const char* file, EquationSystems& es) { input(file, es); }
#include "ref_api.inc"
static void ref_call_check(RefCtx& c) { check_solution(c.es); }
static int ref_call_save(RefCtx& c, const char* csv) {
  std::ofstream f(csv, std::ios::app);
  f.precision(17);
  save_solution(f, c.es);
  return 0;
}
