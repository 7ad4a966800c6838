// ref_pihna.cpp -- the reference's src/pihna.C, unchanged, behind extern "C" entry points.  TEST INFRASTRUCTURE.
#include <cstdio>
#include REF_SOURCE
PerfLog plog("rdcFEs");
void eigen_decomposition(double[3][3], double[3][3], double[3]) {}

#define REF_PREFIX(name) ref_pihna_##name
static const char* ref_main_system() { return "PIHNA"; }
static void ref_setup_systems(EquationSystems& es) {   // pihna.C:28-42
  TransientLinearImplicitSystem& model = es.add_system<TransientLinearImplicitSystem>("PIHNA");
  for (const char* v : {"n", "c", "h", "v", "a"}) model.add_variable(v, FIRST, LAGRANGE);
  model.attach_assemble_function(assemble_pihna);
  model.attach_init_function(initial_pihna);
  ExplicitSystem& ustruct = es.add_system<ExplicitSystem>("uStructure");
  for (const char* v : {"HU", "RT"}) ustruct.add_variable(v, CONSTANT, MONOMIAL);
  ustruct.attach_init_function(initial_structure);
}
struct RefCtx;
static void ref_call_assemble(EquationSystems& es) { assemble_pihna(es, "PIHNA"); }
static void ref_call_input(const char* file, EquationSystems& es) { input(file, es); }
#include "ref_api.inc"
static void ref_call_check(RefCtx& c) { check_solution(c.es); }
static int ref_call_save(RefCtx& c, const char* csv) {
  std::ofstream f(csv, std::ios::app);
  f.precision(17);
  save_solution(f, c.es);
  return 0;
}
