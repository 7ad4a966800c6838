// ref_ripf.cpp -- the reference's src/ripf.C, unchanged, behind extern "C" entry points.  TEST INFRASTRUCTURE.
#include <cstdio>
#include REF_SOURCE
PerfLog plog("rdcFEs");
void eigen_decomposition(double[3][3], double[3][3], double[3]) {}

#define REF_PREFIX(name) ref_ripf_##name
static const char* ref_main_system() { return "RIPF"; }
static void ref_setup_systems(EquationSystems& es) {   // ripf.C:22-41
  TransientLinearImplicitSystem& model = es.add_system<TransientLinearImplicitSystem>("RIPF");
  for (const char* v : {"HU", "cc", "fb"}) model.add_variable(v, FIRST, LAGRANGE);
  model.attach_assemble_function(assemble_ripf);
  model.attach_init_function(initial_ripf);
  System& rates = es.add_system<System>("RIPF-TimeDeriv");
  for (const char* v : {"HU_TimeDeriv", "cc_TimeDeriv", "fb_TimeDeriv"}) rates.add_variable(v, FIRST, LAGRANGE);
  ExplicitSystem& rt = es.add_system<ExplicitSystem>("RT");
  for (const char* v : {"RT_dose/broad", "RT_dose/focus", "RT_dose/total"}) rt.add_variable(v, FIRST, LAGRANGE);
  rt.attach_init_function(initial_radiotherapy);
}
struct RefCtx;
static void ref_call_assemble(EquationSystems& es) { assemble_ripf(es, "RIPF"); }
static void ref_call_input(const char* file, EquationSystems& es) { input(file, es); }
#include "ref_api.inc"
static void ref_call_check(RefCtx& c) { check_solution(c.es, c.prev_soln); }
static int ref_call_save(RefCtx& c, const char* csv) {
  std::ofstream f(csv, std::ios::app);
  f.precision(17);
  save_solution(f, c.es);
  return 0;
}
