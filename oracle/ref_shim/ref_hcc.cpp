// ref_hcc.cpp -- the reference's src/coupled_hcc.C, unchanged, behind extern "C" entry points.  TEST INFRASTRUCTURE.
// Only the reaction-diffusion half (assemble_hcc, check_solution) is run; the SolidSystem members that live in
// solid_system.C get empty bodies here so that the class the driver instantiates links.
#include <cstdio>
#include REF_SOURCE
PerfLog plog("rdcFEs");
void eigen_decomposition(double[3][3], double[3][3], double[3]) {}
void SolidSystem::init_data() {}
void SolidSystem::init_context(DiffContext&) {}
bool SolidSystem::element_time_derivative(bool, DiffContext&) { return false; }
bool SolidSystem::side_time_derivative(bool, DiffContext&) { return false; }
void SolidSystem::update() {}
void SolidSystem::save_initial_mesh() {}
void SolidSystem::run_solver() {}
void SolidSystem::post_process() {}
void SolidSystem::update_data() {}

#define REF_PREFIX(name) ref_hcc_##name
static const char* ref_main_system() { return "HCC"; }
static void ref_setup_systems(EquationSystems& es) {   // coupled_hcc.C:31-37 (the RD system; the solid systems are not built)
  TransientLinearImplicitSystem& model = es.add_system<TransientLinearImplicitSystem>("HCC");
  for (const char* v : {"l", "c", "n"}) model.add_variable(v, FIRST, LAGRANGE);
  model.attach_init_function(initial_hcc);
  model.attach_assemble_function(assemble_hcc);
}
struct RefCtx;
static void ref_call_assemble(EquationSystems& es) { assemble_hcc(es, "HCC"); }
static void ref_call_input(const char* file, EquationSystems& es) { input(file, es); }
#include "ref_api.inc"
static void ref_call_check(RefCtx& c) { check_solution(c.es); }
static int ref_call_save(RefCtx&, const char*) { return -1; }
