// rdc_libmesh_adapter.h -- the reference-side glue: rdcFEs' libMesh callbacks on top of the C ABI (include/rdc.h).
//
// libMesh (+ PETSc, MPI) cannot be installed in the build container, so this header is compiled and RUN against the serial
// stand-in oracle/ref_shim/libmesh (libMesh's class and member names; the reference's own model files compile against it
// unchanged): tests/test_adapter.py builds tests/adapter/adapter_check.cpp, which drives the patched time loop of
// adpm.C:60-84 through this adapter for all five models on a GPU and compares with the oracle.  Against the real libMesh it
// has not been built.  It is the code a maintainer of InSilicoModellingGroup/rdcFEs drops into src/ and includes from
// adpm.C / pihna.C / ripf.C / proteas.C / coupled_hcc.C; INTEGRATION.md walks through it.  Everything
// numerical happens behind rdc.h; this file only flattens libMesh objects once and forwards the per-step calls:
//
//   assemble_<m>(es, name)      -> RdcAdapter::assemble()        (adpm.C:324, pihna.C:318, ripf.C:337, proteas.C:338, coupled_hcc.C:414)
//   model.solve() -> KSPSolve   -> RdcLinearSolver::solve()      (adpm.C:74, ...)
//   *older = *old; *old = *cur  -> RdcAdapter::rotate()          (adpm.C:71-72, ...)
//   check_solution(es)          -> RdcAdapter::check_solution()  (adpm.C:654-688, ...)
//   save_solution / paraview    -> RdcAdapter::pull_solution()   (adpm.C:79-83: output steps only)
//
// The parameter vector is read from es.parameters with the very keys input() stored (adpm.C:130-226 ...), through the
// table shared with the stand-alone driver (driver/param_tables.h); the angles are already in radians there
// (adpm.C:192,212) and RIPF's fraction counts are ints (ripf.C:228-231).
#pragma once
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include "libmesh/elem.h"
#include "libmesh/equation_systems.h"
#include "libmesh/linear_solver.h"
#include "libmesh/mesh_base.h"
#include "libmesh/node.h"
#include "libmesh/numeric_vector.h"
#include "libmesh/transient_system.h"

#include "../driver/param_tables.h"
#include "../include/rdc.h"

namespace rdcfes {

class RdcAdapter {
 public:
  RdcAdapter(libMesh::EquationSystems& es, const std::string& system_name, int model)
      : es_(es), name_(system_name), model_(model) {}
  ~RdcAdapter() { rdc_destroy(ctx_); }
  rdc_ctx* ctx() { return ctx_; }

  // once, after es.init() (adpm.C:43): flattened mesh, dof map, parameters, initial solution
  void hand_over() {
    using namespace libMesh;
    const MeshBase& mesh = es_.get_mesh();
    const System& sys = es_.get_system(name_);
    const unsigned s = sys.number();
    const int nen = (*mesh.active_elements_begin())->n_nodes();          // 4 (TET4) or 8 (HEX8)
    std::vector<int32_t> conn, region;
    for (const auto& elem : mesh.active_element_ptr_range()) {           // element id order == Gmsh file order
      for (unsigned l = 0; l < elem->n_nodes(); l++) conn.push_back(elem->node_id(l));
      region.push_back(elem->subdomain_id());
    }
    std::vector<double> xyz(3 * mesh.n_nodes());
    std::vector<int32_t> base(mesh.n_nodes());
    for (const auto& node : mesh.node_ptr_range()) {
      for (int d = 0; d < 3; d++) xyz[3 * node->id() + d] = (*node)(d);
      base[node->id()] = node->dof_number(s, 0, 0);                      // variables of a node are contiguous
    }
    int rc;
    if (mesh.n_processors() == 1) {
      rc = rdc_create(&ctx_, model_, nen, mesh.n_nodes(), conn.size() / nen, conn.data(), xyz.data(), base.data(), -1);
    } else {                                                             // one MPI rank per GPU
      std::vector<char> uid(128);
      if (mesh.processor_id() == 0) rdc_comm_unique_id(uid.data());
      mesh.comm().broadcast(uid, 0);                                     // Parallel::Communicator (adpm.C:20)
      rc = rdc_create_distributed(&ctx_, model_, nen, mesh.n_nodes(), conn.size() / nen, conn.data(), xyz.data(), base.data(),
                                  -1, mesh.processor_id(), mesh.n_processors(), /* METIS */ 0, uid.data());
    }
    check(rc, "rdc_create");
    set_parameters();
    std::vector<Number> soln;                                            // initial solution, global dof order
    sys.update_global_solution(soln);
    check(rdc_set_solution(ctx_, soln.data()), "rdc_set_solution");
  }

  // es.parameters -> flat vector, in the order of the reads of the callback (rdc.h enums)
  void set_parameters() {
    std::vector<double> p;
    auto fill = [&](const ParamKey* t, size_t n) {
      for (size_t k = 0; k < n; k++) {
        const std::string key = t[k].key;
        const bool is_int = key == "RT_dose/broad/fractions" || key == "RT_dose/focus/fractions";
        p.push_back(is_int ? (double)es_.parameters.get<int>(key) : (double)es_.parameters.get<libMesh::Real>(key));
      }
    };
    switch (model_) {
      case RDC_ADPM: fill(kAdpmTable, sizeof(kAdpmTable) / sizeof(ParamKey)); break;
      case RDC_PIHNA: fill(kPihnaTable, sizeof(kPihnaTable) / sizeof(ParamKey)); break;
      case RDC_RIPF: fill(kRipfTable, sizeof(kRipfTable) / sizeof(ParamKey)); break;
      case RDC_PROTEAS: fill(kProteasTable, sizeof(kProteasTable) / sizeof(ParamKey)); break;
      default: fill(kHccTable, sizeof(kHccTable) / sizeof(ParamKey)); break;
    }
    check(rdc_set_params(ctx_, p.data(), (int)p.size()), "rdc_set_params");
  }

  // element-constant aux system (ADPM "Tracts", adpm.C:32-37,453-458): ncomp values per element, element order
  void set_elem_field(const libMesh::System& aux, int ncomp) {
    std::vector<double> f;
    for (const auto& elem : es_.get_mesh().active_element_ptr_range())
      for (int v = 0; v < ncomp; v++) {
        std::vector<libMesh::dof_id_type> dofs;
        aux.get_dof_map().dof_indices(elem, dofs, v);
        f.push_back((*aux.solution)(dofs[0]));
      }
    check(rdc_set_elem_field(ctx_, 0, f.data(), ncomp), "rdc_set_elem_field");
  }

  // nodal aux system (RIPF "RT" broad/focus, ripf.C:36-41,275-289; PROTEAS "AUX", proteas.C:37-41): 2 values per node
  void set_nodal_field(const libMesh::System& aux) {
    std::vector<libMesh::Number> all;
    aux.update_global_solution(all);
    std::vector<double> f(2 * es_.get_mesh().n_nodes());
    for (const auto& node : es_.get_mesh().node_ptr_range())
      for (int v = 0; v < 2; v++) f[2 * node->id() + v] = all[node->dof_number(aux.number(), v, 0)];
    check(rdc_set_nodal_field(ctx_, 0, f.data(), 2), "rdc_set_nodal_field");
  }

  // body of assemble_<m>: K and F are built on the device; libMesh's matrix/rhs are not touched
  void assemble() {
    auto& system = es_.get_system<libMesh::TransientLinearImplicitSystem>(name_);
    check(rdc_assemble(ctx_, system.time, es_.parameters.get<libMesh::Real>("time_step")), "rdc_assemble");
  }
  void rotate() { check(rdc_rotate(ctx_), "rdc_rotate"); }                 // adpm.C:71-72
  void check_solution() {                                                 // adpm.C:76 (ripf.C:53 also before the loop)
    auto& system = es_.get_system<libMesh::TransientLinearImplicitSystem>(name_);
    check(rdc_set_time(ctx_, system.time), "rdc_set_time");
    check(rdc_set_dt(ctx_, es_.parameters.get<libMesh::Real>("time_step")), "rdc_set_dt");
    check(rdc_clamp(ctx_), "rdc_clamp");
  }
  void update_coords() {                                                  // after SolidSystem::update(), coupled_hcc.C:120-130
    const libMesh::MeshBase& mesh = es_.get_mesh();
    std::vector<double> xyz(3 * mesh.n_nodes());
    for (const auto& node : mesh.node_ptr_range())
      for (int d = 0; d < 3; d++) xyz[3 * node->id() + d] = (*node)(d);
    check(rdc_update_coords(ctx_, xyz.data()), "rdc_update_coords");
  }
  // output steps only (adpm.C:79-83): bring the solution back so that save_solution / paraview.update_pvd see it
  void pull_solution() {
    auto& system = es_.get_system<libMesh::TransientLinearImplicitSystem>(name_);
    std::vector<double> u((size_t)rdc_n_dofs(ctx_));
    check(rdc_get_solution(ctx_, u.data()), "rdc_get_solution");
    for (libMesh::dof_id_type i = system.solution->first_local_index(); i < system.solution->last_local_index(); i++)
      system.solution->set(i, u[i]);
    system.solution->close();
    system.update();
  }

 private:
  void check(int rc, const char* what) {
    if (rc) libmesh_error_msg(std::string(what) + ": " + rdc_last_error(ctx_));   // same convention as ripf.C:773
  }
  libMesh::EquationSystems& es_;
  std::string name_;
  int model_;
  rdc_ctx* ctx_ = nullptr;
};

// installed with  model.linear_solver.reset(new RdcLinearSolver(init.comm(), adapter))  after es.init();
// LinearImplicitSystem::solve() then calls assemble() (our callback) and this solve() instead of PETSc's KSP
// "rdc/ksp" in es.parameters: "bicgstab" (default), "gmres" (libMesh's own default, GMRES(30)) or "cg"
inline int rdc_ksp_from_parameters(const libMesh::EquationSystems& es) {
  if (!es.parameters.have_parameter<std::string>("rdc/ksp")) return RDC_KSP_BICGSTAB;
  const std::string k = es.parameters.get<std::string>("rdc/ksp");
  return k == "gmres" ? RDC_KSP_GMRES : (k == "cg" ? RDC_KSP_CG : RDC_KSP_BICGSTAB);
}

class RdcLinearSolver : public libMesh::LinearSolver<libMesh::Number> {
 public:
  // ksp: pass rdc_ksp_from_parameters(es) -- BiCGStab (the benchmarked method) unless es.parameters "rdc/ksp" says otherwise
  RdcLinearSolver(const libMesh::Parallel::Communicator& comm, RdcAdapter& a, int ksp = RDC_KSP_BICGSTAB)
      : libMesh::LinearSolver<libMesh::Number>(comm), a_(a), ksp_(ksp) {}
  void init(const char* = nullptr) override { this->_is_initialized = true; }
  void clear() override { this->_is_initialized = false; }
  std::pair<unsigned int, libMesh::Real> solve(libMesh::SparseMatrix<libMesh::Number>&, libMesh::SparseMatrix<libMesh::Number>&,
                                              libMesh::NumericVector<libMesh::Number>&, libMesh::NumericVector<libMesh::Number>&,
                                              const std::optional<double> tol, const std::optional<unsigned int> m_its) override {
    int its = 0;
    double res = 0.0;
    // libMesh defaults: "linear solver tolerance" 1e-12, "linear solver maximum iterations" 5000, GMRES restart 30
    const int rc = rdc_solve(a_.ctx(), ksp_, RDC_PC_JACOBI, tol.value_or(1e-12), (int)m_its.value_or(5000), 30, &its, &res);
    if (rc && rc != RDC_E_DIVERGED) libmesh_error_msg(rdc_last_error(a_.ctx()));
    reason_ = rc ? libMesh::DIVERGED_BREAKDOWN : libMesh::CONVERGED_RTOL_NORMAL;
    return {(unsigned)its, res};
  }
  std::pair<unsigned int, libMesh::Real> solve(const libMesh::ShellMatrix<libMesh::Number>&, libMesh::NumericVector<libMesh::Number>&,
                                              libMesh::NumericVector<libMesh::Number>&, const std::optional<double>,
                                              const std::optional<unsigned int>) override { libmesh_not_implemented(); }
  std::pair<unsigned int, libMesh::Real> solve(const libMesh::ShellMatrix<libMesh::Number>&, const libMesh::SparseMatrix<libMesh::Number>&,
                                              libMesh::NumericVector<libMesh::Number>&, libMesh::NumericVector<libMesh::Number>&,
                                              const std::optional<double>, const std::optional<unsigned int>) override { libmesh_not_implemented(); }
  void print_converged_reason() const override {}
  libMesh::LinearConvergenceReason get_converged_reason() const override { return reason_; }

 private:
  RdcAdapter& a_;
  int ksp_;
  libMesh::LinearConvergenceReason reason_ = libMesh::CONVERGED_ITERATING;
};

}  // namespace rdcfes

// ---- how adpm.C changes (the other four models alike) -------------------------------------------------------------
//   static std::unique_ptr<rdcfes::RdcAdapter> rdc;                       // next to pm_ptr (adpm.C:11)
//   void adpm(LibMeshInit& init) {
//     ...                                                                 // unchanged up to es.init()   (adpm.C:17-43)
//     rdc.reset(new rdcfes::RdcAdapter(es, "ADPM", RDC_ADPM));
//     rdc->hand_over();
//     rdc->set_elem_field(es.get_system("Tracts"), 3);
//     model.linear_solver.reset(new rdcfes::RdcLinearSolver(init.comm(), *rdc));
//     ...
//     for (int t = 1; t <= n_t_step; t++) {
//       ...time += dt...                                                  // adpm.C:63-68
//       rdc->rotate();                                                    // instead of adpm.C:71-72
//       model.solve();                                                    // -> assemble_adpm -> rdc->assemble(); -> RdcLinearSolver::solve
//       rdc->check_solution();                                            // instead of check_solution(es)  (adpm.C:76)
//       if (otp.count(t)) { rdc->pull_solution(); save_solution(csv, es); paraview.update_pvd(es, t); }
//     }
//   }
//   void assemble_adpm(EquationSystems&, const std::string&) { rdc->assemble(); }      // body of adpm.C:324-652
