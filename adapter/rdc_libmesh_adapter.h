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
//   SolidSystem::save_initial_mesh / run_solver / post_process -> RdcSolidAdapter (solid_system.C:26-48, 373-538), see below
//
// The parameter vector is read from es.parameters with the very keys input() stored (adpm.C:130-226 ...), through the
// table shared with the stand-alone driver (driver/param_tables.h); the angles are already in radians there
// (adpm.C:192,212) and RIPF's fraction counts are ints (ripf.C:228-231).
#pragma once
#include <algorithm>
#include <memory>
#include <optional>
#include <set>
#include <sstream>
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

// ---- SolidSystem (solid_system.C): the Newton solve and the post-processing behind rdc_solid_* -------------------------
// Used from the three members the drivers call (solid.C:68,96-99, coupled_hcc.C:76,117-128):
//   SolidSystem::save_initial_mesh() -> RdcSolidAdapter::hand_over()     after the reference's own copy into the auxiliary system
//   SolidSystem::run_solver()        -> RdcSolidAdapter::run_solver()    instead of this->solve() (NewtonSolver)
//   SolidSystem::post_process()      -> RdcSolidAdapter::post_process()  instead of the element loop (solid_system.C:422-531)
// `sys` is the SolidSystem ("x","y","z" = current node positions), es holds "SolidSystem::auxiliary" (undeformed positions),
// "SolidSystem::fibre", "SolidSystem::pressure", "SolidSystem::von_mises" and the parameters of solid.C:input().
class RdcSolidAdapter {
 public:
  RdcSolidAdapter(libMesh::EquationSystems& es, libMesh::System& sys) : es_(es), sys_(sys) {}
  ~RdcSolidAdapter() { rdc_destroy(ctx_); }
  rdc_ctx* ctx() { return ctx_; }

  void hand_over() {
    using namespace libMesh;
    const MeshBase& mesh = es_.get_mesh();
    const System& aux = es_.get_system("SolidSystem::auxiliary");
    const System& fib = es_.get_system("SolidSystem::fibre");
    const int nen = (*mesh.active_elements_begin())->n_nodes();
    const dof_id_type N = mesh.n_nodes();
    std::vector<int32_t> conn, base(N);
    std::vector<int> sub;
    for (const auto& elem : mesh.active_element_ptr_range()) {
      for (unsigned l = 0; l < elem->n_nodes(); l++) conn.push_back(elem->node_id(l));
      sub.push_back(elem->subdomain_id());
    }
    const size_t E = sub.size();
    std::vector<double> xund(3 * (size_t)N);
    for (const auto& node : mesh.node_ptr_range()) {
      base[node->id()] = (int32_t)node->dof_number(sys_.number(), 0, 0);     // x, y, z are contiguous at a node
      for (unsigned d = 0; d < 3; d++) xund[3 * (size_t)node->id() + d] = aux.current_solution(node->dof_number(aux.number(), d, 0));
    }
    check(rdc_create(&ctx_, RDC_SOLID, nen, N, (int64_t)E, conn.data(), xund.data(), base.data(), -1), "rdc_create");
    check(rdc_solid_set_reference(ctx_, xund.data()), "rdc_solid_set_reference");
    check(rdc_solid_set_symmetry(ctx_, es_.parameters.have_parameter<bool>("solver/assembly_use_symmetry") &&
                                           es_.parameters.get<bool>("solver/assembly_use_symmetry")), "rdc_solid_set_symmetry");
    push_positions();
    // materials: one row per distinct subdomain id, keys of solid.C:276-291 (read per element at solid_system.C:182-189)
    std::vector<int> ids(sub);
    std::sort(ids.begin(), ids.end());
    ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
    std::vector<double> mats;
    for (int id : ids) {
      const std::string k = "material/" + std::to_string(id) + "/Hyperelastic/";
      for (const char* key : {"Young", "Poisson", "FibreStiffness", "VolumetricStretchRatio/rate_0", "VolumetricStretchRatio/rate_1",
                              "VolumetricStretchRatio/rate_2"})
        mats.push_back(es_.parameters.get<Real>(k + key));
    }
    std::vector<int32_t> mat_of(E);
    for (size_t e = 0; e < E; e++) mat_of[e] = (int32_t)(std::lower_bound(ids.begin(), ids.end(), sub[e]) - ids.begin());
    check(rdc_solid_set_materials(ctx_, (int)ids.size(), mats.data(), mat_of.data()), "rdc_solid_set_materials");
    // reference fibre direction: variables 0-2 of "SolidSystem::fibre" (solid_system.C:205-213)
    std::vector<double> fibres(3 * E);
    std::vector<dof_id_type> di;
    size_t e = 0;
    for (const auto& elem : mesh.active_element_ptr_range()) {
      for (unsigned d = 0; d < 3; d++) {
        fib.get_dof_map().dof_indices(elem, di, d);
        fibres[3 * e + d] = (*fib.current_local_solution)(di[0]);
      }
      e++;
    }
    check(rdc_solid_set_fibres(ctx_, fibres.data()), "rdc_solid_set_fibres");
    // boundary conditions: "BCs" ids, their displacement Points, every (elem, side) that carries one (solid_system.C:288-304)
    std::vector<int> bc_ids;
    {
      std::stringstream ss(es_.parameters.get<std::string>("BCs"));
      std::string tok;
      std::set<int> uniq;
      while (ss >> tok) { int n; if (std::stringstream(tok) >> n) uniq.insert(n); }   // export_integers (utils.h:268-287)
      bc_ids.assign(uniq.begin(), uniq.end());
    }
    std::vector<double> disp;
    for (int id : bc_ids) {
      const Point p = es_.parameters.get<Point>("BC/" + std::to_string(id) + "/displacement");
      for (unsigned d = 0; d < 3; d++) disp.push_back(p(d));
    }
    std::vector<int64_t> side_elem;
    std::vector<int32_t> side_no, side_bc;
    e = 0;
    for (const auto& elem : mesh.active_element_ptr_range()) {
      for (auto s : elem->side_index_range())
        for (size_t b = 0; b < bc_ids.size(); b++)
          if (mesh.get_boundary_info().has_boundary_id(elem, s, (boundary_id_type)bc_ids[b])) {
            side_elem.push_back((int64_t)e); side_no.push_back((int32_t)s); side_bc.push_back((int32_t)b);
          }
      e++;
    }
    check(rdc_solid_set_bcs(ctx_, (int)bc_ids.size(), disp.data(), (int64_t)side_elem.size(), side_elem.data(), side_no.data(), side_bc.data(),
                            es_.parameters.get<Real>("BCs/displacement_penalty")), "rdc_solid_set_bcs");
  }

  // current positions: libMesh solution vector -> device (after anything on the host changed them)
  void push_positions() {
    std::vector<double> x(sys_.solution->size());
    for (libMesh::dof_id_type i = 0; i < sys_.solution->size(); i++) x[i] = (*sys_.solution)(i);
    check(rdc_set_solution(ctx_, x.data()), "rdc_set_solution");
  }
  // device -> libMesh solution vector, then the mesh follows (SolidSystem::update -> mesh_position_set, solid_system.C:101-108)
  void pull_positions() {
    std::vector<double> x((size_t)rdc_n_dofs(ctx_));
    check(rdc_get_solution(ctx_, x.data()), "rdc_get_solution");
    for (libMesh::dof_id_type i = sys_.solution->first_local_index(); i < sys_.solution->last_local_index(); i++) sys_.solution->set(i, x[i]);
    sys_.solution->close();
    sys_.update();
  }

  // SolidSystem::run_solver (solid_system.C:373-392): the Newton solve with the options solid_system.C:80-98 gives NewtonSolver
  bool run_solver(int ksp = RDC_KSP_GMRES) {
    using libMesh::Real;
    const double opts[7] = {(double)es_.parameters.get<int>("solver/nonlinear/max_nonlinear_iterations"),
                            es_.parameters.get<Real>("solver/nonlinear/relative_step_tolerance"),
                            es_.parameters.get<Real>("solver/nonlinear/relative_residual_tolerance"),
                            es_.parameters.get<Real>("solver/nonlinear/absolute_residual_tolerance"),
                            es_.parameters.get<bool>("solver/nonlinear/require_reduction") ? 1.0 : 0.0,
                            (double)es_.parameters.get<int>("solver/linear/max_linear_iterations"),
                            es_.parameters.get<Real>("solver/linear/initial_linear_tolerance")};
    check(rdc_solid_newton(ctx_, es_.parameters.get<Real>("pseudo_time"), opts, ksp, info_), "rdc_solid_newton");
    pull_positions();
    return info_[3] != 0.0;
  }
  const double* last_info() const { return info_; }   // {newton iterations, linear iterations, residual, converged}

  // SolidSystem::post_process (solid_system.C:394-538): fills "SolidSystem::pressure", "::von_mises" and variables 3-5 of "::fibre"
  void post_process() {
    using namespace libMesh;
    const MeshBase& mesh = es_.get_mesh();
    System& ps = es_.get_system("SolidSystem::pressure");
    System& vs = es_.get_system("SolidSystem::von_mises");
    System& fs = es_.get_system("SolidSystem::fibre");
    const size_t E = mesh.n_elem();
    std::vector<double> p(E), vm(E), f(3 * E);
    check(rdc_solid_post_process(ctx_, es_.parameters.get<Real>("pseudo_time"), p.data(), vm.data(), f.data()), "rdc_solid_post_process");
    std::vector<dof_id_type> di;
    size_t e = 0;
    for (const auto& elem : mesh.active_element_ptr_range()) {
      ps.get_dof_map().dof_indices(elem, di); ps.solution->set(di[0], p[e]);
      vs.get_dof_map().dof_indices(elem, di); vs.solution->set(di[0], vm[e]);
      for (unsigned d = 0; d < 3; d++) { fs.get_dof_map().dof_indices(elem, di, d + 3); fs.solution->set(di[0], f[3 * e + d]); }
      e++;
    }
    ps.solution->close(); vs.solution->close(); fs.solution->close();
  }

 private:
  void check(int rc, const char* what) {
    if (rc) libmesh_error_msg(std::string(what) + ": " + rdc_last_error(ctx_));
  }
  libMesh::EquationSystems& es_;
  libMesh::System& sys_;
  rdc_ctx* ctx_ = nullptr;
  double info_[4] = {0, 0, 0, 0};
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
