// ref_solid.cpp -- the reference's src/solid_system.C (+ hyperelastic.h, hyperlastic_inline.h, eig3.C), unchanged, behind
// extern "C" entry points.  TEST INFRASTRUCTURE (oracle/): pins oracle/solid_oracle.c and the CUDA solid path.
//
// What runs is the reference's own SolidSystem::element_time_derivative (solid_system.C:146-271),
// side_time_derivative (:273-371), post_process (:394-538) and save_initial_mesh (:26-48), driven the way
// [upstream] FEMSystem::assembly drives them: per element a zeroed variable-major residual/Jacobian, the element term,
// the side term on every side, then add_vector / add_matrix.  The Newton iteration around it is libMesh's
// ([upstream] NewtonSolver) and is not part of this file.
#include <cstdio>
#include REF_SOURCE
PerfLog plog("rdcFEs");

struct SolidCtx {
  LibMeshInit init;
  Mesh mesh;
  EquationSystems es;
  SolidSystem* model = nullptr;
  std::string error;
  SolidCtx() : mesh(init.comm(), 3), es(mesh) {}
};

#define REF_TRY(...)                                    \
  try { __VA_ARGS__; }                                  \
  catch (const std::exception& e_) {                    \
    if (c) c->error = e_.what();                        \
    return -1;                                          \
  }

extern "C" {

void* ref_solid_create(int etype, int64_t N, int64_t E, const int32_t* conn, const double* xyz, const int32_t* subdomain) {
  SolidCtx* c = new SolidCtx();
  try {
    c->mesh.build(etype == 4 ? TET4 : HEX8, N, E, conn, xyz, subdomain);
    EquationSystems& es = c->es;
    // the add_system / add_variable lines of solid.C:27-60
    SolidSystem& model = es.add_system<SolidSystem>("SolidSystem");
    model.add_variable("x", FIRST, LAGRANGE);
    model.add_variable("y", FIRST, LAGRANGE);
    model.add_variable("z", FIRST, LAGRANGE);
    TransientExplicitSystem& aux_sys = es.add_system<TransientExplicitSystem>("SolidSystem::auxiliary");
    aux_sys.add_variable("undeformed_x", FIRST, LAGRANGE);
    aux_sys.add_variable("undeformed_y", FIRST, LAGRANGE);
    aux_sys.add_variable("undeformed_z", FIRST, LAGRANGE);
    ExplicitSystem& disp_sys = es.add_system<ExplicitSystem>("SolidSystem::displacement");
    disp_sys.add_variable("u_x", FIRST, LAGRANGE);
    disp_sys.add_variable("u_y", FIRST, LAGRANGE);
    disp_sys.add_variable("u_z", FIRST, LAGRANGE);
    ExplicitSystem& fibre_sys = es.add_system<ExplicitSystem>("SolidSystem::fibre");
    for (const char* v : {"fibre_reference_x", "fibre_reference_y", "fibre_reference_z", "fibre_current_x", "fibre_current_y", "fibre_current_z"})
      fibre_sys.add_variable(v, CONSTANT, MONOMIAL);
    ExplicitSystem& press_sys = es.add_system<ExplicitSystem>("SolidSystem::pressure");
    press_sys.add_variable("p", CONSTANT, MONOMIAL);
    ExplicitSystem& von_mises_sys = es.add_system<ExplicitSystem>("SolidSystem::von_mises");
    von_mises_sys.add_variable("VM", CONSTANT, MONOMIAL);
    es.init();
    // the part of SolidSystem::init_data (solid_system.C:50-99) that does not need a DiffSolver
    model.var[0] = model.variable_number("x"); model.var[1] = model.variable_number("y"); model.var[2] = model.variable_number("z");
    model.undefo_var[0] = aux_sys.variable_number("undeformed_x");
    model.undefo_var[1] = aux_sys.variable_number("undeformed_y");
    model.undefo_var[2] = aux_sys.variable_number("undeformed_z");
    // mesh_position_get(): the primary variables start as the node positions
    for (const Node* n : c->mesh.nodes_)
      for (unsigned d = 0; d < 3; d++) model.solution->set(n->dof_number(model.number(), model.var[d], 0), (*n)(d));
    model.System::update();
    model.save_initial_mesh();   // solid_system.C:26-48 (solid.C:68)
    es.parameters.set<Real>("pseudo_time") = 0.0;
    es.parameters.set<bool>("solver/assembly_use_symmetry") = false;
    es.parameters.set<std::string>("BCs") = " ";
    es.parameters.set<Real>("BCs/displacement_penalty") = 1.0e+5;
    c->model = &model;
  } catch (const std::exception& e) {
    fprintf(stderr, "ref_solid create: %s\n", e.what());
    delete c;
    return nullptr;
  }
  return c;
}
void ref_solid_destroy(void* h) { delete (SolidCtx*)h; }
const char* ref_solid_last_error(void* h) { return ((SolidCtx*)h)->error.c_str(); }

int ref_solid_set_real(void* h, const char* key, double v) { SolidCtx* c = (SolidCtx*)h; REF_TRY(c->es.parameters.set<Real>(key) = v); return 0; }
int ref_solid_set_bool(void* h, const char* key, int v) { SolidCtx* c = (SolidCtx*)h; REF_TRY(c->es.parameters.set<bool>(key) = v != 0); return 0; }
int ref_solid_set_string(void* h, const char* key, const char* v) { SolidCtx* c = (SolidCtx*)h; REF_TRY(c->es.parameters.set<std::string>(key) = v); return 0; }
int ref_solid_set_point(void* h, const char* key, double x, double y, double z) {
  SolidCtx* c = (SolidCtx*)h;
  REF_TRY(c->es.parameters.set<Point>(key) = Point(x, y, z));
  return 0;
}
int ref_solid_add_side(void* h, int64_t elem, int side, int id) {
  SolidCtx* c = (SolidCtx*)h;
  REF_TRY(c->mesh.get_boundary_info().add_side(c->mesh.elem_ptr((dof_id_type)elem), (unsigned short)side, (boundary_id_type)id));
  return 0;
}
// reference fibre direction per element (solid.C:303-337 writes variables 0-2 and 3-5 alike)
int ref_solid_set_fibres(void* h, const double* f) {
  SolidCtx* c = (SolidCtx*)h;
  REF_TRY({
    ExplicitSystem& fs = c->es.get_system<ExplicitSystem>("SolidSystem::fibre");
    for (const Elem* e : c->mesh.elems_) {
      std::vector<dof_id_type> di;
      for (unsigned v = 0; v < 6; v++) {
        fs.get_dof_map().dof_indices(e, di, v);
        fs.current_local_solution->set(di[0], f[(size_t)e->id() * 3 + v % 3]);
      }
    }
    *fs.solution = *fs.current_local_solution;
  });
  return 0;
}
// current node positions: what NewtonSolver's iterate holds; SolidSystem::update (solid_system.C:101-120) moves the mesh
// there (mesh_position_set) -- [upstream] FEMContext::elem_position_set does the same per element during assembly
int ref_solid_set_positions(void* h, const double* x) {
  SolidCtx* c = (SolidCtx*)h;
  REF_TRY({
    SolidSystem& m = *c->model;
    for (Node* n : c->mesh.nodes_)
      for (unsigned d = 0; d < 3; d++) {
        const double v = x[(size_t)n->id() * 3 + d];
        m.solution->set(n->dof_number(m.number(), m.var[d], 0), v);
        (*n)(d) = v;
      }
    m.System::update();
  });
  return 0;
}

static void run_element(SolidCtx* c, const Elem* e, bool jac, FEMContext& ctx, FEBase& fe, QGauss& q3, FEBase& sfe, QGauss& q2) {
  SolidSystem& m = *c->model;
  ctx.elem_ = e;
  ctx.elem_fe_ = &fe; ctx.side_fe_ = &sfe; ctx.elem_q_ = &q3; ctx.side_q_ = &q2;
  ctx.resize(3, e->n_nodes());
  fe.reinit(e);
  m.element_time_derivative(jac, ctx);
  for (unsigned s = 0; s < e->n_sides(); s++) {
    bool any = false;   // [upstream] FEMSystem::assembly visits boundary sides only; a side without a boundary id contributes nothing
    for (int bc : export_integers(c->es.parameters.get<std::string>("BCs")))
      any = any || c->mesh.get_boundary_info().has_boundary_id(e, (unsigned short)s, (boundary_id_type)bc);
    if (!any) continue;
    ctx.side_ = (unsigned char)s;
    sfe.reinit(e, s);
    m.side_time_derivative(jac, ctx);
  }
}

// one element: residual [3*nen] and Jacobian [(3*nen)^2], variable-major like FEMContext holds them
int ref_solid_element(void* h, int64_t elem, int want_jac, double* Re, double* Ke) {
  SolidCtx* c = (SolidCtx*)h;
  REF_TRY({
    QGauss q3(3, THIRD), q2(2, THIRD);
    FEBase fe(3, FEType(FIRST, LAGRANGE)), sfe(3, FEType(FIRST, LAGRANGE));
    fe.attach_quadrature_rule(&q3); sfe.attach_quadrature_rule(&q2);
    FEMContext ctx;
    run_element(c, c->mesh.elem_ptr((dof_id_type)elem), want_jac != 0, ctx, fe, q3, sfe, q2);
    std::copy(ctx.residual_.v.begin(), ctx.residual_.v.end(), Re);
    if (want_jac) std::copy(ctx.jacobian_.a.begin(), ctx.jacobian_.a.end(), Ke);
  });
  return 0;
}

// the whole residual and Jacobian ([upstream] FEMSystem::assembly order: elements ascending, add_vector / add_matrix)
int ref_solid_assemble(void* h, int want_jac) {
  SolidCtx* c = (SolidCtx*)h;
  REF_TRY({
    SolidSystem& m = *c->model;
    m.matrix->clear_pattern();
    m.rhs->zero();
    QGauss q3(3, THIRD), q2(2, THIRD);
    FEBase fe(3, FEType(FIRST, LAGRANGE)), sfe(3, FEType(FIRST, LAGRANGE));
    fe.attach_quadrature_rule(&q3); sfe.attach_quadrature_rule(&q2);
    FEMContext ctx;
    std::vector<dof_id_type> di;
    for (const Elem* e : c->mesh.elems_) {
      run_element(c, e, want_jac != 0, ctx, fe, q3, sfe, q2);
      m.get_dof_map().dof_indices(e, di);
      m.rhs->add_vector(ctx.residual_, di);
      if (want_jac) m.matrix->add_matrix(ctx.jacobian_, di);
    }
  });
  return 0;
}
int64_t ref_solid_nnz(void* h) {
  SolidCtx* c = (SolidCtx*)h;
  int64_t n = 0;
  for (auto& r : c->model->matrix->rows) n += (int64_t)r.size();
  return n;
}
int ref_solid_get_csr(void* h, int64_t* rowptr, int32_t* col, double* val, double* rhs) {
  SolidCtx* c = (SolidCtx*)h;
  SolidSystem& m = *c->model;
  int64_t p = 0;
  rowptr[0] = 0;
  for (size_t i = 0; i < m.matrix->rows.size(); i++) {
    for (auto& kvp : m.matrix->rows[i]) { col[p] = (int32_t)kvp.first; val[p] = kvp.second; p++; }
    rowptr[i + 1] = p;
  }
  std::copy(m.rhs->v.begin(), m.rhs->v.end(), rhs);
  return 0;
}

// SolidSystem::post_process (solid_system.C:394-538): per element mean stress (hydrostatic), von Mises stress and the
// current fibre vector
int ref_solid_post_process(void* h, double* press, double* vm, double* fibre /* [E*3] */) {
  SolidCtx* c = (SolidCtx*)h;
  REF_TRY({
    c->model->post_process();
    ExplicitSystem& ps = c->es.get_system<ExplicitSystem>("SolidSystem::pressure");
    ExplicitSystem& vs = c->es.get_system<ExplicitSystem>("SolidSystem::von_mises");
    ExplicitSystem& fs = c->es.get_system<ExplicitSystem>("SolidSystem::fibre");
    std::vector<dof_id_type> di;
    for (const Elem* e : c->mesh.elems_) {
      ps.get_dof_map().dof_indices(e, di); press[e->id()] = (*ps.solution)(di[0]);
      vs.get_dof_map().dof_indices(e, di); vm[e->id()] = (*vs.solution)(di[0]);
      for (unsigned d = 0; d < 3; d++) { fs.get_dof_map().dof_indices(e, di, d + 3); fibre[(size_t)e->id() * 3 + d] = (*fs.solution)(di[0]); }
    }
  });
  return 0;
}

}  // extern "C"
