// rdc_driver.cpp -- stand-alone C++ driver of the ADPM, PIHNA and RIPF models on top of the C ABI (include/rdc.h).
//
// Mirrors what rdcFEs' own drivers do around the hot path, without libMesh: adpm() / pihna() / ripf()
// (adpm.C:15-87, pihna.C:18-96, ripf.C:13-96) read input.dat with GetPot (input()), the Gmsh mesh, the nodal and
// elemental initial fields in node / element order (initial_adpm adpm.C:264-322, initial_tracts adpm.C:230-262,
// initial_pihna pihna.C:272-316, initial_ripf ripf.C:297-335, initial_radiotherapy ripf.C:255-295), then loop
//     time += dt; rotate; solve; check_solution; every output step: save_solution      (adpm.C:60-84)
// and write the CSV of save_solution (adpm.C:690-829, pihna.C:842-976, ripf.C:777-864).  Here the loop body is
// rdc_step and the CSV comes from rdc_region_last_mean / rdc_region_volumes.  It is the C++ counterpart of
// rdcfes_b200/system.py: the same entry points, called from the reference's own language, and what
// tests/test_gpu_driver.py runs end to end.  The input.dat keys and defaults come from param_tables.h, generated from
// the same table the Python mirror uses (driver/gen_tables.py).
//
//   rdc_driver -m adpm|pihna|ripf <input.dat> [ksp=0|1|2] [solution_out=<file>]   (paths relative to input.dat)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <array>
#include <fstream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "param_tables.h"
#include "vtu_writer.h"
#include "rdc.h"

static void die(const std::string& msg) {
  fprintf(stderr, "rdc_driver: %s\n", msg.c_str());
  exit(1);
}

// GetPot subset used by the shipped input files: `key = value`, '#' comments, optional quotes
struct Input {
  std::map<std::string, std::string> kv;
  explicit Input(const std::string& path) {
    std::ifstream f(path);
    if (!f) die("cannot open " + path);
    std::string line;
    while (std::getline(f, line)) {
      const size_t h = line.find('#');
      if (h != std::string::npos) line.erase(h);
      const size_t eq = line.find('=');
      if (eq == std::string::npos) continue;
      auto trim = [](std::string s) {
        const char* ws = " \t\r\n'\"";
        const size_t a = s.find_first_not_of(ws), b = s.find_last_not_of(ws);
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
      };
      kv[trim(line.substr(0, eq))] = trim(line.substr(eq + 1));
    }
  }
  double real(const std::string& k, double dflt) const {
    auto it = kv.find(k);
    return it == kv.end() ? dflt : atof(it->second.c_str());
  }
  int integer(const std::string& k, int dflt) const {
    auto it = kv.find(k);
    return it == kv.end() ? dflt : atoi(it->second.c_str());
  }
  std::string str(const std::string& k, const std::string& dflt) const {
    auto it = kv.find(k);
    return it == kv.end() ? dflt : it->second;
  }
};

struct Mesh {
  int nen = 0;
  std::vector<double> xyz;       // [N*3]
  std::vector<int32_t> conn;     // [E*nen], 0-based
  std::vector<int> subdomain;    // [E] first Gmsh tag (physical id) == libMesh subdomain_id
  // lower-dimensional elements of the file (TRI3 / QUAD4): [upstream] GmshIO turns their physical tag into the boundary id of
  // the volume-element side with the same nodes (BoundaryInfo::add_side) -- what SolidSystem::side_time_derivative asks for
  std::vector<std::array<int32_t, 4>> bface;   // node ids, -1 padded
  std::vector<int> btag;
};

// Gmsh 2.2 ASCII as written by process_mesh.C:22-83; volume elements only (4 = TET4, 5 = HEX8), file order
static Mesh read_gmsh(const std::string& path) {
  std::ifstream f(path);
  if (!f) die("cannot open mesh " + path);
  Mesh m;
  std::string tok;
  std::map<long, int32_t> id2idx;
  while (f >> tok) {
    if (tok == "$Nodes") {
      long n;
      f >> n;
      m.xyz.resize((size_t)n * 3);
      for (long k = 0; k < n; k++) {
        long id;
        f >> id >> m.xyz[3 * k] >> m.xyz[3 * k + 1] >> m.xyz[3 * k + 2];
        id2idx[id] = (int32_t)k;
      }
    } else if (tok == "$Elements") {
      long ne;
      f >> ne;
      std::string line;
      std::getline(f, line);
      for (long k = 0; k < ne; k++) {
        std::getline(f, line);
        std::istringstream ss(line);
        long id, type, ntags;
        ss >> id >> type >> ntags;
        std::vector<long> tags((size_t)ntags);
        for (auto& t : tags) ss >> t;
        if (type == 2 || type == 3) {   // 3-node triangle / 4-node quadrangle
          std::array<int32_t, 4> f = {-1, -1, -1, -1};
          for (int l = 0; l < (type == 2 ? 3 : 4); l++) {
            long v;
            ss >> v;
            f[l] = id2idx.at(v);
          }
          m.bface.push_back(f);
          m.btag.push_back(ntags > 0 ? (int)tags[0] : 0);
          continue;
        }
        const int nen = type == 4 ? 4 : (type == 5 ? 8 : 0);
        if (!nen) continue;
        if (m.nen && m.nen != nen) die("mixed volume element types");
        m.nen = nen;
        for (int l = 0; l < nen; l++) {
          long v;
          ss >> v;
          m.conn.push_back(id2idx.at(v));
        }
        m.subdomain.push_back(ntags > 0 ? (int)tags[0] : 0);
      }
    }
  }
  if (!m.nen) die("no TET4/HEX8 elements in " + path);
  return m;
}

// rows x cols numbers; blank lines and lines starting with '#' are skipped (proteas.C:296-301)
static std::vector<double> read_table(const std::string& path, size_t rows, int cols) {
  std::ifstream f(path);
  if (!f) die("cannot open " + path);
  std::vector<double> v;
  v.reserve(rows * cols);
  std::string line;
  while (v.size() < rows * cols && std::getline(f, line)) {
    const size_t a = line.find_first_not_of(" \t\r");
    if (a == std::string::npos || line[a] == '#') continue;
    std::istringstream iss(line);
    double x;
    while (iss >> x) v.push_back(x);
  }
  if (v.size() < rows * cols) die("short field file " + path);
  v.resize(rows * cols);
  return v;
}

// utils.h:268-288 export_integers: the integers of a blank-separated list
static std::set<int> integers_of(const std::string& s) {
  std::set<int> out;
  std::istringstream iss(s);
  std::string w;
  while (iss >> w) {
    int n;
    if (std::istringstream(w) >> n) out.insert(n);
  }
  return out;
}

template <size_t NK>
static std::vector<double> flat_params(const ParamKey (&table)[NK], const Input& in) {
  std::vector<double> p(NK);
  for (size_t k = 0; k < NK; k++) {
    double v = in.real(table[k].key, table[k].dflt);
    if (table[k].degrees) v *= M_PI / 180.0;   // utils.h:79 degrees_to_radians (adpm.C:192,212)
    p[k] = v;
  }
  return p;
}


// ---- solid mechanics (solid.C, solid_system.C; the other half of coupled_hcc.C) --------------------------------------
// Everything solid.C:input() / coupled_hcc.C:input() read for the SolidSystem, handed to a RDC_SOLID context.
struct SolidModel {
  rdc_ctx* ctx = nullptr;
  double opts[7];        // rdc_solid_newton: solver/nonlinear/*, solver/linear/* (solid.C:226-245)
  double pseudo_time = 0.0;
};

static SolidModel make_solid(const Input& in, const Mesh& mesh, const std::string& dir) {
  SolidModel sm;
  const int64_t N = (int64_t)mesh.xyz.size() / 3, E = (int64_t)mesh.conn.size() / mesh.nen;
  auto ck = [&](int rc, const char* what) {
    if (rc) die(std::string(what) + ": " + rdc_last_error(sm.ctx));
  };
  ck(rdc_create(&sm.ctx, RDC_SOLID, mesh.nen, N, E, mesh.conn.data(), mesh.xyz.data(), nullptr, -1), "rdc_create(solid)");
  // solid.C:243-244 (the shipped files spell it solver/use_symmetry, which is never read: default false)
  ck(rdc_solid_set_symmetry(sm.ctx, in.integer("solver/assembly_use_symmetry", 0) || in.str("solver/assembly_use_symmetry", "false") == "true"),
     "rdc_solid_set_symmetry");
  ck(rdc_set_solution(sm.ctx, mesh.xyz.data()), "rdc_set_solution");            // mesh_position_get (solid_system.C:82)
  ck(rdc_solid_set_reference(sm.ctx, mesh.xyz.data()), "rdc_solid_set_reference");   // save_initial_mesh (solid.C:68)
  // materials = the subdomain ids that carry parameters (solid.C:273-291); every element's subdomain must be one of them
  const std::set<int> mat_ids = integers_of(in.str("materials", " 0 "));
  std::map<int, int> mat_index;
  std::vector<double> mats;
  for (int id : mat_ids) {
    const std::string k = "material/" + std::to_string(id) + "/Hyperelastic/";
    mat_index.emplace(id, (int)mat_index.size());
    mats.push_back(in.real(k + "Young", 1.0e3));
    mats.push_back(in.real(k + "Poisson", 0.3));
    mats.push_back(in.real(k + "FibreStiffness", 0.0));
    for (int d = 0; d < 3; d++) mats.push_back(in.real(k + "VolumetricStretchRatio/rate_" + std::to_string(d), 0.0));
  }
  std::vector<int32_t> mat_of((size_t)E);
  for (int64_t e = 0; e < E; e++) {
    auto it = mat_index.find(mesh.subdomain[e]);
    if (it == mat_index.end())
      die("element subdomain " + std::to_string(mesh.subdomain[e]) + " has no material/<id>/Hyperelastic parameters (solid_system.C:182)");
    mat_of[e] = it->second;
  }
  ck(rdc_solid_set_materials(sm.ctx, (int)mat_ids.size(), mats.data(), mat_of.data()), "rdc_solid_set_materials");
  const std::string fib = in.str("input_fibres", ".");
  if (fib != ".") {   // solid.C:303-337: one direction per element, normalised
    std::vector<double> f = read_table(dir + fib, (size_t)E, 3);
    for (int64_t e = 0; e < E; e++) {
      const double m = sqrt(f[3 * e] * f[3 * e] + f[3 * e + 1] * f[3 * e + 1] + f[3 * e + 2] * f[3 * e + 2]);
      if (m <= 1.0e-6) die("input_fibres: zero fibre vector (solid.C:319)");
      for (int d = 0; d < 3; d++) f[3 * e + d] /= m;
    }
    ck(rdc_solid_set_fibres(sm.ctx, f.data()), "rdc_solid_set_fibres");
  }
  // boundary conditions: "BCs" ids, BC/<id>/displacement/<d> (NAN = free), penalty (solid.C:247-266)
  const std::set<int> bc_ids = integers_of(in.str("BCs", " 0 "));
  std::map<int, int> bc_index;
  std::vector<double> bc_disp;
  for (int id : bc_ids) {
    bc_index.emplace(id, (int)bc_index.size());
    for (int d = 0; d < 3; d++) bc_disp.push_back(in.real("BC/" + std::to_string(id) + "/displacement/" + std::to_string(d), 0.0));
  }
  // sides: the volume-element side with the nodes of each tagged boundary face ([upstream] side_nodes_map)
  static const int tet[4][4] = {{0, 2, 1, -1}, {0, 1, 3, -1}, {1, 2, 3, -1}, {2, 0, 3, -1}};
  static const int hex[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
  const int ns = mesh.nen == 4 ? 3 : 4, nsides = mesh.nen == 4 ? 4 : 6;
  std::map<std::array<int32_t, 4>, std::pair<int64_t, int>> wanted;   // sorted nodes of a tagged face -> slot
  std::vector<std::array<int32_t, 4>> keys(mesh.bface.size());
  for (size_t k = 0; k < mesh.bface.size(); k++) {
    if (!bc_index.count(mesh.btag[k])) continue;
    std::array<int32_t, 4> key = mesh.bface[k];
    std::sort(key.begin(), key.begin() + ns);
    keys[k] = key;
    wanted.emplace(key, std::make_pair((int64_t)-1, -1));
  }
  for (int64_t e = 0; e < E && !wanted.empty(); e++)
    for (int sd = 0; sd < nsides; sd++) {
      std::array<int32_t, 4> key = {-1, -1, -1, -1};
      for (int j = 0; j < ns; j++) key[j] = mesh.conn[(size_t)e * mesh.nen + (mesh.nen == 4 ? tet[sd][j] : hex[sd][j])];
      std::sort(key.begin(), key.begin() + ns);
      auto it = wanted.find(key);
      if (it != wanted.end() && it->second.first < 0) it->second = {e, sd};
    }
  std::vector<int64_t> side_elem;
  std::vector<int32_t> side_no, side_bc;
  for (size_t k = 0; k < mesh.bface.size(); k++) {
    if (!bc_index.count(mesh.btag[k])) continue;
    const auto& hit = wanted.at(keys[k]);
    if (hit.first < 0) die("a tagged boundary face of the mesh is not a side of any volume element");
    side_elem.push_back(hit.first); side_no.push_back(hit.second); side_bc.push_back(bc_index.at(mesh.btag[k]));
  }
  ck(rdc_solid_set_bcs(sm.ctx, (int)bc_ids.size(), bc_disp.data(), (int64_t)side_elem.size(), side_elem.data(), side_no.data(),
                       side_bc.data(), in.real("BCs/displacement_penalty", 1.0e5)), "rdc_solid_set_bcs");
  sm.opts[0] = in.integer("solver/nonlinear/max_nonlinear_iterations", 100);
  sm.opts[1] = in.real("solver/nonlinear/relative_step_tolerance", 1.0e-3);
  sm.opts[2] = in.real("solver/nonlinear/relative_residual_tolerance", 1.0e-8);
  sm.opts[3] = in.real("solver/nonlinear/absolute_residual_tolerance", 1.0e-8);
  sm.opts[4] = in.str("solver/nonlinear/require_reduction", "false") == "true" ? 1.0 : 0.0;
  sm.opts[5] = in.integer("solver/linear/max_linear_iterations", 50000);
  sm.opts[6] = in.real("solver/linear/initial_linear_tolerance", 1.0e-3);
  printf("solid: %lld elements, %zu materials, %zu boundary conditions on %zu sides\n", (long long)E, mat_ids.size(), bc_ids.size(), side_elem.size());
  return sm;
}

// SolidSystem::run_solver + post_process (solid.C:96-99, coupled_hcc.C:117-124) at the model's pseudo-time
static void solid_load_step(SolidModel& sm, int ksp, std::vector<double>* press, std::vector<double>* vm) {
  double info[4] = {0, 0, 0, 0};
  if (rdc_solid_newton(sm.ctx, sm.pseudo_time, sm.opts, ksp, info)) die(std::string("rdc_solid_newton: ") + rdc_last_error(sm.ctx));
  printf("  solid: pseudo-time %g, %d Newton iterations, %d linear iterations, residual %.3e%s\n", sm.pseudo_time, (int)info[0], (int)info[1],
         info[2], info[3] != 0.0 ? "" : "  (NOT converged)");
  if (press && vm && rdc_solid_post_process(sm.ctx, sm.pseudo_time, press->data(), vm->data(), nullptr))
    die(std::string("rdc_solid_post_process: ") + rdc_last_error(sm.ctx));
}

// -m solid: the driver of solid.C:14-112
static int run_solid(const Input& in, const std::string& dir, int ksp, bool vtu_binary, const std::string& sol_out) {
  Mesh mesh = read_gmsh(dir + in.str("input_GMSH", "input.msh"));
  const int64_t N = (int64_t)mesh.xyz.size() / 3, E = (int64_t)mesh.conn.size() / mesh.nen;
  SolidModel sm = make_solid(in, mesh, dir);
  const double loading_step = in.real("loading_step", 1.0);
  const int n_load = (int)(1.0 / loading_step);                                  // solid.C:166-167
  const int out_step = in.integer("output_step", 0);
  std::set<int> out_points;
  if (out_step > 0) for (int l = out_step; l <= n_load; l += out_step) out_points.insert(l);
  else out_points.insert(n_load);                                                // solid.C:171-177 ("output_time_points" of the file is not read)
  if (in.integer("remeshing_step", 0) > 0 && in.integer("remeshing_step", 0) <= n_load && in.integer("mesh/AMR/max_steps", 0) > 0)
    die("remeshing_step asks for adaptive mesh refinement (solid.C:102-103, 339-369), which this driver does not do");
  std::unique_ptr<PvdCollection> pvd;
  if (in.kv.count("output_PARAVIEW")) pvd.reset(new PvdCollection(dir + in.str("output_PARAVIEW", "output4paraview"), vtu_binary));
  const std::vector<std::string> names = {"x", "y", "z", "undeformed_x", "undeformed_y", "undeformed_z", "u_x", "u_y", "u_z"};   // solid.C:27-44
  std::vector<double> x((size_t)N * 3), vals((size_t)N * 9), press((size_t)E), vm((size_t)E);
  auto update_pvd = [&](unsigned l) {
    if (!pvd) return;
    if (rdc_get_solution(sm.ctx, x.data())) die("rdc_get_solution");
    for (int64_t n = 0; n < N; n++)
      for (int d = 0; d < 3; d++) {
        vals[(size_t)n * 9 + d] = x[(size_t)n * 3 + d];
        vals[(size_t)n * 9 + 3 + d] = mesh.xyz[(size_t)n * 3 + d];
        vals[(size_t)n * 9 + 6 + d] = x[(size_t)n * 3 + d] - mesh.xyz[(size_t)n * 3 + d];
      }
    const VtuMesh vmesh = {mesh.nen, &x, &mesh.conn, &mesh.subdomain, nullptr};   // the moved mesh, as Paraview_IO sees it
    if (!pvd->add(vmesh, names, vals, l)) die("cannot write the .vtu file");
  };
  update_pvd(0);
  for (int l = 1; l <= n_load; l++) {
    sm.pseudo_time += loading_step;
    printf(" ==== Step %4d out of %4d (pseudo-time=%g) ====\n", l, n_load, sm.pseudo_time);
    solid_load_step(sm, ksp, &press, &vm);
    if (out_points.count(l)) update_pvd((unsigned)l);
  }
  if (!sol_out.empty()) {   // final positions, then pressure and von Mises stress per element
    if (rdc_get_solution(sm.ctx, x.data())) die("rdc_get_solution");
    FILE* f = fopen(sol_out.c_str(), "wb");
    if (!f || fwrite(x.data(), sizeof(double), x.size(), f) != x.size() || fwrite(press.data(), sizeof(double), press.size(), f) != press.size() ||
        fwrite(vm.data(), sizeof(double), vm.size(), f) != vm.size())
      die("cannot write " + sol_out);
    fclose(f);
  }
  rdc_stats st;
  rdc_get_stats(sm.ctx, &st);
  printf("done: %d load steps, %lld kernel launches\n", n_load, (long long)st.kernel_launches);
  rdc_destroy(sm.ctx);
  return 0;
}

int main(int argc, char** argv) {
  std::string model_name = "adpm", in_path, sol_out;
  int ksp = RDC_KSP_BICGSTAB;
  bool vtu_binary = false, rdc_only = false;
  for (int a = 1; a < argc; a++) {
    if (!strcmp(argv[a], "-m") && a + 1 < argc) model_name = argv[++a];
    else if (!strncmp(argv[a], "ksp=", 4)) ksp = atoi(argv[a] + 4);
    else if (!strncmp(argv[a], "solution_out=", 13)) sol_out = argv[a] + 13;
    else if (!strcmp(argv[a], "vtu=binary")) vtu_binary = true;
    else if (!strcmp(argv[a], "vtu=ascii")) vtu_binary = false;
    else if (!strcmp(argv[a], "solid=off")) rdc_only = true;
    else in_path = argv[a];
  }
  if (in_path.empty())
    die("usage: rdc_driver -m adpm|pihna|ripf|proteas|coupled_hcc|solid <input.dat> [ksp=2] [vtu=binary] [solid=off] [solution_out=file]");
  if (model_name == "solid") {
    const size_t sl = in_path.find_last_of('/');
    return run_solid(Input(in_path), sl == std::string::npos ? std::string() : in_path.substr(0, sl + 1), ksp, vtu_binary, sol_out);
  }
  const int model = model_name == "adpm" ? RDC_ADPM : model_name == "pihna" ? RDC_PIHNA : model_name == "ripf" ? RDC_RIPF
                  : model_name == "proteas" ? RDC_PROTEAS : (model_name == "coupled_hcc" || model_name == "hcc") ? RDC_HCC : -1;
  if (model < 0) die("unknown model " + model_name + " (main.C:24-38 knows adpm, pihna, proteas, ripf; coupled_hcc.C is the fifth)");
  const int nv = rdc_model_nvars(model);
  const size_t slash = in_path.find_last_of('/');
  const std::string dir = slash == std::string::npos ? std::string() : in_path.substr(0, slash + 1);
  Input in(in_path);

  // ---- es.parameters, with the defaults of the model's input() --------------------------------------------------
  std::vector<double> p;
  if (model == RDC_ADPM) p = flat_params(kAdpmTable, in);
  else if (model == RDC_PIHNA) p = flat_params(kPihnaTable, in);
  else if (model == RDC_PROTEAS) p = flat_params(kProteasTable, in);
  else if (model == RDC_HCC) p = flat_params(kHccTable, in);
  else {
    p = flat_params(kRipfTable, in);
    if (!in.kv.count("volume_fraction/max_vacant")) p[RIPF_VF_MAX_VACANT] = 1.0 - p[RIPF_VF_MIN_VACANT];   // ripf.C:181-182
  }
  if ((int)p.size() != rdc_model_nparams(model)) die("parameter table out of date: run driver/gen_tables.py");
  // coupled_hcc.C:184-187 names its step count differently and defaults the step to 1
  const double dt = in.real("time_step", model == RDC_HCC ? 1.0 : 1.0e-9);
  const int n_steps = in.integer(model == RDC_HCC ? "number_of_time_steps" : "time_step_number", 1);
  const int out_step = in.integer("output_step", 0);
  // output steps: every output_step-th, else the integers of "output_time_points" (default: the last step), adpm.C /
  // proteas.C:144-163 / coupled_hcc.C
  std::set<int> out_points;
  if (out_step > 0) for (int t = out_step; t <= n_steps; t += out_step) out_points.insert(t);
  else out_points = integers_of(in.str("output_time_points", std::to_string(n_steps)));
  if (model == RDC_PROTEAS && in.integer("refinement_step", n_steps + 1) <= n_steps)
    die("refinement_step <= time_step_number asks for adaptive mesh refinement (proteas.C:82-83), which this driver does not do");
  // ---- mesh and initial fields ---------------------------------------------------------------------------------
  Mesh mesh = read_gmsh(dir + in.str("input_GMSH", "input.msh"));
  const int64_t N = (int64_t)mesh.xyz.size() / 3, E = (int64_t)mesh.conn.size() / mesh.nen;
  std::vector<double> u0 = read_table(dir + in.str("input_nodal", "input.nodal"), (size_t)N, nv);
  std::set<int> parcellation(mesh.subdomain.begin(), mesh.subdomain.end());   // adpm.C:302-310 (std::set: ascending)
  std::map<int, int> reg_of;
  for (int id : parcellation) reg_of.emplace(id, (int)reg_of.size());
  std::vector<int32_t> region((size_t)E);
  for (int64_t e = 0; e < E; e++) region[e] = reg_of[mesh.subdomain[e]];
  const int n_regions = model == RDC_ADPM ? (int)parcellation.size() : 1;   // only ADPM reports per region
  // coupled_hcc.C:39-74,117-132: the SolidSystem on the same mesh, solved at the loading time points; it moves the mesh
  const bool with_solid = model == RDC_HCC && !rdc_only;
  SolidModel solid;
  std::set<int> load_points;
  double loading_step = 0.0;
  if (with_solid) {
    solid = make_solid(in, mesh, dir);
    const int n_load = in.integer("number_of_loading_steps", 1);
    if (n_load < 1 || n_steps % n_load) die("number_of_time_steps must be a multiple of number_of_loading_steps (coupled_hcc.C:204-208)");
    loading_step = dt * n_steps / n_load;                                       // coupled_hcc.C:194-197
    for (int t = n_steps / n_load; t <= n_steps; t += n_steps / n_load) load_points.insert(t);
  }

  // ---- hand-over (once) -------------------------------------------------------------------------------------------
  rdc_ctx* ctx = nullptr;
  auto ck = [&](int rc, const char* what) {
    if (rc) die(std::string(what) + ": " + rdc_last_error(ctx));
  };
  ck(rdc_create(&ctx, model, mesh.nen, N, E, mesh.conn.data(), mesh.xyz.data(), nullptr, -1), "rdc_create");
  ck(rdc_set_params(ctx, p.data(), (int)p.size()), "rdc_set_params");
  if (model == RDC_ADPM) {
    std::vector<double> tracts = read_table(dir + in.str("input_elemental", "input.elemental"), (size_t)E, 3);
    ck(rdc_set_elem_field(ctx, 0, tracts.data(), 3), "rdc_set_elem_field");
  }
  if (model == RDC_RIPF) {
    std::vector<double> rt = read_table(dir + in.str("input_nodal_RT", "input.nodal~RT"), (size_t)N, 2);
    ck(rdc_set_nodal_field(ctx, 0, rt.data(), 2), "rdc_set_nodal_field");
  }
  std::vector<double> aux;   // PROTEAS: the AUX system (HU, RTD), proteas.C:37-41,218-268; also written to ParaView
  if (model == RDC_PROTEAS) {
    aux = read_table(dir + in.str("input_nodal_aux", "input_aux.nd"), (size_t)N, 2);
    ck(rdc_set_nodal_field(ctx, 0, aux.data(), 2), "rdc_set_nodal_field");
  }
  ck(rdc_set_solution(ctx, u0.data()), "rdc_set_solution");
  ck(rdc_set_subdomains(ctx, n_regions > 1 ? region.data() : nullptr, n_regions), "rdc_set_subdomains");
  if (model == RDC_RIPF) {   // ripf.C:50-53: check_solution once before the loop, at time 0
    ck(rdc_set_time(ctx, 0.0), "rdc_set_time");
    ck(rdc_set_dt(ctx, dt), "rdc_set_dt");
    ck(rdc_clamp(ctx), "rdc_clamp");
  }

  std::ofstream csv(dir + in.str("output_CSV", "output.csv"));
  csv.precision(17);
  auto cond1 = [&](int var, double div, double lo, double hi) {
    rdc_range_cond c;
    memset(&c, 0, sizeof(c));
    c.w[var] = 1.0; c.div = div; c.lo = lo; c.hi = hi;
    return c;
  };
  auto volume = [&](int ncond, const rdc_range_cond* c) {
    std::vector<double> v((size_t)n_regions);
    ck(rdc_region_volumes(ctx, ncond, c, v.data()), "rdc_region_volumes");
    return v;
  };
  auto save_solution = [&](double time) {
    if (model == RDC_ADPM) {   // adpm.C:690-829
      if (time == 0.0) {
        csv << "\"TIME\"";
        for (int id : parcellation) csv << ",\"CONCENTRATION__A_b__" << id << "\",\"CONCENTRATION__Tau__" << id << "\"";
        for (int id : parcellation) csv << ",\"VOLUME__A_b__" << id << "\",\"VOLUME__Tau__" << id << "\"";
        csv << std::endl;
      }
      std::vector<double> cA((size_t)n_regions), cT((size_t)n_regions);
      ck(rdc_region_last_mean(ctx, 1, cA.data()), "rdc_region_last_mean");
      ck(rdc_region_last_mean(ctx, 2, cT.data()), "rdc_region_last_mean");
      const rdc_range_cond a = cond1(1, 1.0, in.real("range/A_b/min", 1.0e-12), in.real("range/A_b/max", 1.0e+12));
      const rdc_range_cond b = cond1(2, 1.0, in.real("range/Tau/min", 1.0e-12), in.real("range/Tau/max", 1.0e+12));
      const std::vector<double> vA = volume(1, &a), vT = volume(1, &b);
      csv << time;
      for (int r = 0; r < n_regions; r++) csv << ',' << cA[r] << ',' << cT[r];
      for (int r = 0; r < n_regions; r++) csv << ',' << vA[r] << ',' << vT[r];
      csv << std::endl;
    } else if (model == RDC_PIHNA) {   // pihna.C:842-976
      if (time == 0.0)
        csv << "\"TIME\",\"DEGREES_OF_FREEDOM\",\"ACTIVE_TUMOR_VOLUME\",\"NECROTIC_VOLUME\",\"VASCULARITY_VOLUME\",\"TOTAL_CELL_VOLUME\"" << std::endl;
      rdc_range_cond act = cond1(1, 1.0, in.real("range/active_tumor/min", 1.0e-12), in.real("range/active_tumor/max", 1.0e+12));
      act.w[2] = 1.0;   // c + h
      const rdc_range_cond nec = cond1(0, 1.0, in.real("range/necrotic/min", 1.0e-12), in.real("range/necrotic/max", 1.0e+12));
      const rdc_range_cond vas = cond1(3, 1.0, in.real("range/vascularity/min", 1.0e-12), in.real("range/vascularity/max", 1.0e+12));
      rdc_range_cond tot = cond1(0, p[PIHNA_KAPPA_K], in.real("range/total_cell/min", 1.0e-12), in.real("range/total_cell/max", 1.0e+12));
      tot.w[1] = tot.w[2] = tot.w[3] = 1.0;   // (n + c + h + v) / Kappa_k
      csv << time << ',' << (long long)nv * N << ',' << volume(1, &act)[0] << ',' << volume(1, &nec)[0] << ',' << volume(1, &vas)[0]
          << ',' << volume(1, &tot)[0] << std::endl;
    } else if (model == RDC_PROTEAS || model == RDC_HCC) {
      // proteas.C:53-55 opens the CSV and never writes to it; coupled_hcc.C has none
    } else {   // ripf.C:777-864 (no header: it is commented out in the reference)
      const double HUmin = p[RIPF_HU_MIN], HUmax = p[RIPF_HU_MAX];
      rdc_range_cond cc[2] = {cond1(0, 1.0, in.real("range_cc/HU/min", HUmin), in.real("range_cc/HU/max", HUmax)),
                              cond1(1, 1.0, in.real("range_cc/min", 1.0e-9), HUGE_VAL)};
      rdc_range_cond fb[2] = {cond1(0, 1.0, in.real("range_fb/HU/min", HUmin), in.real("range_fb/HU/max", HUmax)),
                              cond1(2, 1.0, in.real("range_fb/min", 1.0e-9), HUGE_VAL)};
      csv << time << ',' << volume(2, cc)[0] << ',' << volume(2, fb)[0] << std::endl;
    }
  };
  // ParaView output (paraview.update_pvd, adpm.C:55,82): only when input.dat names output_PARAVIEW
  static const char* kVarNames[5][5] = {{"PrP", "A_b", "Tau", "", ""}, {"n", "c", "h", "v", "a"}, {"HU", "cc", "fb", "", ""},
                                        {"hos", "tum", "nec", "vsc", "oed"}, {"l", "c", "n", "", ""}};   // proteas.C:28-32, coupled_hcc.C:33-35
  std::vector<std::string> var_names;
  const int name_row = model == RDC_ADPM ? 0 : model == RDC_PIHNA ? 1 : model == RDC_RIPF ? 2 : model == RDC_PROTEAS ? 3 : 4;
  for (int a = 0; a < nv; a++) var_names.push_back(kVarNames[name_row][a]);
  const int n_extra = model == RDC_PROTEAS ? 2 : 0;   // Paraview_IO writes every system: PROTEAS_model, then AUX
  if (n_extra) { var_names.push_back("HU"); var_names.push_back("RTD"); }
  const VtuMesh vmesh = {mesh.nen, &mesh.xyz, &mesh.conn, &mesh.subdomain, nullptr};
  std::unique_ptr<PvdCollection> pvd;
  const char* pv_key = model == RDC_PROTEAS ? "output_Paraview" : "output_PARAVIEW";   // proteas.C:126 spells it differently
  if (in.kv.count(pv_key)) {
    pvd.reset(new PvdCollection(dir + in.str(pv_key, "output4paraview"), vtu_binary));
    if (!pvd->ok()) die("cannot open the .pvd collection");
  }
  std::vector<double> u_host((size_t)N * nv), u_out;
  auto update_pvd = [&](unsigned t) {
    if (!pvd) return;
    ck(rdc_get_solution(ctx, u_host.data()), "rdc_get_solution");   // the only device->host copy of the solution
    const std::vector<double>* vals = &u_host;
    if (n_extra) {   // append the AUX columns
      u_out.resize((size_t)N * (nv + n_extra));
      for (int64_t n = 0; n < N; n++) {
        for (int a = 0; a < nv; a++) u_out[(size_t)n * (nv + n_extra) + a] = u_host[(size_t)n * nv + a];
        for (int a = 0; a < n_extra; a++) u_out[(size_t)n * (nv + n_extra) + nv + a] = aux[(size_t)n * n_extra + a];
      }
      vals = &u_out;
    }
    if (!pvd->add(vmesh, var_names, *vals, t)) die("cannot write the .vtu file");
  };
  save_solution(0.0);   // adpm.C:54
  update_pvd(0);        // adpm.C:55

  // ---- the time loop, adpm.C:60-84 ----------------------------------------------------------------------------------
  double time = 0.0;
  long its_total = 0;
  for (int t = 1; t <= n_steps; t++) {
    time += dt;
    if (with_solid && load_points.count(t)) solid.pseudo_time += loading_step;   // coupled_hcc.C:100-101
    int its = 0;
    double res = 0;
    ck(rdc_step(ctx, time, dt, ksp, RDC_PC_JACOBI, 1e-12, 5000, 30, &its, &res), "rdc_step");
    its_total += its;
    printf(" ==== Step %4d out of %4d (Time=%9g) ==== its %d res %.3e\n", t, n_steps, time, its, res);
    if (with_solid && load_points.count(t)) {   // coupled_hcc.C:117-128: equilibrium, then the mesh both systems share has moved
      solid_load_step(solid, ksp, nullptr, nullptr);
      std::vector<double> x((size_t)N * 3);
      if (rdc_get_solution(solid.ctx, x.data())) die("rdc_get_solution(solid)");
      ck(rdc_update_coords(ctx, x.data()), "rdc_update_coords");
      mesh.xyz = x;                                                              // ParaView output shows the moved mesh
    }
    if (out_points.count(t)) { save_solution(time); update_pvd((unsigned)t); }
  }
  if (!sol_out.empty()) {
    std::vector<double> u((size_t)N * nv);
    ck(rdc_get_solution(ctx, u.data()), "rdc_get_solution");
    FILE* f = fopen(sol_out.c_str(), "wb");
    if (!f || fwrite(u.data(), sizeof(double), u.size(), f) != u.size()) die("cannot write " + sol_out);
    fclose(f);
  }
  rdc_stats st;
  rdc_get_stats(ctx, &st);
  printf("done: %d steps, %ld Krylov iterations, %lld kernel launches\n", n_steps, its_total, (long long)st.kernel_launches);
  rdc_destroy(ctx);
  if (with_solid) rdc_destroy(solid.ctx);
  return 0;
}
