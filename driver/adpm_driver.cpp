// adpm_driver.cpp -- stand-alone C++ driver of the ADPM model on top of the C ABI (include/rdc.h).
//
// Mirrors what rdcFEs' own driver does around the hot path, without libMesh: adpm() (adpm.C:15-87) reads input.dat
// with GetPot (input(), adpm.C:89-228), the Gmsh mesh (adpm.C:39), the nodal and elemental initial fields in node /
// element order (initial_adpm adpm.C:264-322, initial_tracts adpm.C:230-262), then loops
//     time += dt; rotate; solve; check_solution; every output step: save_solution      (adpm.C:60-84)
// and writes the per-region CSV of save_solution (adpm.C:690-829).  Here the loop body is rdc_step and the CSV comes
// from rdc_region_last_mean / rdc_region_volumes.  It is the C++ counterpart of rdcfes_b200/system.py: the same entry
// points, called from the reference's own language, and what tests/test_gpu_driver.py runs end to end.
//
//   adpm_driver <input.dat> [ksp=0|1|2] [solution_out=<file>]     (paths in input.dat are relative to its directory)
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "rdc.h"

static void die(const std::string& msg) {
  fprintf(stderr, "adpm_driver: %s\n", msg.c_str());
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

static std::vector<double> read_table(const std::string& path, size_t rows, int cols) {
  std::ifstream f(path);
  if (!f) die("cannot open " + path);
  std::vector<double> v(rows * cols);
  for (auto& x : v)
    if (!(f >> x)) die("short field file " + path);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 2) die("usage: adpm_driver <input.dat> [ksp=2] [solution_out=file]");
  const std::string in_path = argv[1];
  int ksp = RDC_KSP_BICGSTAB;
  std::string sol_out;
  for (int a = 2; a < argc; a++) {
    if (!strncmp(argv[a], "ksp=", 4)) ksp = atoi(argv[a] + 4);
    if (!strncmp(argv[a], "solution_out=", 13)) sol_out = argv[a] + 13;
  }
  const size_t slash = in_path.find_last_of('/');
  const std::string dir = slash == std::string::npos ? std::string() : in_path.substr(0, slash + 1);
  Input in(in_path);

  // ---- es.parameters, with the defaults of input() (adpm.C:130-226) -------------------------------------------
  std::vector<double> p(ADPM_NPARAMS, 0.0);
  auto pulse = [&](int at, const std::string& key) {
    p[at] = in.real(key, 0.);
    p[at + 1] = in.real(key + "/pulse/0", -1.0e-20);
    p[at + 2] = in.real(key + "/pulse/1", +1.0e+20);
  };
  auto sigmoid = [&](int at, const std::string& key) {
    p[at] = in.real(key, 0.);
    p[at + 1] = in.real(key + "/sigmoid/0", +1.0e+20);
    p[at + 2] = in.real(key + "/sigmoid/1", +1.1e+20);
  };
  auto trapezoid = [&](int at, const std::string& key) {
    p[at] = in.real(key, 0.);
    p[at + 1] = in.real(key + "/trapezoid/0", -1.1e-20);
    p[at + 2] = in.real(key + "/trapezoid/1", -1.0e-20);
    p[at + 3] = in.real(key + "/trapezoid/2", +1.0e+20);
    p[at + 4] = in.real(key + "/trapezoid/3", +1.1e+20);
  };
  p[ADPM_GAMMA] = in.real("decay/PrP/time_exponent", 0.);
  pulse(ADPM_DECAY_PRP, "decay/PrP");
  pulse(ADPM_DIFFUSE_AB, "diffuse/A_b"); pulse(ADPM_TAXIS1_AB, "taxis_1/A_b"); pulse(ADPM_TAXIS2_AB, "taxis_2/A_b");
  sigmoid(ADPM_PRODUCE_AB, "produce/A_b"); trapezoid(ADPM_TRANSFORM_AB, "transform/A_b"); pulse(ADPM_DECAY_AB, "decay/A_b");
  pulse(ADPM_DIFFUSE_TAU, "diffuse/Tau"); pulse(ADPM_TAXIS1_TAU, "taxis_1/Tau"); pulse(ADPM_TAXIS2_TAU, "taxis_2/Tau");
  sigmoid(ADPM_PRODUCE_TAU, "produce/Tau"); trapezoid(ADPM_TRANSFORM_TAU, "transform/Tau"); pulse(ADPM_DECAY_TAU, "decay/Tau");
  const double deg = M_PI / 180.0;   // degrees_to_radians, adpm.C:192,212
  p[ADPM_ANGLE_AB] = in.real("taxis/A_b/angle", 89.9) * deg;
  p[ADPM_ANGLE_TAU] = in.real("taxis/Tau/angle", 89.9) * deg;
  const double dt = in.real("time_step", 1.0e-9);
  const int n_steps = in.integer("time_step_number", 1);
  const int out_step = in.integer("output_step", 0);
  const double Ab_min = in.real("range/A_b/min", 1.0e-12), Ab_max = in.real("range/A_b/max", 1.0e+12);
  const double Tau_min = in.real("range/Tau/min", 1.0e-12), Tau_max = in.real("range/Tau/max", 1.0e+12);

  // ---- mesh and initial fields ---------------------------------------------------------------------------------
  Mesh mesh = read_gmsh(dir + in.str("input_GMSH", "input.msh"));
  const int64_t N = (int64_t)mesh.xyz.size() / 3, E = (int64_t)mesh.conn.size() / mesh.nen;
  std::vector<double> u0 = read_table(dir + in.str("input_nodal", "input.nodal"), (size_t)N, 3);
  std::vector<double> tracts = read_table(dir + in.str("input_elemental", "input.elemental"), (size_t)E, 3);
  std::set<int> parcellation(mesh.subdomain.begin(), mesh.subdomain.end());   // adpm.C:302-310 (std::set: ascending)
  std::map<int, int> reg_of;
  for (int id : parcellation) reg_of.emplace(id, (int)reg_of.size());
  std::vector<int32_t> region((size_t)E);
  for (int64_t e = 0; e < E; e++) region[e] = reg_of[mesh.subdomain[e]];
  const int n_regions = (int)parcellation.size();

  // ---- hand-over (once) -------------------------------------------------------------------------------------------
  rdc_ctx* ctx = nullptr;
  auto ck = [&](int rc, const char* what) {
    if (rc) die(std::string(what) + ": " + rdc_last_error(ctx));
  };
  ck(rdc_create(&ctx, RDC_ADPM, mesh.nen, N, E, mesh.conn.data(), mesh.xyz.data(), nullptr, -1), "rdc_create");
  ck(rdc_set_params(ctx, p.data(), (int)p.size()), "rdc_set_params");
  ck(rdc_set_elem_field(ctx, 0, tracts.data(), 3), "rdc_set_elem_field");
  ck(rdc_set_solution(ctx, u0.data()), "rdc_set_solution");
  ck(rdc_set_subdomains(ctx, region.data(), n_regions), "rdc_set_subdomains");

  std::ofstream csv(dir + in.str("output_CSV", "output.csv"));
  csv.precision(17);
  std::vector<double> cA((size_t)n_regions), cT((size_t)n_regions), vA((size_t)n_regions), vT((size_t)n_regions);
  auto save_solution = [&](double time) {   // adpm.C:690-829
    if (time == 0.0) {
      csv << "\"TIME\"";
      for (int id : parcellation) csv << ",\"CONCENTRATION__A_b__" << id << "\",\"CONCENTRATION__Tau__" << id << "\"";
      for (int id : parcellation) csv << ",\"VOLUME__A_b__" << id << "\",\"VOLUME__Tau__" << id << "\"";
      csv << std::endl;
    }
    rdc_range_cond cond;
    memset(&cond, 0, sizeof(cond));
    cond.div = 1.0;
    ck(rdc_region_last_mean(ctx, 1, cA.data()), "rdc_region_last_mean");
    ck(rdc_region_last_mean(ctx, 2, cT.data()), "rdc_region_last_mean");
    cond.w[1] = 1.0; cond.lo = Ab_min; cond.hi = Ab_max;
    ck(rdc_region_volumes(ctx, 1, &cond, vA.data()), "rdc_region_volumes");
    cond.w[1] = 0.0; cond.w[2] = 1.0; cond.lo = Tau_min; cond.hi = Tau_max;
    ck(rdc_region_volumes(ctx, 1, &cond, vT.data()), "rdc_region_volumes");
    csv << time;
    for (int r = 0; r < n_regions; r++) csv << ',' << cA[r] << ',' << cT[r];
    for (int r = 0; r < n_regions; r++) csv << ',' << vA[r] << ',' << vT[r];
    csv << std::endl;
  };
  save_solution(0.0);   // adpm.C:54

  // ---- the time loop, adpm.C:60-84 ----------------------------------------------------------------------------------
  double time = 0.0;
  long its_total = 0;
  for (int t = 1; t <= n_steps; t++) {
    time += dt;
    int its = 0;
    double res = 0;
    ck(rdc_step(ctx, time, dt, ksp, RDC_PC_JACOBI, 1e-12, 5000, 30, &its, &res), "rdc_step");
    its_total += its;
    printf(" ==== Step %4d out of %4d (Time=%9g) ==== its %d res %.3e\n", t, n_steps, time, its, res);
    const bool out = out_step ? (t % out_step == 0) : (t == in.integer("output_time_points", n_steps));
    if (out) save_solution(time);
  }
  if (!sol_out.empty()) {
    std::vector<double> u((size_t)N * 3);
    ck(rdc_get_solution(ctx, u.data()), "rdc_get_solution");
    FILE* f = fopen(sol_out.c_str(), "wb");
    if (!f || fwrite(u.data(), sizeof(double), u.size(), f) != u.size()) die("cannot write " + sol_out);
    fclose(f);
  }
  rdc_stats st;
  rdc_get_stats(ctx, &st);
  printf("done: %d steps, %ld Krylov iterations, %lld kernel launches\n", n_steps, its_total, (long long)st.kernel_launches);
  rdc_destroy(ctx);
  return 0;
}
