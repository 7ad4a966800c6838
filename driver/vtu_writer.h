// vtu_writer.h -- ParaView output of the stand-alone driver: one ASCII .vtu per output step plus a .pvd collection,
// in the layout rdcFEs' Paraview_IO produces (paraview.h:30-150 write_nodal_data, :158-198 open/update/close_pvd):
// Points "position"; PointData node_ID (1-based) and one Float64 array per variable; CellData element_ID (1-based),
// region_ID (subdomain id), processor_ID; Cells connectivity / offsets / types (VTK 10 = tetrahedron, 12 = hexahedron),
// so the visualization.pvsm states shipped with the reference's run directories open these files unchanged.
#pragma once
#include <math.h>
#include <stdint.h>

#include <fstream>
#include <string>
#include <vector>

struct VtuMesh {
  int nen;                                // 4 or 8
  const std::vector<double>* xyz;         // [N*3]
  const std::vector<int32_t>* conn;       // [E*nen], 0-based
  const std::vector<int>* subdomain;      // [E]
  const std::vector<int32_t>* owner;      // [E] processor id per element, or nullptr (all 0)
};

// values: [N*nvar] node-major (the layout of rdc_get_solution); |v| <= tiny is written as 0 like paraview.h:112
inline bool write_vtu(const std::string& path, const VtuMesh& m, const std::vector<std::string>& names,
                      const std::vector<double>& values) {
  const size_t N = m.xyz->size() / 3, E = m.conn->size() / m.nen, nvar = names.size();
  if (values.size() != N * nvar) return false;
  std::ofstream f(path);
  if (!f) return false;
  f.precision(17);
  auto open_array = [&](const char* type, const std::string& name, int ncomp) {
    f << "        <DataArray type=\"" << type << "\" Name=\"" << name << "\" NumberOfComponents=\"" << ncomp << "\" format=\"ascii\">\n";
  };
  auto close_array = [&]() { f << "\n        </DataArray>\n"; };
  f << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n  <UnstructuredGrid>\n";
  f << "    <Piece  NumberOfPoints=\"" << N << "\" NumberOfCells=\"" << E << "\">\n      <Points>\n";
  open_array("Float64", "position", 3);
  for (size_t k = 0; k < 3 * N; k++) f << ' ' << (*m.xyz)[k];
  close_array();
  f << "      </Points>\n      <PointData>\n";
  open_array("Int32", "node_ID", 1);
  for (size_t n = 0; n < N; n++) f << ' ' << n + 1;
  close_array();
  for (size_t j = 0; j < nvar; j++) {
    open_array("Float64", names[j], 1);
    for (size_t n = 0; n < N; n++) {
      const double v = values[n * nvar + j];
      f << ' ' << (fabs(v) <= 1.0e-300 ? 0.0 : v);
    }
    close_array();
  }
  f << "      </PointData>\n      <CellData>\n";
  open_array("Int32", "element_ID", 1);
  for (size_t e = 0; e < E; e++) f << ' ' << e + 1;
  close_array();
  open_array("Int32", "region_ID", 1);
  for (size_t e = 0; e < E; e++) f << ' ' << (*m.subdomain)[e];
  close_array();
  open_array("Int32", "processor_ID", 1);
  for (size_t e = 0; e < E; e++) f << ' ' << (m.owner ? (*m.owner)[e] : 0);
  close_array();
  f << "      </CellData>\n      <Cells>\n";
  open_array("Int32", "connectivity", 1);
  for (size_t k = 0; k < E * m.nen; k++) f << ' ' << (*m.conn)[k];
  close_array();
  open_array("Int32", "offsets", 1);
  for (size_t e = 0; e < E; e++) f << ' ' << (e + 1) * m.nen;
  close_array();
  open_array("Int32", "types", 1);
  for (size_t e = 0; e < E; e++) f << ' ' << (m.nen == 4 ? 10 : 12);
  close_array();
  f << "      </Cells>\n    </Piece>\n  </UnstructuredGrid>\n</VTKFile>\n";
  return (bool)f;
}

// <base>.pvd collecting <base>-<t>.vtu, one DataSet line per output step (paraview.h:158-198)
class PvdCollection {
 public:
  explicit PvdCollection(const std::string& base) : base_(base), f_(base + ".pvd") {
    f_ << "<?xml version=\"1.0\"?>\n<VTKFile type=\"Collection\" version=\"0.1\" byte_order=\"LittleEndian\">\n  <Collection>\n";
  }
  ~PvdCollection() { f_ << "  </Collection>\n</VTKFile>\n"; }
  bool ok() const { return (bool)f_; }
  bool add(const VtuMesh& m, const std::vector<std::string>& names, const std::vector<double>& values, unsigned t) {
    const std::string vtu = base_ + "-" + std::to_string(t) + ".vtu";
    if (!write_vtu(vtu, m, names, values)) return false;
    const size_t slash = vtu.find_last_of('/');
    f_ << "    <DataSet timestep=\"" << t << "\" group=\"\" part=\"0\" file=\"" << (slash == std::string::npos ? vtu : vtu.substr(slash + 1))
       << "\"/>\n" << std::flush;
    return (bool)f_;
  }

 private:
  std::string base_;
  std::ofstream f_;
};
