// vtu_writer.h -- ParaView output of the stand-alone driver: one .vtu per output step plus a .pvd collection,
// in the layout rdcFEs' Paraview_IO produces (paraview.h:30-150 write_nodal_data, :158-198 open/update/close_pvd):
// Points "position"; PointData node_ID (1-based) and one Float64 array per variable; CellData element_ID (1-based),
// region_ID (subdomain id), processor_ID; Cells connectivity / offsets / types (VTK 10 = tetrahedron, 12 = hexahedron),
// so the visualization.pvsm states shipped with the reference's run directories open these files unchanged.
// Two encodings of the SAME arrays in the same order: ASCII (what paraview.h writes) and raw appended binary
// (format="appended", one <AppendedData encoding="raw"> block, UInt64 byte counts) -- ASCII is unusable at 10 M
// tets (~2 GB and minutes per step), the appended form is the arrays' bytes plus a 3 KB header.
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
                      const std::vector<double>& values, bool binary = false) {
  const size_t N = m.xyz->size() / 3, E = m.conn->size() / m.nen, nvar = names.size();
  if (values.size() != N * nvar) return false;
  std::ofstream f(path, binary ? std::ios::out | std::ios::binary : std::ios::out);
  if (!f) return false;
  f.precision(17);
  // the arrays in file order; the ASCII writer streams them, the binary writer lays them out behind the header
  std::vector<int32_t> node_id(N), elem_id(E), region(E), proc(E), offsets(E), types(E);
  for (size_t n = 0; n < N; n++) node_id[n] = (int32_t)(n + 1);
  for (size_t e = 0; e < E; e++) {
    elem_id[e] = (int32_t)(e + 1);
    region[e] = (*m.subdomain)[e];
    proc[e] = m.owner ? (*m.owner)[e] : 0;
    offsets[e] = (int32_t)((e + 1) * m.nen);
    types[e] = m.nen == 4 ? 10 : 12;
  }
  std::vector<std::vector<double>> var(nvar, std::vector<double>(N));
  for (size_t j = 0; j < nvar; j++)
    for (size_t n = 0; n < N; n++) {
      const double v = values[n * nvar + j];
      var[j][n] = fabs(v) <= 1.0e-300 ? 0.0 : v;
    }
  struct Blob { const void* p; uint64_t bytes; };
  std::vector<Blob> blobs;
  uint64_t offset = 0;
  auto array = [&](const char* type, const std::string& name, int ncomp, const void* data, size_t count, bool f64) {
    f << "        <DataArray type=\"" << type << "\" Name=\"" << name << "\" NumberOfComponents=\"" << ncomp << "\" format=\"";
    if (binary) {
      const uint64_t bytes = (uint64_t)count * (f64 ? 8 : 4);
      f << "appended\" offset=\"" << offset << "\"/>\n";
      blobs.push_back({data, bytes});
      offset += 8 + bytes;
    } else {
      f << "ascii\">\n";
      if (f64) for (size_t k = 0; k < count; k++) f << ' ' << static_cast<const double*>(data)[k];
      else for (size_t k = 0; k < count; k++) f << ' ' << static_cast<const int32_t*>(data)[k];
      f << "\n        </DataArray>\n";
    }
  };
  f << "<VTKFile type=\"UnstructuredGrid\" version=\"" << (binary ? "1.0" : "0.1") << "\" byte_order=\"LittleEndian\""
    << (binary ? " header_type=\"UInt64\"" : "") << ">\n  <UnstructuredGrid>\n";
  f << "    <Piece  NumberOfPoints=\"" << N << "\" NumberOfCells=\"" << E << "\">\n      <Points>\n";
  array("Float64", "position", 3, m.xyz->data(), 3 * N, true);
  f << "      </Points>\n      <PointData>\n";
  array("Int32", "node_ID", 1, node_id.data(), N, false);
  for (size_t j = 0; j < nvar; j++) array("Float64", names[j], 1, var[j].data(), N, true);
  f << "      </PointData>\n      <CellData>\n";
  array("Int32", "element_ID", 1, elem_id.data(), E, false);
  array("Int32", "region_ID", 1, region.data(), E, false);
  array("Int32", "processor_ID", 1, proc.data(), E, false);
  f << "      </CellData>\n      <Cells>\n";
  array("Int32", "connectivity", 1, m.conn->data(), E * m.nen, false);
  array("Int32", "offsets", 1, offsets.data(), E, false);
  array("Int32", "types", 1, types.data(), E, false);
  f << "      </Cells>\n    </Piece>\n  </UnstructuredGrid>\n";
  if (binary) {
    f << "  <AppendedData encoding=\"raw\">\n_";
    for (const Blob& b : blobs) {
      f.write(reinterpret_cast<const char*>(&b.bytes), 8);
      f.write(static_cast<const char*>(b.p), (std::streamsize)b.bytes);
    }
    f << "\n  </AppendedData>\n";
  }
  f << "</VTKFile>\n";
  return (bool)f;
}

// <base>.pvd collecting <base>-<t>.vtu, one DataSet line per output step (paraview.h:158-198)
class PvdCollection {
 public:
  explicit PvdCollection(const std::string& base, bool binary = false) : base_(base), binary_(binary), f_(base + ".pvd") {
    f_ << "<?xml version=\"1.0\"?>\n<VTKFile type=\"Collection\" version=\"0.1\" byte_order=\"LittleEndian\">\n  <Collection>\n";
  }
  ~PvdCollection() { f_ << "  </Collection>\n</VTKFile>\n"; }
  bool ok() const { return (bool)f_; }
  bool add(const VtuMesh& m, const std::vector<std::string>& names, const std::vector<double>& values, unsigned t) {
    const std::string vtu = base_ + "-" + std::to_string(t) + ".vtu";
    if (!write_vtu(vtu, m, names, values, binary_)) return false;
    const size_t slash = vtu.find_last_of('/');
    f_ << "    <DataSet timestep=\"" << t << "\" group=\"\" part=\"0\" file=\"" << (slash == std::string::npos ? vtu : vtu.substr(slash + 1))
       << "\"/>\n" << std::flush;
    return (bool)f_;
  }

 private:
  std::string base_;
  bool binary_;
  std::ofstream f_;
};
