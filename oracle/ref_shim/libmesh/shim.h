// shim.h -- a serial, header-only stand-in for the slice of libMesh that the rdcFEs model files use.
//
// TEST INFRASTRUCTURE (oracle/): lets the reference's own src/{adpm,pihna,ripf,proteas,coupled_hcc}.C compile
// UNCHANGED, from where they lie under /root/reference, into oracle/_ref/libref.so (recipe: oracle/build_ref.py), so
// that the in-tree arithmetic of the assemble_* / check_solution / save_solution callbacks runs here without
// libMesh/PETSc/MPI.  What is upstream (not under /root/reference) is restated here, exactly like in the oracle:
// TET4/HEX8 first-order Lagrange tables, QGauss(THIRD), FEMap.  Nothing of the product
// path includes or links this.
//
// The class and member names are libMesh's public API because the unmodified reference sources call them; the
// bodies are written for this repo (serial, std::vector backed, no PETSc).
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <optional>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <tuple>
#include <vector>

#define LIBMESH_DIM 3
#define libmesh_assert(x) ((void)0)
#define libmesh_assert_equal_to(a, b) ((void)0)
#define libmesh_assert_less(a, b) ((void)0)
#define libmesh_assert_greater(a, b) ((void)0)
#define libmesh_assert_greater_equal(a, b) ((void)0)
#define libmesh_assert_less_equal(a, b) ((void)0)
#define libmesh_assert_not_equal_to(a, b) ((void)0)
#define libmesh_dbg_var(x)
#define libmesh_error() throw std::runtime_error(std::string("libmesh_error at ") + __FILE__ + ":" + std::to_string(__LINE__))
#define libmesh_error_msg(m) throw std::runtime_error(std::string("libmesh_error: ") + std::string(m))
#define libmesh_not_implemented() libmesh_error()
#define libmesh_make_unique std::make_unique
#define libmesh_nullptr nullptr
#define LOG_SCOPE(a, b) ((void)0)

namespace libMesh {

typedef double Real;
typedef double Number;
typedef uint32_t dof_id_type;
typedef uint16_t subdomain_id_type;
typedef uint16_t processor_id_type;
typedef int16_t boundary_id_type;
static const Real pi = 3.1415926535897932384626433832795029;
static const Real TOLERANCE = 1.e-6;
static std::ostream& out = std::cout;
static std::ostream& err = std::cerr;
template <class... A> inline void libmesh_ignore(const A&...) {}
template <class T, class U> inline T cast_ref(U& u) { return dynamic_cast<T>(u); }
template <class T, class U> inline T cast_int(U u) { return static_cast<T>(u); }
inline bool libmesh_isnan(double x) { return std::isnan(x); }
inline processor_id_type global_processor_id() { return 0; }
inline processor_id_type global_n_processors() { return 1; }

// ------------------------------------------------------------------------------------------- small vectors
template <class T>
class TypeVector {
 public:
  T c[3];
  TypeVector() : c{0, 0, 0} {}
  TypeVector(T x, T y = 0, T z = 0) : c{x, y, z} {}
  TypeVector(std::initializer_list<T> l) : c{0, 0, 0} { int k = 0; for (T v : l) if (k < 3) c[k++] = v; }
  T& operator()(unsigned i) { return c[i]; }
  const T& operator()(unsigned i) const { return c[i]; }
  T& slice(unsigned i) { return c[i]; }
  // [upstream] TypeVector::add_scaled: _coords[i] += factor * p(i)
  template <class T2> void add_scaled(const TypeVector<T2>& p, const T f) { for (int i = 0; i < 3; i++) c[i] += f * p.c[i]; }
  void add(const TypeVector& p) { for (int i = 0; i < 3; i++) c[i] += p.c[i]; }
  void subtract(const TypeVector& p) { for (int i = 0; i < 3; i++) c[i] -= p.c[i]; }
  TypeVector operator+(const TypeVector& p) const { return TypeVector(c[0] + p.c[0], c[1] + p.c[1], c[2] + p.c[2]); }
  TypeVector operator-(const TypeVector& p) const { return TypeVector(c[0] - p.c[0], c[1] - p.c[1], c[2] - p.c[2]); }
  TypeVector operator-() const { return TypeVector(-c[0], -c[1], -c[2]); }
  TypeVector& operator+=(const TypeVector& p) { add(p); return *this; }
  TypeVector& operator-=(const TypeVector& p) { subtract(p); return *this; }
  TypeVector operator*(const T f) const { return TypeVector(c[0] * f, c[1] * f, c[2] * f); }
  TypeVector operator/(const T f) const { return TypeVector(c[0] / f, c[1] / f, c[2] / f); }
  TypeVector& operator*=(const T f) { for (int i = 0; i < 3; i++) c[i] *= f; return *this; }
  TypeVector& operator/=(const T f) { for (int i = 0; i < 3; i++) c[i] /= f; return *this; }
  // [upstream] dot product: x*x' + y*y' + z*z', left to right
  T operator*(const TypeVector& p) const { return c[0] * p.c[0] + c[1] * p.c[1] + c[2] * p.c[2]; }
  T contract(const TypeVector& p) const { return (*this) * p; }
  TypeVector cross(const TypeVector& p) const {
    return TypeVector(c[1] * p.c[2] - c[2] * p.c[1], -c[0] * p.c[2] + c[2] * p.c[0], c[0] * p.c[1] - c[1] * p.c[0]);
  }
  T norm_sq() const { return c[0] * c[0] + c[1] * c[1] + c[2] * c[2]; }
  T norm() const { return std::sqrt(norm_sq()); }
  // [upstream] TypeVector::unit(): divides every component by the norm
  TypeVector unit() const { const T l = norm(); return TypeVector(c[0] / l, c[1] / l, c[2] / l); }
  void zero() { c[0] = c[1] = c[2] = 0; }
  bool operator==(const TypeVector& p) const { return c[0] == p.c[0] && c[1] == p.c[1] && c[2] == p.c[2]; }
  bool operator!=(const TypeVector& p) const { return !(*this == p); }
  bool is_zero() const { return c[0] == 0 && c[1] == 0 && c[2] == 0; }
  void print(std::ostream& os = libMesh::out) const { os << c[0] << ' ' << c[1] << ' ' << c[2]; }
};
template <class T> inline TypeVector<T> operator*(const T f, const TypeVector<T>& v) { return v * f; }
template <class T> inline std::ostream& operator<<(std::ostream& os, const TypeVector<T>& v) { v.print(os); return os; }
template <class T> using VectorValue = TypeVector<T>;
typedef TypeVector<Real> RealVectorValue;
typedef TypeVector<Real> RealGradient;
typedef TypeVector<Number> Gradient;
typedef TypeVector<Number> NumberVectorValue;
typedef TypeVector<Real> Point;

template <class T>
class TypeTensor {
 public:
  T c[9];
  TypeTensor() { for (T& v : c) v = 0; }
  TypeTensor(T xx, T xy = 0, T xz = 0, T yx = 0, T yy = 0, T yz = 0, T zx = 0, T zy = 0, T zz = 0) : c{xx, xy, xz, yx, yy, yz, zx, zy, zz} {}
  T& operator()(unsigned i, unsigned j) { return c[i * 3 + j]; }
  const T& operator()(unsigned i, unsigned j) const { return c[i * 3 + j]; }
  TypeTensor operator+(const TypeTensor& p) const { TypeTensor r; for (int i = 0; i < 9; i++) r.c[i] = c[i] + p.c[i]; return r; }
  TypeTensor operator-(const TypeTensor& p) const { TypeTensor r; for (int i = 0; i < 9; i++) r.c[i] = c[i] - p.c[i]; return r; }
  TypeTensor& operator+=(const TypeTensor& p) { for (int i = 0; i < 9; i++) c[i] += p.c[i]; return *this; }
  TypeTensor& operator-=(const TypeTensor& p) { for (int i = 0; i < 9; i++) c[i] -= p.c[i]; return *this; }
  TypeTensor operator*(const T f) const { TypeTensor r; for (int i = 0; i < 9; i++) r.c[i] = c[i] * f; return r; }
  TypeTensor operator/(const T f) const { TypeTensor r; for (int i = 0; i < 9; i++) r.c[i] = c[i] / f; return r; }
  TypeTensor& operator*=(const T f) { for (T& v : c) v *= f; return *this; }
  TypeTensor& operator/=(const T f) { for (T& v : c) v /= f; return *this; }
  TypeTensor operator*(const TypeTensor& p) const {
    TypeTensor r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) for (int k = 0; k < 3; k++) r(i, j) += (*this)(i, k) * p(k, j);
    return r;
  }
  TypeVector<T> operator*(const TypeVector<T>& v) const {
    TypeVector<T> r;
    for (int i = 0; i < 3; i++) r(i) = (*this)(i, 0) * v(0) + (*this)(i, 1) * v(1) + (*this)(i, 2) * v(2);
    return r;
  }
  TypeTensor transpose() const { TypeTensor r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = (*this)(j, i); return r; }
  T tr() const { return c[0] + c[4] + c[8]; }
  T det() const {
    return c[0] * (c[4] * c[8] - c[5] * c[7]) - c[1] * (c[3] * c[8] - c[5] * c[6]) + c[2] * (c[3] * c[7] - c[4] * c[6]);
  }
  T contract(const TypeTensor& p) const { T s = 0; for (int i = 0; i < 9; i++) s += c[i] * p.c[i]; return s; }
  T norm_sq() const { return contract(*this); }
  T norm() const { return std::sqrt(norm_sq()); }
  TypeTensor inverse() const;
  void zero() { for (T& v : c) v = 0; }
};
template <class T> inline TypeTensor<T> operator*(const T f, const TypeTensor<T>& v) { return v * f; }
template <class T> inline TypeVector<T> operator*(const TypeVector<T>& a, const TypeTensor<T>& m) {   // row vector * matrix
  TypeVector<T> r;
  for (int j = 0; j < 3; j++) r(j) = a(0) * m(0, j) + a(1) * m(1, j) + a(2) * m(2, j);
  return r;
}
template <class T> using TensorValue = TypeTensor<T>;
typedef TypeTensor<Real> RealTensorValue;
typedef TypeTensor<Real> RealTensor;
typedef TypeTensor<Number> Tensor;
template <class T> TypeTensor<T> TypeTensor<T>::inverse() const {
  const T d = det();
  TypeTensor r;
  r(0, 0) = (c[4] * c[8] - c[5] * c[7]) / d; r(0, 1) = -(c[1] * c[8] - c[2] * c[7]) / d; r(0, 2) = (c[1] * c[5] - c[2] * c[4]) / d;
  r(1, 0) = -(c[3] * c[8] - c[5] * c[6]) / d; r(1, 1) = (c[0] * c[8] - c[2] * c[6]) / d; r(1, 2) = -(c[0] * c[5] - c[2] * c[3]) / d;
  r(2, 0) = (c[3] * c[7] - c[4] * c[6]) / d; r(2, 1) = -(c[0] * c[7] - c[1] * c[6]) / d; r(2, 2) = (c[0] * c[4] - c[1] * c[3]) / d;
  return r;
}

// ------------------------------------------------------------------------------------------- dense algebra
template <class T>
class DenseVector {
 public:
  std::vector<T> v;
  DenseVector(unsigned n = 0) : v(n, T(0)) {}
  void resize(unsigned n) { v.assign(n, T(0)); }
  unsigned size() const { return (unsigned)v.size(); }
  T& operator()(unsigned i) { return v[i]; }
  const T& operator()(unsigned i) const { return v[i]; }
  T el(unsigned i) const { return v[i]; }
  void zero() { std::fill(v.begin(), v.end(), T(0)); }
  std::vector<T>& get_values() { return v; }
  const std::vector<T>& get_values() const { return v; }
  T l2_norm() const { T s = 0; for (T x : v) s += x * x; return std::sqrt(s); }
  DenseVector& operator=(std::initializer_list<T> l) { v.assign(l.begin(), l.end()); return *this; }
  DenseVector& operator+=(const DenseVector& o) { for (size_t i = 0; i < v.size(); i++) v[i] += o.v[i]; return *this; }
  DenseVector& operator*=(T f) { for (T& x : v) x *= f; return *this; }
  void scale(T f) { for (T& x : v) x *= f; }
  void add(T f, const DenseVector& o) { for (size_t i = 0; i < v.size(); i++) v[i] += f * o.v[i]; }
};
template <class T>
class DenseMatrix {
 public:
  unsigned nr, nc;
  std::vector<T> a;  // row major
  DenseMatrix(unsigned m = 0, unsigned n = 0) : nr(m), nc(n), a((size_t)m * n, T(0)) {}
  void resize(unsigned m, unsigned n) { nr = m; nc = n; a.assign((size_t)m * n, T(0)); }
  unsigned m() const { return nr; }
  unsigned n() const { return nc; }
  T& operator()(unsigned i, unsigned j) { return a[(size_t)i * nc + j]; }
  const T& operator()(unsigned i, unsigned j) const { return a[(size_t)i * nc + j]; }
  T el(unsigned i, unsigned j) const { return (*this)(i, j); }
  void zero() { std::fill(a.begin(), a.end(), T(0)); }
  DenseMatrix& operator+=(const DenseMatrix& o) { for (size_t i = 0; i < a.size(); i++) a[i] += o.a[i]; return *this; }
  DenseMatrix& operator*=(T f) { for (T& x : a) x *= f; return *this; }
  void scale(T f) { for (T& x : a) x *= f; }
  void add(T f, const DenseMatrix& o) { for (size_t i = 0; i < a.size(); i++) a[i] += f * o.a[i]; }
  void right_multiply(const DenseMatrix& B) {   // this <- this * B
    DenseMatrix r(nr, B.nc);
    for (unsigned i = 0; i < nr; i++) for (unsigned j = 0; j < B.nc; j++) for (unsigned k = 0; k < nc; k++) r(i, j) += (*this)(i, k) * B(k, j);
    *this = r;
  }
  void right_multiply_transpose(const DenseMatrix& B) {   // this <- this * B^T
    DenseMatrix r(nr, B.nr);
    for (unsigned i = 0; i < nr; i++) for (unsigned j = 0; j < B.nr; j++) for (unsigned k = 0; k < nc; k++) r(i, j) += (*this)(i, k) * B(j, k);
    *this = r;
  }
  void vector_mult(DenseVector<T>& d, const DenseVector<T>& x) const {
    d.resize(nr);
    for (unsigned i = 0; i < nr; i++) for (unsigned j = 0; j < nc; j++) d(i) += (*this)(i, j) * x(j);
  }
};
template <class T>
class DenseSubVector {
 public:
  DenseVector<T>* p;
  unsigned off = 0, len = 0;
  DenseSubVector(DenseVector<T>& parent, unsigned ioff = 0, unsigned n = 0) : p(&parent), off(ioff), len(n) {}
  void reposition(unsigned ioff, unsigned n) { off = ioff; len = n; }
  unsigned size() const { return len; }
  T& operator()(unsigned i) { return (*p)(off + i); }
  const T& operator()(unsigned i) const { return (*p)(off + i); }
  DenseVector<T>& parent() { return *p; }
  void zero() { for (unsigned i = 0; i < len; i++) (*p)(off + i) = 0; }
};
template <class T>
class DenseSubMatrix {
 public:
  DenseMatrix<T>* p;
  unsigned io = 0, jo = 0, nr = 0, nc = 0;
  DenseSubMatrix(DenseMatrix<T>& parent, unsigned ioff = 0, unsigned joff = 0, unsigned m = 0, unsigned n = 0)
      : p(&parent), io(ioff), jo(joff), nr(m), nc(n) {}
  void reposition(unsigned ioff, unsigned joff, unsigned m, unsigned n) { io = ioff; jo = joff; nr = m; nc = n; }
  unsigned m() const { return nr; }
  unsigned n() const { return nc; }
  T& operator()(unsigned i, unsigned j) { return (*p)(io + i, jo + j); }
  const T& operator()(unsigned i, unsigned j) const { return (*p)(io + i, jo + j); }
  DenseMatrix<T>& parent() { return *p; }
  void zero() { for (unsigned i = 0; i < nr; i++) for (unsigned j = 0; j < nc; j++) (*p)(io + i, jo + j) = 0; }
};

// ------------------------------------------------------------------------------------------- enums
enum Order { CONSTANT = 0, FIRST = 1, SECOND = 2, THIRD = 3, FOURTH = 4, FIFTH = 5, SIXTH = 6, SEVENTH = 7 };
enum FEFamily { LAGRANGE = 0, HIERARCHIC = 1, MONOMIAL = 2, L2_LAGRANGE = 6, XYZ = 5, SCALAR = 31 };
enum ElemType { EDGE2 = 0, EDGE3, EDGE4, TRI3, TRI6, QUAD4, QUAD8, QUAD9, TET4, TET10, HEX8, HEX20, HEX27, PRISM6, PRISM15,
                PRISM18, PYRAMID5, PYRAMID13, PYRAMID14, INVALID_ELEM };
enum IOPackage { TECPLOT, VTK, UCD, UNV, DIVA, INVALID_IO_PACKAGE };
enum SolverPackage { PETSC_SOLVERS = 0, TRILINOS_SOLVERS, LASPACK_SOLVERS, SLEPC_SOLVERS, EIGEN_SOLVERS, NOX_SOLVERS, INVALID_SOLVER_PACKAGE };
enum ParallelType { AUTOMATIC = 0, SERIAL, PARALLEL, GHOSTED, INVALID_PARALLELIZATION };
enum XdrMODE { UNKNOWN = -1, ENCODE = 0, DECODE, WRITE, READ };
namespace Utility {
template <class T> inline T string_to_enum(const std::string&) { return T(); }
template <class T> inline std::string enum_to_string(const T) { return std::string(); }
}  // namespace Utility

struct FEType {
  Order order;
  FEFamily family;
  FEType(Order o = FIRST, FEFamily f = LAGRANGE) : order(o), family(f) {}
  FEType(int o, FEFamily f = LAGRANGE) : order((Order)o), family(f) {}
  // [upstream] 2*p + 1
  Order default_quadrature_order() const { return (Order)(2 * (int)order + 1); }
};

// ------------------------------------------------------------------------------------------- parallel (serial)
namespace Parallel {
class Communicator {
 public:
  processor_id_type rank() const { return 0; }
  processor_id_type size() const { return 1; }
  void barrier() const {}
  template <class T> void max(T&) const {}
  template <class T> void min(T&) const {}
  template <class T> void sum(T&) const {}
  template <class T> void broadcast(T&, unsigned = 0) const {}
  template <class T> void allgather(T&) const {}
};
}  // namespace Parallel
class LibMeshInit {
 public:
  Parallel::Communicator c_;
  LibMeshInit() {}
  LibMeshInit(int, char**) {}
  const Parallel::Communicator& comm() const { return c_; }
  Parallel::Communicator& comm() { return c_; }
};
class PerfLog {
 public:
  PerfLog(const std::string& = "", bool = true) {}
  void push(const std::string&, const std::string& = "") {}
  void pop(const std::string&, const std::string& = "") {}
  void print_log() const {}
  void clear() {}
};
inline unsigned n_threads() { return 1; }

// ------------------------------------------------------------------------------------------- GetPot (subset)
// key = value lines ('#' comments, optional quotes), [section] prefixes are not used by the rdcFEs inputs
class GetPot {
  std::map<std::string, std::string> kv;

 public:
  GetPot() {}
  GetPot(const std::string& file) { parse_input_file(file); }
  GetPot(int, char**) {}
  void parse_input_file(const std::string& file) {
    std::ifstream f(file);
    std::string line;
    while (std::getline(f, line)) {
      const size_t h = line.find('#');
      if (h != std::string::npos) line.erase(h);
      const size_t e = line.find('=');
      if (e == std::string::npos) continue;
      auto trim = [](std::string s) {
        const char* ws = " \t\r\n'\"";
        const size_t a = s.find_first_not_of(ws), b = s.find_last_not_of(ws);
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
      };
      kv[trim(line.substr(0, e))] = trim(line.substr(e + 1));
    }
  }
  void set(const std::string& k, const std::string& v) { kv[k] = v; }
  bool have_variable(const std::string& k) const { return kv.count(k) != 0; }
  bool search(const char* k) const { return kv.count(k) != 0; }
  unsigned vector_variable_size(const std::string& k) const {
    auto it = kv.find(k);
    if (it == kv.end()) return 0;
    std::istringstream ss(it->second);
    std::string t;
    unsigned n = 0;
    while (ss >> t) n++;
    return n;
  }
  template <class T> T get(const std::string& k, const T& dflt, unsigned idx = 0) const {
    auto it = kv.find(k);
    if (it == kv.end()) return dflt;
    std::istringstream ss(it->second);
    T v = dflt;
    for (unsigned i = 0; i <= idx; i++) if (!(ss >> v)) return dflt;
    return v;
  }
  double operator()(const std::string& k, double d) const { return get<double>(k, d); }
  double operator()(const std::string& k, double d, unsigned i) const { return get<double>(k, d, i); }
  int operator()(const std::string& k, int d) const { return get<int>(k, d); }
  int operator()(const std::string& k, int d, unsigned i) const { return get<int>(k, d, i); }
  unsigned operator()(const std::string& k, unsigned d) const { return get<unsigned>(k, d); }
  bool operator()(const std::string& k, bool d) const {
    auto it = kv.find(k);
    if (it == kv.end()) return d;
    return it->second == "true" || it->second == "1" || it->second == "TRUE";
  }
  std::string operator()(const std::string& k, const std::string& d) const {
    auto it = kv.find(k);
    return it == kv.end() ? d : it->second;
  }
  std::string operator()(const std::string& k, const char* d) const { return (*this)(k, std::string(d)); }
  std::string operator()(const std::string& k, const char* d, unsigned i) const { return get<std::string>(k, std::string(d), i); }
};

// ------------------------------------------------------------------------------------------- Parameters
class Parameters {
  struct Base { virtual ~Base() {} };
  template <class T> struct Holder : Base { T v{}; };
  std::map<std::string, std::unique_ptr<Base>> m;

 public:
  template <class T> T& set(const std::string& k) {
    auto it = m.find(k);
    if (it == m.end() || !dynamic_cast<Holder<T>*>(it->second.get())) {
      m[k] = std::make_unique<Holder<T>>();
      it = m.find(k);
    }
    return static_cast<Holder<T>*>(it->second.get())->v;
  }
  template <class T> const T& get(const std::string& k) const {
    auto it = m.find(k);
    if (it == m.end()) throw std::runtime_error("Parameters::get: no parameter named " + k);
    auto* h = dynamic_cast<Holder<T>*>(it->second.get());
    if (!h) throw std::runtime_error("Parameters::get: wrong type for " + k);
    return h->v;
  }
  template <class T> bool have_parameter(const std::string& k) const {
    auto it = m.find(k);
    return it != m.end() && dynamic_cast<Holder<T>*>(it->second.get());
  }
};

// ------------------------------------------------------------------------------------------- mesh
class Elem;
class Node : public Point {
 public:
  dof_id_type id_ = 0;
  std::vector<dof_id_type> sys_base;   // per system: global dof of variable 0 at this node (variables are contiguous)
  std::vector<unsigned> sys_nvars;     // per system: number of nodal variables (0: none)
  Node() {}
  Node(const Point& p, dof_id_type i) : Point(p), id_(i) {}
  dof_id_type id() const { return id_; }
  processor_id_type processor_id() const { return 0; }
  // [upstream] variables of one variable group are numbered contiguously at a node (Appendix B-5)
  dof_id_type dof_number(unsigned s, unsigned var, unsigned /*comp*/) const { return sys_base.at(s) + var; }
  unsigned n_comp(unsigned s, unsigned var) const { return (s < sys_nvars.size() && var < sys_nvars[s]) ? 1u : 0u; }
  unsigned n_dofs(unsigned s, unsigned var = 0) const { return n_comp(s, var); }
  Node& operator=(const Point& p) { Point::operator=(p); return *this; }
};

struct IntRange {
  unsigned a, b;
  struct It { unsigned v; unsigned operator*() const { return v; } It& operator++() { ++v; return *this; } bool operator!=(const It& o) const { return v != o.v; } };
  It begin() const { return It{a}; }
  It end() const { return It{b}; }
};

class Elem {
 public:
  ElemType type_ = TET4;
  dof_id_type id_ = 0;
  subdomain_id_type sbd_ = 0;
  std::vector<Node*> nodes;
  ElemType type() const { return type_; }
  dof_id_type id() const { return id_; }
  unsigned dim() const { return 3; }
  subdomain_id_type subdomain_id() const { return sbd_; }
  subdomain_id_type& subdomain_id() { return sbd_; }
  processor_id_type processor_id() const { return 0; }
  unsigned n_nodes() const { return (unsigned)nodes.size(); }
  unsigned n_sides() const { return type_ == TET4 ? 4u : 6u; }
  unsigned n_vertices() const { return n_nodes(); }
  dof_id_type node_id(unsigned i) const { return nodes[i]->id(); }
  const Node* node_ptr(unsigned i) const { return nodes[i]; }
  Node* node_ptr(unsigned i) { return nodes[i]; }
  const Node& node_ref(unsigned i) const { return *nodes[i]; }
  Node& node_ref(unsigned i) { return *nodes[i]; }
  const Point& point(unsigned i) const { return *nodes[i]; }
  Point& point(unsigned i) { return *nodes[i]; }
  const Node* const* get_nodes() const { return nodes.data(); }
  IntRange side_index_range() const { return IntRange{0, n_sides()}; }
  IntRange node_index_range() const { return IntRange{0, n_nodes()}; }
  const Elem* neighbor_ptr(unsigned) const { return nullptr; }   // boundary topology is not modelled (only dead code asks)
  bool active() const { return true; }
  unsigned level() const { return 0; }
  const Elem* parent() const { return nullptr; }
  Point centroid() const {
    Point c;
    for (const Node* n : nodes) c.add(*n);
    return c / (Real)nodes.size();
  }
  Point vertex_average() const { return centroid(); }
  Real hmax() const {
    Real h = 0;
    for (size_t a = 0; a < nodes.size(); a++) for (size_t b = a + 1; b < nodes.size(); b++) h = std::max(h, (*nodes[a] - *nodes[b]).norm());
    return h;
  }
  Real hmin() const {
    Real h = 1e300;
    for (size_t a = 0; a < nodes.size(); a++) for (size_t b = a + 1; b < nodes.size(); b++) h = std::min(h, (*nodes[a] - *nodes[b]).norm());
    return h;
  }
  void connectivity(unsigned, IOPackage, std::vector<unsigned>& c) const {
    c.clear();
    for (const Node* n : nodes) c.push_back(n->id());
  }
  Real volume() const;   // below (needs the quadrature)
};

template <class P>
struct PtrRange {
  const std::vector<P>* v;
  typename std::vector<P>::const_iterator begin() const { return v->begin(); }
  typename std::vector<P>::const_iterator end() const { return v->end(); }
};

class BoundaryInfo {
 public:
  // [upstream] BoundaryInfo: (element, side) -> boundary ids; filled by GmshIO from the lower-dimensional elements of the file
  // (here: by the wrapper through add_side).  Only SolidSystem::side_time_derivative (solid_system.C:296) asks.
  std::set<std::tuple<dof_id_type, unsigned short, boundary_id_type>> sides_;
  std::set<boundary_id_type> ids_;
  bool has_boundary_id(const Elem* e, unsigned short s, boundary_id_type id) const;
  boundary_id_type boundary_id(const Elem*, unsigned short) const { return -1; }
  void boundary_ids(const Elem*, unsigned short, std::vector<boundary_id_type>& v) const { v.clear(); }
  std::size_t n_boundary_conds() const { return sides_.size(); }
  void add_side(const Elem* e, unsigned short s, boundary_id_type id);
  const std::set<boundary_id_type>& get_boundary_ids() const { return ids_; }
};
inline bool BoundaryInfo::has_boundary_id(const Elem* e, unsigned short s, boundary_id_type id) const {
  return sides_.count(std::make_tuple(e->id(), s, id)) != 0;
}
inline void BoundaryInfo::add_side(const Elem* e, unsigned short s, boundary_id_type id) {
  sides_.insert(std::make_tuple(e->id(), s, id));
  ids_.insert(id);
}

class MeshBase {
 public:
  std::vector<Node*> nodes_;
  std::vector<Elem*> elems_;
  BoundaryInfo binfo_;
  Parallel::Communicator comm_;
  typedef std::vector<Node*>::const_iterator const_node_iterator;
  typedef std::vector<Elem*>::const_iterator const_element_iterator;
  typedef std::vector<Node*>::const_iterator node_iterator;
  typedef std::vector<Elem*>::const_iterator element_iterator;
  MeshBase() {}
  MeshBase(const Parallel::Communicator&, unsigned char = 3) {}
  MeshBase(const MeshBase&) = delete;
  virtual ~MeshBase() {
    for (Node* n : nodes_) delete n;
    for (Elem* e : elems_) delete e;
  }
  unsigned mesh_dimension() const { return 3; }
  unsigned spatial_dimension() const { return 3; }
  dof_id_type n_nodes() const { return (dof_id_type)nodes_.size(); }
  dof_id_type n_elem() const { return (dof_id_type)elems_.size(); }
  dof_id_type n_active_elem() const { return n_elem(); }
  dof_id_type n_local_nodes() const { return n_nodes(); }
  dof_id_type max_node_id() const { return n_nodes(); }
  dof_id_type max_elem_id() const { return n_elem(); }
  processor_id_type processor_id() const { return 0; }
  processor_id_type n_processors() const { return 1; }
  const Parallel::Communicator& comm() const { return comm_; }
  PtrRange<Elem*> active_local_element_ptr_range() const { return PtrRange<Elem*>{&elems_}; }
  PtrRange<Elem*> active_element_ptr_range() const { return PtrRange<Elem*>{&elems_}; }
  PtrRange<Elem*> element_ptr_range() const { return PtrRange<Elem*>{&elems_}; }
  PtrRange<Node*> node_ptr_range() const { return PtrRange<Node*>{&nodes_}; }
  PtrRange<Node*> local_node_ptr_range() const { return PtrRange<Node*>{&nodes_}; }
  const_node_iterator nodes_begin() const { return nodes_.begin(); }
  const_node_iterator nodes_end() const { return nodes_.end(); }
  const_node_iterator local_nodes_begin() const { return nodes_.begin(); }
  const_node_iterator local_nodes_end() const { return nodes_.end(); }
  const_element_iterator active_elements_begin() const { return elems_.begin(); }
  const_element_iterator active_elements_end() const { return elems_.end(); }
  const_element_iterator active_local_elements_begin() const { return elems_.begin(); }
  const_element_iterator active_local_elements_end() const { return elems_.end(); }
  const_element_iterator elements_begin() const { return elems_.begin(); }
  const_element_iterator elements_end() const { return elems_.end(); }
  const Node* node_ptr(dof_id_type i) const { return nodes_[i]; }
  Node* node_ptr(dof_id_type i) { return nodes_[i]; }
  const Node& node_ref(dof_id_type i) const { return *nodes_[i]; }
  Node& node_ref(dof_id_type i) { return *nodes_[i]; }
  const Point& point(dof_id_type i) const { return *nodes_[i]; }
  const Elem* elem_ptr(dof_id_type i) const { return elems_[i]; }
  Elem* elem_ptr(dof_id_type i) { return elems_[i]; }
  const Elem& elem_ref(dof_id_type i) const { return *elems_[i]; }
  const BoundaryInfo& get_boundary_info() const { return binfo_; }
  BoundaryInfo& get_boundary_info() { return binfo_; }
  void prepare_for_use(bool = false, bool = false) {}
  void print_info(std::ostream& = libMesh::out) const {}
  void allow_renumbering(bool) {}
  void all_second_order(bool = true) {}
  void read(const std::string&) {}
  void write(const std::string&) {}
  void clear() {}
  // build from flat arrays (the wrapper's entry point)
  void build(ElemType t, int64_t N, int64_t E, const int32_t* conn, const double* xyz, const int32_t* subdomain) {
    const int nen = t == TET4 ? 4 : 8;
    for (int64_t i = 0; i < N; i++) nodes_.push_back(new Node(Point(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), (dof_id_type)i));
    for (int64_t e = 0; e < E; e++) {
      Elem* el = new Elem();
      el->type_ = t; el->id_ = (dof_id_type)e; el->sbd_ = subdomain ? (subdomain_id_type)subdomain[e] : 0;
      for (int k = 0; k < nen; k++) el->nodes.push_back(nodes_[conn[e * nen + k]]);
      elems_.push_back(el);
    }
  }
};
class UnstructuredMesh : public MeshBase { public: using MeshBase::MeshBase; };
class ReplicatedMesh : public UnstructuredMesh { public: using UnstructuredMesh::UnstructuredMesh; };
class Mesh : public ReplicatedMesh { public: using ReplicatedMesh::ReplicatedMesh; };
typedef Mesh SerialMesh;

template <class MT> class MeshInput {
 public:
  MT* mesh_;
  MeshInput(MT& m) : mesh_(&m) {}
  virtual ~MeshInput() {}
  MT& mesh() { return *mesh_; }
};
class EquationSystems;
template <class MT> class MeshOutput {
 public:
  const MT* mesh_;
  MeshOutput(const MT& m) : mesh_(&m) {}
  virtual ~MeshOutput() {}
  const MT& mesh() const { return *mesh_; }
  virtual void write_equation_systems(const std::string&, const EquationSystems&, const std::set<std::string>* = nullptr) {}
  virtual void write_nodal_data(const std::string&, const std::vector<Number>&, const std::vector<std::string>&) {}
};
class GmshIO : public MeshInput<MeshBase>, public MeshOutput<MeshBase> {
 public:
  GmshIO(MeshBase& m) : MeshInput<MeshBase>(m), MeshOutput<MeshBase>(m) {}
  void read(const std::string&) {}
  void write(const std::string&) {}
  bool& binary() { static bool b = false; return b; }
};
class ExodusII_IO : public MeshInput<MeshBase>, public MeshOutput<MeshBase> {
 public:
  ExodusII_IO(MeshBase& m) : MeshInput<MeshBase>(m), MeshOutput<MeshBase>(m) {}
  void read(const std::string&) {}
  void write(const std::string&) {}
  void write_timestep(const std::string&, const EquationSystems&, int, Real) {}
  void append(bool) {}
};

// ------------------------------------------------------------------------------------------- numeric vector / matrix
template <class T>
class NumericVector {
 public:
  std::vector<T> v;
  NumericVector(size_t n = 0) : v(n, T(0)) {}
  virtual ~NumericVector() {}
  void init(size_t n, bool = false) { v.assign(n, T(0)); }
  dof_id_type size() const { return (dof_id_type)v.size(); }
  dof_id_type local_size() const { return size(); }
  dof_id_type first_local_index() const { return 0; }
  dof_id_type last_local_index() const { return size(); }
  T el(dof_id_type i) const { return v[i]; }
  T operator()(dof_id_type i) const { return v[i]; }
  void set(dof_id_type i, T x) { v[i] = x; }
  void add(dof_id_type i, T x) { v[i] += x; }
  void add(T a, const NumericVector& o) { for (size_t i = 0; i < v.size(); i++) v[i] += a * o.v[i]; }
  void add_vector(const DenseVector<T>& d, const std::vector<dof_id_type>& idx) { for (size_t i = 0; i < idx.size(); i++) v[idx[i]] += d(i); }
  void add_vector(const T* d, const std::vector<dof_id_type>& idx) { for (size_t i = 0; i < idx.size(); i++) v[idx[i]] += d[i]; }
  void insert(const std::vector<T>& d, const std::vector<dof_id_type>& idx) { for (size_t i = 0; i < idx.size(); i++) v[idx[i]] = d[i]; }
  void get(const std::vector<dof_id_type>& idx, std::vector<T>& o) const { o.resize(idx.size()); for (size_t i = 0; i < idx.size(); i++) o[i] = v[idx[i]]; }
  void get(const std::vector<dof_id_type>& idx, T* o) const { for (size_t i = 0; i < idx.size(); i++) o[i] = v[idx[i]]; }
  void close() {}
  bool closed() const { return true; }
  void zero() { std::fill(v.begin(), v.end(), T(0)); }
  void scale(T a) { for (T& x : v) x *= a; }
  void localize(std::vector<T>& o) const { o = v; }
  void localize(NumericVector& o) const { o.v = v; }
  void localize(std::vector<T>& o, const std::vector<dof_id_type>&) const { o = v; }
  NumericVector& operator=(const NumericVector& o) { v = o.v; return *this; }
  NumericVector& operator=(const std::vector<T>& o) { v = o; return *this; }
  NumericVector& operator=(T a) { std::fill(v.begin(), v.end(), a); return *this; }
  NumericVector& operator+=(const NumericVector& o) { for (size_t i = 0; i < v.size(); i++) v[i] += o.v[i]; return *this; }
  NumericVector& operator-=(const NumericVector& o) { for (size_t i = 0; i < v.size(); i++) v[i] -= o.v[i]; return *this; }
  NumericVector& operator*=(T a) { scale(a); return *this; }
  std::unique_ptr<NumericVector> clone() const { auto p = std::make_unique<NumericVector>(); p->v = v; return p; }
  std::unique_ptr<NumericVector> zero_clone() const { return std::make_unique<NumericVector>(v.size()); }
  T l2_norm() const { T s = 0; for (T x : v) s += x * x; return std::sqrt(s); }
  T linfty_norm() const { T s = 0; for (T x : v) s = std::max(s, std::fabs(x)); return s; }
  T max() const { return v.empty() ? T(0) : *std::max_element(v.begin(), v.end()); }
  T min() const { return v.empty() ? T(0) : *std::min_element(v.begin(), v.end()); }
  T sum() const { T s = 0; for (T x : v) s += x; return s; }
  T dot(const NumericVector& o) const { T s = 0; for (size_t i = 0; i < v.size(); i++) s += v[i] * o.v[i]; return s; }
  void swap(NumericVector& o) { v.swap(o.v); }
};

// MatSetValues(ADD_VALUES) into one sorted map per row: additions happen in call order (= the serial element loop)
template <class T>
class SparseMatrix {
 public:
  std::vector<std::map<dof_id_type, T>> rows;
  virtual ~SparseMatrix() {}
  void init(size_t n) { rows.assign(n, {}); }
  dof_id_type m() const { return (dof_id_type)rows.size(); }
  dof_id_type n() const { return (dof_id_type)rows.size(); }
  void zero() { for (auto& r : rows) for (auto& kvp : r) kvp.second = T(0); }
  void clear_pattern() { for (auto& r : rows) r.clear(); }
  void add(dof_id_type i, dof_id_type j, T x) { rows[i][j] += x; }
  void set(dof_id_type i, dof_id_type j, T x) { rows[i][j] = x; }
  T operator()(dof_id_type i, dof_id_type j) const { auto it = rows[i].find(j); return it == rows[i].end() ? T(0) : it->second; }
  void add_matrix(const DenseMatrix<T>& K, const std::vector<dof_id_type>& idx) {
    for (size_t i = 0; i < idx.size(); i++) for (size_t j = 0; j < idx.size(); j++) rows[idx[i]][idx[j]] += K(i, j);
  }
  void add_matrix(const DenseMatrix<T>& K, const std::vector<dof_id_type>& ri, const std::vector<dof_id_type>& ci) {
    for (size_t i = 0; i < ri.size(); i++) for (size_t j = 0; j < ci.size(); j++) rows[ri[i]][ci[j]] += K(i, j);
  }
  void close() {}
  bool closed() const { return true; }
  void vector_mult(NumericVector<T>& y, const NumericVector<T>& x) const {
    y.init(rows.size());
    for (size_t i = 0; i < rows.size(); i++) { T s = 0; for (auto& kvp : rows[i]) s += kvp.second * x.v[kvp.first]; y.v[i] = s; }
  }
};

// ------------------------------------------------------------------------------------------- dof map
class System;
class DofConstraints {};
class DofMap {
 public:
  const System* sys = nullptr;
  void dof_indices(const Elem* e, std::vector<dof_id_type>& di) const;
  void dof_indices(const Elem* e, std::vector<dof_id_type>& di, unsigned var) const;
  void dof_indices(const Node* n, std::vector<dof_id_type>& di) const;
  void dof_indices(const Node* n, std::vector<dof_id_type>& di, unsigned var) const;
  template <class... A> void constrain_element_matrix_and_vector(A&&...) const {}   // no constraints in any RDC system
  template <class... A> void constrain_element_matrix(A&&...) const {}
  template <class... A> void constrain_element_vector(A&&...) const {}
  template <class... A> void heterogenously_constrain_element_matrix_and_vector(A&&...) const {}
  template <class... A> void enforce_constraints_exactly(A&&...) const {}
  dof_id_type n_dofs() const;
  dof_id_type n_local_dofs() const { return n_dofs(); }
  dof_id_type first_dof() const { return 0; }
  dof_id_type end_dof() const { return n_dofs(); }
  dof_id_type n_constrained_dofs() const { return 0; }
  bool is_constrained_dof(dof_id_type) const { return false; }
  const std::vector<dof_id_type>& get_send_list() const { static std::vector<dof_id_type> s; return s; }
  FEType variable_type(unsigned v) const;
};

// ------------------------------------------------------------------------------------------- systems
struct VariableInfo { std::string name; FEType type; };
class Variable {
 public:
  VariableInfo info;
  const std::string& name() const { return info.name; }
  const FEType& type() const { return info.type; }
};
class EquationSystems;
class System {
 public:
  EquationSystems* es_;
  std::string name_;
  unsigned number_;
  std::vector<VariableInfo> vars;
  DofMap dofmap_;
  Real time = 0.0;
  std::unique_ptr<NumericVector<Number>> solution, current_local_solution;
  std::map<std::string, std::unique_ptr<NumericVector<Number>>> extra_vectors;
  void (*init_fn)(EquationSystems&, const std::string&) = nullptr;
  void (*assemble_fn)(EquationSystems&, const std::string&) = nullptr;
  bool nodal = true;                 // FIRST LAGRANGE (nodal dofs) or CONSTANT MONOMIAL (one dof per element and variable)
  dof_id_type n_dofs_ = 0;
  bool assemble_before_solve = true;
  System(EquationSystems& es, const std::string& name, unsigned number) : es_(&es), name_(name), number_(number) {
    dofmap_.sys = this;
    solution = std::make_unique<NumericVector<Number>>();
    current_local_solution = std::make_unique<NumericVector<Number>>();
  }
  virtual ~System() {}
  const std::string& name() const { return name_; }
  unsigned number() const { return number_; }
  unsigned n_vars() const { return (unsigned)vars.size(); }
  dof_id_type n_dofs() const { return n_dofs_; }
  dof_id_type n_local_dofs() const { return n_dofs_; }
  unsigned add_variable(const std::string& n, Order o = FIRST, FEFamily f = LAGRANGE, const std::set<subdomain_id_type>* = nullptr) {
    vars.push_back(VariableInfo{n, FEType(o, f)});
    nodal = !(f == MONOMIAL && o == CONSTANT);
    return (unsigned)vars.size() - 1;
  }
  unsigned add_variable(const std::string& n, const FEType& t, const std::set<subdomain_id_type>* = nullptr) { return add_variable(n, t.order, t.family); }
  FEType variable_type(unsigned v) const { return vars[v].type; }
  FEType variable_type(const std::string& n) const { return vars[variable_number(n)].type; }
  unsigned variable_number(const std::string& n) const {
    for (unsigned v = 0; v < vars.size(); v++) if (vars[v].name == n) return v;
    throw std::runtime_error("System::variable_number: no variable " + n);
  }
  const std::string& variable_name(unsigned v) const { return vars[v].name; }
  bool has_variable(const std::string& n) const { for (auto& v : vars) if (v.name == n) return true; return false; }
  const DofMap& get_dof_map() const { return dofmap_; }
  DofMap& get_dof_map() { return dofmap_; }
  const MeshBase& get_mesh() const;
  MeshBase& get_mesh();
  EquationSystems& get_equation_systems() { return *es_; }
  const EquationSystems& get_equation_systems() const { return *es_; }
  const Parallel::Communicator& comm() const { static Parallel::Communicator c; return c; }
  void attach_init_function(void (*f)(EquationSystems&, const std::string&)) { init_fn = f; }
  void attach_assemble_function(void (*f)(EquationSystems&, const std::string&)) { assemble_fn = f; }
  // [upstream] System::update(): current_local_solution <- solution (localize)
  virtual void update() { *current_local_solution = *solution; }
  Number current_solution(dof_id_type d) const { return (*current_local_solution)(d); }
  void update_global_solution(std::vector<Number>& g) const { g = solution->v; }
  void update_global_solution(std::vector<Number>& g, processor_id_type) const { g = solution->v; }
  NumericVector<Number>& add_vector(const std::string& n, bool = true, ParallelType = PARALLEL) {
    auto& p = extra_vectors[n];
    if (!p) p = std::make_unique<NumericVector<Number>>(n_dofs_);
    return *p;
  }
  NumericVector<Number>& get_vector(const std::string& n) { return *extra_vectors.at(n); }
  const NumericVector<Number>& get_vector(const std::string& n) const { return *extra_vectors.at(n); }
  bool have_vector(const std::string& n) const { return extra_vectors.count(n) != 0; }
  virtual void init_data(dof_id_type nd) {
    n_dofs_ = nd;
    solution->init(nd);
    current_local_solution->init(nd);
    for (auto& kvp : extra_vectors) kvp.second->init(nd);
  }
  virtual void reinit() {}
  virtual void solve() {}
  virtual void assemble() { if (assemble_fn) assemble_fn(*es_, name_); }
  virtual std::string system_type() const { return "Basic"; }
  void project_solution(Number (*)(const Point&, const Parameters&, const std::string&, const std::string&), void* = nullptr, const Parameters* = nullptr) {}
  Number point_value(unsigned, const Point&, bool = true) const { return 0; }
};
class ExplicitSystem : public System {
 public:
  std::unique_ptr<NumericVector<Number>> rhs_owner;
  NumericVector<Number>* rhs;
  ExplicitSystem(EquationSystems& es, const std::string& n, unsigned k) : System(es, n, k) {
    rhs_owner = std::make_unique<NumericVector<Number>>();
    rhs = rhs_owner.get();
  }
  void init_data(dof_id_type nd) override { System::init_data(nd); rhs->init(nd); }
  std::string system_type() const override { return "Explicit"; }
};
class ImplicitSystem : public ExplicitSystem {
 public:
  std::unique_ptr<SparseMatrix<Number>> matrix_owner;
  SparseMatrix<Number>* matrix;
  ImplicitSystem(EquationSystems& es, const std::string& n, unsigned k) : ExplicitSystem(es, n, k) {
    matrix_owner = std::make_unique<SparseMatrix<Number>>();
    matrix = matrix_owner.get();
  }
  void init_data(dof_id_type nd) override { ExplicitSystem::init_data(nd); matrix->init(nd); }
  SparseMatrix<Number>& get_system_matrix() { return *matrix; }
  const SparseMatrix<Number>& get_system_matrix() const { return *matrix; }
  SparseMatrix<Number>& get_matrix(const std::string&) { return *matrix; }
};
enum LinearConvergenceReason { CONVERGED_RTOL_NORMAL = 1, CONVERGED_ATOL_NORMAL = 9, CONVERGED_RTOL = 2, CONVERGED_ATOL = 3, CONVERGED_ITS = 4,
                               CONVERGED_ITERATING = 0, DIVERGED_NULL = -2, DIVERGED_ITS = -3, DIVERGED_DTOL = -4, DIVERGED_BREAKDOWN = -5,
                               DIVERGED_NAN = -9, UNKNOWN_FLAG = -128 };
template <class T> class ShellMatrix {};
// [upstream] libMesh::LinearSolver<T>: the pure virtuals a subclass has to provide (signatures of libMesh @d3bda6c)
template <class T> class LinearSolver {
 public:
  LinearSolver(const Parallel::Communicator&) {}
  virtual ~LinearSolver() {}
  virtual void init(const char* name = nullptr) = 0;
  virtual void clear() {}
  bool initialized() const { return _is_initialized; }
  virtual std::pair<unsigned int, Real> solve(SparseMatrix<T>&, SparseMatrix<T>&, NumericVector<T>&, NumericVector<T>&,
                                              const std::optional<double> tol = std::nullopt,
                                              const std::optional<unsigned int> m_its = std::nullopt) = 0;
  virtual std::pair<unsigned int, Real> solve(const ShellMatrix<T>&, NumericVector<T>&, NumericVector<T>&, const std::optional<double> = std::nullopt,
                                              const std::optional<unsigned int> = std::nullopt) = 0;
  virtual std::pair<unsigned int, Real> solve(const ShellMatrix<T>&, const SparseMatrix<T>&, NumericVector<T>&, NumericVector<T>&,
                                              const std::optional<double> = std::nullopt, const std::optional<unsigned int> = std::nullopt) = 0;
  virtual void print_converged_reason() const {}
  virtual LinearConvergenceReason get_converged_reason() const = 0;

 protected:
  bool _is_initialized = false;
};
class LinearImplicitSystem : public ImplicitSystem {
 public:
  using ImplicitSystem::ImplicitSystem;
  std::unique_ptr<LinearSolver<Number>> linear_solver;
  unsigned n_its = 0;
  Real final_res = 0;
  // [upstream] zero K and F, assemble, KSP; the Krylov solve itself is PETSc's and is not part of
  // this shim: the wrapper reads K and F after assemble()
  void solve() override {
    matrix->zero(); rhs->zero();
    assemble();
    if (linear_solver) { auto r = linear_solver->solve(*matrix, *matrix, *solution, *rhs, 1e-12, 5000u); n_its = r.first; final_res = r.second; }
    update();
  }
  unsigned n_linear_iterations() const { return n_its; }
  Real final_linear_residual() const { return final_res; }
  LinearSolver<Number>* get_linear_solver() const { return linear_solver.get(); }
  std::string system_type() const override { return "LinearImplicit"; }
};
template <class Base>
class TransientSystem : public Base {
 public:
  std::unique_ptr<NumericVector<Number>> old_local_solution, older_local_solution;
  TransientSystem(EquationSystems& es, const std::string& n, unsigned k) : Base(es, n, k) {
    old_local_solution = std::make_unique<NumericVector<Number>>();
    older_local_solution = std::make_unique<NumericVector<Number>>();
  }
  void init_data(dof_id_type nd) override { Base::init_data(nd); old_local_solution->init(nd); older_local_solution->init(nd); }
  Number old_solution(dof_id_type d) const { return (*old_local_solution)(d); }
  Number older_solution(dof_id_type d) const { return (*older_local_solution)(d); }
  std::string system_type() const override { return "Transient" + Base::system_type(); }
};
typedef TransientSystem<LinearImplicitSystem> TransientLinearImplicitSystem;
typedef TransientSystem<ImplicitSystem> TransientImplicitSystem;
typedef TransientSystem<ExplicitSystem> TransientExplicitSystem;
typedef TransientSystem<System> TransientBaseSystem;

// ---- FEMSystem family: what src/solid_system.{h,C} touch.  FEMContext is a plain holder that the wrapper
// (ref_solid.cpp) fills per element / side; the Newton driver ([upstream] NewtonSolver) is not modelled. ----
class DiffContext { public: virtual ~DiffContext() {} };
class FEBase;
class QBase;
class FEMContext : public DiffContext {
 public:
  const Elem* elem_ = nullptr;
  unsigned char side_ = 0;
  FEBase* elem_fe_ = nullptr;
  FEBase* side_fe_ = nullptr;
  QBase* elem_q_ = nullptr;
  QBase* side_q_ = nullptr;
  unsigned nvars_ = 0, ndofs_var_ = 0;
  DenseVector<Number> residual_;
  DenseMatrix<Number> jacobian_;
  std::vector<std::unique_ptr<DenseSubVector<Number>>> sub_res_;
  std::vector<std::vector<std::unique_ptr<DenseSubMatrix<Number>>>> sub_jac_;
  Real elem_solution_derivative = 1.0;
  // [upstream] FEMContext::pre_fe_reinit sizes the element residual/Jacobian variable-major (same order as dof_indices)
  void resize(unsigned nvars, unsigned ndofs_var) {
    nvars_ = nvars; ndofs_var_ = ndofs_var;
    residual_.resize(nvars * ndofs_var);
    jacobian_.resize(nvars * ndofs_var, nvars * ndofs_var);
    sub_res_.clear(); sub_jac_.clear();
    for (unsigned a = 0; a < nvars; a++) {
      sub_res_.push_back(std::make_unique<DenseSubVector<Number>>(residual_, a * ndofs_var, ndofs_var));
      sub_jac_.emplace_back();
      for (unsigned b = 0; b < nvars; b++)
        sub_jac_.back().push_back(std::make_unique<DenseSubMatrix<Number>>(jacobian_, a * ndofs_var, b * ndofs_var, ndofs_var, ndofs_var));
    }
  }
  const Elem& get_elem() const { return *elem_; }
  unsigned char get_side() const { return side_; }
  void get_element_fe(unsigned, FEBase*& fe) const { fe = elem_fe_; }
  void get_element_fe(unsigned, FEBase*& fe, unsigned char) const { fe = elem_fe_; }
  void get_side_fe(unsigned, FEBase*& fe) const { fe = side_fe_; }
  void get_side_fe(unsigned, FEBase*& fe, unsigned char) const { fe = side_fe_; }
  const QBase& get_element_qrule() const { return *elem_q_; }
  const QBase& get_side_qrule() const { return *side_q_; }
  unsigned n_dof_indices(unsigned) const { return ndofs_var_; }
  DenseSubVector<Number>& get_elem_residual(unsigned v) { return *sub_res_[v]; }
  DenseSubMatrix<Number>& get_elem_jacobian(unsigned a, unsigned b) { return *sub_jac_[a][b]; }
  DenseVector<Number>& get_elem_residual() { return residual_; }
  DenseMatrix<Number>& get_elem_jacobian() { return jacobian_; }
};
class DiffSolver {
 public:
  bool quiet = true, verbose = false;
  unsigned max_nonlinear_iterations = 0, max_linear_iterations = 0;
  Real relative_step_tolerance = 0, relative_residual_tolerance = 0, absolute_residual_tolerance = 0, initial_linear_tolerance = 0,
       minimum_linear_tolerance = 0;
  bool continue_after_max_iterations = false, continue_after_backtrack_failure = false;
  virtual ~DiffSolver() {}
};
class NewtonSolver : public DiffSolver {
 public:
  bool require_residual_reduction = false;
  Real linear_tolerance_multiplier = 0;
};
class DifferentiableSystem;
class TimeSolver {
 public:
  TimeSolver(DifferentiableSystem&) {}
  virtual ~TimeSolver() {}
  std::unique_ptr<DiffSolver>& diff_solver() { return ds; }
  virtual void advance_timestep() {}
  std::unique_ptr<DiffSolver> ds;
};
class SteadySolver : public TimeSolver { public: using TimeSolver::TimeSolver; };
class DifferentiableSystem : public ImplicitSystem {
 public:
  using ImplicitSystem::ImplicitSystem;
  std::unique_ptr<TimeSolver> time_solver;
  Real deltat = 0;
  bool print_residuals = false, print_jacobians = false, print_element_jacobians = false, print_solutions = false,
       print_residual_norms = false, print_jacobian_norms = false, print_solution_norms = false, verify_analytic_jacobians = false;
  virtual void init_data() {}
  virtual void init_context(DiffContext&) {}
  virtual bool element_time_derivative(bool, DiffContext&) { return false; }
  virtual bool side_time_derivative(bool, DiffContext&) { return false; }
  virtual bool eulerian_residual(bool, DiffContext&) { return false; }
  virtual bool mass_residual(bool, DiffContext&) { return false; }
  void time_evolving(unsigned, unsigned = 1) {}
  using ImplicitSystem::init_data;
};
class FEMSystem : public DifferentiableSystem {
 public:
  using DifferentiableSystem::DifferentiableSystem;
  Real numerical_jacobian_h = 0;
  void mesh_position_get() {}
  void mesh_position_set() {}
  void mesh_x_var(unsigned) {}
  void mesh_y_var(unsigned) {}
  void mesh_z_var(unsigned) {}
  void set_mesh_system(System*) {}
  void set_mesh_x_var(unsigned) {}
  void set_mesh_y_var(unsigned) {}
  void set_mesh_z_var(unsigned) {}
  void postprocess() {}
};

class EquationSystems {
 public:
  MeshBase* mesh_;
  Parameters parameters;
  std::vector<std::unique_ptr<System>> systems;
  EquationSystems(MeshBase& m) : mesh_(&m) {}
  const MeshBase& get_mesh() const { return *mesh_; }
  MeshBase& get_mesh() { return *mesh_; }
  unsigned n_systems() const { return (unsigned)systems.size(); }
  template <class T> T& add_system(const std::string& name) {
    for (auto& s : systems) if (s->name() == name) return dynamic_cast<T&>(*s);
    systems.push_back(std::make_unique<T>(*this, name, (unsigned)systems.size()));
    return static_cast<T&>(*systems.back());
  }
  template <class T> T& get_system(const std::string& name) {
    for (auto& s : systems) if (s->name() == name) return dynamic_cast<T&>(*s);
    throw std::runtime_error("EquationSystems::get_system: no system " + name);
  }
  template <class T> const T& get_system(const std::string& name) const { return const_cast<EquationSystems*>(this)->get_system<T>(name); }
  template <class T> T& get_system(unsigned k) { return dynamic_cast<T&>(*systems.at(k)); }
  template <class T> const T& get_system(unsigned k) const { return dynamic_cast<const T&>(*systems.at(k)); }
  System& get_system(const std::string& name) { return get_system<System>(name); }
  System& get_system(unsigned k) { return *systems.at(k); }
  bool has_system(const std::string& name) const { for (auto& s : systems) if (s->name() == name) return true; return false; }
  // dof numbering: node-blocked per system (dof = base + var); a caller-provided base is used for the system named
  // in `custom_base_system` (libMesh's first-touch numbering is passed in explicitly)
  std::string custom_base_system;
  std::vector<int32_t> custom_base;
  bool run_init_functions = false;
  void init() {
    const dof_id_type N = mesh_->n_nodes(), E = mesh_->n_elem();
    for (Node* n : mesh_->nodes_) { n->sys_base.assign(systems.size(), 0); n->sys_nvars.assign(systems.size(), 0); }
    for (auto& s : systems) {
      const unsigned nv = s->n_vars();
      if (s->nodal) {
        for (Node* n : mesh_->nodes_) {
          n->sys_nvars[s->number()] = nv;
          n->sys_base[s->number()] = (s->name() == custom_base_system && !custom_base.empty()) ? (dof_id_type)custom_base[n->id()] : nv * n->id();
        }
        s->init_data(nv * N);
      } else {
        s->init_data(nv * E);
      }
    }
    if (run_init_functions) for (auto& s : systems) if (s->init_fn) s->init_fn(*this, s->name());
  }
  void reinit() {}
  void update() { for (auto& s : systems) s->update(); }
  void print_info(std::ostream& = libMesh::out) const {}
  void build_variable_names(std::vector<std::string>& names, const FEType* = nullptr, const std::set<std::string>* = nullptr) const {
    names.clear();
    for (auto& s : systems) if (s->nodal) for (auto& v : s->vars) names.push_back(v.name);
  }
  void build_solution_vector(std::vector<Number>& soln, const std::set<std::string>* = nullptr) const {
    std::vector<std::string> names;
    build_variable_names(names);
    const size_t nvt = names.size(), N = mesh_->n_nodes();
    soln.assign(nvt * N, 0.0);
    size_t off = 0;
    for (auto& s : systems) {
      if (!s->nodal) continue;
      for (const Node* n : mesh_->nodes_)
        for (unsigned v = 0; v < s->n_vars(); v++) soln[n->id() * nvt + off + v] = (*s->solution)(n->dof_number(s->number(), v, 0));
      off += s->n_vars();
    }
  }
};
inline const MeshBase& System::get_mesh() const { return es_->get_mesh(); }
inline MeshBase& System::get_mesh() { return es_->get_mesh(); }

inline dof_id_type DofMap::n_dofs() const { return sys->n_dofs(); }
inline FEType DofMap::variable_type(unsigned v) const { return sys->variable_type(v); }
// [upstream] element dof order: variable-major, node order of the element inside a variable
inline void DofMap::dof_indices(const Elem* e, std::vector<dof_id_type>& di) const {
  di.clear();
  for (unsigned v = 0; v < sys->n_vars(); v++) {
    if (sys->nodal) for (unsigned i = 0; i < e->n_nodes(); i++) di.push_back(e->node_ptr(i)->dof_number(sys->number(), v, 0));
    else di.push_back(e->id() * sys->n_vars() + v);
  }
}
inline void DofMap::dof_indices(const Elem* e, std::vector<dof_id_type>& di, unsigned v) const {
  di.clear();
  if (sys->nodal) for (unsigned i = 0; i < e->n_nodes(); i++) di.push_back(e->node_ptr(i)->dof_number(sys->number(), v, 0));
  else di.push_back(e->id() * sys->n_vars() + v);
}
inline void DofMap::dof_indices(const Node* n, std::vector<dof_id_type>& di) const {
  di.clear();
  if (sys->nodal) for (unsigned v = 0; v < sys->n_vars(); v++) di.push_back(n->dof_number(sys->number(), v, 0));
}
inline void DofMap::dof_indices(const Node* n, std::vector<dof_id_type>& di, unsigned v) const {
  di.clear();
  if (sys->nodal) di.push_back(n->dof_number(sys->number(), v, 0));
}

// ------------------------------------------------------------------------------------------- quadrature + FE
// [upstream] QGauss for 3-D TET4/HEX8: THIRD -> 5-point rule with a negative weight on tets,
// 2x2x2 Gauss-Legendre on hexes.  Other orders are not used by the reference (FIRST LAGRANGE everywhere).
class QBase {
 public:
  unsigned dim_;
  Order order_;
  std::vector<Point> pts[2];      // [0] tet, [1] hex
  std::vector<Real> w[2];
  mutable int active = 0;
  QBase(unsigned d, Order o) : dim_(d), order_(o) {
    if (d == 3) {
      const Real s = 1. / 6.;
      pts[0] = {Point(.25, .25, .25), Point(.5, s, s), Point(s, .5, s), Point(s, s, .5), Point(s, s, s)};
      w[0] = {-2. / 15., .075, .075, .075, .075};
      const Real g = 5.7735026918962576450914878050196e-01;
      const Real p1[2] = {-g, g};
      for (int k = 0; k < 2; k++) for (int j = 0; j < 2; j++) for (int i = 0; i < 2; i++) { pts[1].push_back(Point(p1[i], p1[j], p1[k])); w[1].push_back(1.0); }
    } else if (d == 2) {
      // [upstream] QGauss::init_2D, THIRD: triangles get the 4-point rule with a negative centroid weight
      // (allow_rules_with_negative_weights defaults to true), quadrilaterals the 2x2 tensor Gauss rule.
      // Only the side integrals of SolidSystem::side_time_derivative (solid_system.C:273-371) use them.
      pts[0] = {Point(1. / 3., 1. / 3.), Point(.2, .6), Point(.2, .2), Point(.6, .2)};
      w[0] = {-27. / 96., 25. / 96., 25. / 96., 25. / 96.};
      const Real g = 5.7735026918962576450914878050196e-01;
      const Real p1[2] = {-g, g};
      for (int j = 0; j < 2; j++) for (int i = 0; i < 2; i++) { pts[1].push_back(Point(p1[i], p1[j])); w[1].push_back(1.0); }
    }
  }
  virtual ~QBase() {}
  unsigned n_points() const { return (unsigned)pts[active].size(); }
  unsigned get_dim() const { return dim_; }
  Order get_order() const { return order_; }
  const std::vector<Point>& get_points() const { return pts[active]; }
  const std::vector<Real>& get_weights() const { return w[active]; }
  Point qp(unsigned i) const { return pts[active][i]; }
  Real w_(unsigned i) const { return w[active][i]; }
};
class QGauss : public QBase { public: QGauss(unsigned d, Order o = THIRD) : QBase(d, o) {} };

class FEBase {
 public:
  unsigned dim_;
  FEType type_;
  QBase* q = nullptr;
  std::vector<Real> JxW_;
  std::vector<std::vector<Real>> phi_;
  std::vector<std::vector<RealGradient>> dphi_;
  std::vector<Point> xyz_, normals_;
  FEBase(unsigned d, const FEType& t) : dim_(d), type_(t) {}
  virtual ~FEBase() {}
  static std::unique_ptr<FEBase> build(unsigned d, const FEType& t) { return std::make_unique<FEBase>(d, t); }
  void attach_quadrature_rule(QBase* r) { q = r; }
  const std::vector<Real>& get_JxW() const { return JxW_; }
  const std::vector<std::vector<Real>>& get_phi() const { return phi_; }
  const std::vector<std::vector<RealGradient>>& get_dphi() const { return dphi_; }
  const std::vector<Point>& get_xyz() const { return xyz_; }
  const std::vector<Point>& get_normals() const { return normals_; }
  FEType get_fe_type() const { return type_; }
  unsigned n_shape_functions() const { return (unsigned)phi_.size(); }
  unsigned n_quadrature_points() const { return (unsigned)JxW_.size(); }
  // reference shape functions and their local derivatives at one point
  static void shape(ElemType t, const Point& p, std::vector<Real>& N, std::vector<Real>& dxi, std::vector<Real>& deta, std::vector<Real>& dzeta) {
    if (t == TET4) {
      const Real z1 = p(0), z2 = p(1), z3 = p(2), z0 = 1. - z1 - z2 - z3;   // [upstream] zeta0 = 1 - xi - eta - zeta
      N = {z0, z1, z2, z3};
      dxi = {-1., 1., 0., 0.}; deta = {-1., 0., 1., 0.}; dzeta = {-1., 0., 0., 1.};
    } else {
      static const int i0[8] = {0, 1, 1, 0, 0, 1, 1, 0}, i1[8] = {0, 0, 1, 1, 0, 0, 1, 1}, i2[8] = {0, 0, 0, 0, 1, 1, 1, 1};
      const Real xi = p(0), eta = p(1), zeta = p(2);
      const Real Lx[2] = {.5 * (1. - xi), .5 * (1. + xi)}, Ly[2] = {.5 * (1. - eta), .5 * (1. + eta)}, Lz[2] = {.5 * (1. - zeta), .5 * (1. + zeta)};
      const Real dL[2] = {-.5, .5};
      N.resize(8); dxi.resize(8); deta.resize(8); dzeta.resize(8);
      for (int n = 0; n < 8; n++) {
        N[n] = Lx[i0[n]] * Ly[i1[n]] * Lz[i2[n]];
        dxi[n] = dL[i0[n]] * Ly[i1[n]] * Lz[i2[n]];
        deta[n] = Lx[i0[n]] * dL[i1[n]] * Lz[i2[n]];
        dzeta[n] = Lx[i0[n]] * Ly[i1[n]] * dL[i2[n]];
      }
    }
  }
  // [upstream] FEMap::compute_single_point_map + FE::compute_shape_functions for 3-D (Appendix B-4)
  void reinit(const Elem* e) {
    const ElemType t = e->type();
    const unsigned nen = e->n_nodes();
    if (type_.family == MONOMIAL && type_.order == CONSTANT) {   // one constant shape function
      q->active = t == TET4 ? 0 : 1;
      const unsigned nq = q->n_points();
      phi_.assign(1, std::vector<Real>(nq, 1.0));
      dphi_.assign(1, std::vector<RealGradient>(nq));
      map_only(e, t, nen, nq);
      return;
    }
    q->active = t == TET4 ? 0 : 1;
    const unsigned nq = q->n_points();
    phi_.assign(nen, std::vector<Real>(nq));
    dphi_.assign(nen, std::vector<RealGradient>(nq));
    JxW_.assign(nq, 0.0);
    xyz_.assign(nq, Point());
    std::vector<Real> N, dxi, deta, dzeta;
    for (unsigned p = 0; p < nq; p++) {
      shape(t, q->qp(p), N, dxi, deta, dzeta);
      Real dxdxi = 0, dxdeta = 0, dxdzeta = 0, dydxi = 0, dydeta = 0, dydzeta = 0, dzdxi = 0, dzdeta = 0, dzdzeta = 0;
      Point x;
      for (unsigned n = 0; n < nen; n++) {
        const Point& P = e->point(n);
        x.add_scaled(P, N[n]);
        dxdxi += P(0) * dxi[n]; dxdeta += P(0) * deta[n]; dxdzeta += P(0) * dzeta[n];
        dydxi += P(1) * dxi[n]; dydeta += P(1) * deta[n]; dydzeta += P(1) * dzeta[n];
        dzdxi += P(2) * dxi[n]; dzdeta += P(2) * deta[n]; dzdzeta += P(2) * dzeta[n];
      }
      const Real jac = dxdxi * (dydeta * dzdzeta - dzdeta * dydzeta) + dydxi * (dzdeta * dxdzeta - dxdeta * dzdzeta) +
                       dzdxi * (dxdeta * dydzeta - dydeta * dxdzeta);
      const Real inv = 1. / jac;
      const Real xix = (dydeta * dzdzeta - dzdeta * dydzeta) * inv, xiy = (dzdeta * dxdzeta - dxdeta * dzdzeta) * inv,
                 xiz = (dxdeta * dydzeta - dydeta * dxdzeta) * inv;
      const Real etax = (dzdxi * dydzeta - dydxi * dzdzeta) * inv, etay = (dxdxi * dzdzeta - dzdxi * dxdzeta) * inv,
                 etaz = (dydxi * dxdzeta - dxdxi * dydzeta) * inv;
      const Real zex = (dydxi * dzdeta - dzdxi * dydeta) * inv, zey = (dzdxi * dxdeta - dxdxi * dzdeta) * inv,
                 zez = (dxdxi * dydeta - dydxi * dxdeta) * inv;
      JxW_[p] = jac * q->w_(p);
      xyz_[p] = x;
      for (unsigned n = 0; n < nen; n++) {
        phi_[n][p] = N[n];
        dphi_[n][p] = RealGradient(dxi[n] * xix + deta[n] * etax + dzeta[n] * zex, dxi[n] * xiy + deta[n] * etay + dzeta[n] * zey,
                                   dxi[n] * xiz + deta[n] * etaz + dzeta[n] * zez);
      }
    }
  }
  void map_only(const Elem* e, ElemType t, unsigned nen, unsigned nq) {
    JxW_.assign(nq, 0.0);
    xyz_.assign(nq, Point());
    std::vector<Real> N, dxi, deta, dzeta;
    for (unsigned p = 0; p < nq; p++) {
      shape(t, q->qp(p), N, dxi, deta, dzeta);
      Real J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
      for (unsigned n = 0; n < nen; n++)
        for (int c = 0; c < 3; c++) { J[c][0] += e->point(n)(c) * dxi[n]; J[c][1] += e->point(n)(c) * deta[n]; J[c][2] += e->point(n)(c) * dzeta[n]; }
      const Real jac = J[0][0] * (J[1][1] * J[2][2] - J[2][1] * J[1][2]) + J[1][0] * (J[2][1] * J[0][2] - J[0][1] * J[2][2]) +
                       J[2][0] * (J[0][1] * J[1][2] - J[1][1] * J[0][2]);
      JxW_[p] = jac * q->w_(p);
    }
  }
  // [upstream] FE::reinit(elem, side): quadrature on the side element (TRI3 / QUAD4, q must be a 2-D rule), JxW and xyz
  // from the side's own map, phi = the PARENT's shape functions at the side points.  libMesh finds the parent reference
  // point by inverse_map; here it is formed directly from the side's node positions in the parent reference element
  // (exact; the off-side shape functions are then exactly 0 instead of ~1e-16).  dphi is not filled (not requested by
  // SolidSystem::init_context).  Node order of a side = [upstream] Tet4/Hex8::side_nodes_map.
  static const unsigned* side_nodes(ElemType t, unsigned s) {
    static const unsigned tet[4][4] = {{0, 2, 1, 0}, {0, 1, 3, 0}, {1, 2, 3, 0}, {2, 0, 3, 0}};
    static const unsigned hex[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
    return t == TET4 ? tet[s] : hex[s];
  }
  void reinit(const Elem* e, unsigned side) {
    const ElemType t = e->type();
    const unsigned nen = e->n_nodes(), ns = t == TET4 ? 3u : 4u;
    const unsigned* sn = side_nodes(t, side);
    q->active = t == TET4 ? 0 : 1;
    const unsigned nq = q->n_points();
    phi_.assign(nen, std::vector<Real>(nq, 0.0));
    dphi_.assign(nen, std::vector<RealGradient>(nq));
    JxW_.assign(nq, 0.0);
    xyz_.assign(nq, Point());
    normals_.assign(nq, Point());
    for (unsigned p = 0; p < nq; p++) {
      const Real xi = q->qp(p)(0), eta = q->qp(p)(1);
      Real N[4], dxi[4], deta[4];
      if (ns == 3) {
        N[0] = 1. - xi - eta; N[1] = xi; N[2] = eta;
        dxi[0] = -1.; dxi[1] = 1.; dxi[2] = 0.; deta[0] = -1.; deta[1] = 0.; deta[2] = 1.;
      } else {
        const Real Lx[2] = {.5 * (1. - xi), .5 * (1. + xi)}, Ly[2] = {.5 * (1. - eta), .5 * (1. + eta)}, dL[2] = {-.5, .5};
        static const int i0[4] = {0, 1, 1, 0}, i1[4] = {0, 0, 1, 1};
        for (int n = 0; n < 4; n++) { N[n] = Lx[i0[n]] * Ly[i1[n]]; dxi[n] = dL[i0[n]] * Ly[i1[n]]; deta[n] = Lx[i0[n]] * dL[i1[n]]; }
      }
      Point x, dxdxi, dxdeta;
      for (unsigned n = 0; n < ns; n++) {
        const Point& P = e->point(sn[n]);
        x.add_scaled(P, N[n]); dxdxi.add_scaled(P, dxi[n]); dxdeta.add_scaled(P, deta[n]);
        phi_[sn[n]][p] = N[n];
      }
      const Point nrm = dxdxi.cross(dxdeta);
      const Real jac = nrm.norm();   // [upstream] FEMap::compute_face_map: sqrt(g11 g22 - g12 g21) = |dx/dxi x dx/deta|
      JxW_[p] = jac * q->w_(p);
      xyz_[p] = x;
      normals_[p] = nrm / jac;
    }
  }
};
template <unsigned D, FEFamily F> class FE : public FEBase { public: using FEBase::FEBase; };

// [upstream] Elem::volume(): analytic for TET4 (triple product / 6), quadrature of the Jacobian for HEX8
inline Real Elem::volume() const {
  if (type_ == TET4) {
    const Point a = *nodes[1] - *nodes[0], b = *nodes[2] - *nodes[0], c = *nodes[3] - *nodes[0];
    return (a * b.cross(c)) / 6.0;   // triple_product(a, b, c) / 6
  }
  QGauss qr(3, THIRD);
  FEBase fe(3, FEType(FIRST, LAGRANGE));
  fe.attach_quadrature_rule(&qr);
  fe.reinit(this);
  Real v = 0;
  for (Real j : fe.get_JxW()) v += j;
  return v;
}

// ------------------------------------------------------------------------------------------- AMR stubs (never run)
class ErrorVector : public std::vector<float> {
 public:
  Real mean() const { return 0; }
  Real variance() const { return 0; }
  Real median() { return 0; }
};
class SystemNorm {
 public:
  SystemNorm() {}
  template <class... A> SystemNorm(A&&...) {}
};
class ErrorEstimator {
 public:
  typedef std::map<std::pair<const System*, unsigned>, ErrorVector*> ErrorMap;
  SystemNorm error_norm;
  virtual void estimate_errors(const EquationSystems&, ErrorMap&, const std::map<const System*, const NumericVector<Number>*>* = nullptr, bool = false) {}
  virtual ~ErrorEstimator() {}
  virtual void estimate_error(const System&, ErrorVector&, const NumericVector<Number>* = nullptr, bool = false) {}
};
class JumpErrorEstimator : public ErrorEstimator {};
class KellyErrorEstimator : public JumpErrorEstimator {};
class MeshRefinement {
 public:
  MeshRefinement(MeshBase&) {}
  Real r_ = 0, c_ = 0;
  unsigned ml_ = 0;
  Real& refine_fraction() { return r_; }
  Real& coarsen_fraction() { return c_; }
  unsigned& max_h_level() { return ml_; }
  Real& coarsen_threshold() { return c_; }
  Real& absolute_global_tolerance() { return c_; }
  dof_id_type& nelem_target() { static dof_id_type n = 0; return n; }
  template <class... A> void flag_elements_by_error_fraction(A&&...) {}
  template <class... A> void flag_elements_by_mean_stddev(A&&...) {}
  template <class... A> void flag_elements_by_error_tolerance(A&&...) {}
  template <class... A> void flag_elements_by_elem_fraction(A&&...) {}
  bool refine_and_coarsen_elements() { return false; }
  bool refine_elements() { return false; }
  bool coarsen_elements() { return false; }
  void uniformly_refine(unsigned = 1) {}
  void clean_refinement_flags() {}
};

}  // namespace libMesh
