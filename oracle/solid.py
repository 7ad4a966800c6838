"""ctypes front-ends of the solid-mechanics checkers.

TEST INFRASTRUCTURE ONLY -- imported by tests/ and __graft_entry__.smoke().  The product package never imports this.

* `OracleSolid`  : oracle/solid_oracle.c, the CPU restatement of solid_system.C:146-538 + hyperelastic.h +
                   hyperlastic_inline.h + eig3.C and of libMesh's Newton driver.
* `RefSolid`     : the reference's OWN solid_system.C / hyperelastic.h / eig3.C compiled unchanged against the serial
                   libMesh stand-in (oracle/ref_shim/ref_solid.cpp -> oracle/_ref/libref_solid.so).
* `SolidCase`    : one problem description both consume (mesh, materials by subdomain, fibres, boundary sides with
                   their BC ids, BC displacement table, penalty, Newton options).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as O
from . import ref as R

_p = O._p
TET4, HEX8 = 4, 8
# [upstream] Tet4/Hex8::side_nodes_map (libMesh side numbering)
SIDE_NODES = {TET4: [(0, 2, 1), (0, 1, 3), (1, 2, 3), (2, 0, 3)],
              HEX8: [(0, 3, 2, 1), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7), (4, 5, 6, 7)]}
SOLID_SO = os.path.join(R.OUT, "libref_solid.so")


def boundary_faces(elem_type, conn):
    """(elem, side) of every face that belongs to exactly one element."""
    conn = np.asarray(conn)
    seen = {}
    for s, loc in enumerate(SIDE_NODES[elem_type]):
        f = np.sort(conn[:, list(loc)], axis=1)
        for e, key in enumerate(map(tuple, f)):
            seen.setdefault(key, []).append((e, s))
    return [v[0] for v in seen.values() if len(v) == 1]


class SolidCase:
    """A solid-mechanics problem: undeformed mesh, materials, boundary sides, Dirichlet (penalty) data, solver options."""

    def __init__(self, elem_type, conn, xyz, mats=None, mat_of=None, fibres=None, penalty=1.0e5, use_symmetry=False):
        self.elem_type = elem_type
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.xund = np.ascontiguousarray(xyz, dtype=np.float64)
        self.N, self.E = self.xund.shape[0], self.conn.shape[0]
        # {Young, Poisson, FibreStiffness, rate_0, rate_1, rate_2} per material; defaults of solid.C:279-290
        self.mats = np.ascontiguousarray(mats if mats is not None else [[1.0e3, 0.3, 0.0, 0.0, 0.0, 0.0]], dtype=np.float64).reshape(-1, 6)
        self.mat_of = None if mat_of is None else np.ascontiguousarray(mat_of, dtype=np.int32)
        self.fibres = None if fibres is None else np.ascontiguousarray(fibres, dtype=np.float64).reshape(self.E, 3)
        self.penalty = float(penalty)
        self.use_symmetry = bool(use_symmetry)
        self.bc_ids, self.bc_disp = [], []          # BC table: id -> displacement (NaN = free)
        self.side_elem, self.side_no, self.side_bc = [], [], []
        # solid.C:226-245 defaults
        self.opts = dict(max_nonlinear_iterations=100, relative_step_tolerance=1e-3, relative_residual_tolerance=1e-8,
                         absolute_residual_tolerance=1e-8, require_reduction=False, max_linear_iterations=50000,
                         initial_linear_tolerance=1e-3)

    def add_bc(self, bc_id, disp, faces):
        """faces: iterable of (elem, side) that carry boundary id `bc_id`."""
        k = len(self.bc_ids)
        self.bc_ids.append(int(bc_id))
        self.bc_disp.append([float(v) for v in disp])
        for e, s in faces:
            self.side_elem.append(int(e)); self.side_no.append(int(s)); self.side_bc.append(k)

    def faces_where(self, pred):
        """Boundary faces whose nodes all satisfy pred(xyz) (undeformed coordinates)."""
        out = []
        for e, s in boundary_faces(self.elem_type, self.conn):
            nodes = self.conn[e, list(SIDE_NODES[self.elem_type][s])]
            if all(pred(self.xund[n]) for n in nodes):
                out.append((e, s))
        return out

    def arrays(self):
        se = np.ascontiguousarray(self.side_elem, dtype=np.int64)
        sn = np.ascontiguousarray(self.side_no, dtype=np.int32)
        sb = np.ascontiguousarray(self.side_bc, dtype=np.int32)
        bd = np.ascontiguousarray(self.bc_disp if self.bc_disp else [[0.0, 0.0, 0.0]], dtype=np.float64).reshape(-1, 3)
        return se, sn, sb, bd

    def opts_vector(self):
        o = self.opts
        return np.array([o["max_nonlinear_iterations"], o["relative_step_tolerance"], o["relative_residual_tolerance"],
                         o["absolute_residual_tolerance"], 1.0 if o["require_reduction"] else 0.0,
                         o["max_linear_iterations"], o["initial_linear_tolerance"]], dtype=np.float64)


class OracleSolid:
    def __init__(self, case: SolidCase):
        self.c = case
        self.L = O.lib()
        for n in ("orc_solid_element", "orc_solid_side", "orc_solid_assemble", "orc_solid_newton", "orc_solid_post"):
            getattr(self.L, n).restype = C.c_int
        self.L.orc_eig3.restype = None
        nen = 4 if case.elem_type == TET4 else 8
        self.rowptr, self.col = self._pattern(nen)   # (node graph + I) x dense 3x3, what libMesh preallocates

    def _pattern(self, nen):
        c = self.c
        nnz = C.c_int64(0)
        rp, cl = C.c_void_p(), C.c_void_p()
        rc = self.L.orc_build_pattern(C.c_int64(c.N), C.c_int64(c.E), C.c_int(nen), C.c_int(3), _p(c.conn), C.byref(nnz), C.byref(rp), C.byref(cl))
        assert rc == 0
        D = 3 * c.N
        rowptr = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_int64)), shape=(D + 1,)).copy()
        col = np.ctypeslib.as_array(C.cast(cl, C.POINTER(C.c_int32)), shape=(nnz.value,)).copy()
        self.L.orc_free(rp); self.L.orc_free(cl)
        return rowptr, col

    def element(self, e, xcur, t, want_jac=True):
        c = self.c
        nen = c.conn.shape[1]
        nodes = c.conn[e]
        Xc = np.ascontiguousarray(np.asarray(xcur, dtype=np.float64).reshape(-1, 3)[nodes])
        Xu = np.ascontiguousarray(c.xund[nodes])
        mat = np.ascontiguousarray(c.mats[0 if c.mat_of is None else c.mat_of[e]])
        eta = np.ascontiguousarray(c.fibres[e] if c.fibres is not None else np.zeros(3))
        Re = np.zeros(3 * nen); Ke = np.zeros((3 * nen, 3 * nen))
        rc = self.L.orc_solid_element(C.c_int(c.elem_type), _p(Xc), _p(Xu), _p(mat), C.c_double(t), _p(eta), C.c_int(int(want_jac)),
                                      C.c_int(int(c.use_symmetry)), _p(Re), _p(Ke))
        assert rc == 0
        # the side terms of this element
        se, sn, sb, bd = c.arrays()
        for k in np.nonzero(se == e)[0]:
            R2 = np.zeros(3 * nen); K2 = np.zeros((3 * nen, 3 * nen))
            disp = np.ascontiguousarray(bd[sb[k]])
            rc = self.L.orc_solid_side(C.c_int(c.elem_type), C.c_int(int(sn[k])), _p(Xc), _p(Xu), _p(disp), C.c_double(t),
                                       C.c_double(c.penalty), C.c_int(int(want_jac)), _p(R2), _p(K2))
            assert rc == 0
            Re += R2; Ke += K2
        return Re, Ke

    def assemble(self, xcur, t, want_jac=True):
        c = self.c
        x = np.ascontiguousarray(xcur, dtype=np.float64).reshape(-1)
        se, sn, sb, bd = c.arrays()
        val = np.zeros(self.col.shape[0]) if want_jac else None
        rhs = np.zeros(3 * c.N)
        rc = self.L.orc_solid_assemble(C.c_int(c.elem_type), C.c_int64(c.N), C.c_int64(c.E), _p(c.conn), _p(x), _p(c.xund), _p(c.mat_of),
                                       _p(c.mats), _p(c.fibres), C.c_double(t), C.c_int64(se.shape[0]), _p(se), _p(sn), _p(sb), _p(bd),
                                       C.c_double(c.penalty), C.c_int(int(c.use_symmetry)), _p(self.rowptr), _p(self.col), _p(val), _p(rhs))
        assert rc == 0, rc
        return val, rhs

    def newton(self, x, t, pc=O.PC_ILU, nthreads=0):
        """One load step at pseudo-time t from the positions x; returns (x_new, info dict)."""
        c = self.c
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1).copy()
        se, sn, sb, bd = c.arrays()
        info = np.zeros(4)
        ov = c.opts_vector()
        rc = self.L.orc_solid_newton(C.c_int(c.elem_type), C.c_int64(c.N), C.c_int64(c.E), _p(c.conn), _p(x), _p(c.xund), _p(c.mat_of),
                                     _p(c.mats), _p(c.fibres), C.c_double(t), C.c_int64(se.shape[0]), _p(se), _p(sn), _p(sb), _p(bd),
                                     C.c_double(c.penalty), C.c_int(int(c.use_symmetry)), _p(ov), C.c_int(pc),
                                     C.c_int(nthreads or (os.cpu_count() or 1)), _p(info))
        assert rc == 0, rc
        return x, dict(newton_its=int(info[0]), linear_its=int(info[1]), residual=float(info[2]), converged=bool(info[3]))

    def post(self, xcur, t):
        c = self.c
        x = np.ascontiguousarray(xcur, dtype=np.float64).reshape(-1)
        press, vm, fib = np.zeros(c.E), np.zeros(c.E), np.zeros((c.E, 3))
        rc = self.L.orc_solid_post(C.c_int(c.elem_type), C.c_int64(c.N), C.c_int64(c.E), _p(c.conn), _p(x), _p(c.xund), _p(c.mat_of), _p(c.mats),
                                   _p(c.fibres), C.c_double(t), _p(press), _p(vm), _p(fib))
        assert rc == 0
        return press, vm, fib

    def eig3(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64).reshape(9)
        V, d = np.zeros(9), np.zeros(3)
        self.L.orc_eig3(_p(A), _p(V), _p(d))
        return d, V.reshape(3, 3)


def build_ref_solid(force: bool = False) -> bool:
    """g++ on the reference's own solid_system.C + eig3.C (no cmake, no external library)."""
    src = os.path.join(R.REF_ROOT, "src", "solid_system.C")
    eig = os.path.join(R.REF_ROOT, "src", "eig3.C")
    if not (os.path.exists(src) and os.path.exists(eig)):
        return os.path.exists(SOLID_SO)
    os.makedirs(R.OUT, exist_ok=True)
    wrap = os.path.join(R.SHIM, "ref_solid.cpp")
    deps = [wrap, src, eig, os.path.join(R.SHIM, "libmesh", "shim.h")] + [os.path.join(R.REF_ROOT, "src", f) for f in
                                                                        ("hyperelastic.h", "hyperlastic_inline.h", "solid_system.h")]
    if not force and os.path.exists(SOLID_SO) and os.path.getmtime(SOLID_SO) >= max(os.path.getmtime(f) for f in deps):
        return True
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-w", "-I", R.SHIM,
                           f'-DREF_SOURCE="{src}"', "-o", SOLID_SO, wrap, eig])
    return True


def ref_solid_available() -> bool:
    return os.path.exists(SOLID_SO)


class RefSolid:
    """The reference's SolidSystem on the same case (serial)."""

    def __init__(self, case: SolidCase):
        if not ref_solid_available():
            build_ref_solid()
        self.c = case
        L = self.L = C.CDLL(SOLID_SO)
        L.ref_solid_create.restype = C.c_void_p
        L.ref_solid_destroy.restype = None
        L.ref_solid_last_error.restype = C.c_char_p
        L.ref_solid_nnz.restype = C.c_int64
        sub = None
        self.mat_ids = list(range(case.mats.shape[0]))
        if case.mat_of is not None:
            sub = np.ascontiguousarray(case.mat_of, dtype=np.int32)
        self.h = C.c_void_p(L.ref_solid_create(C.c_int(case.elem_type), C.c_int64(case.N), C.c_int64(case.E), _p(case.conn), _p(case.xund), _p(sub)))
        assert self.h.value
        keys = ("Young", "Poisson", "FibreStiffness", "VolumetricStretchRatio/rate_0", "VolumetricStretchRatio/rate_1",
                "VolumetricStretchRatio/rate_2")
        for m in self.mat_ids:
            for k, key in enumerate(keys):
                self._chk(L.ref_solid_set_real(self.h, f"material/{m}/Hyperelastic/{key}".encode(), C.c_double(case.mats[m, k])))
        self._chk(L.ref_solid_set_bool(self.h, b"solver/assembly_use_symmetry", C.c_int(int(case.use_symmetry))))
        self._chk(L.ref_solid_set_real(self.h, b"BCs/displacement_penalty", C.c_double(case.penalty)))
        self._chk(L.ref_solid_set_string(self.h, b"BCs", (" " + " ".join(str(i) for i in case.bc_ids) + " ").encode()))
        for i, d in zip(case.bc_ids, case.bc_disp):
            self._chk(L.ref_solid_set_point(self.h, f"BC/{i}/displacement".encode(), C.c_double(d[0]), C.c_double(d[1]), C.c_double(d[2])))
        for e, s, b in zip(case.side_elem, case.side_no, case.side_bc):
            self._chk(L.ref_solid_add_side(self.h, C.c_int64(e), C.c_int(s), C.c_int(case.bc_ids[b])))
        if case.fibres is not None:
            self._chk(L.ref_solid_set_fibres(self.h, _p(case.fibres)))

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.ref_solid_last_error(self.h).decode())

    def __del__(self):
        try:
            if self.h.value:
                self.L.ref_solid_destroy(self.h)
                self.h = C.c_void_p(None)
        except Exception:
            pass

    def set_state(self, xcur, t):
        x = np.ascontiguousarray(xcur, dtype=np.float64).reshape(-1)
        self._chk(self.L.ref_solid_set_real(self.h, b"pseudo_time", C.c_double(t)))
        self._chk(self.L.ref_solid_set_positions(self.h, _p(x)))

    def element(self, e, xcur, t, want_jac=True):
        self.set_state(xcur, t)
        nen = self.c.conn.shape[1]
        Re = np.zeros(3 * nen); Ke = np.zeros((3 * nen, 3 * nen))
        self._chk(self.L.ref_solid_element(self.h, C.c_int64(e), C.c_int(int(want_jac)), _p(Re), _p(Ke)))
        return Re, Ke

    def assemble(self, xcur, t, want_jac=True):
        """-> (rowptr, col, val, rhs): the reference's global Jacobian as sorted CSR (entries it actually touched)."""
        self.set_state(xcur, t)
        self._chk(self.L.ref_solid_assemble(self.h, C.c_int(int(want_jac))))
        nnz = self.L.ref_solid_nnz(self.h)
        D = 3 * self.c.N
        rowptr = np.zeros(D + 1, dtype=np.int64); col = np.zeros(nnz, dtype=np.int32); val = np.zeros(nnz); rhs = np.zeros(D)
        self._chk(self.L.ref_solid_get_csr(self.h, _p(rowptr), _p(col), _p(val), _p(rhs)))
        return rowptr, col, val, rhs

    def post(self, xcur, t):
        self.set_state(xcur, t)
        E = self.c.E
        press, vm, fib = np.zeros(E), np.zeros(E), np.zeros((E, 3))
        self._chk(self.L.ref_solid_post_process(self.h, _p(press), _p(vm), _p(fib)))
        return press, vm, fib
