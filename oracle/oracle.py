"""ctypes front-end of the CPU oracle (oracle/rdc_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (rdcfes_b200/) never imports this module.
Parity status: pinned to the reference's own compiled sources (oracle/ref.py, tests/test_ref_pin.py) and by the
known-answer tests of tests/test_oracle_kat.py; see the header of rdc_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ADPM, PIHNA, RIPF, PROTEAS, HCC = 0, 1, 2, 3, 4
TET4, HEX8 = 4, 8
PC_ILU, PC_JACOBI, PC_NONE = 0, 1, 2

_f64 = np.float64
_p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "rdc_oracle.c")
    src2 = os.path.join(_HERE, "solid_oracle.c")
    if force or not os.path.exists(so) or any(os.path.exists(s) and os.path.getmtime(so) < os.path.getmtime(s) for s in (src, src2)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_model_nvars.restype = C.c_int
        for name in ("orc_build_pattern", "orc_node_graph", "orc_assemble", "orc_element", "orc_fe_tables",
                     "orc_ripf_check", "orc_gmres", "orc_step", "orc_region_volumes", "orc_region_last_mean"):
            getattr(L, name).restype = C.c_int
        L.orc_free.restype = None
        L.orc_clamp_nonneg.restype = None
        L.orc_spmv.restype = None
        _LIB = L
    return _LIB


def make_conditions(conds, nv):
    """[(weights, div, lo, hi), ...] -> flat [ncond*8] array {w[5], div, lo, hi} (save_solution range tests)."""
    out = np.zeros((len(conds), 8))
    for k, (w, div, lo, hi) in enumerate(conds):
        out[k, :nv] = w
        out[k, 5], out[k, 6], out[k, 7] = div, lo, hi
    return out.ravel()


def region_volumes(elem_type, conn, xyz, u, conds, region=None, n_regions=1):
    """save_solution thresholded volumes (adpm.C:786-812, pihna.C:885-945, ripf.C:822-847): serial reference loop."""
    conn = np.ascontiguousarray(conn, dtype=np.int32); xyz = np.ascontiguousarray(xyz, dtype=_f64)
    u = np.ascontiguousarray(u, dtype=_f64)
    nv = u.size // xyz.shape[0]
    reg = None if region is None else np.ascontiguousarray(region, dtype=np.int32)
    c = make_conditions(conds, nv)
    vol = np.zeros(n_regions)
    rc = lib().orc_region_volumes(C.c_int(elem_type), C.c_int(nv), C.c_int64(xyz.shape[0]), C.c_int64(conn.shape[0]),
                                  _p(conn), _p(xyz), _p(u), _p(reg), C.c_int(n_regions), C.c_int(len(conds)), _p(c), _p(vol))
    assert rc == 0
    return vol


def region_last_mean(elem_type, conn, xyz, u, var, region=None, n_regions=1):
    """adpm.C:763-783: element average of variable `var` in the LAST element of every region."""
    conn = np.ascontiguousarray(conn, dtype=np.int32); xyz = np.ascontiguousarray(xyz, dtype=_f64)
    u = np.ascontiguousarray(u, dtype=_f64)
    nv = u.size // xyz.shape[0]
    reg = None if region is None else np.ascontiguousarray(region, dtype=np.int32)
    mean = np.zeros(n_regions)
    rc = lib().orc_region_last_mean(C.c_int(elem_type), C.c_int(nv), C.c_int64(xyz.shape[0]), C.c_int64(conn.shape[0]),
                                    _p(conn), _p(xyz), _p(u), _p(reg), C.c_int(n_regions), C.c_int(var), _p(mean))
    assert rc == 0
    return mean


def nvars(model: int) -> int:
    return lib().orc_model_nvars(C.c_int(model))


def fe_tables(elem_type: int):
    nen, nqp = C.c_int(), C.c_int()
    w = np.zeros(8)
    phi = np.zeros(64)
    rc = lib().orc_fe_tables(C.c_int(elem_type), C.byref(nen), C.byref(nqp), _p(w), _p(phi))
    assert rc == 0
    return w[: nqp.value].copy(), phi[: nen.value * nqp.value].reshape(nen.value, nqp.value).copy()


def build_pattern(n_nodes: int, conn: np.ndarray, nv: int):
    """Scalar CSR pattern (node graph + I) (x) dense nv x nv in dof = nv*node+var numbering."""
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    E, nen = conn.shape
    nnz = C.c_int64()
    rp = C.c_void_p()
    cl = C.c_void_p()
    rc = lib().orc_build_pattern(C.c_int64(n_nodes), C.c_int64(E), C.c_int(nen), C.c_int(nv), _p(conn),
                                 C.byref(nnz), C.byref(rp), C.byref(cl))
    assert rc == 0
    D = n_nodes * nv
    rowptr = np.ctypeslib.as_array(C.cast(rp, C.POINTER(C.c_int64)), shape=(D + 1,)).copy()
    col = np.ctypeslib.as_array(C.cast(cl, C.POINTER(C.c_int32)), shape=(nnz.value,)).copy()
    lib().orc_free(rp)
    lib().orc_free(cl)
    return rowptr, col


def element(model, elem_type, X, U, params, efield=None, aux=None, rt_max=0, time=0.0, dt=0.1):
    """Dense (Ke, Fe, JxW, dphi) of ONE element.  U is [nen, nv] node-major."""
    nv = nvars(model)
    nen = elem_type
    nqp = 5 if elem_type == TET4 else 8
    X = np.ascontiguousarray(X, dtype=_f64).reshape(nen, 3)
    U = np.ascontiguousarray(U, dtype=_f64).reshape(nen, nv)
    params = np.ascontiguousarray(params, dtype=_f64)
    nd = nen * nv
    Ke = np.zeros((nd, nd))
    Fe = np.zeros(nd)
    JxW = np.zeros(nqp)
    dphi = np.zeros((nen, nqp, 3))
    ef = np.ascontiguousarray(efield, dtype=_f64) if efield is not None else np.zeros(3)
    ax = np.ascontiguousarray(aux, dtype=_f64) if aux is not None else None
    rc = lib().orc_element(C.c_int(model), C.c_int(elem_type), _p(X), _p(U), _p(params), _p(ef), _p(ax),
                           C.c_int(rt_max), C.c_double(time), C.c_double(dt), _p(Ke), _p(Fe), _p(JxW), _p(dphi))
    assert rc == 0
    return Ke, Fe, JxW, dphi


class Problem:
    """State of one oracle run: mesh, pattern, fields.  Mirrors the objects the reference keeps in
    EquationSystems (adpm.C:17-43): the mesh, the TransientLinearImplicitSystem vectors and the aux systems."""

    def __init__(self, model, elem_type, conn, xyz, params, u0, elem_field=None, nodal_field=None, nthreads=1):
        self.model, self.elem_type = model, elem_type
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.xyz = np.ascontiguousarray(xyz, dtype=_f64)
        self.N = self.xyz.shape[0]
        self.E = self.conn.shape[0]
        self.nv = nvars(model)
        self.D = self.N * self.nv
        self.params = np.ascontiguousarray(params, dtype=_f64)
        self.u = np.ascontiguousarray(u0, dtype=_f64).reshape(-1).copy()
        assert self.u.size == self.D
        self.u_old = self.u.copy()
        self.elem_field = None if elem_field is None else np.ascontiguousarray(elem_field, dtype=_f64)
        self.nthreads = nthreads
        self.rowptr, self.col = build_pattern(self.N, self.conn, self.nv)
        self.val = np.zeros(self.rowptr[-1])
        self.rhs = np.zeros(self.D)
        self.time = 0.0
        self.rt_max = C.c_int(0)
        self.prev = None
        self.aux = None
        if model == RIPF:
            # aux = {TD[3], RT_broad, RT_focus, RT_total}; ripf.C:50-53: check_solution before the loop
            self.aux = np.zeros((self.N, 6))
            self.aux[:, 3:5] = np.asarray(nodal_field, dtype=_f64).reshape(self.N, 2)
            self.prev = self.u.copy()
        elif model == PROTEAS:
            self.aux = np.ascontiguousarray(nodal_field, dtype=_f64).reshape(self.N, 2).copy()

    # ripf.C:53 -- the pre-loop check_solution (time = 0, prev = initial solution)
    def ripf_initial_check(self, dt):
        m = lib().orc_ripf_check(C.c_int64(self.N), _p(self.u), _p(self.prev), _p(self.aux), _p(self.params),
                                 C.c_double(0.0), C.c_double(dt))
        if m < 0:
            raise RuntimeError("RT_total_max <= 0 (ripf.C:773)")
        self.rt_max = C.c_int(m)

    def assemble(self, time, dt, u_old=None):
        uo = self.u_old if u_old is None else np.ascontiguousarray(u_old, dtype=_f64)
        rc = lib().orc_assemble(C.c_int(self.model), C.c_int(self.elem_type), C.c_int64(self.N), C.c_int64(self.E),
                                _p(self.conn), _p(self.xyz), _p(uo), _p(self.params), _p(self.elem_field),
                                _p(self.aux), self.rt_max, C.c_double(time), C.c_double(dt), _p(self.rowptr),
                                _p(self.col), _p(self.val), _p(self.rhs), C.c_int(self.nthreads))
        assert rc == 0, rc
        return self.val, self.rhs

    def solve(self, pc=PC_ILU, nblocks=1, restart=30, rtol=1e-12, maxits=5000, x0=None):
        x = self.u if x0 is None else np.ascontiguousarray(x0, dtype=_f64)
        its, res, res0 = C.c_int(), C.c_double(), C.c_double()
        rc = lib().orc_gmres(C.c_int64(self.D), _p(self.rowptr), _p(self.col), _p(self.val), _p(self.rhs), _p(x),
                             C.c_int(pc), C.c_int(nblocks), C.c_int(restart), C.c_double(rtol), C.c_int(maxits),
                             C.c_int(self.nthreads), C.byref(its), C.byref(res), C.byref(res0))
        assert rc >= 0, rc
        return x, its.value, res.value, res0.value

    def step(self, dt, pc=PC_ILU, nblocks=1, restart=30, rtol=1e-12, maxits=5000):
        """adpm.C:63-76: time += dt; rotate; solve (assemble + KSP); check_solution."""
        self.time += dt
        its, res = C.c_int(), C.c_double()
        ta, ts = C.c_double(), C.c_double()
        rc = lib().orc_step(C.c_int(self.model), C.c_int(self.elem_type), C.c_int64(self.N), C.c_int64(self.E),
                            _p(self.conn), _p(self.xyz), _p(self.u), _p(self.u_old), _p(self.params),
                            _p(self.elem_field), _p(self.aux), _p(self.prev), C.byref(self.rt_max),
                            C.c_double(self.time), C.c_double(dt), _p(self.rowptr), _p(self.col), _p(self.val),
                            _p(self.rhs), C.c_int(pc), C.c_int(nblocks), C.c_int(restart), C.c_double(rtol),
                            C.c_int(maxits), C.c_int(self.nthreads), C.byref(its), C.byref(res), C.byref(ta),
                            C.byref(ts))
        assert rc >= 0, rc
        self.t_assemble, self.t_solve = ta.value, ts.value
        return its.value, res.value

    def spmv(self, x):
        y = np.zeros(self.D)
        x = np.ascontiguousarray(x, dtype=_f64)
        lib().orc_spmv(C.c_int64(self.D), _p(self.rowptr), _p(self.col), _p(self.val), _p(x), _p(y),
                       C.c_int(self.nthreads))
        return y

    def scipy_csr(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.val.copy(), self.col.copy(), self.rowptr.copy()), shape=(self.D, self.D))
