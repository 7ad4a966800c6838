"""Host-side mirror of the reference's SolidSystem (solid_system.h/.C, driven by solid.C:81-108 and coupled_hcc.C:117-132)
on top of the C ABI (include/rdc.h, rdc_solid_*).  Same verbs as the reference class -- save_initial_mesh, run_solver,
post_process -- everything numerical happens in librdcgpu.so; this file only marshals buffers.

    for l in 1..n_load_step:                       solid.C:83-108
        pseudo_time += deltat
        model.run_solver()                         Newton on the current node positions
        model.post_process()                       pressure / von Mises / current fibre per element
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib

SOLID = 5
TET4, HEX8 = 4, 8
KSP_GMRES, KSP_CG, KSP_BICGSTAB = 0, 1, 2


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class SolidSystem:
    def __init__(self, elem_type, conn, xyz, device: int = -1, rank: int = 0, nranks: int = 1, partitioner: int = 0, unique_id=None):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self.n_nodes, self.n_elems = xyz.shape[0], self.conn.shape[0]
        if nranks > 1:   # one process per GPU, like the reference's MPI ranks; every rank passes the same replicated mesh
            uid = C.create_string_buffer(unique_id, 128)
            rc = self._L.rdc_create_distributed(C.byref(self._h), SOLID, elem_type, self.n_nodes, self.n_elems, _ptr(self.conn), _ptr(xyz),
                                                None, device, rank, nranks, partitioner, C.cast(uid, C.c_void_p))
        else:
            rc = self._L.rdc_create(C.byref(self._h), SOLID, elem_type, self.n_nodes, self.n_elems, _ptr(self.conn), _ptr(xyz), None, device)
        if rc:
            raise _lib.RdcError(rc, self._L.rdc_last_error(None).decode())
        self.pseudo_time = 0.0
        self.ksp = KSP_GMRES
        # solid.C:226-245 defaults
        self.options = dict(max_nonlinear_iterations=100, relative_step_tolerance=1e-3, relative_residual_tolerance=1e-8,
                            absolute_residual_tolerance=1e-8, require_reduction=False, max_linear_iterations=50000,
                            initial_linear_tolerance=1e-3)
        self.set_positions(xyz)          # mesh_position_get: the unknowns start as the node positions
        self.save_initial_mesh(xyz)

    def _check(self, rc):
        if rc:
            raise _lib.RdcError(rc, self._L.rdc_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._L.rdc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ hand-over
    def save_initial_mesh(self, xyz):
        x = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1)
        self._check(self._L.rdc_solid_set_reference(self._h, _ptr(x)))

    def set_positions(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        assert x.size == 3 * self.n_nodes
        self._check(self._L.rdc_set_solution(self._h, _ptr(x)))

    def get_positions(self):
        x = np.empty(3 * self.n_nodes)
        self._check(self._L.rdc_get_solution(self._h, _ptr(x)))
        return x.reshape(-1, 3)

    def set_materials(self, mats, mat_of=None):
        m = np.ascontiguousarray(mats, dtype=np.float64).reshape(-1, 6)
        mo = None if mat_of is None else np.ascontiguousarray(mat_of, dtype=np.int32)
        self._check(self._L.rdc_solid_set_materials(self._h, m.shape[0], _ptr(m), _ptr(mo)))

    def set_fibres(self, fibres):
        f = None if fibres is None else np.ascontiguousarray(fibres, dtype=np.float64).reshape(-1)
        self._check(self._L.rdc_solid_set_fibres(self._h, _ptr(f)))

    def set_symmetry(self, use_symmetry):
        self._check(self._L.rdc_solid_set_symmetry(self._h, int(bool(use_symmetry))))

    def set_bcs(self, bc_disp, side_elem, side_no, side_bc, penalty):
        bd = np.ascontiguousarray(bc_disp, dtype=np.float64).reshape(-1, 3)
        se = np.ascontiguousarray(side_elem, dtype=np.int64)
        sn = np.ascontiguousarray(side_no, dtype=np.int32)
        sb = np.ascontiguousarray(side_bc, dtype=np.int32)
        self._check(self._L.rdc_solid_set_bcs(self._h, bd.shape[0], _ptr(bd), se.shape[0], _ptr(se), _ptr(sn), _ptr(sb), float(penalty)))

    def set_option(self, name, value):
        self._check(self._L.rdc_set_option(self._h, name.encode(), int(value)))

    # ------------------------------------------------------------------ the path
    def assemble(self, pseudo_time=None):
        t = self.pseudo_time if pseudo_time is None else pseudo_time
        self._check(self._L.rdc_solid_assemble(self._h, float(t)))

    def run_solver(self, pseudo_time=None):
        """One Newton solve at the given pseudo-time (SolidSystem::run_solver); returns the info dict."""
        if pseudo_time is not None:
            self.pseudo_time = float(pseudo_time)
        o = self.options
        ov = np.array([o["max_nonlinear_iterations"], o["relative_step_tolerance"], o["relative_residual_tolerance"],
                       o["absolute_residual_tolerance"], 1.0 if o["require_reduction"] else 0.0, o["max_linear_iterations"],
                       o["initial_linear_tolerance"]], dtype=np.float64)
        info = np.zeros(4)
        self._check(self._L.rdc_solid_newton(self._h, self.pseudo_time, _ptr(ov), self.ksp, _ptr(info)))
        return dict(newton_its=int(info[0]), linear_its=int(info[1]), residual=float(info[2]), converged=bool(info[3]))

    def post_process(self, pseudo_time=None):
        t = self.pseudo_time if pseudo_time is None else pseudo_time
        press, vm, fib = np.zeros(self.n_elems), np.zeros(self.n_elems), np.zeros((self.n_elems, 3))
        self._check(self._L.rdc_solid_post_process(self._h, float(t), _ptr(press), _ptr(vm), _ptr(fib)))
        return press, vm, fib

    # ------------------------------------------------------------------ parity helpers
    def download_csr(self):
        L = self._L
        n_rows, nnz = C.c_int64(), C.c_int64()
        rows, rowptr, col, val, rhs = (C.c_void_p() for _ in range(5))
        self._check(L.rdc_download_csr(self._h, C.byref(n_rows), C.byref(nnz), C.byref(rows), C.byref(rowptr), C.byref(col),
                                       C.byref(val), C.byref(rhs)))
        def take(p, ctype, n):
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(n,)).copy()
            L.rdc_free(p)
            return a
        nr, nz = n_rows.value, nnz.value
        return (take(rows, C.c_int64, nr), take(rowptr, C.c_int64, nr + 1), take(col, C.c_int32, nz), take(val, C.c_double, nz),
                take(rhs, C.c_double, nr))

    def stats(self):
        st = _lib.Stats()
        self._check(self._L.rdc_get_stats(self._h, C.byref(st)))
        return st


def from_case(case, device: int = -1, **dist) -> "SolidSystem":
    """Build a SolidSystem from an oracle.solid.SolidCase-like description (duck-typed: tests only)."""
    s = SolidSystem(case.elem_type, case.conn, case.xund, device=device, **dist)
    s.set_materials(case.mats, case.mat_of)
    s.set_fibres(case.fibres)
    s.set_symmetry(getattr(case, "use_symmetry", False))
    se, sn, sb, bd = case.arrays()
    s.set_bcs(bd, se, sn, sb, case.penalty)
    s.options.update(case.opts)
    return s


# ---- host-only probes (no GPU): the same element arithmetic the kernels run, compiled for the host -------------------
def probe_row(elem_type, Xc, Xu, mat6, pseudo_time, eta, li, use_symmetry=False):
    L = _lib.load()
    nen = 4 if elem_type == TET4 else 8
    Xc = np.ascontiguousarray(Xc, dtype=np.float64); Xu = np.ascontiguousarray(Xu, dtype=np.float64)
    m = np.ascontiguousarray(mat6, dtype=np.float64); e = np.ascontiguousarray(eta, dtype=np.float64)
    R, K = np.zeros(3), np.zeros(9 * nen)
    rc = L.rdc_solid_probe_row(elem_type, _ptr(Xc), _ptr(Xu), _ptr(m), float(pseudo_time), _ptr(e), int(li), int(bool(use_symmetry)), _ptr(R), _ptr(K))
    assert rc == 0
    return R, K.reshape(3, 3, nen)


def probe_bc_row(ns, Xc, Xu, disp, pseudo_time, penalty, i):
    L = _lib.load()
    Xc = np.ascontiguousarray(Xc, dtype=np.float64); Xu = np.ascontiguousarray(Xu, dtype=np.float64)
    d = np.ascontiguousarray(disp, dtype=np.float64)
    R, Kd = np.zeros(3), np.zeros(ns * 3)
    rc = L.rdc_solid_probe_bc_row(ns, _ptr(Xc), _ptr(Xu), _ptr(d), float(pseudo_time), float(penalty), int(i), _ptr(R), _ptr(Kd))
    assert rc == 0
    return R, Kd.reshape(ns, 3)


def probe_post(elem_type, Xc, Xu, mat6, pseudo_time, eta):
    L = _lib.load()
    Xc = np.ascontiguousarray(Xc, dtype=np.float64); Xu = np.ascontiguousarray(Xu, dtype=np.float64)
    m = np.ascontiguousarray(mat6, dtype=np.float64); e = np.ascontiguousarray(eta, dtype=np.float64)
    out = np.zeros(5)
    rc = L.rdc_solid_probe_post(elem_type, _ptr(Xc), _ptr(Xu), _ptr(m), float(pseudo_time), _ptr(e), _ptr(out))
    assert rc == 0
    return out


def probe_bc_rows(elem_type, conn, xyz, rank, nranks, partitioner, side_elem, side_no):
    """Penalty-row lists of one rank (host only): {global node: [(side index, position in the side), ...]} in adding order."""
    L = _lib.load()
    conn = np.ascontiguousarray(conn, dtype=np.int32); xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    se = np.ascontiguousarray(side_elem, dtype=np.int64); sn = np.ascontiguousarray(side_no, dtype=np.int32)
    n = C.c_int32()
    rn, rp, es, ep = (C.c_void_p() for _ in range(4))
    rc = L.rdc_solid_probe_bc_rows(elem_type, xyz.shape[0], conn.shape[0], _ptr(conn), _ptr(xyz), rank, nranks, partitioner, se.shape[0],
                                   _ptr(se), _ptr(sn), C.byref(n), C.byref(rn), C.byref(rp), C.byref(es), C.byref(ep))
    assert rc == 0, rc
    def take(p, count):
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(max(count, 1),))[:count].copy()
        L.rdc_free(p)
        return a
    nr = n.value
    row_ptr = take(rp, nr + 1)
    nodes, side, pos = take(rn, nr), take(es, int(row_ptr[-1]) if nr else 0), take(ep, int(row_ptr[-1]) if nr else 0)
    return {int(nodes[k]): list(zip(side[row_ptr[k]:row_ptr[k + 1]].tolist(), pos[row_ptr[k]:row_ptr[k + 1]].tolist())) for k in range(nr)}
