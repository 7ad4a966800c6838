"""Host-side mirror of how the reference drives its RDC systems, on top of the C ABI (include/rdc.h).

The reference keeps a libMesh TransientLinearImplicitSystem per model and runs (adpm.C:60-84)

    time += dt; *older = *old; *old = *current; model.solve(); check_solution(es)

`TransientRdcSystem` exposes the same verbs (solve / check_solution / the time-level rotation) and the same
defaults (GMRES restart 30, rtol 1e-12, 5000 iterations -- libMesh's "linear solver tolerance" /
"linear solver maximum iterations"), so that the parity tests read like a driver
of the reference.  Everything numerical happens in librdcgpu.so; this file only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import lib as _lib
from . import params as P

ADPM, PIHNA, RIPF, PROTEAS, HCC = 0, 1, 2, 3, 4
TET4, HEX8 = 4, 8
KSP_GMRES, KSP_CG, KSP_BICGSTAB = 0, 1, 2
PC_JACOBI, PC_NONE = 0, 1


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_unique_id() -> bytes:
    """ncclUniqueId for rdc_create_distributed (rank 0 makes it, the caller broadcasts it)."""
    L = _lib.load()
    buf = C.create_string_buffer(128)
    rc = L.rdc_comm_unique_id(buf)
    if rc:
        raise _lib.RdcError(rc, L.rdc_last_error(None).decode())
    return buf.raw


class TransientRdcSystem:
    def __init__(self, model: int, elem_type: int, conn, xyz, node_dof_base=None, device: int = -1,
                 rank: int = 0, nranks: int = 1, partitioner: int = 0, unique_id: Optional[bytes] = None):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self.model, self.elem_type = model, elem_type
        conn = np.ascontiguousarray(conn, dtype=np.int32)
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        assert conn.ndim == 2 and conn.shape[1] == elem_type and xyz.ndim == 2 and xyz.shape[1] == 3
        base = None if node_dof_base is None else np.ascontiguousarray(node_dof_base, dtype=np.int32)
        self.n_nodes, self.n_elems = xyz.shape[0], conn.shape[0]
        self.nv = P.NVARS[model]
        if nranks > 1:
            uid = C.create_string_buffer(unique_id, 128)
            rc = self._L.rdc_create_distributed(C.byref(self._h), model, elem_type, self.n_nodes, self.n_elems,
                                                _ptr(conn), _ptr(xyz), _ptr(base), device, rank, nranks, partitioner,
                                                C.cast(uid, C.c_void_p))
        else:
            rc = self._L.rdc_create(C.byref(self._h), model, elem_type, self.n_nodes, self.n_elems, _ptr(conn),
                                    _ptr(xyz), _ptr(base), device)
        if rc:
            raise _lib.RdcError(rc, self._L.rdc_last_error(None).decode())
        self.n_dofs = int(self._L.rdc_n_dofs(self._h))
        self.time = 0.0
        self.dt = 0.0
        # libMesh defaults (Appendix B-7/8); Jacobi replaces ILU(0) on the device
        self.ksp, self.pc, self.restart = KSP_GMRES, PC_JACOBI, 30
        self.rtol, self.maxits = 1e-12, 5000
        self.iterations, self.resnorm = 0, 0.0

    # ------------------------------------------------------------------ plumbing
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
        except Exception:  # noqa: BLE001
            pass

    # ------------------------------------------------------------------ hand-over
    def set_parameters(self, flat):
        flat = np.ascontiguousarray(flat, dtype=np.float64)
        self._check(self._L.rdc_set_params(self._h, _ptr(flat), flat.size))

    def set_elem_field(self, f, slot=0):
        f = np.ascontiguousarray(f, dtype=np.float64)
        self._check(self._L.rdc_set_elem_field(self._h, slot, _ptr(f), f.shape[1]))

    def set_nodal_field(self, f, slot=0):
        f = np.ascontiguousarray(f, dtype=np.float64)
        self._check(self._L.rdc_set_nodal_field(self._h, slot, _ptr(f), f.shape[1]))

    def update_coords(self, xyz):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self._check(self._L.rdc_update_coords(self._h, _ptr(xyz)))

    def set_solution(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
        assert u.size == self.n_dofs
        self._check(self._L.rdc_set_solution(self._h, _ptr(u)))

    def get_solution(self, out=None):
        out = np.empty(self.n_dofs) if out is None else out
        self._check(self._L.rdc_get_solution(self._h, _ptr(out)))
        return out

    def get_solution_owned(self, out):
        """Refresh only the entries owned by this rank in the global-dof-indexed array `out` (no all-gather)."""
        self._check(self._L.rdc_get_solution_owned(self._h, _ptr(out)))
        return out

    def get_old_solution(self):
        out = np.empty(self.n_dofs)
        self._check(self._L.rdc_get_old_solution(self._h, _ptr(out)))
        return out

    # ------------------------------------------------------------------ the reference's verbs
    def rotate(self):
        """*older = *old; *old = *current (adpm.C:71-72)."""
        self._check(self._L.rdc_rotate(self._h))

    def assemble(self, time=None, dt=None):
        self.time = self.time if time is None else time
        self.dt = self.dt if dt is None else dt
        self._check(self._L.rdc_assemble(self._h, self.time, self.dt))

    def linear_solve(self):
        its, res = C.c_int(), C.c_double()
        rc = self._L.rdc_solve(self._h, self.ksp, self.pc, self.rtol, self.maxits, self.restart, C.byref(its),
                               C.byref(res))
        self.iterations, self.resnorm = its.value, res.value
        self._check(rc)
        return its.value, res.value

    def solve(self):
        """TransientLinearImplicitSystem::solve(): zero K,F -> assemble callback -> KSP (adpm.C:74)."""
        self.assemble()
        return self.linear_solve()

    def check_solution(self):
        """The model's check_solution at the current time (adpm.C:76; ripf.C:53,85)."""
        self._check(self._L.rdc_set_time(self._h, self.time))
        self._check(self._L.rdc_set_dt(self._h, self.dt))
        self._check(self._L.rdc_clamp(self._h))

    def step(self, dt):
        """One pass of the loop body adpm.C:63-76 through the fused entry point rdc_step."""
        self.time += dt
        self.dt = dt
        its, res = C.c_int(), C.c_double()
        rc = self._L.rdc_step(self._h, self.time, dt, self.ksp, self.pc, self.rtol, self.maxits, self.restart,
                              C.byref(its), C.byref(res))
        self.iterations, self.resnorm = its.value, res.value
        self._check(rc)
        return its.value, res.value

    # ------------------------------------------------------------------ save_solution reductions
    def set_subdomains(self, region, n_regions):
        """elem->subdomain_id() renumbered 0..n_regions-1 (adpm.C:298-320 builds the parcellation set)."""
        region = None if region is None else np.ascontiguousarray(region, dtype=np.int32)
        self._n_regions = int(n_regions)
        self._check(self._L.rdc_set_subdomains(self._h, _ptr(region), int(n_regions)))

    def region_volumes(self, conds):
        """conds = [(weights[nv], div, lo, hi), ...]; returns vol[n_regions] (adpm.C:786-812, pihna.C:885-945, ripf.C:822-847)."""
        n_regions = getattr(self, "_n_regions", 1)
        flat = np.zeros((len(conds), 8))
        for k, (w, div, lo, hi) in enumerate(conds):
            flat[k, :self.nv] = w
            flat[k, 5], flat[k, 6], flat[k, 7] = div, lo, hi
        vol = np.zeros(n_regions)
        self._check(self._L.rdc_region_volumes(self._h, len(conds), _ptr(flat), _ptr(vol)))
        return vol

    def region_last_mean(self, var):
        """adpm.C:763-783: element average of variable `var` in the last element of every region."""
        mean = np.zeros(getattr(self, "_n_regions", 1))
        self._check(self._L.rdc_region_last_mean(self._h, int(var), _ptr(mean)))
        return mean

    # ------------------------------------------------------------------ parity / measurement
    def get_rhs(self):
        """system.rhs after the assemble callback (global dof order)."""
        f = np.empty(self.n_dofs)
        self._check(self._L.rdc_get_rhs(self._h, _ptr(f)))
        return f

    def spmv(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty(self.n_dofs)
        self._check(self._L.rdc_spmv(self._h, _ptr(x), _ptr(y)))
        return y

    def bench_spmv(self, reps=20):
        ms = C.c_double()
        self._check(self._L.rdc_bench_spmv(self._h, reps, C.byref(ms)))
        return ms.value

    def bench_dfma(self):
        """measured fp64-pipe rate of this GPU, TFLOP/s"""
        t = C.c_double()
        self._check(self._L.rdc_bench_dfma(self._h, C.byref(t)))
        return t.value

    def bench_stream(self, reps=20, ctas_per_sm=8):
        """(mean ms, bytes) of a read-only pass over the stored operator values."""
        ms, nb = C.c_double(), C.c_int64()
        self._check(self._L.rdc_bench_stream(self._h, reps, ctas_per_sm, C.byref(ms), C.byref(nb)))
        return ms.value, nb.value

    def download_csr(self):
        """(rows, rowptr, col, val, rhs): scalar CSR of the rows owned by this rank, global dof numbering."""
        n_rows, nnz = C.c_int64(), C.c_int64()
        rows, rowptr, col, val, rhs = (C.c_void_p() for _ in range(5))
        self._check(self._L.rdc_download_csr(self._h, C.byref(n_rows), C.byref(nnz), C.byref(rows), C.byref(rowptr),
                                             C.byref(col), C.byref(val), C.byref(rhs)))
        nr, nz = n_rows.value, nnz.value

        def take(p, ctype, n):
            a = np.ctypeslib.as_array(C.cast(p, C.POINTER(ctype)), shape=(n,)).copy()
            self._L.rdc_free(p)
            return a

        return (take(rows, C.c_int64, nr), take(rowptr, C.c_int64, nr + 1), take(col, C.c_int32, nz),
                take(val, C.c_double, nz), take(rhs, C.c_double, nr))

    def stats(self) -> "_lib.Stats":
        s = _lib.Stats()
        self._check(self._L.rdc_get_stats(self._h, C.byref(s)))
        return s

    def set_option(self, name: str, value: int):
        self._check(self._L.rdc_set_option(self._h, name.encode(), int(value)))

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self._L.rdc_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))


def probe_partition(elem_type, nvars, conn, xyz, rank, nranks, partitioner=0):
    """Host-only view of the partition / halo lists of one rank (rdc_probe_partition); needs no GPU."""
    L = _lib.load()
    conn = np.ascontiguousarray(conn, dtype=np.int32)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    n_owned, n_ghost, n_nbr = C.c_int32(), C.c_int32(), C.c_int32()
    n_el = C.c_int64()
    ptrs = [C.c_void_p() for _ in range(6)]
    L.rdc_probe_partition.restype = C.c_int
    rc = L.rdc_probe_partition(C.c_int(elem_type), C.c_int(nvars), C.c_int64(xyz.shape[0]), C.c_int64(conn.shape[0]),
                               _ptr(conn), _ptr(xyz), C.c_int(rank), C.c_int(nranks), C.c_int(partitioner),
                               C.byref(n_owned), C.byref(n_ghost), C.byref(n_el), C.byref(ptrs[0]), C.byref(n_nbr),
                               C.byref(ptrs[1]), C.byref(ptrs[2]), C.byref(ptrs[3]), C.byref(ptrs[4]), C.byref(ptrs[5]))
    if rc:
        raise _lib.RdcError(rc, L.rdc_last_error(None).decode())

    def take(p, n):
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(max(n, 1),))[:n].copy()
        L.rdc_free(p)
        return a

    nn = n_nbr.value
    owner = take(ptrs[0], xyz.shape[0])
    nbr = take(ptrs[1], nn)
    send_ptr = take(ptrs[2], nn + 1)
    send_glob = take(ptrs[3], int(send_ptr[-1]) if nn else 0)
    recv_ptr = take(ptrs[4], nn + 1)
    recv_glob = take(ptrs[5], n_ghost.value)
    return dict(n_owned=n_owned.value, n_ghost=n_ghost.value, n_elems_local=n_el.value, owner=owner, nbr=nbr,
                send_ptr=send_ptr, send_glob=send_glob, recv_ptr=recv_ptr, recv_glob=recv_glob)
