"""oracle/_ref -- the reference's OWN model sources, compiled from where they lie, behind ctypes.

TEST INFRASTRUCTURE ONLY (tests/, tests/golden/make_ref_golden.py).  The product package never imports this.

`build_ref()` compiles /root/reference/src/{adpm,pihna,ripf,proteas,coupled_hcc}.C UNCHANGED with plain g++ against
the serial libMesh stand-in under oracle/ref_shim/ (the reference's own build needs libMesh + PETSc + MPI, which are
not installable here) into oracle/_ref/libref_<model>.so (git-ignored, travels to the GPU box with gpurun).  What
runs is therefore the reference's in-tree arithmetic -- assemble_* (adpm.C:324-652, pihna.C:318-758, ripf.C:337-673,
proteas.C:338-705, coupled_hcc.C:414-649), check_solution, save_solution, input() and utils.h -- on top of a
restatement of the upstream pieces (FE tables, FEMap, dof numbering, ADD_VALUES).  The Krylov solve is PETSc's and
is not part of it.

`RefProblem` mirrors oracle.Problem so that tests can run both on the same inputs and compare K, F, the clamped
solution and the RIPF state.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

from rdcfes_b200 import params as P

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("RDC_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(_HERE, "_ref")
SHIM = os.path.join(_HERE, "ref_shim")
MODELS = {P.ADPM: ("adpm", "adpm.C", "ADPM"), P.PIHNA: ("pihna", "pihna.C", "PIHNA"), P.RIPF: ("ripf", "ripf.C", "RIPF"),
          P.PROTEAS: ("proteas", "proteas.C", "PROTEAS_model"), P.HCC: ("hcc", "coupled_hcc.C", "HCC")}
TET4, HEX8 = 4, 8
_LIBS = {}
_p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None


def have_reference_sources() -> bool:
    return all(os.path.exists(os.path.join(REF_ROOT, "src", src)) for _, src, _ in MODELS.values())


def so_path(model: int) -> str:
    return os.path.join(OUT, f"libref_{MODELS[model][0]}.so")


def available() -> bool:
    return all(os.path.exists(so_path(m)) for m in MODELS)


def build_ref(force: bool = False) -> bool:
    """g++ on the reference's own files (no cmake, no external library).  Returns False when the reference
    sources are not on this machine (the GPU box): the prebuilt .so files are used then."""
    if not have_reference_sources():
        return available()
    os.makedirs(OUT, exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    deps = [os.path.join(SHIM, "ref_api.inc"), os.path.join(SHIM, "libmesh", "shim.h")]
    for m, (name, src, _) in MODELS.items():
        so, wrap, ref = so_path(m), os.path.join(SHIM, f"ref_{name}.cpp"), os.path.join(REF_ROOT, "src", src)
        newest = max(os.path.getmtime(f) for f in deps + [wrap, ref])
        if not force and os.path.exists(so) and os.path.getmtime(so) >= newest:
            continue
        # -ffp-contract=off: the reference is built for baseline x86-64 (no FMA); keep the same rounding
        cmd = [cxx, "-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-shared", "-w", "-I", SHIM,
               f'-DREF_SOURCE="{ref}"', "-o", so, wrap]
        subprocess.check_call(cmd)
    return True


def _lib(model: int):
    if model not in _LIBS:
        if not available():
            build_ref()
        L = C.CDLL(so_path(model))
        pre = f"ref_{MODELS[model][0]}_"
        fn = lambda n: getattr(L, pre + n)
        fn("create").restype = C.c_void_p
        fn("destroy").restype = None
        fn("last_error").restype = C.c_char_p
        fn("n_dofs").restype = C.c_int64
        fn("nnz").restype = C.c_int64
        _LIBS[model] = (L, pre)
    return _LIBS[model]


class RefProblem:
    """The reference's EquationSystems for one model on a flattened mesh (serial)."""

    def __init__(self, model, elem_type, conn, xyz, params, u0, elem_field=None, nodal_field=None, subdomain=None,
                 node_dof_base=None, dt=None):
        self.model, self.elem_type = model, elem_type
        self.L, self.pre = _lib(model)
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self.N, self.E = self.xyz.shape[0], self.conn.shape[0]
        self.nv = P.NVARS[model]
        self.D = self.N * self.nv
        self.sysname = MODELS[model][2].encode()
        sub = None if subdomain is None else np.ascontiguousarray(subdomain, dtype=np.int32)
        base = None if node_dof_base is None else np.ascontiguousarray(node_dof_base, dtype=np.int32)
        self.h = C.c_void_p(self._f("create")(C.c_int(elem_type), C.c_int64(self.N), C.c_int64(self.E), _p(self.conn),
                                               _p(self.xyz), _p(sub), _p(base)))
        assert self.h.value, "reference context could not be created"
        self.set_params(params)
        self.time = 0.0
        self.u = np.ascontiguousarray(u0, dtype=np.float64).reshape(-1).copy()
        self.set_vector(self.sysname, 0, self.u)
        self.set_vector(self.sysname, 1, self.u)
        if model == P.ADPM and elem_field is not None:
            self.set_vector(b"Tracts", 0, np.ascontiguousarray(elem_field, dtype=np.float64).reshape(-1))
        if model == P.RIPF:
            rt = np.zeros((self.N, 3))
            rt[:, :2] = np.asarray(nodal_field, dtype=np.float64).reshape(self.N, 2)
            self.set_vector(b"RT", 0, rt.ravel())
            self._chk(self._f("set_prev")(self.h, _p(self.u), C.c_int64(self.D)))   # ripf.C:50-51
        if model == P.PROTEAS:
            self.set_vector(b"AUX", 0, np.ascontiguousarray(nodal_field, dtype=np.float64).reshape(-1))
        if dt is not None:
            self.set_real("time_step", dt)

    def _f(self, name):
        return getattr(self.L, self.pre + name)

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("reference call failed: " + self._f("last_error")(self.h).decode())

    def close(self):
        if self.h:
            self._f("destroy")(self.h)
            self.h = None

    def set_real(self, key, v):
        self._chk(self._f("set_real")(self.h, key.encode(), C.c_double(v)))

    def set_int(self, key, v):
        self._chk(self._f("set_int")(self.h, key.encode(), C.c_int(int(v))))

    def set_string(self, key, v):
        self._chk(self._f("set_string")(self.h, key.encode(), v.encode()))

    def get_real(self, key):
        v = C.c_double()
        self._chk(self._f("get_real")(self.h, key.encode(), C.byref(v)))
        return v.value

    def get_int(self, key):
        v = C.c_int()
        self._chk(self._f("get_int")(self.h, key.encode(), C.byref(v)))
        return v.value

    def set_params(self, flat):
        """flat vector of include/rdc.h -> es.parameters by GetPot key (rdcfes_b200/params.py TABLES; angles are
        already radians in both)."""
        flat = np.asarray(flat, dtype=np.float64)
        table = P.TABLES[self.model]
        assert flat.size == len(table)
        for (key, _), v in zip(table, flat):
            if key in ("RT_dose/broad/fractions", "RT_dose/focus/fractions"):   # ints in the reference (ripf.C:173-174)
                self.set_int(key, v)
            else:
                self.set_real(key, float(v))

    def input(self, path):
        """The reference's own input(): run in a scratch directory (it mkdir's and cp's, adpm.C:95-101)."""
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(tmp)
            try:
                with open(path) as src, open("input.dat", "w") as dst:
                    dst.write(src.read())
                self._chk(self._f("input")(self.h, b"input.dat"))
            finally:
                os.chdir(cwd)

    def set_vector(self, system, which, v):
        v = np.ascontiguousarray(v, dtype=np.float64)
        self._chk(self._f("set_vector")(self.h, system, C.c_int(which), _p(v), C.c_int64(v.size)))

    def get_vector(self, system, which, n=None):
        n = int(self._f("n_dofs")(self.h, system)) if n is None else n
        out = np.zeros(n)
        self._chk(self._f("get_vector")(self.h, system, C.c_int(which), _p(out), C.c_int64(n)))
        return out

    def rotate(self):
        self._chk(self._f("rotate")(self.h))

    def assemble(self, time, dt):
        """zero K, F + assemble_<model>() as LinearImplicitSystem::solve() does; returns sorted CSR + rhs."""
        self.set_real("time_step", dt)
        self._chk(self._f("assemble")(self.h, C.c_double(time)))
        nnz = int(self._f("nnz")(self.h))
        self.rowptr = np.zeros(self.D + 1, dtype=np.int64)
        self.col = np.zeros(nnz, dtype=np.int32)
        self.val = np.zeros(nnz)
        self.rhs = np.zeros(self.D)
        self._chk(self._f("get_csr")(self.h, _p(self.rowptr), _p(self.col), _p(self.val), _p(self.rhs)))
        return self.val, self.rhs

    def check_solution(self, time, dt=None):
        if dt is not None:
            self.set_real("time_step", dt)
        self._chk(self._f("check_solution")(self.h, C.c_double(time)))
        self.u = self.get_vector(self.sysname, 0)
        return self.u

    def set_solution(self, u):
        self.set_vector(self.sysname, 0, u)

    def save_solution(self, time):
        """Runs the reference's save_solution into a scratch CSV; returns the parsed numeric rows."""
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "out.csv")
            rc = self._f("save_solution")(self.h, C.c_double(time), path.encode())
            self._chk(rc)
            with open(path) as fh:
                lines = [ln.strip() for ln in fh if ln.strip()]
        header = [c.strip('"') for c in lines[0].split(",")] if lines and lines[0].startswith('"') else None
        rows = [np.array([float(x) for x in ln.split(",")]) for ln in lines if not ln.startswith('"')]
        return header, rows

    def run_init(self):
        self._chk(self._f("run_init")(self.h))


if __name__ == "__main__":
    print("built" if build_ref(force=True) else "reference sources not found; prebuilt libraries " +
          ("present" if available() else "missing"))
