"""Golden vectors produced by the REFERENCE'S OWN SOURCES (run from the repo root, in the build container):

    python tests/golden/make_ref_golden.py

oracle/ref.py compiles /root/reference/src/{adpm,pihna,ripf,proteas,coupled_hcc}.C unchanged (g++, serial libMesh
stand-in oracle/ref_shim/) and this script drives them: for every model, on the distorted n=4 Kuhn-tet cube and the
n=3 hex cube,
  * K and F of the first step exactly as assemble_<model>() + add_matrix/add_vector produce them,
  * the state after `nsteps` steps of the reference's time-loop body (rotate, assemble_<model>, K u = F,
    check_solution) where the linear solve -- PETSc's in the reference, not in-tree -- is a sparse direct solve
    (scipy SuperLU; residual <= 1e-15), i.e. the answer every converged Krylov method must reproduce,
  * RIPF: TD, RT_total and int(max RT_total) after the last check_solution; the run crosses a day boundary.
The files are small (<= 125 nodes) and committed, so tests/ on the GPU box (where /root/reference does not exist)
compare the CUDA path and the oracle with the reference's own arithmetic: tests/test_ref_pin.py (CPU, oracle) and
tests/test_gpu_ref_golden.py (GPU, through the C ABI)."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import ref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
NSTEPS = {cases.ADPM: 3, cases.PIHNA: 3, cases.RIPF: 12, cases.PROTEAS: 3, cases.HCC: 3}
MESHES = {"tet": (cases.TET4, 4), "hex": (cases.HEX8, 3)}


def inputs(model, tag):
    et, n = MESHES[tag]
    conn, xyz = cases.mesh(et, n, distort=0.2, length=50.0 if model == cases.RIPF else 1.0)
    p, u0, ef, nf = cases.case(model, conn, xyz, "full")
    return et, conn, xyz, p, u0, ef, nf


def reference_run(model, et, conn, xyz, p, u0, ef, nf, nsteps):
    """The reference's time loop (adpm.C:60-84 and siblings) with a direct solve in place of KSPSolve."""
    dt = cases.DT[model]
    rp = R.RefProblem(model, et, conn, xyz, p, u0, ef, nf, dt=dt)
    t = 0.0
    if model == cases.RIPF:
        rp.check_solution(0.0, dt)                     # ripf.C:53
    out = {}
    for k in range(nsteps):
        t += dt
        rp.rotate()
        val, rhs = rp.assemble(t, dt)
        if k == 0:
            out.update(rowptr=rp.rowptr.copy(), col=rp.col.copy(), val=val.copy(), rhs=rhs.copy())
        K = sp.csr_matrix((val, rp.col, rp.rowptr), shape=(rp.D, rp.D))
        u = spla.spsolve(K.tocsc(), rhs)
        u = u + spla.spsolve(K.tocsc(), rhs - K @ u)   # one refinement step: residual at round-off
        assert np.linalg.norm(rhs - K @ u) <= 1e-14 * np.linalg.norm(rhs)
        rp.set_solution(u)
        u = rp.check_solution(t, dt)
        if k == 0:
            out["u1"] = u.copy()
    out["uN"] = u.copy()
    out["nsteps"] = nsteps
    if model == cases.RIPF:
        out["td"] = rp.get_vector(b"RIPF-TimeDeriv", 0)
        out["rt"] = rp.get_vector(b"RT", 0)
        out["rt_max"] = rp.get_int("RT_dose/total/max")
    rp.close()
    return out


def main():
    assert R.build_ref(), "the reference sources are needed to (re)generate these files"
    for model in range(5):
        for tag in MESHES:
            et, conn, xyz, p, u0, ef, nf = inputs(model, tag)
            out = reference_run(model, et, conn, xyz, p, u0, ef, nf, NSTEPS[model])
            path = os.path.join(HERE, f"ref_{cases.NAMES[model]}_{tag}.npz")
            np.savez_compressed(path, **out)
            print(path, "nnz", out["val"].size, "|uN|", np.linalg.norm(out["uN"]))


if __name__ == "__main__":
    main()
