"""Generates the committed golden fixtures from the CPU oracle (run from the repo root):

    python tests/golden/make_golden.py

One file per model: Kuhn-tet cube n=4 (distorted), the `full` parameter set, one time step
(assemble at t = dt, GMRES(30)+ILU(0) rtol 1e-12, check_solution).  The reference itself cannot be run here
(no libMesh/PETSc), so these vectors pin the ORACLE (tests/test_golden.py) and, through it, the CUDA path
(tests/test_gpu_parity.py::test_golden_fixture_adpm)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    for model in range(5):
        length = 50.0 if model == cases.RIPF else 1.0
        conn, xyz = cases.mesh(cases.TET4, 4, distort=0.2, length=length)
        p, u0, ef, nf = cases.case(model, conn, xyz, "full")
        pr = cases.oracle_problem(model, cases.TET4, conn, xyz, p, u0, ef, nf, nthreads=1)
        dt = cases.DT[model]
        pr.u_old = pr.u.copy()
        val, rhs = pr.assemble(dt, dt)
        val, rhs = val.copy(), rhs.copy()
        pr.time = 0.0
        its, res = pr.step(dt, pc=O.PC_ILU)
        out = os.path.join(HERE, f"{cases.NAMES[model]}_tet_n4.npz")
        np.savez_compressed(out, rowptr=pr.rowptr, col=pr.col, val=val, rhs=rhs, u1=pr.u, its=its)
        print(out, "its", its, "nnz", val.size)


if __name__ == "__main__":
    main()
