"""RIPF per-step solution error of the device solvers against the oracle, split into its two sources (needs a GPU).

Every linear solve is compared with a DIRECT solve of the very system it was given (downloaded operator and rhs), on
both sides.  Result (profiles/r2_ripf_step_error_analysis.log): each solver is within ~1e-12 (HU) of the exact solution
of its own system; what differs from step 2 on is the SYSTEM, because the convergence test is relative to the norm of
the whole (HU ~ 1e3 dominated) vector and leaves the cell fractions (~1e-1) converged to only ~1e-8 relative in step 1
-- in the oracle's GMRES as much as in the device solvers -- and TD = (u - prev)/dt feeds that into the next operator.
There is no drift of the BiCGStab recurrence residual: the true residuals match GMRES's."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    model = cases.RIPF
    conn, xyz = cases.mesh(4, n, distort=0.2, length=50.0)
    p, u0, ef, nf = cases.case(model, conn, xyz, "full")
    dt = cases.DT[model]
    for label, ksp, persist in (("bicgstab five-launch", 2, 0), ("bicgstab persistent", 2, 1), ("gmres(30)", 0, 0)):
        orc = cases.oracle_problem(model, 4, conn, xyz, p, u0, ef, nf)
        gpu = cases.gpu_system(model, 4, conn, xyz, p, u0, ef, nf)
        gpu.ksp = ksp
        gpu.set_option("bicg_persist", persist)
        nv = orc.nv
        for step in range(4):
            orc.step(dt, pc=O.PC_ILU)
            gpu.time += dt
            gpu.dt = dt
            gpu.rotate()
            gpu.assemble(gpu.time, dt)
            its, res = gpu.linear_solve()
            x = gpu.get_solution().copy()
            b = gpu.get_rhs()
            r = b - gpu.spmv(x)
            import scipy.sparse as sp
            import scipy.sparse.linalg as spl
            rows, rowptr, col, val, rhs = gpu.download_csr()
            A = sp.csr_matrix((val, col, rowptr), shape=(orc.D, orc.D))
            xs = spl.spsolve(A.tocsc(), rhs)
            xo = spl.spsolve(orc.scipy_csr().tocsc(), orc.rhs)
            eg = [np.linalg.norm(x[a::nv] - xs[a::nv]) / max(np.linalg.norm(xs[a::nv]), 1e-300) for a in range(nv)]
            eo = [np.linalg.norm(orc.prev[a::nv] - xo[a::nv]) / max(np.linalg.norm(xo[a::nv]), 1e-300) for a in range(nv)]
            ex = [np.linalg.norm(xs[a::nv] - xo[a::nv]) / max(np.linalg.norm(xo[a::nv]), 1e-300) for a in range(nv)]
            print(f"   vs direct solve of its own system: device {' '.join(f'{v:.1e}' for v in eg)} | oracle {' '.join(f'{v:.1e}' for v in eo)} "
                  f"| the two exact solutions differ by {' '.join(f'{v:.1e}' for v in ex)}")
            gpu.check_solution()
            u = gpu.get_solution()
            per = [np.linalg.norm(u[a::nv] - orc.u[a::nv]) / max(np.linalg.norm(orc.u[a::nv]), 1e-300) for a in range(nv)]
            rres = [np.linalg.norm(r[a::nv]) / max(np.linalg.norm(b[a::nv]), 1e-300) for a in range(nv)]
            print(f"{label} step {step + 1}: its {its} reported res {res:.2e} | "
                  f"|b-Ax|/|b| per species {' '.join(f'{v:.1e}' for v in rres)} | rel err vs oracle {' '.join(f'{v:.1e}' for v in per)} "
                  f"| all {np.linalg.norm(u - orc.u) / np.linalg.norm(orc.u):.2e}")
        gpu.close()


if __name__ == "__main__":
    main()
