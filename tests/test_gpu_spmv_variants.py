"""The SpMV kernels (TMA bulk-copy staged, LDG at several register budgets) and the BiCGStab that uses their
fused dot products must agree with each other and with a CSR matvec of the downloaded operator; long block
rows exercise the host-side tile cutter (<= 16 rows, <= 256 blocks per tile)."""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
from cases import ADPM, HCC, PIHNA, PROTEAS, RIPF, TET4

pytestmark = pytest.mark.gpu


def _system(model, n=6):
    length = 50.0 if model == RIPF else 1.0
    conn, xyz = cases.mesh(TET4, n, distort=0.2, length=length)
    p, u0, ef, nf = cases.case(model, conn, xyz, "full")
    return cases.gpu_system(model, TET4, conn, xyz, p, u0, ef, nf), conn, xyz


@pytest.mark.parametrize("model", [ADPM, PIHNA, RIPF, PROTEAS, HCC])
def test_spmv_variants_match_csr(model):
    gpu, conn, xyz = _system(model)
    dt = cases.DT[model]
    gpu.rotate()
    gpu.assemble(dt, dt)
    rows, rowptr, col, val, rhs = gpu.download_csr()
    A = sp.csr_matrix((val, col, rowptr), shape=(gpu.n_dofs, gpu.n_dofs))
    x = np.random.default_rng(7).standard_normal(gpu.n_dofs)
    ref = A @ x
    scale = np.abs(A).dot(np.abs(x)).max()
    out = {}
    for name, opts in {"tma": {"spmv_tma": 1}, "tma3": {"spmv_tma": 1, "tma_stages": 3, "tma_ctas_per_sm": 4},
                       "ldg4": {"spmv_tma": 0, "spmv_minb": 4}, "ldg8": {"spmv_tma": 0, "spmv_minb": 8},
                       "ldg6": {"spmv_tma": 0, "spmv_minb": 6}}.items():
        for k, v in opts.items():
            gpu.set_option(k, v)
        y = gpu.spmv(x)
        assert np.abs(y - ref).max() <= 1e-13 * scale, name
        out[name] = y
    # same lane -> block mapping and the same shuffle tree in every variant: identical bits
    for name in out:
        assert np.array_equal(out[name], out["tma"]), name
    gpu.close()


@pytest.mark.parametrize("ksp", [0, 2])
def test_solve_independent_of_spmv_variant(ksp):
    sols = []
    for tma in (1, 0):
        gpu, _, _ = _system(ADPM, 7)
        gpu.set_option("spmv_tma", tma)
        gpu.ksp = ksp
        for _ in range(2):
            its, res = gpu.step(cases.DT[ADPM])
        sols.append((gpu.get_solution(), its))
        gpu.close()
    assert sols[0][1] == sols[1][1]
    assert np.array_equal(sols[0][0], sols[1][0])


def _fan_mesh(m):
    """m tets around a common edge A-B: the rows of A and B have m + 2 blocks (long-row tiles)."""
    ang = 2 * np.pi * np.arange(m) / m
    ring = np.stack([np.cos(ang), np.sin(ang), 0.5 + 0.1 * np.sin(3 * ang)], axis=1)
    xyz = np.vstack([[0.0, 0.0, 0.0], [0.0, 0.0, 1.0], ring])
    conn = np.array([[0, 2 + i, 2 + (i + 1) % m, 1] for i in range(m)], dtype=np.int32)
    # positive orientation as libMesh expects
    a, b, c, d = (xyz[conn[:, k]] for k in range(4))
    vol = np.einsum("ij,ij->i", np.cross(b - a, c - a), d - a)
    flip = vol < 0
    conn[flip, 1], conn[flip, 2] = conn[flip, 2].copy(), conn[flip, 1].copy()
    return conn, xyz


@pytest.mark.parametrize("m", [40, 100, 127])
def test_long_rows_against_oracle(m):
    conn, xyz = _fan_mesh(m)
    p, u0, ef, nf = cases.case(HCC, conn, xyz, "full")
    orc = cases.oracle_problem(HCC, TET4, conn, xyz, p, u0, ef, nf)
    gpu = cases.gpu_system(HCC, TET4, conn, xyz, p, u0, ef, nf)
    dt = cases.DT[HCC]
    orc.u_old = orc.u.copy()
    val_o, rhs_o = orc.assemble(dt, dt)
    gpu.rotate()
    gpu.assemble(dt, dt)
    rows, rowptr, col, val, rhs = gpu.download_csr()
    assert np.array_equal(rowptr, orc.rowptr) and np.array_equal(col, orc.col)
    assert (np.abs(val - val_o) / cases.csr_tolerance(val_o)).max() <= 1.0
    A = sp.csr_matrix((val, col, rowptr), shape=(gpu.n_dofs, gpu.n_dofs))
    x = np.random.default_rng(3).standard_normal(gpu.n_dofs)
    scale = np.abs(A).dot(np.abs(x)).max()
    for tma in (1, 0):
        gpu.set_option("spmv_tma", tma)
        assert np.abs(gpu.spmv(x) - A @ x).max() <= 1e-13 * scale
    gpu.time = 0.0
    from oracle import oracle as O
    orc.step(dt, pc=O.PC_ILU)
    gpu.ksp = 2
    gpu.step(dt)
    rel = np.linalg.norm(gpu.get_solution() - orc.u) / np.linalg.norm(orc.u)
    assert rel <= 1e-8, rel
    gpu.close()


def test_owned_download_equals_full_download_on_one_rank():
    gpu, _, _ = _system(ADPM, 5)
    gpu.step(cases.DT[ADPM])
    full = gpu.get_solution()
    out = np.full(gpu.n_dofs, np.nan)
    gpu.get_solution_owned(out)
    assert np.array_equal(out, full)
    gpu.close()


def test_unknown_option_is_rejected():
    from rdcfes_b200.lib import RdcError
    gpu, _, _ = _system(ADPM, 4)
    with pytest.raises(RdcError):
        gpu.set_option("no_such_switch", 1)
    gpu.close()
