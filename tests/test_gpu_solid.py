"""CUDA solid-mechanics path (rdcfes_b200/csrc/solid.cu) against the CPU oracle, through the C ABI.

The oracle (oracle/solid_oracle.c) is pinned to the reference's own solid_system.C / hyperelastic.h / eig3.C
(tests/test_solid_pin.py).  Tolerances: Jacobian entries |err| <= 1e-12 max|J| (the penalty entries are 1e5..1e8 times the
elastic ones and are added in a different order), residual 1e-12 max|R|, converged positions 1e-8 relative when both
Newton drivers run with tight tolerances, post-processing 1e-10 of the stress scale."""
import numpy as np
import pytest
import scipy.sparse as sp

import solid_cases as SC
from oracle import oracle as O
from oracle import solid as S
from rdcfes_b200 import solid as G

pytestmark = pytest.mark.gpu


def _csr(rowptr, col, val, D):
    return sp.csr_matrix((val, col, rowptr), shape=(D, D))


@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
@pytest.mark.parametrize("n", [2, 5])
def test_jacobian_and_residual_match_oracle(et, n):
    c = SC.general_case(et, n=n)
    x = SC.perturbed(c, amp=0.03 / n)
    orc = S.OracleSolid(c)
    val_o, rhs_o = orc.assemble(x, 0.3)
    g = G.from_case(c)
    g.set_positions(x)
    g.assemble(0.3)
    rows, rowptr, col, val, rhs = g.download_csr()
    D = 3 * c.N
    assert np.array_equal(rowptr, orc.rowptr) and np.array_equal(col, orc.col), "sparsity pattern differs"
    assert np.abs(val - val_o).max() <= 1e-12 * np.abs(val_o).max()
    assert np.abs(rhs - rhs_o).max() <= 1e-12 * np.abs(rhs_o).max()
    # the elastic part alone (no penalty): relative to the elastic entries
    c0 = SC.general_case(et, n=n, penalty=0.0)
    v0, r0 = S.OracleSolid(c0).assemble(x, 0.3)
    g0 = G.from_case(c0)
    g0.set_positions(x)
    g0.assemble(0.3)
    _, _, _, val0, rhs0 = g0.download_csr()
    assert np.abs(val0 - v0).max() <= 1e-12 * np.abs(v0).max()
    assert np.abs(rhs0 - r0).max() <= 1e-12 * np.abs(r0).max()
    # bit-reproducible
    g.assemble(0.3)
    _, _, _, val2, rhs2 = g.download_csr()
    assert np.array_equal(val, val2) and np.array_equal(rhs, rhs2)
    g.close(); g0.close()


@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
def test_post_process_matches_oracle(et):
    c = SC.general_case(et, n=4)
    x = SC.perturbed(c, amp=0.01)
    po, vo, fo = S.OracleSolid(c).post(x, 0.4)
    g = G.from_case(c)
    g.set_positions(x)
    p, v, f = g.post_process(0.4)
    scale = np.abs(po).max() + vo.max()
    assert np.abs(p - po).max() <= 1e-10 * scale and np.abs(v - vo).max() <= 1e-10 * scale
    assert np.abs(f - fo).max() <= 1e-12
    g.close()


@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
@pytest.mark.parametrize("ksp", [G.KSP_GMRES, G.KSP_BICGSTAB])
def test_load_steps_match_oracle_with_tight_tolerances(et, ksp):
    # run/Solid/uniaxial_compression in miniature, three load steps; both Newton drivers converge their iterates far below
    # the comparison tolerance, so the linear solvers (GMRES+ILU(0) on the host, Krylov+Jacobi on the device) do not matter
    c = SC.compression_case(et, n=4, penalty=1.0e6)
    c.opts.update(max_nonlinear_iterations=25, relative_step_tolerance=1e-11, relative_residual_tolerance=1e-13,
                  absolute_residual_tolerance=1e-9, initial_linear_tolerance=1e-10)
    orc = S.OracleSolid(c)
    g = G.from_case(c)
    g.ksp = ksp
    xo = c.xund.copy().ravel()
    for step in (1, 2, 3):
        t = 0.1 * step
        xo, io = orc.newton(xo, t)
        ig = g.run_solver(t)
        assert io["converged"] and ig["converged"], (io, ig)
        xg = g.get_positions().ravel()
        rel = np.linalg.norm(xg - xo) / np.linalg.norm(xo)
        assert rel <= 1e-8, (step, rel, io, ig)
    X = g.get_positions()
    top = np.abs(c.xund[:, 2] - 1.5) < 1e-9
    assert np.abs(X[top, 2] - (1.5 - 0.75 * 0.3 * 1.000001)).max() < 1e-3
    g.close()


@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
def test_load_steps_with_the_shipped_tolerances(et):
    # the tolerances of run/Solid/*/input.dat (relative step 1e-3, initial linear tolerance 1e-3): the iterates of two
    # correct drivers then agree only to those tolerances -- same Newton iteration counts, positions to 1e-4 of the size
    c = SC.compression_case(et, n=4)
    orc = S.OracleSolid(c)
    g = G.from_case(c)
    xo = c.xund.copy().ravel()
    for step in (1, 2):
        xo, io = orc.newton(xo, 0.1 * step)
        ig = g.run_solver(0.1 * step)
        assert io["converged"] and ig["converged"]
        assert abs(io["newton_its"] - ig["newton_its"]) <= 1
        assert np.abs(g.get_positions().ravel() - xo).max() <= 1e-4 * 1.5
    g.close()


def test_growing_inclusion_matches_oracle():
    # the solid half of the coupled HCC model (coupled_hcc.C:117-132): a region growing with rate 0.3 in a passive matrix
    c = SC.growth_case(SC.TET4, n=4)
    c.opts.update(max_nonlinear_iterations=25, relative_step_tolerance=1e-11, relative_residual_tolerance=1e-13,
                  absolute_residual_tolerance=1e-9, initial_linear_tolerance=1e-10)
    orc = S.OracleSolid(c)
    g = G.from_case(c)
    g.ksp = G.KSP_GMRES
    xo = c.xund.copy().ravel()
    for step in (1, 2):
        t = 0.025 * step
        xo, io = orc.newton(xo, t)
        ig = g.run_solver(t)
        assert io["converged"] and ig["converged"], (io, ig)
        rel = np.linalg.norm(g.get_positions().ravel() - xo) / np.linalg.norm(xo)
        assert rel <= 1e-8, (step, rel)
    # the inclusion grew: its elements are larger than at the start, pressure inside is compressive
    p, v, f = g.post_process(0.05)
    po, vo, fo = orc.post(xo, 0.05)
    assert np.abs(p - po).max() <= 1e-6 * (np.abs(po).max() + vo.max())
    assert p[c.mat_of == 1].mean() < 0.0
    g.close()


@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
def test_assembly_use_symmetry(et):
    # solver/assembly_use_symmetry = true (solid_system.C:248-262): with growth and fibres the mirrored matrix differs from the
    # exact tangent; the device reproduces the reference's choice
    c = SC.general_case(et, n=3, use_symmetry=True)
    x = SC.perturbed(c, amp=0.01)
    val_o, rhs_o = S.OracleSolid(c).assemble(x, 0.3)
    g = G.from_case(c)
    g.set_positions(x)
    g.assemble(0.3)
    _, _, _, val, rhs = g.download_csr()
    assert np.abs(val - val_o).max() <= 1e-12 * np.abs(val_o).max()
    assert np.abs(rhs - rhs_o).max() <= 1e-12 * np.abs(rhs_o).max()
    g.set_symmetry(False)
    g.assemble(0.3)
    _, _, _, val2, _ = g.download_csr()
    assert np.abs(val2 - val).max() > 1e-6 * np.abs(val).max()
    g.close()


def test_wrong_calls_are_refused():
    c = SC.compression_case(SC.TET4, n=2)
    g = G.from_case(c)
    L = g._L
    assert L.rdc_clamp(g._h) != 0 and L.rdc_assemble(g._h, 0.1, 0.1) != 0
    g.close()
