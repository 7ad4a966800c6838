"""Pins oracle/solid_oracle.c to the reference's OWN solid-mechanics sources (CPU only).

oracle/_ref/libref_solid.so = /root/reference/src/solid_system.C (+ hyperelastic.h, hyperlastic_inline.h, eig3.C) compiled
unchanged against the serial libMesh stand-in (oracle/solid.py build_ref_solid).  Compared: every element's residual and
tangent (element term + penalty side terms), the global residual and Jacobian, and post_process (mean stress, von Mises
stress, current fibre vector) on TET4 and HEX8 with two materials, fibres, anisotropic growth and NaN-masked boundary
conditions.  Tolerance 1e-12 relative to the largest entry (the oracle evaluates the 3^8 push-forward of the tangent as
four single-index contractions); post_process bit for bit."""
import numpy as np
import pytest
import scipy.sparse as sp

import solid_cases as SC
from oracle import solid as S

pytestmark = pytest.mark.skipif(not (S.ref_solid_available() or S.build_ref_solid()), reason="reference sources not on this machine")


@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
@pytest.mark.parametrize("sym", [False, True])
def test_elements_match_reference(et, sym):
    c = SC.general_case(et, use_symmetry=sym)
    x = SC.perturbed(c)
    orc, ref = S.OracleSolid(c), S.RefSolid(c)
    for e in range(c.E):
        Ro, Ko = orc.element(e, x, 0.3)
        Rr, Kr = ref.element(e, x, 0.3)
        assert np.abs(Ro - Rr).max() <= 1e-12 * np.abs(Rr).max()
        assert np.abs(Ko - Kr).max() <= 1e-12 * np.abs(Kr).max()


def test_use_symmetry_changes_the_tangent_of_a_growing_element():
    # with anisotropic growth the spatial tangent loses its major symmetry: mirroring the upper triangle (solid_system.C:258-262)
    # is then a different matrix -- the flag must reach both implementations
    a, b = SC.general_case(SC.TET4, use_symmetry=False), SC.general_case(SC.TET4, use_symmetry=True)
    x = SC.perturbed(a)
    Ka = S.OracleSolid(a).element(0, x, 0.3)[1]
    Kb = S.OracleSolid(b).element(0, x, 0.3)[1]
    assert np.abs(Ka - Kb).max() > 1e-6 * np.abs(Ka).max()
    nen = 4
    node = np.tile(np.arange(nen), 3)                      # variable-major: entry a*nen + i belongs to node i
    off = node[:, None] != node[None, :]                   # blocks between different nodes are mirrored, i == j blocks are not
    assert np.abs(Kb - Kb.T)[off].max() <= 1e-12 * np.abs(Kb).max()
    assert np.abs(Ka - Ka.T)[off].max() > 1e-6 * np.abs(Ka).max()


@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
def test_global_system_matches_reference(et):
    c = SC.general_case(et, n=3)
    x = SC.perturbed(c)
    orc, ref = S.OracleSolid(c), S.RefSolid(c)
    val, rhs = orc.assemble(x, 0.7)
    rp, cl, vr, rr = ref.assemble(x, 0.7)
    D = 3 * c.N
    assert np.abs(rhs - rr).max() <= 1e-12 * np.abs(rr).max()
    Ao = sp.csr_matrix((val, orc.col, orc.rowptr), shape=(D, D))
    Ar = sp.csr_matrix((vr, cl, rp), shape=(D, D))
    assert abs(Ao - Ar).max() <= 1e-12 * abs(Ar).max()
    # the reference touches exactly the (node graph + I) x dense 3x3 pattern the oracle preallocates
    assert np.array_equal(rp, orc.rowptr) and np.array_equal(cl, orc.col)


@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
def test_post_process_matches_reference(et):
    c = SC.general_case(et, n=3)
    x = SC.perturbed(c)
    po, vo, fo = S.OracleSolid(c).post(x, 0.4)
    pr, vr, fr = S.RefSolid(c).post(x, 0.4)
    assert np.array_equal(po, pr) and np.array_equal(vo, vr) and np.array_equal(fo, fr)


def test_eig3_known_answers():
    orc = S.OracleSolid(SC.general_case(SC.TET4))
    A = np.array([[2.0, -1.0, 0.0], [-1.0, 2.0, -1.0], [0.0, -1.0, 2.0]])
    d, V = orc.eig3(A)
    assert np.allclose(d, [2 - np.sqrt(2), 2.0, 2 + np.sqrt(2)], rtol=0, atol=1e-14)
    assert np.allclose(A @ V, V * d, atol=1e-13)
    rng = np.random.default_rng(0)
    for _ in range(20):
        B = rng.normal(size=(3, 3)); B = B + B.T
        d, V = orc.eig3(B)
        assert np.allclose(d, np.linalg.eigvalsh(B), atol=1e-13)


def test_jacobian_is_the_derivative_of_the_residual():
    # independent of both transcriptions: central differences of the element residual reproduce the analytic tangent (growth
    # and fibres included).  The penalty term is left out: solid_system.C:357-360 ignores the dependence of the side's JxW
    # on the positions, so with it the reference's Jacobian is an approximation (0.3 % of max|J| at penalty 1e3) by design.
    c = SC.general_case(SC.HEX8, penalty=0.0)
    x = SC.perturbed(c, amp=0.01).ravel()
    orc = S.OracleSolid(c)
    val, _ = orc.assemble(x, 0.2)
    D = x.size
    A = sp.csr_matrix((val, orc.col, orc.rowptr), shape=(D, D)).toarray()
    h = 1e-6
    for j in np.random.default_rng(5).choice(D, 12, replace=False):
        xp, xm = x.copy(), x.copy()
        xp[j] += h; xm[j] -= h
        fd = (orc.assemble(xp, 0.2, want_jac=False)[1] - orc.assemble(xm, 0.2, want_jac=False)[1]) / (2 * h)
        assert np.abs(fd - A[:, j]).max() <= 1e-8 * np.abs(A).max()


@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
def test_newton_load_steps_converge(et):
    # run/Solid/uniaxial_compression in miniature: the top face follows the prescribed displacement, the bottom stays put
    c = SC.compression_case(et, n=3)
    orc = S.OracleSolid(c)
    x = c.xund.copy().ravel()
    for step in (1, 2):
        x, info = orc.newton(x, 0.1 * step)
        assert info["converged"] and info["newton_its"] <= 10
    X = x.reshape(-1, 3)
    top = np.abs(c.xund[:, 2] - 1.5) < 1e-9
    bot = np.abs(c.xund[:, 2]) < 1e-9
    assert np.abs(X[top, 2] - (1.5 - 0.75 * 0.2 * 1.000001)).max() < 1e-4      # ratio = pseudo_time * 1.000001 (solid_system.C:285)
    assert np.abs(X[bot] - c.xund[bot]).max() < 1e-4


@pytest.mark.parametrize("name", ["solid_uniaxial", "solid_hydrogel"])
def test_shipped_solid_cases_match_reference(name):
    """run/Solid/uniaxial_compression (cube.msh, HEX8) and run/Solid/hydrogel_tension (hydrogel_model.msh, TET4) as shipped:
    mesh, tagged boundary faces, BC table and penalty of the input.dat; Jacobian, residual and post-processing of a state
    away from equilibrium -- oracle against the reference's own sources."""
    c, kv, d = SC.shipped_case(name)
    h = float(np.linalg.norm(c.xund.max(axis=0) - c.xund.min(axis=0)))
    x = c.xund + 2e-3 * h * np.random.default_rng(4).normal(size=c.xund.shape)
    orc, ref = S.OracleSolid(c), S.RefSolid(c)
    val, rhs = orc.assemble(x, 0.4)
    rp, cl, vr, rr = ref.assemble(x, 0.4)
    assert np.array_equal(rp, orc.rowptr) and np.array_equal(cl, orc.col)
    assert np.abs(val - vr).max() <= 1e-12 * np.abs(vr).max()
    assert np.abs(rhs - rr).max() <= 1e-12 * np.abs(rr).max()
    po, vo, fo = orc.post(x, 0.4)
    pr, vr2, fr = ref.post(x, 0.4)
    assert np.array_equal(po, pr) and np.array_equal(vo, vr2) and np.array_equal(fo, fr)
