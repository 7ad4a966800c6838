"""Known-answer tests that pin the CPU oracle (oracle/rdc_oracle.c) WITHOUT libMesh/PETSc.

The reference ships no tests or golden vectors, so the restatement is pinned by the
independent checks listed there: analytic P1/Q1 element matrices, quadrature exactness, constant
preservation, row sums, a finite-difference check of every Jacobian block against its load vector
(reproducing -- not fixing -- the reference's inconsistencies, Appendix C-3) and a sparse direct solve.
"""
import itertools

import numpy as np
import pytest

from oracle import oracle as O
from rdcfes_b200 import params as P
from rdcfes_b200 import synth

TET = np.array([[0.1, 0.0, 0.2], [1.3, 0.2, 0.1], [0.2, 1.1, 0.0], [0.3, 0.2, 0.9]])


def tet_volume(X):
    return np.linalg.det(X[1:] - X[0]) / 6.0


def tet_grads(X):
    """Independent derivation: rows of inv([1 x y z]) give the barycentric gradients."""
    A = np.hstack([np.ones((4, 1)), X])
    return np.linalg.inv(A)[1:, :].T  # [4,3]


# ------------------------------------------------------------------------------------- quadrature KATs
def test_tet_rule_weights_and_exactness():
    w, phi = O.fe_tables(O.TET4)
    assert w.shape == (5,) and phi.shape == (4, 5)
    assert abs(w.sum() - 1.0 / 6.0) < 1e-16
    assert w[0] < 0  # the negative-weight Keast rule matters for parity (Appendix B-2)
    pts = phi[1:].T  # (xi, eta, zeta) = (phi1, phi2, phi3)
    # exact integrals of x^a y^b z^c over the unit tet: a! b! c! / (a+b+c+3)!
    from math import factorial as f
    for a, b, c in itertools.product(range(4), repeat=3):
        if a + b + c > 3:
            continue
        num = (w * pts[:, 0] ** a * pts[:, 1] ** b * pts[:, 2] ** c).sum()
        assert abs(num - f(a) * f(b) * f(c) / f(a + b + c + 3)) < 1e-15
    np.testing.assert_allclose(phi.sum(0), 1.0, atol=1e-15)  # partition of unity


def test_hex_rule():
    w, phi = O.fe_tables(O.HEX8)
    assert w.shape == (8,) and abs(w.sum() - 8.0) < 1e-14
    np.testing.assert_allclose(phi.sum(0), 1.0, atol=1e-15)
    g = 1.0 / np.sqrt(3.0)
    # x fastest ordering: phi of node 0 (-,-,-) is largest at qp 0
    assert np.argmax(phi[0]) == 0 and np.argmax(phi[1]) == 1 and np.argmax(phi[6]) == 7
    assert abs(phi[0, 0] - ((1 + g) / 2) ** 3) < 1e-15


# -------------------------------------------------------------------------- analytic element matrices
def test_p1_mass_and_stiffness_tet():
    V = tet_volume(TET)
    G = tet_grads(TET)
    rng = np.random.default_rng(0)
    U = rng.random((4, 3))
    # all rates zero -> Ke = block-diag(M), Fe = M u   (Appendix A with f = q = r = s = t = 0)
    Ke, Fe, JxW, dphi = O.element(O.ADPM, O.TET4, TET, U, P.flat_params(P.ADPM), efield=[0.1, 0.2, 0.3])
    M = V / 20.0 * (np.ones((4, 4)) + np.eye(4))
    assert abs(JxW.sum() - V) < 1e-15
    np.testing.assert_allclose(dphi[:, 0, :], G, rtol=0, atol=1e-13)
    for a in range(3):
        for b in range(3):
            blk = Ke[a * 4:(a + 1) * 4, b * 4:(b + 1) * 4]
            np.testing.assert_allclose(blk, M if a == b else 0.0, rtol=0, atol=1e-16)
        np.testing.assert_allclose(Fe[a * 4:(a + 1) * 4], M @ U[:, a], rtol=1e-13)
    # diffusion only: Ke_AA = M + dt/2 * D * V * G G^T  (sign: K = M - dt/2 * (-D grad.grad))
    D, dt = 0.37, 0.2
    p = P.flat_params(P.ADPM, {"diffuse/A_b": D})
    Ke, Fe, _, _ = O.element(O.ADPM, O.TET4, TET, U, p, efield=[0, 0, 0], dt=dt)
    S = V * G @ G.T
    np.testing.assert_allclose(Ke[4:8, 4:8], M + dt / 2 * D * S, rtol=1e-12)
    np.testing.assert_allclose(S.sum(1), 0.0, atol=1e-13)  # row sums of the diffusion block vanish
    np.testing.assert_allclose(Fe[4:8], M @ U[:, 1] - dt / 2 * D * S @ U[:, 1], rtol=1e-12)


def test_q1_mass_hex():
    h = 0.1875  # cell size of run/Solid/uniaxial_compression/cube.msh
    conn, xyz = synth.hex_cube(1, h)
    U = np.random.default_rng(1).random((8, 3))
    Ke, Fe, JxW, _ = O.element(O.HCC, O.HEX8, xyz[conn[0]], U, P.flat_params(P.HCC))
    assert abs(JxW.sum() - h ** 3) < 1e-15
    c = xyz[conn[0]] / h
    dist = np.abs(c[:, None, :] - c[None, :, :]).sum(-1).round().astype(int)
    M = h ** 3 / 216.0 * np.array([8.0, 4.0, 2.0, 1.0])[dist]
    np.testing.assert_allclose(Ke[16:24, 16:24], M, rtol=1e-13)


# ------------------------------------------------------------------- global: volume, constants, solve
@pytest.mark.parametrize("model", [O.ADPM, O.PIHNA, O.HCC])
def test_constant_preservation_and_volume(model):
    conn, xyz = synth.kuhn_cube(3, 1.5)
    xyz = synth.distort(xyz, 0.1)
    nv = O.nvars(model)
    u0 = np.random.default_rng(2).random((xyz.shape[0], nv)) + 0.5
    pm = {O.ADPM: P.ADPM, O.PIHNA: P.PIHNA, O.HCC: P.HCC}[model]
    pr = O.Problem(model, O.TET4, conn, xyz, P.flat_params(pm), u0, elem_field=np.zeros((conn.shape[0], 3)))
    val, rhs = pr.assemble(0.1, 0.1)
    A = pr.scipy_csr()
    if model != O.HCC:  # HCC carries capacity on off-diagonal blocks (Appendix C-3)
        assert abs(val.sum() - nv * 1.5 ** 3) < 1e-12  # sum_ij M_ij = volume, per variable
        x, its, res, res0 = pr.solve(rtol=1e-14)
        np.testing.assert_allclose(x, u0.ravel(), rtol=1e-10)  # K u = F returns u_old when all rates vanish
    else:
        cap = np.array([[1, 1, 1], [1, 2, 0], [0, 0, 1]])
        assert abs(val.sum() - cap.sum() * 1.5 ** 3) < 1e-12
    assert (np.diff(pr.rowptr) % nv == 0).all()
    assert abs(A - A.T).max() < 1e-15 or model == O.HCC


def test_pattern_is_nodegraph_kron_dense():
    conn, xyz = synth.kuhn_cube(2)
    rowptr, col = O.build_pattern(xyz.shape[0], conn, 3)
    N = xyz.shape[0]
    adj = [set([n]) for n in range(N)]
    for e in conn:
        for a in e:
            adj[a].update(int(b) for b in e)
    for n in range(N):
        exp = sorted(3 * m + b for m in adj[n] for b in range(3))
        for a in range(3):
            r = 3 * n + a
            assert list(col[rowptr[r]:rowptr[r + 1]]) == exp


def test_gmres_ilu_against_direct_solve():
    conn, xyz = synth.kuhn_cube(4)
    u0, tr = synth.adpm_fields(conn, xyz, smooth=True)
    pr = O.Problem(O.ADPM, O.TET4, conn, xyz, synth.adpm_params("full"), u0, elem_field=tr)
    pr.assemble(0.05, 0.05)
    import scipy.sparse.linalg as sla
    xs = sla.spsolve(pr.scipy_csr().tocsc(), pr.rhs)
    for pc, nb in ((O.PC_ILU, 1), (O.PC_ILU, 4), (O.PC_JACOBI, 1), (O.PC_NONE, 1)):
        x, its, res, res0 = pr.solve(pc=pc, nblocks=nb, x0=u0.ravel().copy())
        assert res <= 1e-12 * res0 * 1.0001 and its < 200
        assert np.linalg.norm(x - xs) / np.linalg.norm(xs) < 1e-9, (pc, nb)


def test_ilu0_exact_for_tridiagonal():
    # ILU(0) of a tridiagonal matrix is its exact LU -> preconditioned GMRES converges in one iteration
    n = 50
    rowptr = np.zeros(n + 1, dtype=np.int64)
    col, val = [], []
    for i in range(n):
        for j in (i - 1, i, i + 1):
            if 0 <= j < n:
                col.append(j)
                val.append(4.0 if i == j else (-1.0 if j < i else -2.0))
        rowptr[i + 1] = len(col)
    col = np.asarray(col, dtype=np.int32)
    val = np.asarray(val)
    b = np.arange(1.0, n + 1)
    x = np.zeros(n)
    import ctypes as C
    its, res, res0 = C.c_int(), C.c_double(), C.c_double()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = O.lib().orc_gmres(C.c_int64(n), p(rowptr), p(col), p(val), p(b), p(x), C.c_int(0), C.c_int(1), C.c_int(30),
                           C.c_double(1e-12), C.c_int(100), C.c_int(1), C.byref(its), C.byref(res), C.byref(res0))
    assert rc == 0 and its.value <= 1
    import scipy.sparse as sp
    A = sp.csr_matrix((val, col, rowptr), shape=(n, n))
    np.testing.assert_allclose(A @ x, b, rtol=1e-12)


# ------------------------------------------------------------ finite-difference Jacobian consistency
def _fd_blocks(model, X, U, params, efield, aux, rt_max, time, dt, elem=O.TET4, rel=1e-6):
    """d Fe / d U by central differences, returned in the var-major layout of Ke."""
    nen = X.shape[0]
    nv = U.shape[1]
    nd = nen * nv
    J = np.zeros((nd, nd))
    for b in range(nv):
        for j in range(nen):
            h = rel * (np.abs(U[:, b]).max() or 1.0)  # relative to the variable's own scale (a ~ 1e-9 in PIHNA)
            Up, Um = U.copy(), U.copy()
            Up[j, b] += h
            Um[j, b] -= h
            Fp = O.element(model, elem, X, Up, params, efield, aux, rt_max, time, dt)[1]
            Fm = O.element(model, elem, X, Um, params, efield, aux, rt_max, time, dt)[1]
            J[:, b * nen + j] = (Fp - Fm) / (2 * h)
    return J


def _mass(model, X, nv, elem=O.TET4):
    pm = {O.ADPM: P.ADPM, O.PIHNA: P.PIHNA, O.RIPF: P.RIPF, O.PROTEAS: P.PROTEAS, O.HCC: P.HCC}[model]
    zero = np.zeros(len(P.TABLES[pm]))
    zero_params = P.flat_params(pm)
    if model == O.PROTEAS:
        zero_params = zero.copy()
        zero_params[0] = 1.0  # T_max
        zero_params[1] = 1.0  # RT_max
        zero_params[25] = 1.0
    Ke = O.element(model, elem, X, np.zeros((X.shape[0], nv)) + 0.1, zero_params,
                   efield=[0, 0, 0], aux=np.zeros((X.shape[0], 6 if model == O.RIPF else 2)), rt_max=1)[0]
    return Ke[:X.shape[0], :X.shape[0]].copy()


def _check_fd(model, X, U, params, efield=None, aux=None, rt_max=1, time=0.3, dt=0.1, cap=None, fold=None,
              elem=O.TET4, tol=2e-6):
    nen, nv = X.shape[0], U.shape[1]
    Ke = O.element(model, elem, X, U, params, efield, aux, rt_max, time, dt)[0]
    J = _fd_blocks(model, X, U, params, efield, aux, rt_max, time, dt, elem)
    M = _mass(model, X, nv, elem)
    cap = np.eye(nv) if cap is None else cap
    scale = np.abs(Ke).max()
    for a in range(nv):
        for b in range(nv):
            Jab = J[a * nen:(a + 1) * nen, b * nen:(b + 1) * nen] - (M if a == b else 0.0)  # = dt/2 * dg_a/dU_b
            if fold is not None:
                Jab = fold(a, b, J, M, nen)
            expect = cap[a, b] * M - Jab
            got = Ke[a * nen:(a + 1) * nen, b * nen:(b + 1) * nen]
            err = np.abs(got - expect).max()
            assert err <= tol * scale, f"block ({a},{b}): err {err:.3e} scale {scale:.3e}"


def test_fd_jacobian_adpm():
    rng = np.random.default_rng(5)
    U = np.stack([0.6 + 0.3 * rng.random(4), 0.2 + 0.5 * rng.random(4), 0.3 + 0.4 * rng.random(4)], axis=1)
    # tract vector aligned with grad(A_b) so that both taxis branches are active
    G = tet_grads(TET)
    gA = G.T @ U[:, 1]
    gT = G.T @ U[:, 2]
    tr = 0.1 * (gA / np.linalg.norm(gA) - 0.8 * gT / np.linalg.norm(gT))
    p = synth.adpm_params("full")
    p[P.TABLES[P.ADPM].index(("produce/A_b/sigmoid/0", 1e20))] = 0.1  # inside the ramp of SD_
    p[[k for k, (key, _) in enumerate(P.TABLES[P.ADPM]) if key == "produce/Tau/sigmoid/0"][0]] = 0.1
    p[[k for k, (key, _) in enumerate(P.TABLES[P.ADPM]) if key == "transform/A_b/trapezoid/1"][0]] = 0.9  # ramp
    p[[k for k, (key, _) in enumerate(P.TABLES[P.ADPM]) if key == "transform/Tau/trapezoid/2"][0]] = 0.2  # down-ramp
    _check_fd(O.ADPM, TET, U, p, efield=tr, time=0.7)


def test_fd_jacobian_pihna():
    rng = np.random.default_rng(6)
    U = np.stack([2e3 * rng.random(4), 4e4 + 2e4 * rng.random(4), 1e4 + 1e4 * rng.random(4),
                  7e3 + 3e3 * rng.random(4), 4e-9 + 4e-9 * rng.random(4)], axis=1)
    _check_fd(O.PIHNA, TET * 10.0, U, synth.pihna_params("full"), tol=5e-6)


def test_fd_jacobian_ripf():
    rng = np.random.default_rng(7)
    U = np.stack([-500 + 100 * rng.random(4), 0.2 + 0.1 * rng.random(4), 0.1 + 0.1 * rng.random(4)], axis=1)
    aux = np.zeros((4, 6))
    aux[:, 1] = 0.01 + 0.01 * rng.random(4)    # TD cc  > tol -> build branch
    aux[:, 2] = -0.02 - 0.01 * rng.random(4)   # TD fb  < -tol -> decay branch
    aux[:, 5] = 20 + 10 * rng.random(4)        # RT_total
    p = synth.ripf_params("full")
    p[18] = 0.0  # fb/lambda/RT/r = 0 -> falls back to int(RT_total_max) (ripf.C:398-399)
    _check_fd(O.RIPF, TET, U, p, aux=aux, rt_max=31, tol=5e-6)


def test_fd_jacobian_proteas():
    """The reference's PROTEAS Jacobian is not the exact derivative in two blocks, and parity means
    reproducing that: [3][1] omits rho_v*Kappa*vsc (proteas.C:661-665 vs 507,551) and [1][1] omits the two
    D_c_h terms of d/d tum (proteas.C:605-615 vs 533).  All other blocks match finite differences."""
    rng = np.random.default_rng(8)
    U = np.stack([0.3 + 0.1 * rng.random(4), 0.2 + 0.1 * rng.random(4), 0.05 + 0.05 * rng.random(4),
                  0.12 + 0.05 * rng.random(4), 0.1 + 0.1 * rng.random(4)], axis=1)
    aux = np.zeros((4, 2))
    aux[:, 0] = 20 + 10 * rng.random(4)
    p = synth.proteas_params()
    dt = 0.1
    _, _, JxW, dphi = O.element(O.PROTEAS, O.TET4, TET, U, p, aux=aux, dt=dt)
    _, phi = O.fe_tables(O.TET4)
    q = phi.T @ U                                      # [nqp, 5] values at the quadrature points
    Kq = np.clip(1.0 - q[:, :4].sum(1) / p[0], 0.0, 1.0)
    dK = -1.0 / p[0]
    rho_v, D_c_h = p[19], p[9]
    Gh = np.einsum("iqd,qd->iq", dphi, np.einsum("lqd,l->qd", dphi, U[:, 0]))   # grad(hos) . grad(phi_i)
    miss31 = np.einsum("q,q,q,jq,iq->ij", JxW, rho_v * Kq, q[:, 3], phi, phi)
    miss11 = np.einsum("q,jq,iq->ij", JxW * (-D_c_h * dK * q[:, 1] - D_c_h * Kq), phi, Gh)

    def fold(a, b, J, M, nen):
        blk = J[a * nen:(a + 1) * nen, b * nen:(b + 1) * nen] - (M if a == b else 0.0)
        if (a, b) == (3, 1):
            return blk - dt / 2 * miss31
        if (a, b) == (1, 1):
            return blk - dt / 2 * miss11
        return blk

    assert np.abs(miss31).max() > 1e-5 and np.abs(miss11).max() > 1e-7
    _check_fd(O.PROTEAS, TET, U, p, aux=aux, dt=dt, fold=fold, tol=5e-6)


def test_fd_jacobian_hcc_reproduces_reference_quirks():
    """coupled_hcc.C:577-619: capacity on [0][1],[0][2],[1][0]; the d/dn block of row c is added to
    [1][1] a second time and [1][2] stays zero."""
    rng = np.random.default_rng(9)
    U = np.stack([0.3 + 0.1 * rng.random(4), 0.2 + 0.1 * rng.random(4), 0.1 + 0.1 * rng.random(4)], axis=1)
    cap = np.array([[1.0, 1.0, 1.0], [1.0, 2.0, 0.0], [0.0, 0.0, 1.0]])

    def fold(a, b, J, M, nen):
        blk = lambda r, c: J[r * nen:(r + 1) * nen, c * nen:(c + 1) * nen] - (M if r == c else 0.0)
        if (a, b) == (1, 1):
            return blk(1, 1) + blk(1, 2)
        if (a, b) == (1, 2):
            return np.zeros((nen, nen))
        return blk(a, b)

    _check_fd(O.HCC, TET, U, synth.hcc_params(), cap=cap, fold=fold, tol=5e-6)


def test_fd_jacobian_adpm_hex():
    conn, xyz = synth.hex_cube(1, 0.7)
    X = xyz[conn[0]] + 0.05 * np.random.default_rng(10).random((8, 3))
    rng = np.random.default_rng(11)
    U = np.stack([0.6 + 0.3 * rng.random(8), 0.2 + 0.5 * rng.random(8), 0.3 + 0.4 * rng.random(8)], axis=1)
    p = synth.adpm_params("full")
    _check_fd(O.ADPM, X, U, p, efield=[0.0, 0.0, 0.0], elem=O.HEX8, time=0.4)


# ------------------------------------------------------------------------------- check_solution KATs
def test_ripf_check_solution_schedule():
    """ripf.C:705,739-741,753-759,770-772: day = floor(time); TD from the UNCLAMPED previous vector;
    RT_total schedule; RT_total_max truncated to int."""
    N = 4
    p = synth.ripf_params("ref")
    u = np.array([[-1200.0, -0.1, 0.2], [1500.0, 0.3, -0.2], [10.0, 0.5, 0.5], [-20.0, 0.0, 0.0]])
    rt = np.array([[56.0, 6.4], [28.0, 3.2], [0.0, 0.0], [67.0, 6.7]])
    pr = O.Problem(O.RIPF, O.TET4, np.array([[0, 1, 2, 3]], dtype=np.int32), TET, p, u, nodal_field=rt)
    pr.ripf_initial_check(0.1)
    np.testing.assert_allclose(pr.u.reshape(N, 3), [[-1000, 0, 0.2], [1000, 0.3, 0], [10, .5, .5], [-20, 0, 0]])
    np.testing.assert_allclose(pr.prev.reshape(N, 3), u)  # unclamped
    np.testing.assert_allclose(pr.aux[:, :3], (pr.u.reshape(N, 3) - u) / 0.1)
    np.testing.assert_allclose(pr.aux[:, 5], rt[:, 0] / 28 * 1)  # day 0 of the broad phase
    assert pr.rt_max.value == int(67.0 / 28)  # = 2
    import ctypes as C
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    for time, exp in ((27.5, rt[:, 0]), (30.2, rt[:, 1] / 8 * (31 - 28) + rt[:, 0]), (40.0, rt.sum(1))):
        m = O.lib().orc_ripf_check(C.c_int64(N), ptr(pr.u), ptr(pr.prev), ptr(pr.aux), ptr(pr.params),
                                   C.c_double(time), C.c_double(0.1))
        np.testing.assert_allclose(pr.aux[:, 5], exp, rtol=1e-15)
        assert m == int(exp.max())


def test_gmsh_reader_on_shipped_meshes():
    import os
    base = "/root/reference/run/Solid"
    if not os.path.isdir(base):
        pytest.skip("reference tree not present (GPU box)")
    et, conn, xyz, sub = synth.read_gmsh(f"{base}/uniaxial_compression/cube.msh")
    assert et == O.HEX8 and conn.shape == (512, 8) and xyz.shape == (729, 3)
    et, conn, xyz, sub = synth.read_gmsh(f"{base}/hydrogel_tension/hydrogel_model.msh")
    assert et == O.TET4 and conn.shape == (5504, 4) and xyz.shape == (1569, 3)
    X = xyz[conn]
    assert (np.linalg.det(X[:, 1:] - X[:, :1]) > 0).all()  # libMesh needs positive Jacobians


# ---- save_solution reductions (oracle restatement of adpm.C:690-829 / pihna.C:842-976 / ripf.C:777-864) ----
def test_region_reductions_known_answers():
    import cases
    conn, xyz = cases.mesh(cases.TET4, 5, distort=0.2)
    N, E = xyz.shape[0], conn.shape[0]
    u = np.zeros((N, 3))
    u[:, 0] = 2.0 + 3.0 * xyz[:, 0] - xyz[:, 1] + 0.5 * xyz[:, 2]       # linear field
    u[:, 1] = (xyz[:, 0] > 0.5).astype(float)
    region = (np.arange(E) % 4).astype(np.int32)
    # open range: every element counts, region volumes add up to the mesh volume, each equals the sum of its tets
    vol = O.region_volumes(4, conn, xyz, u, [([1, 0, 0], 1.0, -1e300, 1e300)], region, 4)
    a, b, c, d = (xyz[conn[:, k]] for k in range(4))
    tet = np.einsum("ij,ij->i", np.cross(b - a, c - a), d - a) / 6.0
    assert abs(vol.sum() - 1.0) < 1e-13
    for r in range(4):
        assert abs(vol[r] - tet[region == r].sum()) < 1e-13
    # a range test is an AND over the nodes of an element and over the conditions
    inside = np.all(u[conn, 1] >= 0.5, axis=1)
    vol1 = O.region_volumes(4, conn, xyz, u, [([0, 1, 0], 1.0, 0.5, 1e300)])
    assert abs(vol1[0] - tet[inside].sum()) < 1e-13
    both = inside & np.all(u[conn, 0] <= 4.0, axis=1)
    vol2 = O.region_volumes(4, conn, xyz, u, [([0, 1, 0], 1.0, 0.5, 1e300), ([1, 0, 0], 1.0, -1e300, 4.0)])
    assert abs(vol2[0] - tet[both].sum()) < 1e-13
    # the element average of a linear field is its value at the centroid; the LAST element of a region wins
    mean = O.region_last_mean(4, conn, xyz, u, 0, region, 4)
    for r in range(4):
        e = np.nonzero(region == r)[0][-1]
        assert abs(mean[r] - u[conn[e], 0].mean()) < 1e-12
    # HEX8: volumes from the 2x2x2 rule
    conn8, xyz8 = cases.mesh(cases.HEX8, 4, distort=0.2)
    v8 = O.region_volumes(8, conn8, xyz8, np.ones((xyz8.shape[0], 3)), [([1, 0, 0], 1.0, 0.0, 2.0)])
    assert abs(v8[0] - 1.0) < 1e-13
