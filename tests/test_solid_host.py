"""The element arithmetic of the CUDA solid path (rdcfes_b200/csrc/solid_dev.cuh) compiled for the host and held to the
CPU oracle (CPU only, no device): the closed-form tangent row, the penalty row and post_process of one element.  The
kernels of solid.cu call exactly these functions; the GPU tests (test_gpu_solid.py) then cover staging and assembly."""
import numpy as np
import pytest

import solid_cases as SC
from oracle import solid as S
from rdcfes_b200 import solid as G


@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
def test_tangent_rows_match_oracle(et, sym):
    # sym: solver/assembly_use_symmetry -- blocks j >= i evaluated, the others mirrored (solid_system.C:248-262)
    c = SC.general_case(et, penalty=0.0, use_symmetry=sym)
    x = SC.perturbed(c)
    orc = S.OracleSolid(c)
    nen = c.conn.shape[1]
    worst_r = worst_k = 0.0
    for e in range(c.E):
        Ro, Ko = orc.element(e, x, 0.3)
        Ko = Ko.reshape(3, nen, 3, nen)                     # variable-major: (a, i, c, j)
        nodes = c.conn[e]
        for li in range(nen):
            R, K = G.probe_row(et, x[nodes], c.xund[nodes], c.mats[c.mat_of[e]], 0.3, c.fibres[e], li, use_symmetry=sym)
            worst_r = max(worst_r, np.abs(R - Ro.reshape(3, nen)[:, li]).max() / np.abs(Ro).max())
            worst_k = max(worst_k, np.abs(K - Ko[:, li, :, :]).max() / np.abs(Ko).max())
    assert worst_r <= 1e-12 and worst_k <= 1e-12, (worst_r, worst_k)


@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
def test_penalty_rows_match_oracle(et):
    c = SC.general_case(et, penalty=1.0e5)
    x = SC.perturbed(c)
    nen = c.conn.shape[1]
    se, sn, sb, bd = c.arrays()
    L = S.O.lib()
    import ctypes as C
    for k in range(se.shape[0]):
        nodes = c.conn[se[k]]
        Xc, Xu = np.ascontiguousarray(x[nodes]), np.ascontiguousarray(c.xund[nodes])
        Re, Ke = np.zeros(3 * nen), np.zeros((3 * nen, 3 * nen))
        disp = np.ascontiguousarray(bd[sb[k]])
        assert L.orc_solid_side(C.c_int(et), C.c_int(int(sn[k])), S._p(Xc), S._p(Xu), S._p(disp), C.c_double(0.3), C.c_double(c.penalty),
                                C.c_int(1), S._p(Re), S._p(Ke)) == 0
        Ke = Ke.reshape(3, nen, 3, nen)
        loc = S.SIDE_NODES[et][sn[k]]
        for i, li in enumerate(loc):
            R, Kd = G.probe_bc_row(len(loc), Xc[list(loc)], Xu[list(loc)], disp, 0.3, c.penalty, i)
            assert np.abs(R - Re.reshape(3, nen)[:, li]).max() <= 1e-12 * np.abs(Re).max()
            for j, lj in enumerate(loc):
                for d in range(3):
                    assert abs(Kd[j, d] - Ke[d, li, d, lj]) <= 1e-12 * np.abs(Ke).max()


@pytest.mark.parametrize("et", [SC.TET4, SC.HEX8])
def test_post_process_matches_oracle(et):
    c = SC.general_case(et)
    x = SC.perturbed(c)
    po, vo, fo = S.OracleSolid(c).post(x, 0.4)
    scale = np.abs(po).max() + vo.max()
    for e in range(c.E):
        nodes = c.conn[e]
        out = G.probe_post(et, x[nodes], c.xund[nodes], c.mats[c.mat_of[e]], 0.4, c.fibres[e])
        assert abs(out[0] - po[e]) <= 1e-12 * scale and abs(out[1] - vo[e]) <= 1e-12 * scale
        assert np.abs(out[2:] - fo[e]).max() <= 1e-13


def test_tangent_rows_on_random_elements():
    """Random affine-plus-noise elements, deformation gradients, materials (Young 1e1..1e5, Poisson 0..0.45, with and without
    fibres) and anisotropic growth: the closed form equals the oracle's literal 3^8 evaluation to 1e-13 of the largest entry."""
    import ctypes as C
    from oracle import oracle as O
    L = O.lib()
    L.orc_solid_element.restype = C.c_int
    rng = np.random.default_rng(0)
    base = {SC.TET4: np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float),
            SC.HEX8: np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0], [0, 0, 1], [1, 0, 1], [1, 1, 1], [0, 1, 1]], float)}
    done = 0
    for trial in range(120):
        et = SC.TET4 if trial % 2 == 0 else SC.HEX8
        A = np.eye(3) + 0.3 * rng.normal(size=(3, 3))
        Fd = np.eye(3) + 0.25 * rng.normal(size=(3, 3))
        if np.linalg.det(A) < 0.2 or np.linalg.det(Fd) < 0.3:
            continue
        Xu = np.ascontiguousarray(base[et] @ A.T * rng.uniform(0.01, 10) + 0.05 * rng.normal(size=base[et].shape))
        Xc = np.ascontiguousarray(Xu @ Fd.T + 0.02 * rng.normal(size=Xu.shape) * np.abs(Xu).max())
        mat = np.array([10 ** rng.uniform(1, 5), rng.uniform(0.0, 0.45), rng.choice([0.0, 10 ** rng.uniform(0, 3)]), *rng.uniform(-0.5, 0.8, 3)])
        t, eta = rng.uniform(0, 1), rng.normal(size=3)
        Re, Ke = np.zeros(3 * et), np.zeros((3 * et, 3 * et))
        assert L.orc_solid_element(C.c_int(et), S._p(Xc), S._p(Xu), S._p(mat), C.c_double(t), S._p(eta), C.c_int(1), C.c_int(0), S._p(Re), S._p(Ke)) == 0
        if not np.isfinite(Ke).all():
            continue
        K4, R2 = Ke.reshape(3, et, 3, et), Re.reshape(3, et)
        for li in range(et):
            R, K = G.probe_row(et, Xc, Xu, mat, t, eta, li)
            assert np.abs(R - R2[:, li]).max() <= 1e-13 * np.abs(Re).max()
            assert np.abs(K - K4[:, li]).max() <= 1e-13 * np.abs(Ke).max()
        done += 1
    assert done >= 60
