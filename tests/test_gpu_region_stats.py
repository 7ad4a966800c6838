"""Device-side save_solution reductions (rdc_region_volumes / rdc_region_last_mean, csrc/reduce.cu) against the
oracle's serial restatement of adpm.C:690-829, pihna.C:842-976 and ripf.C:777-864."""
import numpy as np
import pytest

import cases
from cases import ADPM, HEX8, PIHNA, RIPF, TET4
from oracle import oracle as O

pytestmark = pytest.mark.gpu
BIG = 1e300


def _setup(model, etype, n, nreg, steps=1):
    length = 50.0 if model == RIPF else 1.0
    conn, xyz = cases.mesh(etype, n, distort=0.2, length=length)
    p, u0, ef, nf = cases.case(model, conn, xyz, "full")
    gpu = cases.gpu_system(model, etype, conn, xyz, p, u0, ef, nf)
    for _ in range(steps):
        gpu.step(cases.DT[model])
    u = gpu.get_solution()
    region = None
    if nreg > 1:
        region = np.random.default_rng(5).integers(0, nreg, conn.shape[0]).astype(np.int32)
        region[-1] = 0  # a region whose last element is the very last one
    gpu.set_subdomains(region, nreg)
    return gpu, conn, xyz, u, region


def _check(got, ref):
    assert np.abs(got - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1e-300), (got, ref)


@pytest.mark.parametrize("etype,n", [(TET4, 7), (HEX8, 5)])
def test_adpm_parcellation_outputs(etype, n):
    nreg = 7
    gpu, conn, xyz, u, region = _setup(ADPM, etype, n, nreg)
    U = u.reshape(-1, 3)
    for var in (1, 2):  # A_b, Tau (adpm.C:704-707 ranges)
        lo, hi = np.quantile(U[:, var], 0.3), np.quantile(U[:, var], 0.95)
        w = [0, 0, 0]
        w[var] = 1.0
        conds = [(w, 1.0, lo, hi)]
        _check(gpu.region_volumes(conds), O.region_volumes(etype, conn, xyz, u, conds, region, nreg))
        _check(gpu.region_last_mean(var), O.region_last_mean(etype, conn, xyz, u, var, region, nreg))
    # every element passes an open range: the region volumes add up to the mesh volume
    vol = gpu.region_volumes([([1, 0, 0], 1.0, -BIG, BIG)])
    assert abs(vol.sum() - 1.0) <= 1e-12
    gpu.close()


def test_pihna_volumes_single_region():
    gpu, conn, xyz, u, region = _setup(PIHNA, TET4, 7, 1)
    U = u.reshape(-1, 5)
    kappa = 2.39e5
    ch = U[:, 1] + U[:, 2]
    tot = (U[:, 0] + U[:, 1] + U[:, 2] + U[:, 3]) / kappa
    for conds in ([([0, 1, 1, 0, 0], 1.0, np.quantile(ch, 0.5), BIG)],                      # active tumour: c + h
                  [([1, 0, 0, 0, 0], 1.0, np.quantile(U[:, 0], 0.6), BIG)],                # necrotic
                  [([0, 0, 0, 1, 0], 1.0, 0.0, np.quantile(U[:, 3], 0.7))],                # vascularity
                  [([1, 1, 1, 1, 0], kappa, np.quantile(tot, 0.2), np.quantile(tot, 0.9))]):  # total cell / Kappa_k
        got = gpu.region_volumes(conds)
        ref = O.region_volumes(TET4, conn, xyz, u, conds)
        assert 0.0 < ref[0] < 1.0
        _check(got, ref)
    gpu.close()


def test_ripf_two_conditions_per_node():
    gpu, conn, xyz, u, region = _setup(RIPF, TET4, 6, 3, steps=2)
    U = u.reshape(-1, 3)
    conds = [([1, 0, 0], 1.0, np.quantile(U[:, 0], 0.1), np.quantile(U[:, 0], 0.9)),   # HU window
             ([0, 1, 0], 1.0, np.quantile(U[:, 1], 0.3), BIG)]                         # cc >= min  (ripf.C:826-829)
    _check(gpu.region_volumes(conds), O.region_volumes(TET4, conn, xyz, u, conds, region, 3))
    gpu.close()


def test_argument_errors():
    from rdcfes_b200.lib import RdcError
    gpu, conn, xyz, u, region = _setup(ADPM, TET4, 4, 1, steps=0)
    with pytest.raises(RdcError):
        gpu.region_last_mean(9)
    with pytest.raises(RdcError):
        gpu.region_volumes([([1, 0, 0], 0.0, 0, 1)])   # div == 0
    bad = np.full(conn.shape[0], 5, dtype=np.int32)
    with pytest.raises(RdcError):
        gpu.set_subdomains(bad, 2)
    gpu.close()
