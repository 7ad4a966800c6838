"""The CUDA path (through the C ABI) against vectors produced by the reference's own sources
(tests/golden/ref_*.npz, tests/golden/make_ref_golden.py): operator, load vector, one step, N steps, RIPF state.
north_star tolerances: K/F entries 1e-12 relative, per-step solution 1e-8 relative L2, species after N steps 1e-6."""
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MESHES = {"tet": (cases.TET4, 4), "hex": (cases.HEX8, 3)}


@pytest.mark.parametrize("ksp", [2, 0])
@pytest.mark.parametrize("tag", ["tet", "hex"])
@pytest.mark.parametrize("model", range(5))
def test_gpu_matches_reference_vectors(model, tag, ksp):
    g = np.load(os.path.join(GOLD, f"ref_{cases.NAMES[model]}_{tag}.npz"))
    et, n = MESHES[tag]
    conn, xyz = cases.mesh(et, n, distort=0.2, length=50.0 if model == cases.RIPF else 1.0)
    p, u0, ef, nf = cases.case(model, conn, xyz, "full")
    gpu = cases.gpu_system(model, et, conn, xyz, p, u0, ef, nf)
    gpu.ksp = ksp
    dt = cases.DT[model]
    gpu.rotate()
    gpu.assemble(dt, dt)
    rows, rowptr, col, val, rhs = gpu.download_csr()
    gpu.time = 0.0
    assert np.array_equal(rowptr, g["rowptr"]) and np.array_equal(col, g["col"])      # pattern bit-exact
    ref = g["val"]
    nz = ref != 0.0
    assert np.abs(val[~nz]).max(initial=0.0) == 0.0                                    # explicit zeros stay zeros
    worst = (np.abs(val[nz] - ref[nz]) / np.abs(ref[nz])).max()
    # pure relative 1e-12 on every entry of ordinary size; entries that are the remainder of a cancellation between
    # element contributions (|K_ij| < 1e-3 max|K|) are held to 1e-12 of that scale
    tol = 1e-12 * np.maximum(np.abs(ref), 1e-3 * np.abs(ref).max())
    assert (np.abs(val - ref) <= tol).all(), worst
    assert np.abs(rhs - g["rhs"]).max() <= 1e-12 * np.abs(g["rhs"]).max()
    nsteps = int(g["nsteps"])
    for k in range(nsteps):
        gpu.step(dt)
        if k == 0:
            u1 = gpu.get_solution()
            assert np.linalg.norm(u1 - g["u1"]) <= 1e-8 * np.linalg.norm(g["u1"])
    uN = gpu.get_solution()
    nv = cases.P.NVARS[model]
    for a in range(nv):   # every species separately; a species that is numerically absent (PIHNA's n stays ~1e-15 next
        # to cell densities of 1e4) is held to 1e-6 of 1e-12 x the state's norm instead of its own round-off-sized norm
        ra, ga = uN.reshape(-1, nv)[:, a], g["uN"].reshape(-1, nv)[:, a]
        assert np.linalg.norm(ra - ga) <= 1e-6 * max(np.linalg.norm(ga), 1e-12 * np.linalg.norm(g["uN"]))
    assert np.linalg.norm(uN - g["uN"]) <= 1e-8 * np.linalg.norm(g["uN"])
    if model == cases.RIPF:
        assert gpu.stats().ripf_rt_total_max == int(g["rt_max"])
    print(f"{cases.NAMES[model]} {tag} ksp={ksp}: worst pure relative K entry error {worst:.2e}")
    gpu.close()
