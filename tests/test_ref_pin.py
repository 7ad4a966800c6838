"""Pins the CPU oracle to the reference's OWN model sources.

Two layers, both CPU-only:
  * committed vectors (tests/golden/ref_*.npz, made by tests/golden/make_ref_golden.py from the compiled reference
    sources): K, F, one step and N steps of every model on TET4 and HEX8 -- runs anywhere, also on the GPU box;
  * live runs of oracle/_ref (the reference's src/*.C compiled unchanged against the serial libMesh stand-in) next
    to the oracle on more inputs: permuted dof numbering, the two meshes the reference ships, later times,
    check_solution, save_solution and input().  Skipped when neither the prebuilt libraries nor /root/reference exist.
Tolerances: F and the clamped solution bit-for-bit or 1e-15; K entries 1e-12 pure relative on every entry the
reference produces as non-zero (summation order inside Ke differs, nothing else)."""
import os

import numpy as np
import pytest

import cases
from oracle import oracle as O
from oracle import ref as R

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
MESHES = {"tet": (cases.TET4, 4), "hex": (cases.HEX8, 3)}
REF_RUN = "/root/reference/run"


def _inputs(model, tag):
    et, n = MESHES[tag]
    conn, xyz = cases.mesh(et, n, distort=0.2, length=50.0 if model == cases.RIPF else 1.0)
    p, u0, ef, nf = cases.case(model, conn, xyz, "full")
    return et, conn, xyz, p, u0, ef, nf


def _k_close(val, ref, tol=1e-12):
    """pure relative error on the reference's non-zero entries, exact zeros where the reference has zeros"""
    nz = ref != 0.0
    assert np.array_equal(val[~nz], ref[~nz])
    rel = np.abs(val[nz] - ref[nz]) / np.abs(ref[nz])
    assert rel.max() <= tol, rel.max()
    return rel.max()


@pytest.mark.parametrize("tag", ["tet", "hex"])
@pytest.mark.parametrize("model", range(5))
def test_oracle_matches_reference_vectors(model, tag):
    g = np.load(os.path.join(GOLD, f"ref_{cases.NAMES[model]}_{tag}.npz"))
    et, conn, xyz, p, u0, ef, nf = _inputs(model, tag)
    pr = cases.oracle_problem(model, et, conn, xyz, p, u0, ef, nf, nthreads=2)
    dt = cases.DT[model]
    pr.u_old = pr.u.copy()
    val, rhs = pr.assemble(dt, dt)
    assert np.array_equal(pr.rowptr, g["rowptr"]) and np.array_equal(pr.col, g["col"])   # pattern bit-exact
    _k_close(val, g["val"])
    assert np.abs(rhs - g["rhs"]).max() <= 1e-15 * np.abs(g["rhs"]).max()
    n = int(g["nsteps"])
    for k in range(n):
        pr.step(dt, pc=O.PC_ILU)
        if k == 0:
            assert np.linalg.norm(pr.u - g["u1"]) <= 1e-10 * np.linalg.norm(g["u1"])
    assert np.linalg.norm(pr.u - g["uN"]) <= 1e-9 * np.linalg.norm(g["uN"])
    if model == cases.RIPF:
        assert pr.rt_max.value == int(g["rt_max"])
        td, rt = g["td"].reshape(-1, 3), g["rt"].reshape(-1, 3)
        assert np.abs(pr.aux[:, 5] - rt[:, 2]).max() <= 1e-14 * np.abs(rt[:, 2]).max()
        assert np.abs(pr.aux[:, :3] - td).max() <= 1e-8 * max(np.abs(td).max(), 1e-300)


needs_ref = pytest.mark.skipif(not (R.available() or R.have_reference_sources()),
                               reason="oracle/_ref is not built and /root/reference is absent")


@needs_ref
@pytest.mark.parametrize("model", range(5))
def test_live_reference_permuted_dofs_and_later_time(model):
    """node-blocked dofs with an arbitrary base (libMesh numbers nodes in first-touch order, Appendix B-5), t > 0
    (ADPM's decay/PrP * time^gamma, adpm.C:369), a solution that needs clamping"""
    R.build_ref()
    et, conn, xyz, p, u0, ef, nf = _inputs(model, "tet")
    if model == cases.ADPM:
        p = p.copy()
        p[0] = 0.7   # decay/PrP/time_exponent
    nv = cases.P.NVARS[model]
    N = xyz.shape[0]
    base = (np.random.default_rng(5).permutation(N) * nv).astype(np.int32)
    u0 = np.asarray(u0, dtype=float).reshape(N, nv)
    u_perm = np.zeros(N * nv)
    for a in range(nv):
        u_perm[base + a] = u0[:, a]
    dt, t = cases.DT[model], 3.3
    rp = R.RefProblem(model, et, conn, xyz, p, u_perm, ef, nf, node_dof_base=base, dt=dt)
    pr = cases.oracle_problem(model, et, conn, xyz, p, u0, ef, nf)
    if model == cases.RIPF:
        rp.check_solution(0.0, dt)
        rp.rotate()
        pr.u_old = pr.u.copy()
    vo, fo = pr.assemble(t, dt)
    vr, fr = rp.assemble(t, dt)
    # compare through dense matrices in node-major ordering: entry (node i var a, node j var b)
    import scipy.sparse as sp
    Ko = sp.csr_matrix((vo, pr.col, pr.rowptr), shape=(N * nv, N * nv)).toarray()
    Kr = sp.csr_matrix((vr, rp.col, rp.rowptr), shape=(N * nv, N * nv)).toarray()
    idx = (base[:, None] + np.arange(nv)[None, :]).ravel()       # reference dof id of (node, var) in node-major order
    Kr = Kr[np.ix_(idx, idx)]
    nz = Kr != 0
    assert np.array_equal(Ko != 0, nz) or np.abs(Ko[~nz]).max() == 0.0
    assert (np.abs(Ko - Kr)[nz] / np.abs(Kr[nz])).max() <= 1e-12
    assert np.abs(fo - fr[idx]).max() <= 1e-15 * np.abs(fr).max()
    # check_solution on a vector with negative entries
    w = np.cos(np.arange(N * nv) * 0.9) * np.abs(u0).max()
    wp = np.zeros(N * nv)
    wp[idx] = w
    rp.set_solution(wp)
    got = rp.check_solution(t, dt)[idx]
    if model == cases.RIPF:
        keys = [k for k, _ in cases.P.TABLES[cases.RIPF]]
        hu_min, hu_max = p[keys.index("HU/min")], p[keys.index("HU/max")]
        exp = w.reshape(N, 3).copy()
        exp[:, 0] = np.clip(exp[:, 0], hu_min, hu_max)
        exp[:, 1:] = np.maximum(exp[:, 1:], 0.0)
        assert np.array_equal(got, exp.ravel())
    else:
        assert np.array_equal(got, np.maximum(w, 0.0))
    rp.close()


@needs_ref
@pytest.mark.parametrize("name,etype", [("hydrogel_tet4", cases.TET4), ("cube_hex8", cases.HEX8)])
@pytest.mark.parametrize("model", [cases.ADPM, cases.PIHNA, cases.HCC])
def test_live_reference_on_the_shipped_meshes(name, etype, model):
    """run/Solid/hydrogel_tension/hydrogel_model.msh and run/Solid/uniaxial_compression/cube.msh"""
    R.build_ref()
    d = np.load(os.path.join(GOLD, name + ".npz"))
    conn, xyz = d["conn"], d["xyz"]
    p, u0, ef, nf = cases.case(model, conn, xyz, "full")
    dt = cases.DT[model]
    rp = R.RefProblem(model, etype, conn, xyz, p, u0, ef, nf, dt=dt)
    pr = cases.oracle_problem(model, etype, conn, xyz, p, u0, ef, nf)
    vo, fo = pr.assemble(dt, dt)
    vr, fr = rp.assemble(dt, dt)
    assert np.array_equal(rp.rowptr, pr.rowptr) and np.array_equal(rp.col, pr.col)
    _k_close(vo, vr)
    assert np.abs(fo - fr).max() <= 1e-15 * np.abs(fr).max()
    rp.close()


@needs_ref
def test_live_reference_ripf_state_machine():
    """ripf.C:675-775 over a day boundary: TD, RT_total(day), int(max) and the unclamped `prev` vector, fed with the
    same post-solve vectors on both sides"""
    R.build_ref()
    et, conn, xyz, p, u0, ef, nf = _inputs(cases.RIPF, "tet")
    dt = 0.4
    rp = R.RefProblem(cases.RIPF, et, conn, xyz, p, u0, ef, nf, dt=dt)
    pr = cases.oracle_problem(cases.RIPF, et, conn, xyz, p, u0, ef, nf)   # runs the pre-loop check with DT[RIPF]
    pr = O.Problem(cases.RIPF, et, conn, xyz, p, u0, nodal_field=nf)
    pr.ripf_initial_check(dt)
    rp.check_solution(0.0, dt)
    rng = np.random.default_rng(3)
    t = 0.0
    for k in range(7):
        t += dt
        w = pr.u + rng.normal(0, 0.05, pr.u.size) * np.abs(pr.u).max()
        rp.set_solution(w)
        ur = rp.check_solution(t, dt)
        pr.u[:] = w
        m = O.lib().orc_ripf_check(O.C.c_int64(pr.N), O._p(pr.u), O._p(pr.prev), O._p(pr.aux), O._p(pr.params),
                                   O.C.c_double(t), O.C.c_double(dt))
        assert np.array_equal(pr.u, ur)
        assert m == rp.get_int("RT_dose/total/max")
        assert np.array_equal(pr.aux[:, :3].ravel(), rp.get_vector(b"RIPF-TimeDeriv", 0))
        assert np.array_equal(pr.aux[:, 5], rp.get_vector(b"RT", 0).reshape(-1, 3)[:, 2])
    rp.close()


@needs_ref
def test_live_reference_save_solution_adpm(tmp_path):
    """adpm.C:690-829 (the CSV behind rdc_region_volumes / rdc_region_last_mean), with initial_adpm / initial_tracts
    reading the ASCII field files in node / element order (adpm.C:241-246, 279-284)"""
    R.build_ref()
    et, conn, xyz, p, u0, ef, nf = _inputs(cases.ADPM, "tet")
    nreg = 4
    region = np.random.default_rng(11).integers(1, nreg + 1, conn.shape[0]).astype(np.int32)   # subdomain ids 1..4
    rp = R.RefProblem(cases.ADPM, et, conn, xyz, p, np.zeros_like(u0), None, None, subdomain=region, dt=0.05)
    nodal, elemental = tmp_path / "nodal.dat", tmp_path / "elemental.dat"
    np.savetxt(nodal, np.asarray(u0).reshape(-1, 3), fmt="%.17g")
    np.savetxt(elemental, np.asarray(ef).reshape(-1, 3), fmt="%.17g")
    rp.set_string("input_nodal", str(nodal))
    rp.set_string("input_elemental", str(elemental))
    rp.run_init()
    assert np.array_equal(rp.get_vector(b"ADPM", 0), np.asarray(u0).ravel())      # file order = node order
    assert np.array_equal(rp.get_vector(b"Tracts", 0), np.asarray(ef).ravel())    # file order = element order
    u = np.asarray(u0, dtype=float).reshape(-1, 3)
    lo_a, hi_a = float(np.quantile(u[:, 1], 0.3)), float(np.quantile(u[:, 1], 0.9))
    lo_t, hi_t = float(np.quantile(u[:, 2], 0.2)), float(np.quantile(u[:, 2], 0.8))
    for k, v in (("range/A_b/min", lo_a), ("range/A_b/max", hi_a), ("range/Tau/min", lo_t), ("range/Tau/max", hi_t)):
        rp.set_real(k, v)
    header, rows = rp.save_solution(0.0)
    row = rows[-1]
    ids = sorted(set(region.tolist()))
    conc = row[1:1 + 2 * len(ids)].reshape(-1, 2)
    vol = row[1 + 2 * len(ids):].reshape(-1, 2)
    reg0 = region - 1
    va = O.region_volumes(et, conn, xyz, u.ravel(), [([0, 1, 0], 1.0, lo_a, hi_a)], reg0, nreg)
    vt = O.region_volumes(et, conn, xyz, u.ravel(), [([0, 0, 1], 1.0, lo_t, hi_t)], reg0, nreg)
    ma = O.region_last_mean(et, conn, xyz, u.ravel(), 1, reg0, nreg)
    mt = O.region_last_mean(et, conn, xyz, u.ravel(), 2, reg0, nreg)
    assert np.allclose(vol[:, 0], va, rtol=1e-13, atol=0) and np.allclose(vol[:, 1], vt, rtol=1e-13, atol=0)
    assert np.allclose(conc[:, 0], ma, rtol=1e-12, atol=0) and np.allclose(conc[:, 1], mt, rtol=1e-12, atol=0)
    rp.close()


@needs_ref
@pytest.mark.parametrize("model,path", [(cases.ADPM, "HCP102513/input.dat"), (cases.PIHNA, "PIHNA/input.dat"),
                                         (cases.RIPF, "RIPF133/input.dat"), (cases.HCC, "Coupled/HCC/input.dat")])
def test_live_reference_input_parsing(model, path):
    """the reference's input() on the input.dat files it ships == rdcfes_b200/params.py (keys, defaults, unit
    conversions, ignored keys: SURVEY Appendix C-1)"""
    full = os.path.join(REF_RUN, path)
    if not os.path.exists(full):
        pytest.skip("reference run directory not on this machine")
    R.build_ref()
    conn, xyz = cases.mesh(cases.TET4, 2)
    p0, u0, ef, nf = cases.case(model, conn, xyz, "full")
    rp = R.RefProblem(model, cases.TET4, conn, xyz, p0, u0, ef, nf)
    rp.input(full)
    flat, kv = cases.P.params_from_input(model, full)
    for (key, _), v in zip(cases.P.TABLES[model], flat):
        if key in ("RT_dose/broad/fractions", "RT_dose/focus/fractions"):
            assert rp.get_int(key) == int(v)
        elif key == "volume_fraction/max_vacant":
            assert rp.get_real(key) == v
        else:
            got = rp.get_real(key)
            assert got == v or (np.isnan(got) and np.isnan(v)), (key, got, v)
    assert rp.get_real("time_step") == float(kv["time_step"])
    rp.close()
