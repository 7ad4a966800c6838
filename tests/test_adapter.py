"""adapter/rdc_libmesh_adapter.h -- the reference-side glue of INTEGRATION.md -- compiled and run.

libMesh cannot be installed here, so the adapter is built against the serial stand-in oracle/ref_shim/libmesh (the
interface the reference's own model files compile against unchanged).  CPU: it compiles and links against
librdcgpu.so.  GPU: tests/adapter/adapter_check.cpp runs the patched time loop of adpm.C:60-84 (assemble callback ->
RdcAdapter::assemble, RdcLinearSolver on model.linear_solver, rotate / check_solution / pull_solution) for every model
and the result is compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import cases
from rdcfes_b200 import build as B
from rdcfes_b200 import params as P

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.path.join(HERE, "adapter", "adapter_check")


def build_adapter_check():
    so = B.build()
    src = os.path.join(HERE, "adapter", "adapter_check.cpp")
    deps = [src, os.path.join(ROOT, "adapter", "rdc_libmesh_adapter.h"), os.path.join(ROOT, "oracle", "ref_shim", "libmesh", "shim.h"), so]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "oracle", "ref_shim"), "-o", EXE, src,
                           "-L", os.path.dirname(so), "-lrdcgpu", "-Wl,-rpath," + os.path.dirname(so)])
    return EXE


def test_adapter_compiles_against_the_libmesh_interface():
    assert os.path.exists(build_adapter_check())


def _write_input(path, model, conn, xyz, p, u0, ef, nf, dt, nsteps, ksp):
    table = P.TABLES[model]
    with open(path, "w") as f:
        N, E = xyz.shape[0], conn.shape[0]
        f.write(f"{model} {conn.shape[1]} {N} {E} {dt!r} {nsteps} {ksp} {len(table)} {0 if ef is None else 3} {0 if nf is None else 2}\n")
        for (key, _), v in zip(table, p):
            kind = "int" if key in ("RT_dose/broad/fractions", "RT_dose/focus/fractions") else "real"
            f.write(f"{key} {kind} {float(v)!r}\n")
        for arr in (xyz, conn, np.asarray(u0).reshape(-1)) + (() if ef is None else (np.asarray(ef).reshape(-1),)) + \
                (() if nf is None else (np.asarray(nf).reshape(-1),)):
            np.savetxt(f, np.asarray(arr).reshape(1, -1), fmt="%d" if np.asarray(arr).dtype.kind == "i" else "%.17g")


@pytest.mark.gpu
@pytest.mark.parametrize("ksp", ["bicgstab", "gmres"])
@pytest.mark.parametrize("model", range(5))
def test_adapter_runs_the_patched_time_loop(model, ksp, tmp_path):
    from oracle import oracle as O
    exe = build_adapter_check()
    conn, xyz = cases.mesh(cases.TET4, 5, distort=0.2, length=50.0 if model == cases.RIPF else 1.0)
    p, u0, ef, nf = cases.case(model, conn, xyz, "full")
    if model == cases.PIHNA and ksp == "gmres":
        pytest.skip("covered by the BiCGStab run")
    dt, nsteps = cases.DT[model], 3
    inp, out = tmp_path / "in.txt", tmp_path / "out.txt"
    _write_input(inp, model, conn, xyz, p, u0, ef, nf, dt, nsteps, ksp)
    res = subprocess.run([exe, str(inp), str(out)], capture_output=True, text=True, timeout=300)
    print(res.stdout, res.stderr[-2000:])
    assert res.returncode == 0
    u = np.loadtxt(out)
    orc = cases.oracle_problem(model, cases.TET4, conn, xyz, p, u0, ef, nf)
    for _ in range(nsteps):
        orc.step(dt, pc=O.PC_ILU)
    # north_star: species after N steps within 1e-6; 1e-8 holds for every model but RIPF, whose source terms switch on the
    # sign of a time derivative that is round-off noise where nothing changes (ripf.C:491-496) -- two converged Krylov
    # solvers may take different branches there from the second step on
    tol = 1e-6 if model == cases.RIPF else 1e-8
    assert np.linalg.norm(u - orc.u) <= tol * np.linalg.norm(orc.u)


# ---- the SolidSystem glue (RdcSolidAdapter) next to the reference's own solid_system.C ------------------------------------
SOLID_EXE = os.path.join(HERE, "adapter", "solid_adapter_check")
REF_SOLID = "/root/reference/src/solid_system.C"


def build_solid_adapter_check():
    """One program with both sides: the reference's solid_system.C + eig3.C compiled unchanged against the libMesh stand-in,
    and the device path reached through adapter/rdc_libmesh_adapter.h.  Needs the reference sources; the binary travels."""
    so = B.build()
    src = os.path.join(HERE, "adapter", "solid_adapter_check.cpp")
    if not os.path.exists(REF_SOLID):
        return SOLID_EXE if os.path.exists(SOLID_EXE) else None
    deps = [src, os.path.join(ROOT, "adapter", "rdc_libmesh_adapter.h"), os.path.join(ROOT, "oracle", "ref_shim", "libmesh", "shim.h"),
            os.path.join(ROOT, "oracle", "ref_shim", "ref_solid.cpp"), so]
    if os.path.exists(SOLID_EXE) and all(os.path.getmtime(SOLID_EXE) >= os.path.getmtime(d) for d in deps):
        return SOLID_EXE
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++17", "-O1", "-w", "-ffp-contract=off", "-I", os.path.join(ROOT, "oracle", "ref_shim"),
                           f'-DREF_SOURCE="{REF_SOLID}"', "-o", SOLID_EXE, src, "/root/reference/src/eig3.C",
                           "-L", os.path.dirname(so), "-lrdcgpu", "-Wl,-rpath," + os.path.dirname(so)])
    return SOLID_EXE


def test_solid_adapter_compiles_against_the_libmesh_interface():
    exe = build_solid_adapter_check()
    if exe is None:
        pytest.skip("reference sources not on this machine and no prebuilt binary")
    assert os.path.exists(exe)


def _write_solid_case(path, c, xpert, pseudo_time, nload, loading_step):
    se, sn, sb, bd = c.arrays()
    with open(path, "w") as f:
        f.write(f"{c.elem_type} {c.N} {c.E}\n")
        for arr in (c.xund, xpert, c.conn, c.mat_of if c.mat_of is not None else np.zeros(c.E, dtype=np.int32)):
            a = np.asarray(arr).reshape(1, -1)
            np.savetxt(f, a, fmt="%d" if a.dtype.kind == "i" else "%.17g")
        f.write(f"{c.mats.shape[0]}\n")
        np.savetxt(f, c.mats.reshape(1, -1), fmt="%.17g")
        np.savetxt(f, (c.fibres if c.fibres is not None else np.ones((c.E, 3))).reshape(1, -1), fmt="%.17g")
        f.write(f"{len(c.bc_ids)}\n")
        for i, d in zip(c.bc_ids, c.bc_disp):
            f.write(f"{i} " + " ".join("nan" if v != v else repr(float(v)) for v in d) + "\n")
        f.write(f"{se.shape[0]}\n")
        for k in range(se.shape[0]):
            f.write(f"{se[k]} {sn[k]} {sb[k]}\n")
        f.write(f"{c.penalty!r} {pseudo_time!r} {nload} {loading_step!r}\n")
        f.write(" ".join(repr(float(v)) for v in c.opts_vector()) + "\n")


@pytest.mark.gpu
@pytest.mark.parametrize("et", [cases.TET4, cases.HEX8])
def test_solid_adapter_against_the_reference_solid_system(et, tmp_path):
    """Device Jacobian / residual / post-processing, reached from libMesh objects through RdcSolidAdapter, against the
    reference's own element_time_derivative + side_time_derivative + post_process in the same process; then load steps through
    RdcSolidAdapter::run_solver against the oracle's Newton driver."""
    import solid_cases as SC
    from oracle import solid as S
    exe = build_solid_adapter_check()
    if exe is None:
        pytest.skip("no solid_adapter_check binary")
    # (A), (B): every term switched on
    c = SC.general_case(et, n=3)
    inp, out = tmp_path / "case.txt", tmp_path / "out.txt"
    _write_solid_case(inp, c, SC.perturbed(c, amp=0.01), 0.3, 0, 0.1)
    res = subprocess.run([exe, str(inp), str(out)], capture_output=True, text=True, timeout=300)
    print(res.stdout, res.stderr[-2000:])
    assert res.returncode == 0
    head = np.loadtxt(out, max_rows=1)
    assert head[0] == 1, "sparsity pattern differs from the one the reference touches"
    assert head[1] <= 1e-12 and head[2] <= 1e-12, head          # Jacobian, residual
    assert head[3] <= 1e-10 and head[4] <= 1e-12, head          # stresses, fibre vector
    # (C): load steps of the compression problem, tight tolerances
    c = SC.compression_case(et, n=3, penalty=1.0e6)
    c.opts.update(max_nonlinear_iterations=25, relative_step_tolerance=1e-11, relative_residual_tolerance=1e-13,
                  absolute_residual_tolerance=1e-9, initial_linear_tolerance=1e-10)
    c.mat_of = np.zeros(c.E, dtype=np.int32)
    _write_solid_case(inp, c, c.xund, 0.0, 3, 0.1)
    res = subprocess.run([exe, str(inp), str(out)], capture_output=True, text=True, timeout=300)
    print(res.stdout, res.stderr[-2000:])
    assert res.returncode == 0
    head = np.loadtxt(out, max_rows=1)
    assert head[5] == 1
    x = np.loadtxt(out, skiprows=1)
    orc = S.OracleSolid(c)
    xo = c.xund.copy().ravel()
    for l in (1, 2, 3):
        xo, info = orc.newton(xo, 0.1 * l)
        assert info["converged"]
    assert np.linalg.norm(x - xo) <= 1e-8 * np.linalg.norm(xo)
