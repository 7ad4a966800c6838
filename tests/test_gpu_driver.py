"""The stand-alone C++ driver (driver/rdc_driver.cpp) end to end: Gmsh mesh + input.dat + field files in, the
reference's time loop over the C ABI, save_solution CSV out -- compared with the Python mirror (same library, so the
numbers must be identical) and with the oracle."""
import os
import subprocess

import numpy as np
import pytest

import cases
from oracle import oracle as O
from rdcfes_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_gmsh(path, conn, xyz, ids):
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % xyz.shape[0])
        for k, (x, y, z) in enumerate(xyz):
            f.write("%d %.17g %.17g %.17g\n" % (k + 1, x, y, z))
        f.write("$EndNodes\n$Elements\n%d\n" % (conn.shape[0] + 2))
        f.write("1 2 2 99 1 1 2 3\n2 2 2 99 1 2 3 4\n")                      # two boundary triangles: skipped
        for e, c in enumerate(conn):
            f.write("%d 4 2 %d %d %s\n" % (e + 3, ids[e], ids[e], " ".join(str(v + 1) for v in c)))
        f.write("$EndElements\n")


def _write_input(path, model, flat, extra):
    """input.dat with every key of the model's table (angles back in degrees, like a user would write them)."""
    from rdcfes_b200 import params as P
    with open(path, "w") as f:
        f.write("# written by tests/test_gpu_driver.py\n" + extra)
        for (key, _), v in zip(P.TABLES[model], flat):
            f.write(f"{key} = {float(np.degrees(v) if key in P._ANGLE_KEYS else v)!r}\n")


@pytest.mark.parametrize("model", [cases.PIHNA, cases.RIPF])
def test_cpp_driver_pihna_ripf(tmp_path, model):
    """The same end-to-end check for the other two models with shipped run directories (run/PIHNA, run/RIPF133)."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "driver"), "-s"])
    length = 50.0 if model == cases.RIPF else 1.0
    conn, xyz = cases.mesh(cases.TET4, 6, distort=0.2, length=length)
    p, u0, ef, nf = cases.case(model, conn, xyz, "full")
    nv = cases.P.NVARS[model]
    d = str(tmp_path)
    _write_gmsh(os.path.join(d, "cube.msh"), conn, xyz, np.ones(conn.shape[0], dtype=int))
    np.savetxt(os.path.join(d, "nodal.dat"), np.asarray(u0).reshape(-1, nv), fmt="%.17g")
    extra = "input_GMSH = cube.msh\ninput_nodal = nodal.dat\noutput_CSV = out.csv\n"
    if model == cases.RIPF:
        np.savetxt(os.path.join(d, "rt.dat"), nf, fmt="%.17g")
        extra += "input_nodal_RT = rt.dat\n"
    nsteps, dt = 4, cases.DT[model]
    U0 = np.asarray(u0).reshape(-1, nv)
    if model == cases.PIHNA:
        rng = {"range/active_tumor/min": 50.0, "range/necrotic/min": 1.0, "range/vascularity/max": 7000.0,
               "range/total_cell/min": 0.03, "range/total_cell/max": 0.2}
    else:
        rng = {"range_cc/HU/min": -900.0, "range_cc/HU/max": -100.0, "range_cc/min": float(np.quantile(U0[:, 1], 0.5)),
               "range_fb/min": -1.0}
    extra += f"time_step_number = {nsteps}\ntime_step = {dt}\noutput_step = 2\n"
    extra += "".join(f"{k} = {float(v)!r}\n" for k, v in rng.items())
    _write_input(os.path.join(d, "input.dat"), model, p, extra)
    sol = os.path.join(d, "u.bin")
    out = subprocess.run([os.path.join(ROOT, "driver", "rdc_driver"), "-m", cases.NAMES[model], os.path.join(d, "input.dat"),
                          "ksp=2", "solution_out=" + sol], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    u_drv = np.fromfile(sol)
    rows = [ln.split(",") for ln in open(os.path.join(d, "out.csv")).read().strip().splitlines()]
    if model == cases.PIHNA:
        assert rows[0][1] == '"DEGREES_OF_FREEDOM"'
        rows = rows[1:]
    table = np.array([[float(x) for x in r] for r in rows])

    gpu = cases.gpu_system(model, cases.TET4, conn, xyz, p, u0, ef, nf)
    gpu.ksp = 2
    BIG = float("inf")

    def line(time):
        if model == cases.PIHNA:
            kappa = p[1]
            v = [gpu.region_volumes([c])[0] for c in (([0, 1, 1, 0, 0], 1.0, 50.0, 1e12), ([1, 0, 0, 0, 0], 1.0, 1.0, 1e12),
                                                      ([0, 0, 0, 1, 0], 1.0, 1e-12, 7000.0),
                                                      ([1, 1, 1, 1, 0], kappa, 0.03, 0.2))]
            return np.array([time, 5 * xyz.shape[0]] + v)
        hu_min, hu_max = p[27], p[28]
        cc = gpu.region_volumes([([1, 0, 0], 1.0, -900.0, -100.0), ([0, 1, 0], 1.0, rng["range_cc/min"], BIG)])[0]
        fb = gpu.region_volumes([([1, 0, 0], 1.0, hu_min, hu_max), ([0, 0, 1], 1.0, -1.0, BIG)])[0]
        return np.array([time, cc, fb])

    ref = [line(0.0)]
    orc = cases.oracle_problem(model, cases.TET4, conn, xyz, p, u0, ef, nf)
    for t in range(1, nsteps + 1):
        gpu.step(dt)
        orc.step(dt, pc=O.PC_ILU)
        if t % 2 == 0:
            ref.append(line(gpu.time))
    ref = np.array(ref)
    assert table.shape == ref.shape
    assert np.array_equal(table, ref), (table, ref)
    assert np.array_equal(u_drv, gpu.get_solution())
    assert np.linalg.norm(u_drv - orc.u) <= 1e-8 * np.linalg.norm(orc.u)
    assert 0.0 < table[-1, -1]           # the thresholds select something
    gpu.close()


def _read_appended_vtu(path):
    """{name: array} of a raw-appended .vtu written by driver/vtu_writer.h"""
    import xml.etree.ElementTree as ET
    raw = open(path, "rb").read()
    cut = raw.index(b'<AppendedData encoding="raw">')
    start = raw.index(b"_", cut) + 1
    root = ET.fromstring(raw[:cut] + b"</VTKFile>")
    out = {}
    for a in root.iter("DataArray"):
        off = start + int(a.get("offset"))
        nbytes = int(np.frombuffer(raw[off:off + 8], dtype="<u8")[0])
        out[a.get("Name")] = np.frombuffer(raw[off + 8:off + 8 + nbytes], dtype="<f8" if a.get("type") == "Float64" else "<i4")
    return out


@pytest.mark.parametrize("model", [cases.PROTEAS, cases.HCC])
def test_cpp_driver_proteas_hcc(tmp_path, model):
    """The last two models through the stand-alone driver (proteas.C:16-90, coupled_hcc.C:16-142 with the solid solver
    switched off): '#'-commented nodal files, the AUX system's file, output_time_points as a list, the model's own
    key names (time_step_number / number_of_time_steps, output_Paraview / output_PARAVIEW), raw-appended VTU output."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "driver"), "-s"])
    conn, xyz = cases.mesh(cases.TET4, 6, distort=0.2)
    p, u0, ef, nf = cases.case(model, conn, xyz, "full")
    nv = cases.P.NVARS[model]
    d = str(tmp_path)
    _write_gmsh(os.path.join(d, "cube.msh"), conn, xyz, np.ones(conn.shape[0], dtype=int))
    U0 = np.asarray(u0).reshape(-1, nv)
    with open(os.path.join(d, "nodal.dat"), "w") as f:
        f.write("# initial state\n\n")
        for k, row in enumerate(U0):
            if k == 3 and model == cases.PROTEAS:
                f.write("# a comment between the rows\n")
            f.write(" ".join("%.17g" % v for v in row) + "\n")
    nsteps, dt = 5, cases.DT[model]
    extra = "input_GMSH = cube.msh\ninput_nodal = nodal.dat\n"
    if model == cases.PROTEAS:
        np.savetxt(os.path.join(d, "aux.dat"), np.asarray(nf).reshape(-1, 2), fmt="%.17g", header="HU RTD")
        extra += f"input_nodal_aux = aux.dat\noutput_Paraview = view\ntime_step_number = {nsteps}\ntime_step = {dt}\n"
    else:
        extra += f"output_PARAVIEW = view\nnumber_of_time_steps = {nsteps}\ntime_step = {dt}\n"
    extra += "output_time_points = '2 5'\n"
    _write_input(os.path.join(d, "input.dat"), model, p, extra)
    sol = os.path.join(d, "u.bin")
    name = "proteas" if model == cases.PROTEAS else "coupled_hcc"
    args = [os.path.join(ROOT, "driver", "rdc_driver"), "-m", name, os.path.join(d, "input.dat"), "ksp=2", "vtu=binary",
            "solution_out=" + sol]
    if model == cases.HCC:
        refused = subprocess.run(args, capture_output=True, text=True, timeout=300)
        # without solid=off the driver sets the SolidSystem up as well (coupled_hcc.C:39-74); this input names no material for
        # subdomain 1, which the reference would hit at solid_system.C:182 (Parameters::get throws): refused, said loudly
        assert refused.returncode != 0 and "Hyperelastic" in refused.stderr
        args.append("solid=off")
    out = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    u_drv = np.fromfile(sol)

    gpu = cases.gpu_system(model, cases.TET4, conn, xyz, p, u0, ef, nf)
    gpu.ksp = 2
    orc = cases.oracle_problem(model, cases.TET4, conn, xyz, p, u0, ef, nf)
    snap = {}
    for t in range(1, nsteps + 1):
        gpu.step(dt)
        orc.step(dt, pc=O.PC_ILU)
        if t in (2, 5):
            snap[t] = gpu.get_solution().reshape(-1, nv).copy()
    assert np.array_equal(u_drv, gpu.get_solution())
    assert np.linalg.norm(u_drv - orc.u) <= 1e-8 * np.linalg.norm(orc.u)
    gpu.close()
    # ParaView output: steps 0, 2 and 5; every variable of the model (and PROTEAS' AUX system) as its own array
    assert sorted(f for f in os.listdir(d) if f.endswith(".vtu")) == ["view-0.vtu", "view-2.vtu", "view-5.vtu"]
    names = ["hos", "tum", "nec", "vsc", "oed"] if model == cases.PROTEAS else ["l", "c", "n"]
    for t in (2, 5):
        arr = _read_appended_vtu(os.path.join(d, f"view-{t}.vtu"))
        for a, nm in enumerate(names):
            want = snap[t][:, a].copy()
            want[np.abs(want) <= 1e-300] = 0.0
            assert np.array_equal(arr[nm], want), nm
        assert np.array_equal(arr["position"].reshape(-1, 3), xyz) and np.array_equal(arr["connectivity"].reshape(-1, 4), conn)
        if model == cases.PROTEAS:
            assert np.array_equal(arr["HU"], np.asarray(nf).reshape(-1, 2)[:, 0]) and np.array_equal(arr["RTD"], np.asarray(nf).reshape(-1, 2)[:, 1])


def test_cpp_driver_matches_python_mirror_and_oracle(tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "driver"), "-s"])
    conn, xyz = cases.mesh(cases.TET4, 6, distort=0.2)
    u0, tracts = synth.adpm_fields(conn, xyz, smooth=True)
    E = conn.shape[0]
    gmsh_ids = np.array([12, 3, 7])[np.random.default_rng(4).integers(0, 3, E)]     # subdomain ids; std::set order 3,7,12
    region = np.searchsorted(np.array([3, 7, 12]), gmsh_ids).astype(np.int32)
    d = str(tmp_path)
    _write_gmsh(os.path.join(d, "cube.msh"), conn, xyz, gmsh_ids)
    np.savetxt(os.path.join(d, "nodal.dat"), u0, fmt="%.17g")
    np.savetxt(os.path.join(d, "elemental.dat"), tracts, fmt="%.17g")
    nsteps, dt = 4, 0.05
    kv = synth.adpm_param_dict("full")
    lo, hi = 0.02, 0.5
    with open(os.path.join(d, "input.dat"), "w") as f:
        f.write("# written by tests/test_gpu_driver.py\ninput_GMSH = 'cube.msh'\ninput_nodal = 'nodal.dat'\n"
                "input_elemental = 'elemental.dat'\noutput_CSV = out.csv\noutput_PARAVIEW = 'view'\n")
        f.write(f"time_step_number = {nsteps}\ntime_step = {dt}\noutput_step = 2\n")
        f.write(f"range/A_b/min = {lo}\nrange/A_b/max = {hi}\nrange/Tau/min = {lo}\nrange/Tau/max = {hi}\n")
        for k, v in kv.items():
            f.write(f"{k} = {v!r}\n")
        f.write("taxis/A_b = 999.0   # ignored key, like in run/HCP102513/input.dat\n")
    sol = os.path.join(d, "u.bin")
    out = subprocess.run([os.path.join(ROOT, "driver", "rdc_driver"), "-m", "adpm", os.path.join(d, "input.dat"), "ksp=2",
                          "solution_out=" + sol], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    u_drv = np.fromfile(sol)
    rows = [ln.split(",") for ln in open(os.path.join(d, "out.csv")).read().strip().splitlines()]
    assert rows[0][0] == '"TIME"' and rows[0][1] == '"CONCENTRATION__A_b__3"' and rows[0][-1] == '"VOLUME__Tau__12"'
    table = np.array([[float(x) for x in r] for r in rows[1:]])
    assert table.shape == (1 + nsteps // 2, 1 + 4 * 3)

    # the Python mirror drives the same library: identical numbers
    gpu = cases.gpu_system(cases.ADPM, cases.TET4, conn, xyz, synth.adpm_params("full"), u0, tracts, None)
    gpu.ksp = 2
    gpu.set_subdomains(region, 3)

    def line(time):
        cA, cT = gpu.region_last_mean(1), gpu.region_last_mean(2)
        vA = gpu.region_volumes([([0, 1, 0], 1.0, lo, hi)])
        vT = gpu.region_volumes([([0, 0, 1], 1.0, lo, hi)])
        return np.concatenate([[time], np.stack([cA, cT], 1).ravel(), np.stack([vA, vT], 1).ravel()])

    ref = [line(0.0)]
    orc = cases.oracle_problem(cases.ADPM, cases.TET4, conn, xyz, synth.adpm_params("full"), u0, tracts, None)
    for t in range(1, nsteps + 1):
        gpu.step(dt)
        orc.step(dt, pc=O.PC_ILU)
        if t % 2 == 0:
            ref.append(line(gpu.time))
    ref = np.array(ref)
    assert np.array_equal(table, ref), np.abs(table - ref).max()
    assert np.array_equal(u_drv, gpu.get_solution())
    # ParaView collection: steps 0, 2, 4; the point data of the last file is the final solution
    import xml.etree.ElementTree as ET
    sets = ET.parse(os.path.join(d, "view.pvd")).getroot().find("Collection").findall("DataSet")
    assert [x.get("timestep") for x in sets] == ["0", "2", "4"]
    piece = ET.parse(os.path.join(d, "view-4.vtu")).getroot().find("UnstructuredGrid").find("Piece")
    arr = {a.get("Name"): np.array(a.text.split(), dtype=float) for a in piece.iter("DataArray")}
    for j, name in enumerate(("PrP", "A_b", "Tau")):
        assert np.array_equal(arr[name], u_drv[j::3])
    assert np.array_equal(arr["region_ID"], gmsh_ids) and np.array_equal(arr["connectivity"], conn.ravel())
    # and the oracle: solution to 1e-8, CSV quantities from the oracle's serial loops on ITS solution to 1e-6
    assert np.linalg.norm(u_drv - orc.u) <= 1e-8 * np.linalg.norm(orc.u)
    cA = O.region_last_mean(cases.TET4, conn, xyz, orc.u, 1, region, 3)
    vT = O.region_volumes(cases.TET4, conn, xyz, orc.u, [([0, 0, 1], 1.0, lo, hi)], region, 3)
    assert np.allclose(table[-1, 1:7:2], cA, rtol=1e-6)
    assert np.allclose(table[-1, 8::2], vT, rtol=1e-6, atol=1e-12)
    gpu.close()


def _write_gmsh_with_faces(path, conn, xyz, ids, faces):
    """Gmsh 2.2 with tagged boundary faces (type 2 triangle / 3 quadrangle) in front of the volume elements, like a mesh
    written by Gmsh with physical surfaces: `faces` = [(tag, node ids)]."""
    vtype = 4 if conn.shape[1] == 4 else 5
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % xyz.shape[0])
        for k, (x, y, z) in enumerate(xyz):
            f.write("%d %.17g %.17g %.17g\n" % (k + 1, x, y, z))
        f.write("$EndNodes\n$Elements\n%d\n" % (conn.shape[0] + len(faces)))
        k = 1
        for tag, nodes in faces:
            f.write("%d %d 2 %d %d %s\n" % (k, 2 if len(nodes) == 3 else 3, tag, tag, " ".join(str(v + 1) for v in nodes)))
            k += 1
        for e, c in enumerate(conn):
            f.write("%d %d 2 %d %d %s\n" % (k, vtype, ids[e], ids[e], " ".join(str(v + 1) for v in c)))
            k += 1
        f.write("$EndElements\n")


def _tagged_faces(case):
    from oracle import solid as S
    out = []
    for e, s, b in zip(case.side_elem, case.side_no, case.side_bc):
        out.append((case.bc_ids[b], [int(case.conn[e, l]) for l in S.SIDE_NODES[case.elem_type][s]]))
    return out


_TIGHT = ("solver/nonlinear/max_nonlinear_iterations = 25\nsolver/nonlinear/relative_step_tolerance = 1e-11\n"
          "solver/nonlinear/relative_residual_tolerance = 1e-13\nsolver/nonlinear/absolute_residual_tolerance = 1e-9\n"
          "solver/linear/initial_linear_tolerance = 1e-10\n")


@pytest.mark.parametrize("et", [cases.TET4, cases.HEX8])
def test_cpp_driver_solid(tmp_path, et):
    """rdc_driver -m solid = the loop of solid.C:81-108 over the C ABI: mesh with tagged boundary faces, input.dat in the
    layout of run/Solid/uniaxial_compression, final positions and element stresses against the oracle's load steps."""
    import solid_cases as SC
    from oracle import solid as S
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "driver"), "-s"])
    c = SC.compression_case(et, n=3, penalty=1.0e6)
    d = str(tmp_path)
    _write_gmsh_with_faces(os.path.join(d, "cube.msh"), c.conn, c.xund, np.zeros(c.E, dtype=int), _tagged_faces(c))
    with open(os.path.join(d, "input.dat"), "w") as f:
        f.write("input_GMSH = cube.msh\nloading_step = 0.25\noutput_PARAVIEW = out\n" + _TIGHT)
        f.write("BCs = ' 0 5 '\nBC/0/displacement/0 = +0.000\nBC/0/displacement/1 = +0.000\nBC/0/displacement/2 = +0.000\n")
        f.write("BC/5/displacement/0 = NAN\nBC/5/displacement/1 = NAN\nBC/5/displacement/2 = -0.750\nBCs/displacement_penalty = 1.e+6\n")
        f.write("materials = ' 0 '\nmaterial/0/Neohookean/Young = 1.0e+4\n")     # misspelt like the shipped file: defaults apply
    sol = os.path.join(d, "x.bin")
    out = subprocess.run([os.path.join(ROOT, "driver", "rdc_driver"), "-m", "solid", os.path.join(d, "input.dat"), "ksp=0",
                          "vtu=binary", "solution_out=" + sol], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = np.fromfile(sol)
    x_drv, p_drv, vm_drv = raw[:3 * c.N], raw[3 * c.N:3 * c.N + c.E], raw[3 * c.N + c.E:]
    c.opts.update(max_nonlinear_iterations=25, relative_step_tolerance=1e-11, relative_residual_tolerance=1e-13,
                  absolute_residual_tolerance=1e-9, initial_linear_tolerance=1e-10)
    orc = S.OracleSolid(c)
    x = c.xund.copy().ravel()
    for l in range(1, 5):
        x, info = orc.newton(x, 0.25 * l)
        assert info["converged"]
    assert np.linalg.norm(x_drv - x) <= 1e-8 * np.linalg.norm(x)
    po, vo, _ = orc.post(x, 1.0)
    scale = np.abs(po).max() + vo.max()
    assert np.abs(p_drv - po).max() <= 1e-6 * scale and np.abs(vm_drv - vo).max() <= 1e-6 * scale
    arrays = _read_appended_vtu(os.path.join(d, "out-4.vtu"))      # solid.C:27-44 variable names, the mesh at its moved position
    assert np.allclose(arrays["u_z"], x_drv.reshape(-1, 3)[:, 2] - c.xund[:, 2], atol=1e-12)
    assert np.allclose(arrays["position"].reshape(-1, 3), x_drv.reshape(-1, 3), atol=1e-12)


def test_cpp_driver_coupled_hcc_with_solid(tmp_path):
    """coupled_hcc.C:93-140 with the solid solves switched on: the mesh moves at the loading time points and the
    reaction-diffusion system is assembled on the moved mesh afterwards (rdc_update_coords).  Checked against the same
    sequence driven from Python: oracle Newton for the mesh positions, oracle RD steps on those positions."""
    import solid_cases as SC
    from oracle import solid as S
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "driver"), "-s"])
    sc = SC.growth_case(cases.TET4, n=3)
    sc.mats[1, 3:] = 4.0      # strong growth: the moved mesh changes the RD solution by 9e-8, well above the 1e-9 parity bar
    conn, xyz = sc.conn, sc.xund
    p, u0, ef, nf = cases.case(cases.HCC, conn, xyz, "full")
    d = str(tmp_path)
    _write_gmsh_with_faces(os.path.join(d, "m.msh"), conn, xyz, 3000 + sc.mat_of, _tagged_faces(sc))
    np.savetxt(os.path.join(d, "nodal.dat"), np.asarray(u0).reshape(-1, 3), fmt="%.17g")
    nsteps, dt, nload = 4, cases.DT[cases.HCC], 2
    extra = f"input_GMSH = m.msh\ninput_nodal = nodal.dat\ntime_step = {dt}\nnumber_of_time_steps = {nsteps}\nnumber_of_loading_steps = {nload}\n" + _TIGHT
    extra += "BCs = ' 2000 2002 '\nBC/2000/displacement/0 = 0\nBC/2000/displacement/1 = 0\nBC/2000/displacement/2 = 0\n"
    extra += "BC/2002/displacement/0 = NAN\nBC/2002/displacement/1 = NAN\nBC/2002/displacement/2 = 0\nBCs/displacement_penalty = 1.e+8\n"
    extra += "materials = ' 3000 3001 '\n"
    for m, row in zip((3000, 3001), sc.mats):
        for key, v in zip(("Young", "Poisson", "FibreStiffness", "VolumetricStretchRatio/rate_0", "VolumetricStretchRatio/rate_1",
                           "VolumetricStretchRatio/rate_2"), row):
            extra += f"material/{m}/Hyperelastic/{key} = {float(v)!r}\n"
    _write_input(os.path.join(d, "input.dat"), cases.HCC, p, extra)
    sol = os.path.join(d, "u.bin")
    out = subprocess.run([os.path.join(ROOT, "driver", "rdc_driver"), "-m", "coupled_hcc", os.path.join(d, "input.dat"), "ksp=0",
                          "solution_out=" + sol], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("Newton iterations") == nload
    u_drv = np.fromfile(sol)
    # the same sequence with the oracles
    sc.opts.update(max_nonlinear_iterations=25, relative_step_tolerance=1e-11, relative_residual_tolerance=1e-13,
                   absolute_residual_tolerance=1e-9, initial_linear_tolerance=1e-10)
    so = S.OracleSolid(sc)
    x = xyz.copy().ravel()
    orc = cases.oracle_problem(cases.HCC, cases.TET4, conn, xyz, p, u0, ef, nf)
    pseudo, loading_step = 0.0, dt * nsteps / nload
    for t in range(1, nsteps + 1):
        if t % (nsteps // nload) == 0:
            pseudo += loading_step
        orc.step(dt, pc=O.PC_ILU)
        if t % (nsteps // nload) == 0:
            x, info = so.newton(x, pseudo)
            assert info["converged"]
            orc.xyz = np.ascontiguousarray(x.reshape(-1, 3))
    assert np.linalg.norm(u_drv - orc.u) <= 1e-9 * np.linalg.norm(orc.u)
    # and the mesh really moved: the run on the fixed mesh differs
    out2 = subprocess.run([os.path.join(ROOT, "driver", "rdc_driver"), "-m", "coupled_hcc", os.path.join(d, "input.dat"), "ksp=0", "solid=off",
                           "solution_out=" + sol], capture_output=True, text=True, timeout=300)
    assert out2.returncode == 0, out2.stdout + out2.stderr
    assert np.linalg.norm(np.fromfile(sol) - u_drv) > 3e-8 * np.linalg.norm(u_drv)


@pytest.mark.parametrize("name", ["solid_uniaxial", "solid_hydrogel"])
def test_cpp_driver_on_the_shipped_solid_cases(tmp_path, name):
    """The two solid-mechanics run directories the reference ships (run/Solid/uniaxial_compression: 512 HEX8, the cube pressed
    to half its height in 10 load steps; run/Solid/hydrogel_tension: 5504 TET4, symmetry planes + a pulled face), rebuilt
    verbatim from tests/golden/solid_*.npz (mesh with its tagged faces, the input.dat text) and run by `rdc_driver -m solid`.
    (a) as shipped: relative step tolerance 1e-3 and linear tolerance 1e-3 -- two correct drivers agree to those tolerances;
    (b) the same files with the solver tolerances tightened: positions equal to the oracle's to 1e-8."""
    import re
    import solid_cases as SC
    from oracle import solid as S
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "driver"), "-s"])
    c, kv, fx = SC.shipped_case(name)
    d = str(tmp_path)
    faces = [(int(t), [int(v) for v in n if v >= 0]) for t, n in zip(fx["face_tag"], fx["face_nodes"])]
    _write_gmsh_with_faces(os.path.join(d, str(fx["mesh_name"])), c.conn, c.xund, fx["sub"], faces)
    text = str(fx["input_dat"])
    nload = int(1.0 / float(kv["loading_step"]))

    def run(input_text, steps):
        with open(os.path.join(d, "input.dat"), "w") as f:
            f.write(input_text)
        sol = os.path.join(d, "x.bin")
        out = subprocess.run([os.path.join(ROOT, "driver", "rdc_driver"), "-m", "solid", os.path.join(d, "input.dat"), "ksp=0",
                              "vtu=binary", "solution_out=" + sol], capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stdout + out.stderr
        assert out.stdout.count("Newton iterations") == steps and "NOT converged" not in out.stdout
        return np.fromfile(sol)[:3 * c.N]

    # (a) verbatim
    x_drv = run(text, nload)
    orc = S.OracleSolid(c)
    x = c.xund.copy().ravel()
    for l in range(1, nload + 1):
        x, info = orc.newton(x, float(kv["loading_step"]) * l)
        assert info["converged"]
    disp = np.abs(x - c.xund.ravel()).max()
    assert np.abs(x_drv - x).max() <= 5e-3 * disp, (np.abs(x_drv - x).max(), disp)
    assert os.path.exists(os.path.join(d, "out-%d.vtu" % nload))        # output_PARAVIEW = out, last load step
    # (b) tightened, three load steps
    tight = re.sub(r"(?m)^loading_step.*$", "loading_step = 0.34", text) + _TIGHT + "solver/nonlinear/absolute_residual_tolerance = 1e-14\n"
    x_drv = run(tight, 2)
    ct, _, _ = SC.shipped_case(name, tight=True)
    ct.opts.update(absolute_residual_tolerance=1e-14)
    orc = S.OracleSolid(ct)
    x = ct.xund.copy().ravel()
    for l in (1, 2):
        x, info = orc.newton(x, 0.34 * l)
        assert info["converged"]
    assert np.linalg.norm(x_drv - x) <= 1e-8 * np.linalg.norm(x)
