"""The oracle against the committed golden vectors (tests/golden/make_golden.py), plus the C-ABI checks
that need no GPU: the library loads, exports every symbol include/rdc.h declares, and refuses to run
without a CUDA device (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest

import cases
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("model", range(5))
def test_oracle_reproduces_golden(model):
    g = np.load(os.path.join(GOLD, f"{cases.NAMES[model]}_tet_n4.npz"))
    length = 50.0 if model == cases.RIPF else 1.0
    conn, xyz = cases.mesh(cases.TET4, 4, distort=0.2, length=length)
    p, u0, ef, nf = cases.case(model, conn, xyz, "full")
    for nthreads in (1, 3):  # the threaded assembly is bit-identical to the serial one
        pr = cases.oracle_problem(model, cases.TET4, conn, xyz, p, u0, ef, nf, nthreads=nthreads)
        dt = cases.DT[model]
        pr.u_old = pr.u.copy()
        val, rhs = pr.assemble(dt, dt)
        assert np.array_equal(pr.rowptr, g["rowptr"]) and np.array_equal(pr.col, g["col"])
        np.testing.assert_allclose(val, g["val"], rtol=1e-13, atol=1e-16 * np.abs(g["val"]).max())
        np.testing.assert_allclose(rhs, g["rhs"], rtol=1e-13, atol=1e-16 * np.abs(g["rhs"]).max())
        pr.step(dt, pc=O.PC_ILU)
        assert np.linalg.norm(pr.u - g["u1"]) <= 1e-10 * np.linalg.norm(g["u1"])


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rdc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rdc_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_everything_the_header_declares():
    from rdcfes_b200 import build, lib
    so = build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (rdc_[a-z_0-9]+)", out))
    declared = _declared_symbols()
    assert len(declared) >= 25
    missing = [s for s in declared if s not in exported]
    assert not missing, missing
    L = lib.load()  # argtypes for every binding resolve
    assert L.rdc_model_nvars(cases.PIHNA) == 5 and L.rdc_model_nparams(cases.ADPM) == 46
    assert sorted(lib.EXPORTS) == declared


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rdcfes_b200 import lib
    from rdcfes_b200.system import TransientRdcSystem
    conn, xyz = cases.mesh(cases.TET4, 2)
    with pytest.raises(lib.RdcError) as ei:
        TransientRdcSystem(cases.ADPM, cases.TET4, conn, xyz)
    assert ei.value.code == -2 and "no CUDA device" in str(ei.value)


def test_parameter_tables_match_header_enums():
    text = open(os.path.join(ROOT, "include", "rdc.h")).read()
    for name, model in (("ADPM", cases.ADPM), ("PIHNA", cases.PIHNA), ("RIPF", cases.RIPF), ("PROTEAS", cases.PROTEAS),
                        ("HCC", cases.HCC)):
        n = int(re.search(rf"{name}_NPARAMS\s*=\s*(\d+)", text).group(1))
        assert len(cases.P.TABLES[model]) == n


def test_input_dat_parsing_of_the_shipped_cases():
    base = "/root/reference/run"
    if not os.path.isdir(base):
        pytest.skip("reference tree not present (GPU box)")
    p, kv = cases.P.params_from_input(cases.ADPM, f"{base}/HCP102513/input.dat")
    ref = cases.synth.adpm_params("ref")
    np.testing.assert_array_equal(p, ref)  # the taxis/* keys of the shipped file are ignored (Appendix C-1)
    assert float(kv["time_step"]) == 0.05
    p, _ = cases.P.params_from_input(cases.PIHNA, f"{base}/PIHNA/input.dat")
    np.testing.assert_array_equal(p, cases.synth.pihna_params("ref"))
    p, _ = cases.P.params_from_input(cases.RIPF, f"{base}/RIPF133/input.dat")
    np.testing.assert_array_equal(p, cases.synth.ripf_params("ref"))


# ---- the two meshes the reference ships (fixtures made by tests/golden/make_mesh_fixtures.py) ----
def _pins(pr, dt, O):
    pr.u_old = pr.u.copy()
    val, rhs = pr.assemble(dt, dt)
    pr.time = 0.0
    pr.step(dt, pc=O.PC_ILU)
    w = np.cos(np.arange(val.size) * 0.37)
    return np.array([val.sum(), (val * w).sum(), np.abs(val).max(), rhs.sum(), np.linalg.norm(pr.u), pr.u.sum()])


@pytest.mark.parametrize("name,etype", [("hydrogel_tet4", 4), ("cube_hex8", 8)])
def test_oracle_on_the_shipped_meshes(name, etype):
    """Unstructured TET4 (5 504 elements, valence up to 40) and the HEX8 cube of run/Solid: the oracle's operator
    checksums and one-step solutions are pinned."""
    import cases
    from oracle import oracle as O
    d = np.load(os.path.join(HERE, "golden", name + ".npz"))
    conn, xyz = d["conn"], d["xyz"]
    for model in (cases.ADPM, cases.PIHNA, cases.HCC):
        p, u0, ef, nf = cases.case(model, conn, xyz, "full")
        pr = cases.oracle_problem(model, etype, conn, xyz, p, u0, ef, nf)
        got = _pins(pr, cases.DT[model], O)
        ref = d["pin_" + cases.NAMES[model]]
        assert np.allclose(got, ref, rtol=1e-10, atol=0), (cases.NAMES[model], got, ref)


def test_vtu_writer_of_the_cpp_driver(tmp_path):
    """driver/vtu_writer.h needs no GPU: a tiny program writes a two-tet mesh; the XML is read back and compared with
    the layout of the reference's Paraview_IO (paraview.h:60-150, 158-198)."""
    import xml.etree.ElementTree as ET
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include "vtu_writer.h"
int main(int argc, char** argv) {
  std::vector<double> xyz = {0,0,0, 1,0,0, 0,1,0, 0,0,1, 1,1,1};
  std::vector<int32_t> conn = {0,1,2,3, 1,2,3,4};
  std::vector<int> sub = {7, 3};
  VtuMesh m = {4, &xyz, &conn, &sub, nullptr};
  std::vector<double> u = {1,10, 2,20, 3,30, 4,1e-320, 5,50};
  PvdCollection pvd(argv[1]);
  PvdCollection bin(std::string(argv[1]) + "_bin", true);
  return pvd.add(m, {"A", "B"}, u, 0) && pvd.add(m, {"A", "B"}, u, 20) && bin.add(m, {"A", "B"}, u, 20) ? 0 : 1;
}
''')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(os.path.dirname(HERE), "driver"), "-o", str(exe), str(src)])
    base = str(tmp_path / "out")
    subprocess.check_call([str(exe), base])
    pvd = ET.parse(base + ".pvd").getroot()
    sets = pvd.find("Collection").findall("DataSet")
    assert [s.get("timestep") for s in sets] == ["0", "20"] and sets[1].get("file") == "out-20.vtu"
    piece = ET.parse(base + "-20.vtu").getroot().find("UnstructuredGrid").find("Piece")
    assert piece.get("NumberOfPoints") == "5" and piece.get("NumberOfCells") == "2"
    arr = {a.get("Name"): np.array(a.text.split(), dtype=float) for a in piece.iter("DataArray")}
    assert np.array_equal(arr["position"], [0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 1, 1, 1])
    assert np.array_equal(arr["node_ID"], [1, 2, 3, 4, 5]) and np.array_equal(arr["element_ID"], [1, 2])
    assert np.array_equal(arr["A"], [1, 2, 3, 4, 5]) and np.array_equal(arr["B"], [10, 20, 30, 0, 50])   # denormal -> 0
    assert np.array_equal(arr["region_ID"], [7, 3]) and np.array_equal(arr["processor_ID"], [0, 0])
    assert np.array_equal(arr["connectivity"], [0, 1, 2, 3, 1, 2, 3, 4])
    assert np.array_equal(arr["offsets"], [4, 8]) and np.array_equal(arr["types"], [10, 10])
    # the raw appended form: same arrays, same order, UInt64 byte counts in front of every block
    raw = open(base + "_bin-20.vtu", "rb").read()
    cut = raw.index(b'<AppendedData encoding="raw">')
    start = raw.index(b"_", cut) + 1
    root = ET.fromstring(raw[:cut] + b"</VTKFile>")
    assert root.get("header_type") == "UInt64" and root.get("byte_order") == "LittleEndian"
    names = []
    for a in root.iter("DataArray"):
        assert a.get("format") == "appended"
        off = start + int(a.get("offset"))
        nbytes = int(np.frombuffer(raw[off:off + 8], dtype="<u8")[0])
        dt = "<f8" if a.get("type") == "Float64" else "<i4"
        got = np.frombuffer(raw[off + 8:off + 8 + nbytes], dtype=dt)
        assert np.array_equal(got, arr[a.get("Name")]), a.get("Name")
        names.append(a.get("Name"))
    assert names == [a.get("Name") for a in piece.iter("DataArray")]
    assert raw[start + int(list(root.iter("DataArray"))[-1].get("offset")) + 8 + 8:].strip().startswith(b"</AppendedData>")
