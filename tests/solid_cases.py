"""Solid-mechanics test problems shared by the CPU pin tests and the GPU parity tests."""
import numpy as np

from rdcfes_b200 import synth
from oracle import solid as S

TET4, HEX8 = 4, 8


def mesh(et, n, length=1.0, distort=0.0):
    conn, xyz = synth.kuhn_cube(n, length) if et == TET4 else synth.hex_cube(n, length)
    if distort:
        xyz = synth.distort(xyz, distort)
    return conn, xyz


def general_case(et, n=2, use_symmetry=False, penalty=1.0e5):
    """Every term switched on: two materials, fibres, anisotropic growth, a clamped and a partly constrained face."""
    conn, xyz = mesh(et, n, distort=0.05)
    E = conn.shape[0]
    mats = [[1.0e3, 0.3, 50.0, 0.3, 0.1, -0.2], [2.0e3, 0.4, 0.0, 0.0, 0.0, 0.0]]
    mat_of = (np.arange(E) % 2).astype(np.int32)
    fib = np.random.default_rng(3).normal(size=(E, 3))
    c = S.SolidCase(et, conn, xyz, mats=mats, mat_of=mat_of, fibres=fib, penalty=penalty, use_symmetry=use_symmetry)
    zmin, zmax = xyz[:, 2].min(), xyz[:, 2].max()
    c.add_bc(0, [0.0, 0.0, 0.0], c.faces_where(lambda p: abs(p[2] - zmin) < 1e-9))
    c.add_bc(5, [np.nan, np.nan, -0.2], c.faces_where(lambda p: abs(p[2] - zmax) < 1e-9))
    return c


def perturbed(c, amp=0.03, seed=1):
    return c.xund + amp * np.random.default_rng(seed).normal(size=c.xund.shape)


def compression_case(et, n=4, length=1.5, penalty=1.0e8, disp=-0.75):
    """run/Solid/uniaxial_compression/input.dat: bottom face (id 0) clamped, top face (id 5) pushed down; the material
    keys of that file are spelled 'Neohookean' and never read, so the defaults E = 1e3, nu = 0.3 apply (solid.C:279-283)."""
    conn, xyz = mesh(et, n, length)
    c = S.SolidCase(et, conn, xyz, penalty=penalty)
    zmax = xyz[:, 2].max()
    c.add_bc(0, [0.0, 0.0, 0.0], c.faces_where(lambda p: abs(p[2]) < 1e-9))
    c.add_bc(5, [np.nan, np.nan, disp], c.faces_where(lambda p: abs(p[2] - zmax) < 1e-9))
    c.opts.update(max_nonlinear_iterations=10)
    return c


def growth_case(et, n=3):
    """The solid half of run/Coupled/HCC/input.dat in miniature: a growing inclusion (rate 0.3 in every direction) inside a
    passive matrix, bottom clamped, top free to slide."""
    conn, xyz = mesh(et, n)
    cen = xyz[conn].mean(axis=1)
    mat_of = (np.linalg.norm(cen - 0.5, axis=1) < 0.3).astype(np.int32)
    mats = [[2.0e3, 0.4, 0.0, 0.0, 0.0, 0.0], [2.0e3, 0.4, 0.0, 0.3, 0.3, 0.3]]
    c = S.SolidCase(et, conn, xyz, mats=mats, mat_of=mat_of, penalty=1.0e8)
    c.add_bc(2000, [0.0, 0.0, 0.0], c.faces_where(lambda p: abs(p[2]) < 1e-9))
    c.add_bc(2002, [np.nan, np.nan, 0.0], c.faces_where(lambda p: abs(p[2] - 1.0) < 1e-9))
    c.opts.update(max_nonlinear_iterations=30, relative_residual_tolerance=1e-6)
    return c


# ---- the two solid cases the reference ships (run/Solid/*), from tests/golden/solid_*.npz --------------------------------
def parse_input_dat(text):
    """GetPot-style `key = value` lines ('#' comments, quotes stripped)."""
    kv = {}
    for ln in text.splitlines():
        ln = ln.split("#")[0].strip()
        if "=" in ln:
            k, v = ln.split("=", 1)
            kv[k.strip()] = v.strip().strip("'").strip()
    return kv


def shipped_case(name, tight=False):
    """SolidCase of tests/golden/<name>.npz exactly as solid.C:input() reads the shipped input.dat: only the keys it asks for
    count -- the files spell the material keys 'Neohookean' and the symmetry key 'solver/use_symmetry', which are never read,
    so the defaults E = 1e3, nu = 0.3 (solid.C:279-283) and assembly_use_symmetry = false apply.  -> (case, kv, fixture)"""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    kv = parse_input_dat(str(d["input_dat"]))
    conn, xyz = d["conn"], d["xyz"]
    et = conn.shape[1]
    mat_ids = sorted(int(v) for v in kv.get("materials", "0").split())
    mats = []
    for m in mat_ids:
        k = f"material/{m}/Hyperelastic/"
        mats.append([float(kv.get(k + "Young", 1.0e3)), float(kv.get(k + "Poisson", 0.3)), float(kv.get(k + "FibreStiffness", 0.0))] +
                    [float(kv.get(k + f"VolumetricStretchRatio/rate_{r}", 0.0)) for r in range(3)])
    mat_of = np.array([mat_ids.index(int(s)) for s in d["sub"]], dtype=np.int32)
    c = S.SolidCase(et, conn, xyz, mats=mats, mat_of=mat_of, penalty=float(kv.get("BCs/displacement_penalty", 1.0e5)))
    key2side = {}
    for s, loc in enumerate(S.SIDE_NODES[et]):
        for e, key in enumerate(map(tuple, np.sort(conn[:, list(loc)], axis=1))):
            key2side.setdefault(key, (e, s))
    ns = 3 if et == TET4 else 4
    for bc in sorted(int(v) for v in kv.get("BCs", "0").split()):
        disp = [float(kv.get(f"BC/{bc}/displacement/{k}", 0.0)) for k in range(3)]
        faces = [key2side[tuple(sorted(int(v) for v in n[:ns]))] for t, n in zip(d["face_tag"], d["face_nodes"]) if t == bc]
        c.add_bc(bc, disp, faces)
    g = lambda k, dflt: float(kv.get(k, dflt))
    c.opts.update(max_nonlinear_iterations=int(g("solver/nonlinear/max_nonlinear_iterations", 100)),
                  relative_step_tolerance=g("solver/nonlinear/relative_step_tolerance", 1e-3),
                  relative_residual_tolerance=g("solver/nonlinear/relative_residual_tolerance", 1e-8),
                  absolute_residual_tolerance=g("solver/nonlinear/absolute_residual_tolerance", 1e-8),
                  require_reduction=kv.get("solver/nonlinear/require_reduction", "false") == "true",
                  max_linear_iterations=int(g("solver/linear/max_linear_iterations", 50000)),
                  initial_linear_tolerance=g("solver/linear/initial_linear_tolerance", 1e-3))
    if tight:
        c.opts.update(max_nonlinear_iterations=25, relative_step_tolerance=1e-11, relative_residual_tolerance=1e-13,
                      absolute_residual_tolerance=1e-11 * c.penalty * 1e-6, initial_linear_tolerance=1e-10)
    return c, kv, d
