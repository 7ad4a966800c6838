#!/usr/bin/env python3
"""Turns the two meshes the reference ships (run/Solid/hydrogel_tension/hydrogel_model.msh: 5 504 TET4 / 1 569 nodes,
unstructured; run/Solid/uniaxial_compression/cube.msh: 512 HEX8 / 729 nodes) into small fixtures for the parity
tests, and pins the ORACLE on them: per model the assembled operator's checksums and the solution after one step.

    python tests/golden/make_mesh_fixtures.py            (needs /root/reference; run once in the build container)

Gmsh 2.2 ASCII (the dialect process_mesh.C:22-83 writes): only the volume elements (type 4 = TET4, 5 = HEX8) are
kept, in file order, which is the element id order libMesh's GmshIO gives them; node ids are made 0-based.  The
hydrogel geometry is in metres (2.5 mm plate); it is scaled to unit size so that the models' per-length parameters
stay in the range the other tests use.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference/run/Solid"


def read_gmsh22(path, vol_type):
    with open(path) as fh:
        lines = fh.read().split("\n")
    i = lines.index("$Nodes")
    n = int(lines[i + 1])
    ids = np.empty(n, dtype=np.int64)
    xyz = np.empty((n, 3))
    for k in range(n):
        t = lines[i + 2 + k].split()
        ids[k] = int(t[0]); xyz[k] = [float(t[1]), float(t[2]), float(t[3])]
    remap = {int(g): k for k, g in enumerate(ids)}
    j = lines.index("$Elements")
    m = int(lines[j + 1])
    nen = {4: 4, 5: 8}[vol_type]
    conn = []
    for k in range(m):
        t = lines[j + 2 + k].split()
        if int(t[1]) != vol_type:
            continue
        ntags = int(t[2])
        conn.append([remap[int(v)] for v in t[3 + ntags:3 + ntags + nen]])
    return np.asarray(conn, dtype=np.int32), xyz


def main():
    import cases
    from oracle import oracle as O
    out = {}
    for name, rel, vt, et in (("hydrogel_tet4", "hydrogel_tension/hydrogel_model.msh", 4, cases.TET4),
                              ("cube_hex8", "uniaxial_compression/cube.msh", 5, cases.HEX8)):
        conn, xyz = read_gmsh22(os.path.join(REF, rel), vt)
        xyz = (xyz - xyz.min(0)) / (xyz.max(0) - xyz.min(0)).max()      # unit size
        used = np.unique(conn)
        assert used.size == xyz.shape[0], "unused nodes"
        pins = {}
        for model in (cases.ADPM, cases.PIHNA, cases.HCC):
            p, u0, ef, nf = cases.case(model, conn, xyz, "full")
            pr = cases.oracle_problem(model, et, conn, xyz, p, u0, ef, nf)
            dt = cases.DT[model]
            pr.u_old = pr.u.copy()
            val, rhs = pr.assemble(dt, dt)
            pr.time = 0.0
            pr.step(dt, pc=O.PC_ILU)
            w = np.cos(np.arange(val.size) * 0.37)       # position-sensitive checksums
            pins[cases.NAMES[model]] = np.array([val.sum(), (val * w).sum(), np.abs(val).max(), rhs.sum(),
                                                 np.linalg.norm(pr.u), pr.u.sum()])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), conn=conn, xyz=xyz,
                            **{"pin_" + k: v for k, v in pins.items()})
        out[name] = (conn.shape, xyz.shape)
    print(out)


if __name__ == "__main__":
    main()
