"""Makes tests/golden/solid_{uniaxial,hydrogel}.npz from the two solid-mechanics cases the reference ships
(run/Solid/uniaxial_compression: cube.msh, 512 HEX8 + 6 x 64 tagged quadrangles; run/Solid/hydrogel_tension:
hydrogel_model.msh, 5504 TET4 + 2352 tagged triangles) -- the mesh with its tagged boundary faces and the input.dat TEXT,
so that the GPU box (no /root/reference) can rebuild both files verbatim.  Input data only: no expected values inside
(those come from the oracle at test time, and from the compiled reference in tests/test_solid_pin.py).

    python tests/golden/make_solid_fixtures.py
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RUN = "/root/reference/run/Solid"
CASES = {"solid_uniaxial": ("uniaxial_compression", "cube.msh"), "solid_hydrogel": ("hydrogel_tension", "hydrogel_model.msh")}


def read_msh(path):
    L = open(path).read().split("\n")
    i = L.index("$Nodes")
    n = int(L[i + 1])
    ids, xyz = {}, np.zeros((n, 3))
    for k in range(n):
        t = L[i + 2 + k].split()
        ids[int(t[0])] = k
        xyz[k] = [float(v) for v in t[1:4]]
    j = L.index("$Elements")
    conn, sub, ftag, fnodes = [], [], [], []
    for k in range(int(L[j + 1])):
        t = [int(v) for v in L[j + 2 + k].split()]
        ty, nt = t[1], t[2]
        nodes = [ids[v] for v in t[3 + nt:]]
        if ty in (4, 5):
            conn.append(nodes); sub.append(t[3])
        elif ty in (2, 3):
            ftag.append(t[3]); fnodes.append(nodes + [-1] * (4 - len(nodes)))
    return np.array(conn, dtype=np.int32), xyz, np.array(sub, dtype=np.int32), np.array(ftag, dtype=np.int32), np.array(fnodes, dtype=np.int32)


if __name__ == "__main__":
    for name, (d, msh) in CASES.items():
        conn, xyz, sub, ftag, fnodes = read_msh(os.path.join(RUN, d, msh))
        text = open(os.path.join(RUN, d, "input.dat")).read()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), conn=conn, xyz=xyz, sub=sub, face_tag=ftag, face_nodes=fnodes,
                            input_dat=np.array(text), mesh_name=np.array(msh))
        print(name, conn.shape, xyz.shape, ftag.shape)
