"""Synthetic meshes, fields and parameter sets of SURVEY.md section 8(d) (cases S1-S5) plus a Gmsh 2.x reader.

Host-side utilities (numpy); nothing here is on the device path.  The four production meshes of the
reference are missing (.MISSING_LARGE_BLOBS), so the unit-cube Kuhn-tet mesh stands in for them; its
value distributions follow the shipped field files (run/HCP102513/*.dat etc.).
"""
from __future__ import annotations

import itertools
from typing import Dict, Tuple

import numpy as np

from . import params as P

TET4, HEX8 = 4, 8


# --------------------------------------------------------------------------------------------- meshes
def kuhn_cube(n: int, length: float = 1.0, permute_seed: int | None = None) -> Tuple[np.ndarray, np.ndarray]:
    """[0,L]^3 split into n^3 cells x 6 Kuhn tets (conforming), all positively oriented.
    Nodes lexicographic (x fastest); elements lexicographic by cell, then tet.  n=119 -> 10 110 954 tets."""
    m = n + 1
    ax = np.linspace(0.0, length, m)
    zz, yy, xx = np.meshgrid(ax, ax, ax, indexing="ij")
    xyz = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    base = ((k * m + j) * m + i).ravel().astype(np.int64)
    step = np.array([1, m, m * m], dtype=np.int64)
    tets = []
    for perm in itertools.permutations(range(3)):
        v0 = base
        v1 = v0 + step[perm[0]]
        v2 = v1 + step[perm[1]]
        v3 = v2 + step[perm[2]]
        # det(e_a, e_a+e_b, e_a+e_b+e_c) = sign(perm): swap two vertices of the odd ones
        inv = sum(1 for a in range(3) for b in range(a + 1, 3) if perm[a] > perm[b])
        tets.append(np.stack([v0, v1, v2, v3] if inv % 2 == 0 else [v0, v2, v1, v3], axis=1))
    conn = np.stack(tets, axis=1).reshape(-1, 4).astype(np.int32)
    if permute_seed is not None:
        conn, xyz = permute_nodes(conn, xyz, permute_seed)
    return conn, xyz


def hex_cube(n: int, length: float = 1.0) -> Tuple[np.ndarray, np.ndarray]:
    """Structured HEX8 cube, libMesh/Gmsh node order."""
    m = n + 1
    ax = np.linspace(0.0, length, m)
    zz, yy, xx = np.meshgrid(ax, ax, ax, indexing="ij")
    xyz = np.stack([xx.ravel(), yy.ravel(), zz.ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    b = ((k * m + j) * m + i).ravel().astype(np.int64)
    conn = np.stack([b, b + 1, b + 1 + m, b + m, b + m * m, b + 1 + m * m, b + 1 + m + m * m, b + m + m * m], axis=1)
    return conn.astype(np.int32), xyz


def permute_nodes(conn, xyz, seed):
    """Seeded random node renumbering (exposes gather locality; SURVEY.md 8d uses seed 12345)."""
    rng = np.random.default_rng(seed)
    N = xyz.shape[0]
    new_of_old = rng.permutation(N)
    xyz2 = np.empty_like(xyz)
    xyz2[new_of_old] = xyz
    return new_of_old[conn].astype(np.int32), xyz2


def distort(xyz: np.ndarray, amp: float, seed: int = 7) -> np.ndarray:
    """Move interior nodes by a seeded random displacement (keeps elements valid for small amp)."""
    rng = np.random.default_rng(seed)
    lo, hi = xyz.min(0), xyz.max(0)
    interior = np.all((xyz > lo + 1e-12) & (xyz < hi - 1e-12), axis=1)
    out = xyz.copy()
    out[interior] += amp * (rng.random((int(interior.sum()), 3)) - 0.5)
    return out


def read_gmsh(path: str):
    """Gmsh 2.x ASCII reader for the dialect of process_mesh.C:22-83 and the two shipped meshes.
    Keeps TET4 (type 4) or HEX8 (type 5); lower-dimensional elements are dropped like libMesh's GmshIO
    does (Appendix B-9).  node id = file id - 1.  Returns (elem_type, conn, xyz, subdomain)."""
    with open(path) as fh:
        tok = fh.read().split()
    i = tok.index("$Nodes") + 1
    nn = int(tok[i]); i += 1
    ids = np.empty(nn, dtype=np.int64)
    xyz = np.empty((nn, 3))
    for r in range(nn):
        ids[r] = int(tok[i]); xyz[r] = [float(tok[i + 1]), float(tok[i + 2]), float(tok[i + 3])]
        i += 4
    remap = {int(g): r for r, g in enumerate(ids)}
    i = tok.index("$Elements") + 1
    ne = int(tok[i]); i += 1
    tets, hexs, sub_t, sub_h = [], [], [], []
    nnodes = {1: 2, 2: 3, 3: 4, 4: 4, 5: 8, 15: 1}
    for _ in range(ne):
        et, ntags = int(tok[i + 1]), int(tok[i + 2])
        tags = tok[i + 3:i + 3 + ntags]
        k = nnodes[et]
        nodes = [remap[int(t)] for t in tok[i + 3 + ntags:i + 3 + ntags + k]]
        if et == 4:
            tets.append(nodes); sub_t.append(int(tags[0]) if ntags else 0)
        elif et == 5:
            hexs.append(nodes); sub_h.append(int(tags[0]) if ntags else 0)
        i += 3 + ntags + k
    if tets and not hexs:
        return TET4, np.asarray(tets, dtype=np.int32), xyz, np.asarray(sub_t, dtype=np.int32)
    if hexs and not tets:
        return HEX8, np.asarray(hexs, dtype=np.int32), xyz, np.asarray(sub_h, dtype=np.int32)
    raise ValueError("mixed or empty 3D element set")


# ------------------------------------------------------------------------------------ parameter sets
def adpm_param_dict(which: str = "full") -> Dict[str, float]:
    """GetPot key -> value of the two ADPM parameter sets (what an input.dat would hold)."""
    if which == "ref":
        return {"decay/PrP": 1.0e-4, "decay/PrP/pulse/0": 0.01, "decay/PrP/pulse/1": 10.0,
                "decay/Tau": 10.0, "decay/Tau/pulse/0": 0.0005}
    v: Dict[str, float] = {"decay/PrP": 0.1, "decay/PrP/pulse/0": 0.01, "decay/PrP/pulse/1": 10.0,
                           "decay/PrP/time_exponent": 0.5}
    for s in ("A_b", "Tau"):
        v.update({f"diffuse/{s}": 1e-3, f"taxis_1/{s}": 1e-3, f"taxis_2/{s}": 5e-4, f"taxis/{s}/angle": 45.0,
                  f"produce/{s}": 0.1, f"produce/{s}/sigmoid/0": 0.5, f"produce/{s}/sigmoid/1": 0.9,
                  f"transform/{s}": 0.05, f"transform/{s}/trapezoid/0": 1e-4, f"transform/{s}/trapezoid/1": 1e-3,
                  f"transform/{s}/trapezoid/2": 1.0, f"transform/{s}/trapezoid/3": 10.0, f"decay/{s}": 0.1})
    return v


def adpm_params(which: str = "full") -> np.ndarray:
    """P-ref = effective run/HCP102513/input.dat (reaction only, Appendix C-1); P-full exercises every term."""
    return P.flat_params(P.ADPM, adpm_param_dict(which))


def pihna_params(which: str = "ref") -> np.ndarray:
    """run/PIHNA/input.dat:24-49; 'full' additionally switches on the c/h diffusion and all taxis terms."""
    v = {"cells_min_capacity": 1.0, "cells_max_capacity": 2.39e5, "cells_max_capacity/exponent": 3.0,
         "cytokines_max_capacity": 1.0e-8, "necrosis/c": 500.0, "necrosis/h": 200.0, "necrosis/v": 300.0,
         "produce/c": -2.5, "switch/c/to/h": 1.0, "switch/h/to/c": 1.82, "switch/h/to/n": 0.5,
         "diffuse/v": 0.5, "produce/v": 10.0, "secrete/a/from/c": 2.77e-13, "secrete/a/from/h": 5.22e-10,
         "decay/a": 5678.4}
    if which == "full":
        v.update({"diffuse/c": 0.05, "taxis/c": 1e-6, "diffuse/h": 0.02, "taxis/h": 2e-6, "taxis/v": 1e-3,
                  "uptake/a/from/v": 1e-4})
    return P.flat_params(P.PIHNA, v)


def ripf_params(which: str = "ref") -> np.ndarray:
    """run/RIPF133/input.dat:12-36; 'full' activates kappa, omicro, HU coupling and radiotaxis as well."""
    v = {"RT_dose/broad/fractions": 28, "RT_dose/focus/fractions": 8, "volume_fraction/stroma": 0.30,
         "volume_fraction/parenchyma": 0.20, "volume_fraction/exponent": 2.5, "volume_fraction/min_vacant": 1e-5,
         "HU/min": -1000.0, "HU/max": 1000.0, "HU/phi/tolerance": 1e-3, "cc/delta": 0.0864, "cc/delta/RT/a": 0.3,
         "cc/delta/RT/b": 0.03, "fb/lambda": 0.01, "fb/lambda/RT/r": 1.0, "fb/omega": 0.1, "fb/diffusion": 1e-20,
         "fb/haptotaxis": 0.05}
    if which == "full":
        v.update({"cc/kappa": 0.01, "cc/kappa/RT/c": 0.02, "fb/omicro": 0.02, "fb/omicro/fb/b": 0.1,
                  "fb/lambda/HU/r": -600.0, "fb/diffusion": 1e-3, "fb/radiotaxis": 0.02,
                  "HU/phi/cc/build": 5.0, "HU/phi/cc/decay": -3.0, "HU/phi/cc/rate": 2.0,
                  "HU/phi/fb/build": 7.0, "HU/phi/fb/decay": -2.0, "HU/phi/fb/rate": 4.0, "HU/phi/tolerance": 1e-4})
    return P.flat_params(P.RIPF, v)


def proteas_params() -> np.ndarray:
    """No run/ case ships for PROTEAS; a plausible set in which no term vanishes."""
    v = {"cells/total_capacity": 1.0, "radiotherapy/max_dosage": 60.0,
         "host/proliferation": 0.3, "host/vsc_threshold": 0.05, "host/RT_death_rate": 0.2, "host/RT_exp_a": 0.03,
         "host/RT_exp_b": 0.003, "host/necrosis_rate": 0.1,
         "tumour/diffusion": 2e-3, "tumour/diffusion_host": 1e-3, "tumour/proliferation": 0.8,
         "tumour/vsc_threshold": 0.02, "tumour/RT_death_rate": 0.5, "tumour/RT_exp_a": 0.05, "tumour/RT_exp_b": 0.005,
         "tumour/necrosis_rate": 0.2,
         "necrosis/clearance": 0.05, "necrosis/slope": 10.0, "necrosis/vsc_threshold": 0.5,
         "vascular/proliferation": 0.4, "vascular/necrosis_rate": 0.15,
         "oedema/diffusion": 5e-3, "oedema/proliferation": 0.6, "oedema/vsc_threshold": 0.1, "oedema/RT_coeff": 0.2,
         "oedema/RT_exp": 1.5, "oedema/reabsorption_rate": 0.3}
    return P.flat_params(P.PROTEAS, v)


def hcc_params() -> np.ndarray:
    """run/Coupled/HCC/input.dat:28-30 + non-zero produce/diffuse/necrosis."""
    return P.flat_params(P.HCC, {"cells/min_capacity": 0.0, "cells/max_capacity": 1.0,
                                 "cells/max_capacity/exponent": 3.0, "produce/l": 0.2, "diffuse/c": 2e-3,
                                 "mechano/c": 1e-3, "produce/c": 0.9, "necrosis/l": 0.3, "necrosis/c": 0.4,
                                 "necrosis/pressure": 0.1})


# -------------------------------------------------------------------------------------------- fields
def _blob(xyz, centre, width):
    d2 = ((xyz - np.asarray(centre)) ** 2).sum(1)
    return np.exp(-d2 / (2.0 * width * width))


def adpm_fields(conn, xyz, smooth: bool = False):
    """PrP = 1; A_b = Tau = 0.01 on a seeded 0.24 % of the nodes (62 / 25 935 in the shipped field file);
    tract vectors ~ N(0, 0.1^2) per component (seed 2).  smooth=True adds smooth seeds so that gradients,
    thresholds and taxis alignment are exercised on every element."""
    N, E = xyz.shape[0], conn.shape[0]
    rng = np.random.default_rng(1)
    u = np.zeros((N, 3))
    u[:, 0] = 1.0
    seeds = rng.random(N) < 0.0024
    u[seeds, 1] = 0.01
    u[seeds, 2] = 0.01
    if smooth:
        L = xyz.max(0) - xyz.min(0)
        c = xyz.min(0) + 0.5 * L
        u[:, 0] = 0.6 + 0.5 * _blob(xyz, c + 0.1 * L, 0.3 * L.max())
        u[:, 1] = 0.95 * _blob(xyz, c - 0.15 * L, 0.25 * L.max()) + 1e-5
        u[:, 2] = 0.9 * _blob(xyz, c + [0.2 * L[0], -0.1 * L[1], 0.0], 0.2 * L.max()) + 2e-5
    tracts = np.random.default_rng(2).normal(0.0, 0.1, size=(E, 3))
    return u, tracts


def pihna_fields(xyz, smooth: bool = False):
    """n=c=h=a=0, v=7170, c=1000 on 0.09 % of the nodes (23 / 24 903 in run/PIHNA/*.dat)."""
    N = xyz.shape[0]
    u = np.zeros((N, 5))
    u[:, 3] = 7170.0
    rng = np.random.default_rng(3)
    u[rng.random(N) < 0.0009, 1] = 1000.0
    if smooth:
        L = xyz.max(0) - xyz.min(0)
        c = xyz.min(0) + 0.5 * L
        u[:, 0] = 50.0 * _blob(xyz, c, 0.15 * L.max())
        u[:, 1] = 1.0e3 * _blob(xyz, c + 0.1 * L, 0.2 * L.max())
        u[:, 2] = 3.0e2 * _blob(xyz, c - 0.1 * L, 0.25 * L.max())
        u[:, 3] = 7170.0 + 5.0e2 * _blob(xyz, c + [0.0, 0.2 * L[1], 0.0], 0.3 * L.max())
        u[:, 4] = 5.0e-9 * _blob(xyz, c, 0.3 * L.max())
    return u


def ripf_fields(xyz):
    """HU smooth in [-800,-500] + U(-20,20) noise (the shipped lung field spans -1019..1094 but is spatially
    coherent; i.i.d. U(-1000,0) noise would make the haptotaxis term dominate at unit-cube scale), cc blob in
    [0,1], fb small blob; RT_broad <= 67, RT_focus <= 6.7 Gaussians (shipped maxima)."""
    N = xyz.shape[0]
    rng = np.random.default_rng(4)
    L = xyz.max(0) - xyz.min(0)
    c = xyz.min(0) + 0.5 * L
    u = np.zeros((N, 3))
    u[:, 0] = -800.0 + 300.0 * _blob(xyz, c - 0.1 * L, 0.3 * L.max()) + rng.uniform(-20.0, 20.0, N)
    u[:, 1] = 0.8 * _blob(xyz, c, 0.2 * L.max())
    u[:, 2] = 0.3 * _blob(xyz, c + 0.15 * L, 0.25 * L.max())
    rt = np.zeros((N, 2))
    rt[:, 0] = 67.0 * _blob(xyz, c, 0.35 * L.max())
    rt[:, 1] = 6.7 * _blob(xyz, c, 0.15 * L.max())
    return u, rt


def proteas_fields(xyz):
    N = xyz.shape[0]
    L = xyz.max(0) - xyz.min(0)
    c = xyz.min(0) + 0.5 * L
    u = np.zeros((N, 5))
    u[:, 0] = 0.55 - 0.3 * _blob(xyz, c, 0.2 * L.max())
    u[:, 1] = 0.35 * _blob(xyz, c, 0.18 * L.max())
    u[:, 2] = 0.1 * _blob(xyz, c, 0.08 * L.max())
    u[:, 3] = 0.03 + 0.1 * _blob(xyz, c + 0.1 * L, 0.3 * L.max())
    u[:, 4] = 0.2 * _blob(xyz, c - 0.1 * L, 0.25 * L.max())
    aux = np.zeros((N, 2))
    aux[:, 0] = 30.0 + 25.0 * _blob(xyz, c, 0.3 * L.max())   # var 0 is what the reference reads (App. C-4)
    aux[:, 1] = 50.0 * _blob(xyz, c, 0.2 * L.max())
    return u, aux


def hcc_fields(xyz):
    N = xyz.shape[0]
    L = xyz.max(0) - xyz.min(0)
    c = xyz.min(0) + 0.5 * L
    u = np.zeros((N, 3))
    u[:, 0] = 0.5 - 0.3 * _blob(xyz, c, 0.2 * L.max())
    u[:, 1] = 0.4 * _blob(xyz, c, 0.2 * L.max())
    u[:, 2] = 0.15 * _blob(xyz, c, 0.1 * L.max())
    return u


# [upstream] Tet4/Hex8::side_nodes_map: local nodes of side s in libMesh's side numbering
SIDE_NODES = {4: ((0, 2, 1), (0, 1, 3), (1, 2, 3), (2, 0, 3)),
              8: ((0, 3, 2, 1), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7), (4, 5, 6, 7))}


def boundary_sides(conn, xyz, axis, value, tol=1e-9):
    """(elem, side) of the element sides that lie in the plane x[axis] = value (what a tagged physical surface of a Gmsh file
    becomes in libMesh's BoundaryInfo): the boundary-condition sides of the solid-mechanics cases."""
    conn = np.asarray(conn)
    on = np.abs(np.asarray(xyz)[:, axis] - value) < tol
    elems, sides = [], []
    for s, loc in enumerate(SIDE_NODES[conn.shape[1]]):
        hit = np.nonzero(on[conn[:, list(loc)]].all(axis=1))[0]
        elems.append(hit)
        sides.append(np.full(hit.shape, s, dtype=np.int32))
    return np.concatenate(elems).astype(np.int64), np.concatenate(sides)
