"""Shared builders for the parity tests: the same seeded inputs are handed to the CPU oracle and to the
CUDA path (through the C ABI)."""
import numpy as np

from rdcfes_b200 import params as P
from rdcfes_b200 import synth

ADPM, PIHNA, RIPF, PROTEAS, HCC = 0, 1, 2, 3, 4
TET4, HEX8 = 4, 8
NAMES = {ADPM: "adpm", PIHNA: "pihna", RIPF: "ripf", PROTEAS: "proteas", HCC: "hcc"}
DT = {ADPM: 0.05, PIHNA: 0.1, RIPF: 0.1, PROTEAS: 0.05, HCC: 0.01}


def mesh(elem_type, n, distort=0.0, permute=None, length=1.0):
    if elem_type == TET4:
        conn, xyz = synth.kuhn_cube(n, length)
    else:
        conn, xyz = synth.hex_cube(n, length)
    if distort:
        xyz = synth.distort(xyz, distort * length / n)
    if permute is not None:
        conn, xyz = synth.permute_nodes(conn, xyz, permute)
    return conn, xyz


def case(model, conn, xyz, variant="full"):
    """(params, u0 [N,nv], elem_field or None, nodal_field or None) for a model on a mesh of size ~ 1."""
    L = xyz.max(0) - xyz.min(0)
    if model == ADPM:
        u0, tr = synth.adpm_fields(conn, xyz, smooth=True)
        return synth.adpm_params(variant), u0, tr, None
    if model == PIHNA:
        p = synth.pihna_params(variant)
        u0 = synth.pihna_fields(xyz, smooth=True)
        if variant == "full":
            # brain-sized diffusion numbers on a unit cube: scale the transport coefficients with L^2
            for key in ("diffuse/c", "diffuse/h", "diffuse/v", "taxis/c", "taxis/h", "taxis/v"):
                k = [i for i, (kk, _) in enumerate(P.TABLES[P.PIHNA]) if kk == key][0]
                p[k] *= 1e-3 * float(L.max()) ** 2
        return p, u0, None, None
    if model == RIPF:
        u0, rt = synth.ripf_fields(xyz)
        return synth.ripf_params(variant), u0, None, rt
    if model == PROTEAS:
        u0, aux = synth.proteas_fields(xyz)
        return synth.proteas_params(), u0, None, aux
    if model == HCC:
        return synth.hcc_params(), synth.hcc_fields(xyz), None, None
    raise ValueError(model)


def oracle_problem(model, elem_type, conn, xyz, params, u0, ef, nf, nthreads=4):
    from oracle import oracle as O
    pr = O.Problem(model, elem_type, conn, xyz, params, u0, elem_field=ef, nodal_field=nf, nthreads=nthreads)
    if model == RIPF:
        pr.ripf_initial_check(DT[RIPF])
    return pr


def gpu_system(model, elem_type, conn, xyz, params, u0, ef, nf, node_dof_base=None, **kw):
    from rdcfes_b200.system import TransientRdcSystem
    s = TransientRdcSystem(model, elem_type, conn, xyz, node_dof_base=node_dof_base, **kw)
    s.set_parameters(params)
    if ef is not None:
        s.set_elem_field(ef)
    if nf is not None:
        s.set_nodal_field(nf)
    nv = P.NVARS[model]
    if node_dof_base is None:
        s.set_solution(u0)
    else:
        u = np.zeros(s.n_dofs)
        for a in range(nv):
            u[np.asarray(node_dof_base) + a] = np.asarray(u0).reshape(-1, nv)[:, a]
        s.set_solution(u)
    if model == RIPF:  # ripf.C:53: check_solution once before the loop, at time 0
        s.time, s.dt = 0.0, DT[RIPF]
        s.check_solution()
    return s


def csr_tolerance(val_ref):
    """|err_ij| <= 1e-12 * max(|K_ij|, 1e-3 * max|K|): relative 1e-12 for every entry within three orders of
    magnitude of the largest one, absolute 1e-15 * max|K| below that."""
    scale = np.abs(val_ref).max()
    return 1e-12 * np.maximum(np.abs(val_ref), 1e-3 * scale)
