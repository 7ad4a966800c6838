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
