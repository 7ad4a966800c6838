"""Timing of the solid-mechanics path on one GPU next to the CPU oracle (not a bench line; numbers for DESIGN.md).
usage: python tools/solid_probe.py [n=40] [elem=4|8]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import solid_cases as SC
from oracle import solid as S
from rdcfes_b200 import solid as G

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
et = int(sys.argv[2]) if len(sys.argv) > 2 else 4
c = SC.compression_case(et, n=n, penalty=1.0e8)
print(f"solid probe: elem {et}, n {n}: {c.E} elements, {c.N} nodes, {len(c.side_elem)} boundary sides")
g = G.from_case(c)
g.ksp = G.KSP_GMRES
g.assemble(0.1)
t0 = time.time()
for _ in range(10):
    g.assemble(0.1)
t_asm = (time.time() - t0) / 10
orc = S.OracleSolid(c)
t0 = time.time()
val, rhs = orc.assemble(c.xund.ravel(), 0.1)
t_cpu = time.time() - t0
rows, rowptr, col, v, r = g.download_csr()
print(f"assembly: device {t_asm * 1e3:.3f} ms (wall, incl. sync), oracle (1 thread) {t_cpu * 1e3:.1f} ms, "
      f"max|dJ|/max|J| {np.abs(v - val).max() / np.abs(val).max():.2e}, max|dR|/max|R| {np.abs(r - rhs).max() / max(np.abs(rhs).max(), 1e-300):.2e}")
for ksp, name in ((G.KSP_GMRES, "gmres30"), (G.KSP_BICGSTAB, "bicgstab")):
    g.set_positions(c.xund)
    g.ksp = ksp
    for step in (1, 2):
        t0 = time.time()
        info = g.run_solver(0.1 * step)
        print(f"load step {step} [{name}]: {info}, {1e3 * (time.time() - t0):.1f} ms")
if c.N <= 30000:
    xo = c.xund.copy().ravel()
    t0 = time.time()
    xo, io = orc.newton(xo, 0.1)
    print(f"oracle load step 1: {io}, {1e3 * (time.time() - t0):.1f} ms")
st = g.stats()
print("kernel launches", st.kernel_launches)
