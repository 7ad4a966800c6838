"""Cost of the persistent solver's grid barriers (rdc_bench_barrier) for 1..6 CTAs per SM."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402

conn, xyz = cases.mesh(cases.TET4, 6)
p, u0, ef, nf = cases.case(cases.ADPM, conn, xyz, "full")
gpu = cases.gpu_system(cases.ADPM, cases.TET4, conn, xyz, p, u0, ef, nf)
for mode in (0, 1, 2):
    for per_sm in (1, 2, 4, 6):
        us = C.c_double()
        rc = gpu._L.rdc_bench_barrier(gpu._h, 2000, per_sm, mode, C.byref(us))
        print(f"mode {mode} ({('barrier', 'reduce by the last block + barrier', 'barrier + every CTA reduces')[mode]}) {148 * per_sm} CTAs: {us.value:.2f} us rc={rc}")
gpu.close()
