"""True cost of one BiCGStab iteration: solve time (CUDA events) for maxits = 4 and 12 on the same system."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from rdcfes_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
conn, xyz = synth.kuhn_cube(n)
u0, tr = synth.adpm_fields(conn, xyz, smooth=True)
gpu = cases.gpu_system(cases.ADPM, cases.TET4, conn, xyz, synth.adpm_params("full"), u0, tr, None)
gpu.ksp = 2
gpu.rotate()
gpu.assemble(0.05, 0.05)
u_start = gpu.get_solution().copy()
for persist in (1, 0):
    gpu.set_option("bicg_persist", persist)
    res = {}
    for maxits in (4, 12, 4, 12):
        gpu.maxits = maxits
        ts = []
        for rep in range(6):
            gpu.set_solution(u_start)
            its, _ = gpu.linear_solve()
            ts.append(gpu.stats().ms_solve)
        res[maxits] = (its, float(np.median(ts[1:])))
    (i4, t4), (i12, t12) = res[4], res[12]
    print(f"n={n} persist={persist}: {i4} its {t4*1e3:.1f} us, {i12} its {t12*1e3:.1f} us -> {(t12 - t4) / (i12 - i4) * 1e3:.1f} us per iteration, "
          f"fixed {t4*1e3 - i4 * (t12 - t4) / (i12 - i4) * 1e3:.1f} us")
gpu.close()
