"""Times the assembly kernel alone (rdc_assemble, CUDA events of rdc_stats) on the bench meshes.

    python tools/asm_probe.py [--n 119] [--models adpm,pihna] [--reps 10]

Prints one line per model: mean / min ms over `reps` assemblies after 2 warm-ups, algorithmic GB/s.  Tuning knobs come
from the environment (RDC_ASM_MINB, RDC_ASM_PAIRS)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=119)
    ap.add_argument("--models", default="adpm,pihna")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    import cases
    from rdcfes_b200 import synth
    conn, xyz = synth.kuhn_cube(args.n)
    for name in args.models.split(","):
        model = {"adpm": cases.ADPM, "pihna": cases.PIHNA, "ripf": cases.RIPF, "proteas": cases.PROTEAS, "hcc": cases.HCC}[name]
        if model == cases.ADPM:
            u0, tr = synth.adpm_fields(conn, xyz, smooth=True)
            p, ef, nf = synth.adpm_params("full"), tr, None
        else:
            p, u0, ef, nf = cases.case(model, conn, xyz, "full")
        gpu = cases.gpu_system(model, cases.TET4, conn, xyz, p, u0, ef, nf)
        dt = cases.DT[model]
        gpu.rotate()
        ts = []
        for k in range(args.reps + 2):
            gpu.assemble(dt, dt)
            t = gpu.stats().ms_assemble
            if k >= 2:
                ts.append(t)
        st = gpu.stats()
        print(f"{name} n={args.n} MINB={os.environ.get('RDC_ASM_MINB', '-')} PAIRS={os.environ.get('RDC_ASM_PAIRS', '-')}: "
              f"assemble mean {np.mean(ts):.3f} ms min {np.min(ts):.3f} ms  ({st.bytes_assemble / np.mean(ts) / 1e6:.0f} GB/s algorithmic, "
              f"index {st.bytes_index / 1e6:.0f} MB)", flush=True)
        gpu.close()


if __name__ == "__main__":
    main()
