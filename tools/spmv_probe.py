"""Stand-alone SpMV time (rdc_bench_spmv) and assembly time for the node orderings: RDC_NODE_ORDER=0/1, lexicographic or
randomly permuted input numbering."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from rdcfes_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 119
for seed in (None, 12345):
    conn, xyz = synth.kuhn_cube(n, permute_seed=seed)
    u0, tr = synth.adpm_fields(conn, xyz, smooth=True)
    gpu = cases.gpu_system(cases.ADPM, cases.TET4, conn, xyz, synth.adpm_params("full"), u0, tr, None)
    gpu.rotate()
    for _ in range(3):
        gpu.assemble(0.05, 0.05)
    st = gpu.stats()
    ms = gpu.bench_spmv(30)
    print(f"n={n} input numbering {'lexicographic' if seed is None else 'random'} node_order={os.environ.get('RDC_NODE_ORDER', 'default')}: "
          f"spmv {ms*1e3:.1f} us ({st.bytes_spmv / ms / 1e6:.0f} GB/s), assemble {st.ms_assemble:.3f} ms", flush=True)
    gpu.close()
