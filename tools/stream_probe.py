"""Read-only streaming bandwidth over the assembled operator of the headline case, next to the SpMV time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from rdcfes_b200 import system as rs
conn, xyz, params, u0, tracts = bench.workload(119)
s = rs.TransientRdcSystem(rs.ADPM, rs.TET4, conn, xyz)
s.set_parameters(params); s.set_elem_field(tracts); s.set_solution(u0.ravel())
s.rotate(); s.assemble(0.05, 0.05)
for per_sm in (2, 4, 8, 16):
    ms, nb = s.bench_stream(20, per_sm)
    print(f"stream read {per_sm:2d} CTAs/SM: {ms:.4f} ms  {nb / ms / 1e6:.0f} GB/s")
st = s.stats()
for tma in (1, 0):
    s.set_option("spmv_tma", tma)
    ms = s.bench_spmv(20)
    print(f"spmv plain tma={tma}: {ms:.4f} ms  {st.bytes_spmv / ms / 1e6:.0f} GB/s")
