#!/usr/bin/env python
"""bench.py -- RDC time-steps/s of the rdcFEs hot path (FE assembly + Krylov solve) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--n CELLS]

Workload (BASELINE.json metric, SURVEY.md 8d case S1): ADPM operators (parameter set P-full: every term
active) on the synthetic unit-cube Kuhn-tet mesh n=119 -> 10 110 954 tets, 1 728 000 nodes, 5.18 M dofs,
dt = 0.05.  One "step" = one pass of the time-loop body adpm.C:63-76: rotate time levels, assemble K and F,
Krylov solve (BiCGStab + Jacobi by default, --ksp 0 = libMesh's GMRES(30)) to rtol 1e-12, check_solution.
  * value : device-resident steps/s (inputs in HBM when the timed region starts), CUDA events, max over ranks
  * e2e   : the SAME steps (state restored in between) through the C ABI with HOST buffers every step:
            rdc_set_solution (pinned H2D) -> rdc_step -> rdc_get_solution_owned (D2H; distributed: every rank moves
            its own dofs only)
  * roofline : the dominant kernel (block-CSR SpMV): algorithmic bytes / mean launch time, timed live with an
            event pair around every SpMV launch of the timed steps; assembly reported next to it
  * cpu_baseline : the CPU oracle (port of the reference path: element loop + scalar CSR + GMRES(30) +
            block-Jacobi/ILU(0)) on the box's host cores: a few steps of the SAME mesh (about 10 s per step on 16
            cores); --cpu-n picks a smaller sample mesh, then scaled per element
N > 1 (torchrun): strong scaling of the same mesh, METIS node partition, ghost exchange + all-reduce over NVLink
peer memory inside the Krylov kernels (NCCL for set-up and as fallback).  --model pihna: secondary 5-species case.
--impl reference: times the CPU port only (the real rdcFEs binary needs libMesh/PETSc/MPI: not installable).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

DT = 0.05
METRIC = "rdc_time_steps_per_s_10Mtet_adpm"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Number of samples read so far: brackets the timed region inside a sampler that started earlier."""
        return len(self.rows)

    def stop(self, first=0, last=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if last is not None and last > first:
            self.rows = self.rows[first:last]       # the samples taken DURING the timed region
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def workload(n, model="adpm"):
    """(conn, xyz, params, u0, tract vectors or None) of the synthetic case; "adpm" is the BASELINE.json workload,
    "pihna" (SURVEY 8d case S2: run/PIHNA/input.dat parameters, 5 species) is a secondary measurement."""
    from rdcfes_b200 import synth
    conn, xyz = synth.kuhn_cube(n)
    if model == "pihna":
        import cases
        p, u0, ef, nf = cases.case(cases.PIHNA, conn, xyz, "full")
        return conn, xyz, p, u0, None
    u0, tracts = synth.adpm_fields(conn, xyz, smooth=True)
    return conn, xyz, synth.adpm_params("full"), u0, tracts


def cpu_port_run(n, steps, warmup, nthreads, keep_state=None):
    """The CPU port on a bounded sample mesh: returns (seconds per step, elements, its per step, phases).
    keep_state (a dict) receives the oracle's solution after the warmup + steps steps from u0: the parity check
    of the bench line compares the GPU path with it on the SAME mesh."""
    from oracle import oracle as O
    conn, xyz, params, u0, tracts = workload(n)
    pr = O.Problem(O.ADPM, O.TET4, conn, xyz, params, u0, elem_field=tracts, nthreads=nthreads)
    times, its_all, ta, ts = [], [], [], []
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        its, _ = pr.step(DT, pc=O.PC_ILU, nblocks=nthreads, restart=30, rtol=1e-12, maxits=5000)
        t1 = time.perf_counter()
        if k >= warmup:
            times.append(t1 - t0); its_all.append(its); ta.append(pr.t_assemble); ts.append(pr.t_solve)
    if keep_state is not None:
        keep_state["u"] = pr.u.copy()
        keep_state["steps"] = warmup + steps
    return float(np.mean(times)), conn.shape[0], float(np.mean(its_all)), float(np.mean(ta)), float(np.mean(ts))


def run_reference(args):
    """--impl reference: the reference's CPU path (our libMesh-free port, oracle/) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ncores = os.cpu_count() or 1
    n_sample = args.cpu_n
    sec, E, its, ta, ts = cpu_port_run(n_sample, args.steps, args.warmup, ncores)
    E_full = 6 * args.n ** 3
    value = 1.0 / (sec * E_full / E)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"S1 ADPM P-full, unit-cube Kuhn tets n={args.n} ({E_full} tets, {(args.n + 1) ** 3} nodes, "
                                  f"{3 * (args.n + 1) ** 3} dofs), dt={DT}, GMRES(30)+BJacobi/ILU(0) rtol 1e-12 (libMesh defaults)",
                      "parallelism": f"CPU port of the reference path, {ncores} OpenMP threads (one ILU(0) block per thread)"},
           "cpu_baseline": {"value": value, "unit": "steps/s", "cores": ncores, "kind": "port",
                            "sample": f"n={n_sample} ({E} tets) timed {sec:.3f} s/step ({ta:.3f} assemble + {ts:.3f} solve, "
                                      f"{its:.0f} GMRES(30)+BJacobi/ILU0 its), scaled x{E_full / E:.1f} by element count"},
           "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", "--cells", dest="n", type=int, default=119,
                    help="cells per edge (119 -> 10.1 M tets; under torchrun spell it --cells: --n is ambiguous there)")
    ap.add_argument("--cpu-n", type=int, default=119,
                    help="mesh of the CPU baseline: by default the workload itself (a few steps of it are the bounded sample)")
    ap.add_argument("--ksp", type=int, default=2, help="0 GMRES(30) (libMesh default), 1 CG, 2 BiCGStab; all Jacobi, rtol 1e-12")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--partitioner", type=int, default=0)
    ap.add_argument("--model", default="adpm", choices=["adpm", "pihna"],
                    help="adpm = the BASELINE.json workload; pihna = secondary 5-species measurement (no CPU baseline)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from rdcfes_b200 import system as rs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    uid = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(rs.make_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())

    t_setup0 = time.perf_counter()
    conn, xyz, params, u0, tracts = workload(args.n, args.model)
    N, E = xyz.shape[0], conn.shape[0]
    nv = 3 if args.model == "adpm" else 5
    dt = DT if args.model == "adpm" else 0.1
    sysm = rs.TransientRdcSystem(rs.ADPM if args.model == "adpm" else rs.PIHNA, rs.TET4, conn, xyz, device=local,
                                 rank=rank, nranks=world, partitioner=args.partitioner, unique_id=uid)
    sysm.set_parameters(params)
    if tracts is not None:
        sysm.set_elem_field(tracts)
    sysm.ksp = args.ksp
    stream = torch.cuda.current_stream()
    sysm.set_stream(stream.cuda_stream)  # torch.cuda.Event on this stream brackets the library's kernels
    u_host = torch.empty(nv * N, dtype=torch.float64).pin_memory()
    u_np = u_host.numpy()
    u_np[:] = u0.ravel()
    sysm.set_solution(u_np)
    t_setup = time.perf_counter() - t_setup0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident steps ("value")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                  # started before the warm-up so that it is already delivering samples
    for _ in range(args.warmup):
        sysm.step(dt)
    # value and e2e time the SAME steps: the state after the warm-up is kept and restored in between
    u_start = sysm.get_solution().copy()
    t_start = sysm.time
    st0 = sysm.stats()                   # running totals before the timed steps (the call synchronises: untimed)
    launches0 = st0.kernel_launches
    barrier()
    mark0 = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        sysm.step(dt)                    # nothing but the step inside the timed region
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(mark0, sampler.mark()) if rank == 0 else None
    # untimed: the state after warm-up + timed steps from u0 is the same problem at every N -- its norm and sum let the
    # 1/2/4/8-GPU lines of a scaling run be compared with each other (they must agree to the solver tolerance)
    u_end = sysm.get_solution()
    solution_check = {"l2": float(np.linalg.norm(u_end)), "sum": float(u_end.sum()), "steps_from_u0": args.warmup + args.steps,
                      "min": float(u_end.min())}
    st1 = sysm.stats()
    launches = st1.kernel_launches - launches0
    acc = {"its": st1.sum_iterations - st0.sum_iterations, "ms_asm": st1.sum_ms_assemble - st0.sum_ms_assemble,
           "ms_solve": st1.sum_ms_solve - st0.sum_ms_solve, "ms_clamp": st1.sum_ms_clamp - st0.sum_ms_clamp,
           "ms_spmv": st1.sum_ms_spmv - st0.sum_ms_spmv, "n_spmv": st1.sum_n_spmv - st0.sum_n_spmv}
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    steps_per_s = args.steps / (ms * 1e-3)

    # ---------------------------------------------------------------- end to end through host buffers
    e2e_steps = args.steps
    u_np[:] = u_start                    # back to the state the value steps started from (untimed)
    sysm.set_solution(u_np)
    sysm.time = t_start
    barrier()
    e0.record(stream)
    for _ in range(e2e_steps):
        sysm.set_solution(u_np)          # H2D of the step's input (pinned)
        sysm.step(dt)
        sysm.get_solution_owned(u_np)    # D2H of the step's result (distributed: every rank reads back its own dofs)
    e1.record(stream)
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = e2e_steps / (ms_e2e * 1e-3)

    st = sysm.stats()
    if rank != 0:
        sysm.close()
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    spmv_ms = acc["ms_spmv"] / max(acc["n_spmv"], 1)
    spmv_gbs = st.bytes_spmv / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else 0.0
    asm_ms = acc["ms_asm"] / args.steps
    asm_gbs = st.bytes_assemble / (asm_ms * 1e-3) / 1e9 if asm_ms > 0 else 0.0
    out = {
        "metric": METRIC if args.model == "adpm" else "rdc_time_steps_per_s_10Mtet_pihna", "value": steps_per_s, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"S1 ADPM P-full" if args.model == "adpm" else "S2 PIHNA (run/PIHNA parameters + c/h transport)") +
                               f", unit-cube Kuhn tets n={args.n} ({E} tets, {N} nodes, {nv * N} dofs), "
                               f"dt={dt}, {('GMRES(30)', 'CG', 'BiCGStab')[args.ksp]}+Jacobi rtol 1e-12",
                   "parallelism": f"node partition x{world} (METIS), ghost exchange + all-reduce over NVLink peer memory" if world > 1 else "single GPU",
                   "l2": f"operator ({st.nnzb_local * (7 if args.model == 'adpm' else 21) * 8 / 1e9:.2f} GB per rank) and vectors exceed the 126 MB L2; no flush needed between steps",
                   "setup_s": round(t_setup, 2)},
        "e2e": {"value": e2e_value, "unit": "steps/s",
                "h2d_bytes_per_step": int(8 * nv * (st.n_nodes_local + st.n_nodes_ghost) if world > 1 else 8 * nv * N),
                "d2h_bytes_per_step": int(8 * nv * st.n_nodes_local if world > 1 else 8 * nv * N),
                "per": "rank" if world > 1 else "job", "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": ("k_spmv_tma<3,KMASK_ADPM,*,2>" if args.model == "adpm" else "k_spmv_tma<5,KMASK_PIHNA,*,2>") + " (row-local block-CSR SpMV, TMA-staged, fused Jacobi scaling "
                               "and BiCGStab dot products)", "bound": "hbm",
                     "achieved": spmv_gbs, "peak": peak, "unit": "GB/s", "frac": spmv_gbs / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full captures of this
                     # workload on one GPU (profiles/r1h_assemble_spmv_full.csv: 1.672 + 0.043 GB; r1d: 1.630 + 0.043 GB);
                     # other sizes were not captured
                     "traffic": 1.716e9 if (args.n == 119 and world == 1 and args.model == "adpm") else None,
                     "peak_source": peak_src, "bytes_per_launch": int(st.bytes_spmv), "ms_per_launch": spmv_ms,
                     "launches_timed": acc["n_spmv"], "frac_of_nominal_8TBs": spmv_gbs / 8000.0},
        "roofline_assembly": {"kernel": ("k_assemble<Adpm,4,128,4>" if args.model == "adpm" else "k_assemble<Pihna,4,128,2>") + " (fp64-pipe bound, see DESIGN.md 4.1)", "bound": "hbm", "achieved": asm_gbs, "peak": peak,
                              "unit": "GB/s", "frac": asm_gbs / peak, "bytes_per_launch": int(st.bytes_assemble),
                              "index_bytes_per_launch": int(st.bytes_index), "ms_per_launch": asm_ms},
        "phases_ms_per_step": {"assemble": asm_ms, "solve": acc["ms_solve"] / args.steps,
                               "clamp": acc["ms_clamp"] / args.steps, "spmv_in_solve": acc["ms_spmv"] / args.steps},
        "krylov_its_per_step": acc["its"] / args.steps,
        "solution_check": solution_check,
    }
    # parity on the bench mesh itself (N = 1, untimed): the same k steps from u0 on the GPU and with the CPU oracle
    cpu_steps, cpu_warm = 2, 1
    want_cpu = not args.no_cpu_baseline and args.model == "adpm" and world == 1
    u_gpu_k = None
    if want_cpu and args.cpu_n == args.n:
        u_np[:] = u0.ravel()
        sysm.set_solution(u_np)
        sysm.time = 0.0
        for _ in range(cpu_steps + cpu_warm):
            sysm.step(dt)
        u_gpu_k = sysm.get_solution().copy()
    sysm.close()
    if world > 1:
        dist.destroy_process_group()
    if want_cpu:  # reported at N = 1 only
        ncores = os.cpu_count() or 1
        state = {}
        sec, Es, its, ta, ts = cpu_port_run(args.cpu_n, cpu_steps, cpu_warm, ncores, keep_state=state)
        if u_gpu_k is not None:
            uo, ug = state["u"].reshape(-1, nv), u_gpu_k.reshape(-1, nv)
            out["parity"] = {"against": "CPU oracle (oracle/rdc_oracle.c, pinned to the reference's own sources by "
                                        "tests/test_ref_pin.py), same mesh, same u0, GMRES(30)+BJacobi/ILU(0) vs the GPU solver",
                             "steps": state["steps"],
                             "rel_l2": float(np.linalg.norm(ug - uo) / np.linalg.norm(uo)),
                             "per_species_rel_l2": [float(np.linalg.norm(ug[:, a] - uo[:, a]) / max(np.linalg.norm(uo[:, a]), 1e-300))
                                                    for a in range(nv)],
                             "max_abs": float(np.abs(ug - uo).max()), "tolerance": 1e-6}
        val = 1.0 / (sec * E / Es)
        out["cpu_baseline"] = {"value": val, "unit": "steps/s", "cores": ncores, "kind": "port",
                               "sample": f"n={args.cpu_n} ({Es} tets) {sec:.3f} s/step ({ta:.3f} assemble + {ts:.3f} solve, "
                                         f"{its:.0f} GMRES(30)+BJacobi/ILU0 its), scaled x{E / Es:.1f} by element count"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
