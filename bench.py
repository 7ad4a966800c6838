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
  * parity : (N = 1) the same k steps from u0 on the GPU and with the CPU oracle on the bench mesh itself, relative L2
  * solution_check : norm / sum of the state after the timed steps: the same problem at every N, so the lines of a
            scaling run can be compared with each other
  * weak_scaling : (N > 1) a second, shorter measurement on a mesh with N x 10.1 M tets (--no-weak skips it)
  * ksp_gmres30 : (N = 1) the same steps with libMesh's default Krylov method, GMRES(30), for reference
N > 1 (torchrun): strong scaling of the same mesh, METIS node partition, ghost exchange + all-reduce over NVLink
peer memory inside the Krylov kernels (NCCL for set-up and as fallback).  --model pihna: secondary 5-species case.
--impl reference: times the CPU port on all host cores (the real rdcFEs binary needs libMesh/PETSc/MPI: not installable);
its line also carries `reference_assembly_sample`: the reference's OWN assemble_adpm (oracle/_ref, src/adpm.C compiled
unchanged, serial) timed next to the port on one thread on a bounded sample, with the two operators compared.
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


def kernel_counters(key):
    """ncu counters of one launch of a kernel on a given workload, from the committed summary of a --set full capture
    (profiles/r2_kernel_counters.json, written by profiles/counters.py from the .ncu-rep): None when that workload was
    not captured -- nothing here is a constant typed into this file."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_kernel_counters.json")) as fh:
            return json.load(fh).get(key)
    except Exception:  # noqa: BLE001
        return None


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


def reference_assembly_sample(n_s=24):
    """The reference's OWN assemble_adpm (oracle/_ref: src/adpm.C compiled unchanged against the serial libMesh stand-in)
    timed on a bounded sample next to the port on ONE thread, with the two operators compared: shows what the port's
    per-element cost is relative to the reference's own code.  None when oracle/_ref is not built."""
    try:
        from oracle import oracle as O
        from oracle import ref as R
        if not R.available():
            return None
        from rdcfes_b200 import params as P
        conn, xyz, params, u0, tracts = workload(n_s)
        rp = R.RefProblem(P.ADPM, 4, conn, xyz, params, u0, elem_field=tracts)
        t0 = time.perf_counter()
        val_r, rhs_r = rp.assemble(DT, DT)
        t_ref = time.perf_counter() - t0
        pr = O.Problem(O.ADPM, O.TET4, conn, xyz, params, u0, elem_field=tracts, nthreads=1)
        pr.u_old = pr.u.copy()
        t0 = time.perf_counter()
        val_p, rhs_p = pr.assemble(DT, DT)
        t_port = time.perf_counter() - t0
        E = conn.shape[0]
        same = bool(val_r.shape == val_p.shape)
        diff = float(np.abs(val_r - val_p).max() / np.abs(val_r).max()) if same else None
        rp.close()
        return {"kind": "reference", "what": "assemble_adpm of the reference (src/adpm.C unchanged, serial libMesh stand-in with a "
                                             "std::map-per-row matrix) against the port on one thread, same mesh and fields",
                "sample": f"n={n_s} ({E} tets)", "cores": 1, "reference_s": t_ref, "port_s": t_port,
                "reference_us_per_tet": 1e6 * t_ref / E, "port_us_per_tet": 1e6 * t_port / E,
                "max_abs_diff_over_max": diff}
    except Exception as exc:   # a checker must never break the reference arm
        return {"kind": "reference", "error": f"{type(exc).__name__}: {exc}"}


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
    ras = reference_assembly_sample()
    if ras is not None:
        out["reference_assembly_sample"] = ras
    print(json.dumps(out))


def measure(args, n, world, rank, local, steps, warmup, ksp, with_e2e=True, sampler=None):
    """Builds the system on the n^3-cell mesh and times `steps` steps after `warmup`: returns a dict (every rank) with the
    device-resident time, the end-to-end time, phase totals and the state check; leaves nothing allocated."""
    import torch
    import torch.distributed as dist
    from rdcfes_b200 import system as rs

    uid = None
    if world > 1:
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(rs.make_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())
    t_setup0 = time.perf_counter()
    conn, xyz, params, u0, tracts = workload(n, args.model)
    N, E = xyz.shape[0], conn.shape[0]
    nv = 3 if args.model == "adpm" else 5
    dt = DT if args.model == "adpm" else 0.1
    sysm = rs.TransientRdcSystem(rs.ADPM if args.model == "adpm" else rs.PIHNA, rs.TET4, conn, xyz, device=local,
                                 rank=rank, nranks=world, partitioner=args.partitioner, unique_id=uid)
    sysm.set_parameters(params)
    if tracts is not None:
        sysm.set_elem_field(tracts)
    del conn, xyz, tracts
    sysm.ksp = ksp
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

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- device-resident steps ("value")
    for _ in range(warmup):
        sysm.step(dt)
    u_start = sysm.get_solution().copy()   # value and e2e time the SAME steps: the state is restored in between
    t_start = sysm.time
    st0 = sysm.stats()                     # running totals before the timed steps (the call synchronises: untimed)
    barrier()
    mark0 = sampler.mark() if sampler else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        sysm.step(dt)                      # nothing but the step inside the timed region
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    mark1 = sampler.mark() if sampler else 0
    st1 = sysm.stats()
    res = {"N": N, "E": E, "nv": nv, "dt": dt, "setup_s": t_setup, "ms": ms, "marks": (mark0, mark1),
           "launches": st1.kernel_launches - st0.kernel_launches,
           "its": st1.sum_iterations - st0.sum_iterations, "ms_asm": st1.sum_ms_assemble - st0.sum_ms_assemble,
           "ms_solve": st1.sum_ms_solve - st0.sum_ms_solve, "ms_clamp": st1.sum_ms_clamp - st0.sum_ms_clamp,
           "ms_spmv": st1.sum_ms_spmv - st0.sum_ms_spmv, "n_spmv": st1.sum_n_spmv - st0.sum_n_spmv}
    # untimed: the state after warm-up + timed steps from u0 is the same problem at every N -- its norm and sum let the
    # 1/2/4/8-GPU lines of a scaling run be compared with each other (they must agree to the solver tolerance)
    u_end = sysm.get_solution()
    res["solution_check"] = {"l2": float(np.linalg.norm(u_end)), "sum": float(u_end.sum()), "min": float(u_end.min()),
                             "steps_from_u0": warmup + steps}
    # ---- end to end through host buffers
    if with_e2e:
        u_np[:] = u_start                  # back to the state the value steps started from (untimed)
        sysm.set_solution(u_np)
        sysm.time = t_start
        barrier()
        e0.record(stream)
        for _ in range(steps):
            sysm.set_solution(u_np)        # H2D of the step's input (pinned)
            sysm.step(dt)
            sysm.get_solution_owned(u_np)  # D2H of the step's result (distributed: every rank reads back its own dofs)
        e1.record(stream)
        barrier()
        res["ms_e2e"] = max_over_ranks(e0.elapsed_time(e1))
    st = sysm.stats()
    res.update(bytes_spmv=st.bytes_spmv, bytes_assemble=st.bytes_assemble, bytes_index=st.bytes_index,
               nnzb_local=st.nnzb_local, n_nodes_local=st.n_nodes_local, n_nodes_ghost=st.n_nodes_ghost,
               p2p_on=st.p2p_on, p2p_fused=st.p2p_fused, persistent=st.bicg_persistent)
    res["_sys"], res["_u_np"], res["_u0"] = sysm, u_np, u0
    return res


def solid_secondary(n=60):
    """SURVEY 8(f) rank 3 next to the headline: the solid-mechanics path (solid_system.C) on a 1.3 M-tet cube --
    uniaxial compression like run/Solid/uniaxial_compression (bottom clamped, top pushed down, penalty 1e8, the shipped
    Newton tolerances), two load steps from the undeformed state.  Device path only (rdcfes_b200/solid.py over the C ABI);
    wall-clock around calls that synchronise; parity of this path is the business of tests/test_gpu_solid.py."""
    from rdcfes_b200 import solid as G
    from rdcfes_b200 import synth
    conn, xyz = synth.kuhn_cube(n, 1.5)
    g = G.SolidSystem(4, conn, xyz)
    e0, s0 = synth.boundary_sides(conn, xyz, 2, 0.0)
    e1, s1 = synth.boundary_sides(conn, xyz, 2, xyz[:, 2].max())
    g.set_bcs([[0.0, 0.0, 0.0], [float("nan"), float("nan"), -0.75]], np.concatenate([e0, e1]), np.concatenate([s0, s1]),
              np.concatenate([np.zeros(e0.size, dtype=np.int32), np.ones(e1.size, dtype=np.int32)]), 1.0e8)
    g.options.update(max_nonlinear_iterations=10)
    g.ksp = G.KSP_BICGSTAB
    g.assemble(0.1)
    t0 = time.perf_counter()
    for _ in range(10):
        g.assemble(0.1)
    asm_ms = (time.perf_counter() - t0) / 10 * 1e3
    steps = []
    for l in (1, 2):
        t0 = time.perf_counter()
        info = g.run_solver(0.1 * l)
        steps.append({"ms": (time.perf_counter() - t0) * 1e3, **info})
    z = g.get_positions()[:, 2]
    g.close()
    return {"what": "SolidSystem load steps (neo-Hookean, penalty BCs), k_solid_assemble + k_solid_bc + BiCGStab/Jacobi Newton systems",
            "workload": f"unit-cube Kuhn tets n={n} ({conn.shape[0]} tets, {xyz.shape[0]} nodes), pseudo-time 0.1 and 0.2",
            "assemble_ms": asm_ms, "load_steps": steps, "top_face_z": float(z.max())}


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
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the weak-scaling measurement (N x 10.1 M tets)")
    ap.add_argument("--no-solid", action="store_true", help="N = 1: skip the secondary solid-mechanics measurement (SURVEY 8(f) rank 3)")
    ap.add_argument("--partitioner", type=int, default=0)
    ap.add_argument("--model", default="adpm", choices=["adpm", "pihna"],
                    help="adpm = the BASELINE.json workload; pihna = secondary 5-species measurement (no CPU baseline)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                  # started before the warm-up so that it is already delivering samples
    R = measure(args, args.n, world, rank, local, args.steps, args.warmup, args.ksp, True, sampler if rank == 0 else None)
    clocks = sampler.stop(*R["marks"]) if rank == 0 else None
    sysm, u_np, u0 = R.pop("_sys"), R.pop("_u_np"), R.pop("_u0")
    N, E, nv, dt = R["N"], R["E"], R["nv"], R["dt"]

    # ---- N = 1 extras (untimed region of the main measurement is over): parity state, GMRES(30) reference point, fp64 peak
    want_cpu = not args.no_cpu_baseline and args.model == "adpm" and world == 1
    cpu_steps, cpu_warm = 2, 1
    u_gpu_k, gm, dfma = None, None, None
    if world == 1:
        dfma = sysm.bench_dfma()
        if want_cpu and args.cpu_n == args.n:
            u_np[:] = u0.ravel()
            sysm.set_solution(u_np)
            sysm.time = 0.0
            for _ in range(cpu_steps + cpu_warm):
                sysm.step(dt)
            u_gpu_k = sysm.get_solution().copy()
        if args.ksp != 0:   # libMesh's default Krylov method on the same steps, for reference
            u_np[:] = u0.ravel()
            sysm.set_solution(u_np)
            sysm.time = 0.0
            sysm.ksp = 0
            for _ in range(args.warmup):
                sysm.step(dt)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0 = sysm.stats()
            e0.record(torch.cuda.current_stream())
            k_g = max(3, args.steps // 2)
            for _ in range(k_g):
                sysm.step(dt)
            e1.record(torch.cuda.current_stream())
            torch.cuda.synchronize()
            s1 = sysm.stats()
            gm = {"value": k_g / (e0.elapsed_time(e1) * 1e-3), "unit": "steps/s", "ms_per_step": e0.elapsed_time(e1) / k_g,
                  "steps": k_g, "its_per_step": (s1.sum_iterations - s0.sum_iterations) / k_g,
                  "what": "same workload, GMRES(30)+Jacobi (libMesh's default KSP), device-resident"}
    sysm.close()
    del sysm, u_np

    # ---- weak scaling (N > 1): N x the tets on N GPUs, fewer steps
    weak = None
    if world > 1 and not args.no_weak and args.model == "adpm":
        n_w = int(round((world * args.n ** 3) ** (1.0 / 3.0)))
        W = measure(args, n_w, world, rank, local, max(3, args.steps // 2), 3, args.ksp, False, None)
        W.pop("_sys").close()
        W.pop("_u_np"); W.pop("_u0")
        k_w = max(3, args.steps // 2)
        weak = {"cells": n_w, "tets": int(W["E"]), "tets_per_gpu": int(W["E"] // world), "steps": k_w,
                "value": k_w / (W["ms"] * 1e-3), "unit": "steps/s", "ms_per_step": W["ms"] / k_w,
                "phases_ms_per_step": {"assemble": W["ms_asm"] / k_w, "solve": W["ms_solve"] / k_w},
                "krylov_its_per_step": W["its"] / k_w, "setup_s": round(W["setup_s"], 1),
                "note": "compare ms_per_step x (its of the N = 1 line / its here) with the N = 1 line's ms_per_step: "
                        "the Jacobi-scaled spectrum widens with the mesh, so the larger mesh needs more iterations per step"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    steps = args.steps
    ms = R["ms"]
    spmv_ms = R["ms_spmv"] / max(R["n_spmv"], 1)
    spmv_gbs = R["bytes_spmv"] / (spmv_ms * 1e-3) / 1e9 if spmv_ms > 0 else 0.0
    asm_ms = R["ms_asm"] / steps
    asm_gbs = R["bytes_assemble"] / (asm_ms * 1e-3) / 1e9 if asm_ms > 0 else 0.0
    adpm = args.model == "adpm"
    kname_spmv = "k_spmv_tma<3,KMASK_ADPM,*,2>" if adpm else "k_spmv_tma<5,KMASK_PIHNA,*,2>"
    kname_asm = "k_assemble<Adpm,4,128,4>" if adpm else "k_assemble<Pihna,4,128,2>"
    persistent = bool(R["persistent"])
    key = f"n{args.n}_{args.model}_1gpu" if world == 1 else None
    cs = kernel_counters(f"spmv_{key}") if key else None
    ca = kernel_counters(f"assemble_{key}") if key else None
    out = {
        "metric": METRIC if adpm else "rdc_time_steps_per_s_10Mtet_pihna", "value": steps / (ms * 1e-3), "unit": "steps/s",
        "n_gpus": world, "steps": steps, "warmup": args.warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("S1 ADPM P-full (every term of adpm.C:497-590 active), smooth seeded fields (synth.adpm_fields smooth=True: "
                                "gradients, thresholds and tract alignment exercised on every element; harder than S1's 0.24 % point seeds)"
                                if adpm else "S2 PIHNA (run/PIHNA parameters + c/h transport)") +
                               f", unit-cube Kuhn tets n={args.n} ({E} tets, {N} nodes, {nv * N} dofs), "
                               f"dt={dt}, {('GMRES(30)', 'CG', 'BiCGStab')[args.ksp]}+Jacobi rtol 1e-12",
                   "parallelism": (f"node partition x{world} (METIS), Morton-ordered rows per rank, ghost exchange + all-reduce over NVLink peer "
                                   f"memory inside one cooperative BiCGStab kernel per solve (p2p_on={R['p2p_on']})") if world > 1 else
                                  "single GPU, Morton-ordered rows",
                   "l2": f"operator ({R['nnzb_local'] * (7 if adpm else 21) * 8 / 1e9:.2f} GB per rank) and vectors exceed the 126 MB L2; no flush needed between steps",
                   "setup_s": round(R["setup_s"], 2)},
        "e2e": {"value": steps / (R["ms_e2e"] * 1e-3), "unit": "steps/s",
                "h2d_bytes_per_step": int(8 * nv * (R["n_nodes_local"] + R["n_nodes_ghost"]) if world > 1 else 8 * nv * N),
                "d2h_bytes_per_step": int(8 * nv * R["n_nodes_local"] if world > 1 else 8 * nv * N),
                "per": "rank" if world > 1 else "job", "steps": steps},
        "gpu_launches": int(R["launches"]),
        "clocks": clocks,
        "roofline": {"kernel": kname_spmv + " (row-local block-CSR SpMV, TMA-staged, fused Jacobi scaling and BiCGStab dot products)",
                     "bound": "hbm", "achieved": spmv_gbs, "peak": peak, "unit": "GB/s", "frac": spmv_gbs / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full summary of
                     # this workload (profiles/r2_kernel_counters.json); null when this size / GPU count was not captured
                     "traffic": (cs["dram_read"] + cs["dram_write"]) if cs else None,
                     "traffic_source": cs["source"] if cs else None,
                     "peak_source": peak_src, "bytes_per_launch": int(R["bytes_spmv"]), "ms_per_launch": spmv_ms,
                     "launches_timed": int(R["n_spmv"]), "frac_of_nominal_8TBs": spmv_gbs / 8000.0,
                     "timed_with": ("%globaltimer around the SpMV phases (incl. their reduction barrier) inside the cooperative BiCGStab kernel"
                                    if persistent else "cudaEvent pair around every SpMV launch of the timed steps")},
        "roofline_assembly": {"kernel": kname_asm + " (pair-parallel, atomic-free; bound by the fp64 pipe and the issue slots, not by HBM: DESIGN.md 4.1)",
                              "bound": "hbm", "achieved": asm_gbs, "peak": peak,
                              "unit": "GB/s", "frac": asm_gbs / peak, "bytes_per_launch": int(R["bytes_assemble"]),
                              "index_bytes_per_launch": int(R["bytes_index"]), "ms_per_launch": asm_ms,
                              "traffic": (ca["dram_read"] + ca["dram_write"]) if ca else None,
                              "fp64": ({"peak_tflops_measured": dfma, "peak_source": "rdc_bench_dfma: independent DFMA chains, full occupancy, this run",
                                        "fp64_thread_inst_per_launch": ca.get("fp64_thread_inst") if ca else None,
                                        "achieved_tinst_per_s": (ca["fp64_thread_inst"] / (asm_ms * 1e-3) / 1e12) if ca and ca.get("fp64_thread_inst") else None,
                                        "frac_of_fp64_issue_peak": (ca["fp64_thread_inst"] / (asm_ms * 1e-3)) / (dfma * 1e12 / 2.0)
                                        if ca and ca.get("fp64_thread_inst") and dfma else None,
                                        "warp_inst_per_launch": ca.get("inst_executed") if ca else None} if dfma else None)},
        "phases_ms_per_step": {"assemble": asm_ms, "solve": R["ms_solve"] / steps,
                               "clamp": R["ms_clamp"] / steps, "spmv_in_solve": R["ms_spmv"] / steps},
        "krylov_its_per_step": R["its"] / steps,
        "solution_check": R["solution_check"],
    }
    if weak:
        out["weak_scaling"] = weak
    if gm:
        out["ksp_gmres30"] = gm
    if world > 1:
        dist.destroy_process_group()
    if world == 1 and not args.no_solid and args.model == "adpm" and args.n >= 60:
        try:
            out["solid_mechanics"] = solid_secondary()
        except Exception as exc:  # noqa: BLE001 -- a secondary number must not take the headline line down
            out["solid_mechanics"] = {"error": str(exc)[:300]}
    if want_cpu:  # reported at N = 1 only
        ncores = os.cpu_count() or 1
        state = {}
        sec, Es, its, ta, ts = cpu_port_run(args.cpu_n, cpu_steps, cpu_warm, ncores, keep_state=state)
        val = 1.0 / (sec * E / Es)
        out["cpu_baseline"] = {"value": val, "unit": "steps/s", "cores": ncores, "kind": "port",
                               "sample": f"n={args.cpu_n} ({Es} tets) {sec:.3f} s/step ({ta:.3f} assemble + {ts:.3f} solve, "
                                         f"{its:.0f} GMRES(30)+BJacobi/ILU0 its), scaled x{E / Es:.1f} by element count"}
        if u_gpu_k is not None:
            uo, ug = state["u"].reshape(-1, nv), u_gpu_k.reshape(-1, nv)
            out["parity"] = {"against": "CPU oracle (oracle/rdc_oracle.c, pinned to the reference's own sources by "
                                        "tests/test_ref_pin.py), same mesh, same u0, GMRES(30)+BJacobi/ILU(0) vs the GPU solver",
                             "steps": state["steps"],
                             "rel_l2": float(np.linalg.norm(ug - uo) / np.linalg.norm(uo)),
                             "per_species_rel_l2": [float(np.linalg.norm(ug[:, a] - uo[:, a]) / max(np.linalg.norm(uo[:, a]), 1e-300))
                                                    for a in range(nv)],
                             "max_abs": float(np.abs(ug - uo).max()), "tolerance": 1e-6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
