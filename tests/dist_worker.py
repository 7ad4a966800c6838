"""torchrun worker of the multi-GPU parity test: every rank drives the same problem through
rdc_create_distributed (METIS or RCB node partition, NCCL halo exchange + all-reduce); rank 0 compares the
gathered solution and the owned operator rows with the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import oracle as O
    from rdcfes_b200 import system as rs
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    failures = []
    # mesh large enough that every rank owns more nodes than it has ghosts (the peer-memory staging area needs that)
    n = {2: 7, 4: 10}.get(world, 12)
    # BASELINE.json configs: RIPF on 2 ranks (BiCGStab here), coupled_hcc with GMRES on 4 ranks (ksp 0 = GMRES(30))
    for model, ksp, partitioner, nsteps in ((cases.ADPM, 2, 0, 3), (cases.PIHNA, 0, 1, 2), (cases.RIPF, 2, 0, 12),
                                            (cases.HCC, 0, 1, 2)):
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(rs.make_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())
        length = 50.0 if model == cases.RIPF else 1.0
        conn, xyz = cases.mesh(cases.TET4, n, distort=0.2, length=length)
        p, u0, ef, nf = cases.case(model, conn, xyz, "full")
        gpu = cases.gpu_system(model, cases.TET4, conn, xyz, p, u0, ef, nf, device=local, rank=rank, nranks=world,
                               partitioner=partitioner, unique_id=uid)
        gpu.ksp = ksp
        st = gpu.stats()
        if not (st.p2p_on and st.p2p_fused) and os.environ.get("RDC_P2P", "1") != "0":
            failures.append(f"{cases.NAMES[model]}: the NVLink peer-memory path is not active on rank {rank} "
                            f"(p2p_on={st.p2p_on}, fused={st.p2p_fused}): the test would only exercise the NCCL fallback")
        orc = cases.oracle_problem(model, cases.TET4, conn, xyz, p, u0, ef, nf) if rank == 0 else None
        dt = cases.DT[model]
        # operator rows owned by this rank against the oracle (every rank checks its own rows)
        gpu.rotate()
        gpu.assemble(dt, dt)
        rows, rowptr, col, val, rhs = gpu.download_csr()
        gpu.time = 0.0  # assemble(time, dt) set system.time; the time loop below starts again from t = 0
        o2 = cases.oracle_problem(model, cases.TET4, conn, xyz, p, u0, ef, nf)
        o2.u_old = o2.u.copy()
        val_o, rhs_o = o2.assemble(dt, dt)
        for k, r in enumerate(rows):
            a, b = o2.rowptr[r], o2.rowptr[r + 1]
            if not np.array_equal(col[rowptr[k]:rowptr[k + 1]], o2.col[a:b]):
                failures.append(f"{cases.NAMES[model]}: pattern of row {r}")
                break
        ref = np.concatenate([val_o[o2.rowptr[r]:o2.rowptr[r + 1]] for r in rows])
        tol = 1e-12 * np.maximum(np.abs(ref), 1e-3 * np.abs(val_o).max())   # cases.csr_tolerance, row subset
        if (np.abs(val - ref) > tol).any():
            failures.append(f"{cases.NAMES[model]}: K rows differ on rank {rank}: worst {(np.abs(val - ref) / tol).max():.2e} x tol")
        if np.abs(rhs - rhs_o[rows]).max() > 1e-12 * np.abs(rhs_o).max():
            failures.append(f"{cases.NAMES[model]}: F differs on rank {rank}")
        trace = []
        for _ in range(nsteps):
            its, res = gpu.step(dt)
            if rank == 0:
                oits, ores = orc.step(dt, pc=O.PC_ILU)
                trace.append((its, oits, float(np.linalg.norm(gpu.get_solution() - orc.u) / np.linalg.norm(orc.u))))
            else:
                gpu.get_solution()
        if rank == 0:
            print("   trace (gpu its, oracle its, rel err):", trace, flush=True)
        u = gpu.get_solution()
        # save_solution reductions: every element is counted by exactly one rank, the result is the same on all ranks
        nreg = 5
        region = np.random.default_rng(11).integers(0, nreg, conn.shape[0]).astype(np.int32)
        gpu.set_subdomains(region, nreg)
        nv = cases.P.NVARS[model]
        w = [0.0] * nv
        w[1] = 1.0
        conds = [(w, 1.0, float(np.quantile(u.reshape(-1, nv)[:, 1], 0.4)), 1e300)]
        vol_ref = O.region_volumes(cases.TET4, conn, xyz, u, conds, region, nreg)
        mean_ref = O.region_last_mean(cases.TET4, conn, xyz, u, 1, region, nreg)
        vol = gpu.region_volumes(conds)
        mean = gpu.region_last_mean(1)
        if np.abs(vol - vol_ref).max() > 1e-12 * max(np.abs(vol_ref).max(), 1e-300) or \
                np.abs(mean - mean_ref).max() > 1e-12 * max(np.abs(mean_ref).max(), 1e-300):
            failures.append(f"{cases.NAMES[model]}: region reductions differ on rank {rank}: {vol} {vol_ref} {mean} {mean_ref}")
        # owned-only download: the ranks' pieces are disjoint and together give the gathered vector; the pinned
        # buffer takes the zero-copy path, the pageable one the staged path
        pinned = torch.full((gpu.n_dofs,), float("nan"), dtype=torch.float64).pin_memory()
        for target in (pinned.numpy(), np.full(gpu.n_dofs, np.nan)):
            gpu.get_solution_owned(target)
            mine = ~np.isnan(target)
            cnt = torch.tensor([int(mine.sum())], device="cuda")
            dist.all_reduce(cnt)
            if int(cnt.item()) != gpu.n_dofs or not np.array_equal(target[mine], u[mine]):
                failures.append(f"{cases.NAMES[model]}: owned download differs on rank {rank}")
        if model != cases.RIPF:  # RIPF keeps TD / prev state across steps and re-primes on upload: no replay there
            # upload from a pinned buffer (zero-copy gather of the local entries) == upload from pageable memory
            pinned.copy_(torch.from_numpy(u))
            gpu.set_solution(pinned.numpy())
            if not np.array_equal(gpu.get_solution(), u):
                failures.append(f"{cases.NAMES[model]}: pinned upload differs on rank {rank}")
            # the fused and the stand-alone peer-memory exchanges must produce the same step (fixed sum order in both)
            its_a, _ = gpu.step(dt)
            ua = gpu.get_solution()
            gpu.set_solution(u)
            gpu.time -= dt
            gpu.set_option("p2p_fused_ar", 0)
            gpu.set_option("p2p_fused_halo", 0)
            its_b, _ = gpu.step(dt)
            ub = gpu.get_solution()
            if its_a != its_b or np.linalg.norm(ua - ub) > 1e-12 * np.linalg.norm(ua):
                failures.append(f"{cases.NAMES[model]}: fused vs unfused exchange differ on rank {rank}: {its_a} {its_b} "
                                f"{np.linalg.norm(ua - ub) / np.linalg.norm(ua):.2e}")
        if rank == 0:
            rel = np.linalg.norm(u - orc.u) / np.linalg.norm(orc.u)
            st = gpu.stats()
            print(f"{cases.NAMES[model]} x{world}: rel L2 vs oracle after {nsteps} steps {rel:.2e}; owned {st.n_nodes_local} "
                  f"ghost {st.n_nodes_ghost}", flush=True)
            # north_star tolerances: every single step 1e-8, the state after N steps 1e-6.  RIPF switches source terms on the
            # sign of a time derivative that is round-off noise where nothing changes (ripf.C:491-496), so two converged
            # Krylov solvers may take different branches there after the first step: 2e-8 on the 8-rank mesh
            if not trace[0][2] <= 1e-8:
                failures.append(f"{cases.NAMES[model]}: first-step solution rel L2 {trace[0][2]:.3e}")
            if not rel <= (1e-6 if model == cases.RIPF else 1e-8):
                failures.append(f"{cases.NAMES[model]}: solution rel L2 {rel:.3e}")
        gpu.close()
    # ---- solid mechanics, distributed: owned Jacobian rows, Newton load steps, post-processing ----------
    import solid_cases as SC
    from oracle import solid as S
    from rdcfes_b200 import solid as G

    def fresh_uid():
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(rs.make_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    c = SC.general_case(SC.TET4, n=n)
    x = SC.perturbed(c, amp=0.03 / n)
    g = G.from_case(c, device=local, rank=rank, nranks=world, partitioner=0, unique_id=fresh_uid())
    g.set_positions(x)
    g.assemble(0.3)
    rows, rowptr, col, val, rhs = g.download_csr()
    so = S.OracleSolid(c)
    val_o, rhs_o = so.assemble(x, 0.3)
    ref = np.concatenate([val_o[so.rowptr[r]:so.rowptr[r + 1]] for r in rows])
    cols_ok = all(np.array_equal(col[rowptr[k]:rowptr[k + 1]], so.col[so.rowptr[r]:so.rowptr[r + 1]]) for k, r in enumerate(rows))
    if not cols_ok or np.abs(val - ref).max() > 1e-12 * np.abs(val_o).max() or np.abs(rhs - rhs_o[rows]).max() > 1e-12 * np.abs(rhs_o).max():
        failures.append(f"solid: Jacobian / residual rows differ on rank {rank}")
    p, v, f = g.post_process(0.3)
    po, vo, fo = so.post(x, 0.3)
    if np.abs(p - po).max() > 1e-10 * (np.abs(po).max() + vo.max()) or np.abs(v - vo).max() > 1e-10 * (np.abs(po).max() + vo.max()) or \
            np.abs(f - fo).max() > 1e-12:
        failures.append(f"solid: post-processing differs on rank {rank}")
    g.close()
    c = SC.compression_case(SC.TET4, n=n, penalty=1.0e6)
    c.opts.update(max_nonlinear_iterations=25, relative_step_tolerance=1e-11, relative_residual_tolerance=1e-13,
                  absolute_residual_tolerance=1e-9, initial_linear_tolerance=1e-10)
    for ksp in (G.KSP_GMRES, G.KSP_BICGSTAB):
        g = G.from_case(c, device=local, rank=rank, nranks=world, partitioner=1 if ksp else 0, unique_id=fresh_uid())
        g.ksp = ksp
        so = S.OracleSolid(c)
        xo = c.xund.copy().ravel()
        for step in (1, 2):
            info = g.run_solver(0.1 * step)
            xg = g.get_positions().ravel()
            if rank == 0:
                xo, io = so.newton(xo, 0.1 * step)
                rel = float(np.linalg.norm(xg - xo) / np.linalg.norm(xo))
                print(f"solid x{world} ksp {ksp} load step {step}: {info}  rel L2 vs oracle {rel:.2e}", flush=True)
                if not (info["converged"] and io["converged"] and rel <= 1e-8):
                    failures.append(f"solid: load step {step} (ksp {ksp}): {info} {io} rel {rel:.3e}")
        g.close()
    flag = torch.tensor([len(failures)], device="cuda")
    dist.all_reduce(flag)
    if failures:
        print(f"rank {rank} FAILURES: {failures}", flush=True)
    dist.destroy_process_group()
    sys.exit(1 if flag.item() else 0)


if __name__ == "__main__":
    main()
