"""world_size-2 (and 3) tests of the N>1 host logic on CPU with the gloo backend: node partition, owned/ghost
numbering and halo lists (rdc_probe_partition, the same build_setup code the GPU path runs), exercised by a
distributed SpMV + dot product in which every rank only ever touches its own rows and the ghost values it
receives through the halo exchange -- the communication pattern of the device Krylov solver."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, partitioner, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from rdcfes_b200 import system as rs
        conn, xyz = cases.mesh(cases.TET4, 6, distort=0.2)
        p, u0, ef, nf = cases.case(cases.ADPM, conn, xyz, "full")
        info = rs.probe_partition(cases.TET4, 3, conn, xyz, rank, world, partitioner)
        owner = info["owner"]
        N = xyz.shape[0]
        owned = np.flatnonzero(owner == rank)
        assert owned.size == info["n_owned"]
        # every node has exactly one owner and the owned sets tile the mesh
        cnt = torch.zeros(N, dtype=torch.int64)
        cnt[owned] = 1
        dist.all_reduce(cnt)
        assert bool((cnt == 1).all())
        # local elements = elements with an owned node; ghosts = their other nodes
        mine = (owner[conn] == rank).any(axis=1)
        assert int(mine.sum()) == info["n_elems_local"]
        ghosts = np.setdiff1d(np.unique(conn[mine]), owned)
        assert np.array_equal(np.sort(info["recv_glob"]), ghosts)
        # the operator (replicated on the CPU here; each rank uses only ITS rows)
        pr = O.Problem(O.ADPM, O.TET4, conn, xyz, p, u0, elem_field=ef)
        pr.assemble(0.05, 0.05)
        A = pr.scipy_csr()
        x_glob = np.random.default_rng(3).standard_normal(3 * N)
        # local vector: owned values are known, ghost values arrive through the halo exchange
        x_loc = {int(g): x_glob[3 * g:3 * g + 3].copy() for g in owned}
        reqs, bufs = [], []
        for k, qn in enumerate(info["nbr"]):
            sg = info["send_glob"][info["send_ptr"][k]:info["send_ptr"][k + 1]]
            assert (owner[sg] == rank).all()
            sb = torch.from_numpy(np.concatenate([x_glob[3 * g:3 * g + 3] for g in sg]) if sg.size else np.zeros(0))
            rg = info["recv_glob"][info["recv_ptr"][k]:info["recv_ptr"][k + 1]]
            assert (owner[rg] == qn).all()
            rb = torch.zeros(3 * rg.size, dtype=torch.float64)
            if sg.size:
                reqs.append(dist.isend(sb, int(qn)))
            if rg.size:
                reqs.append(dist.irecv(rb, int(qn)))
            bufs.append((rg, rb))
        for r in reqs:
            r.wait()
        for rg, rb in bufs:
            for i, g in enumerate(rg):
                x_loc[int(g)] = rb[3 * i:3 * i + 3].numpy()
        # y = A x on the owned rows, using ONLY locally available x
        xl = np.full(3 * N, np.nan)
        for g, v in x_loc.items():
            xl[3 * g:3 * g + 3] = v
        rows = (3 * owned[:, None] + np.arange(3)[None, :]).ravel()
        Ar = A[rows]
        assert not np.isnan(xl[np.unique(Ar.indices)]).any(), "a needed ghost value was not delivered"
        y_loc = Ar @ np.nan_to_num(xl)
        y_ref = (A @ x_glob)[rows]
        assert np.abs(y_loc - y_ref).max() <= 1e-13 * np.abs(y_ref).max()
        # distributed dot product = all-reduce of the owned partial sums
        d = torch.tensor([float(x_glob[rows] @ y_loc)], dtype=torch.float64)
        dist.all_reduce(d)
        assert abs(d.item() - float(x_glob @ (A @ x_glob))) <= 1e-10 * abs(d.item())
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,partitioner", [(2, 0), (2, 1), (3, 0)])
def test_halo_exchange_spmv_gloo(world, partitioner):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, partitioner, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
