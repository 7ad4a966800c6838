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


# ---- host-only set-up logic probed without a device ------------------------------------------------------------
def _probe_tiles(rowptr, max_rows=16, max_blocks=256):
    import ctypes as C
    from rdcfes_b200 import lib
    L = lib.load()
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    n, tiles = C.c_int32(), C.c_void_p()
    rc = L.rdc_probe_spmv_tiles(C.c_int32(rowptr.size - 1), rowptr.ctypes.data_as(C.c_void_p), C.c_int(max_rows),
                                C.c_int(max_blocks), C.byref(n), C.byref(tiles))
    assert rc == 0
    out = None
    if n.value >= 0:
        out = np.ctypeslib.as_array(C.cast(tiles, C.POINTER(C.c_int32)), shape=(max(n.value, 1) * 4,))[: n.value * 4].copy()
        out = out.reshape(-1, 4)
    L.rdc_free(tiles)
    return out


def test_spmv_tile_cutter_invariants():
    rng = np.random.default_rng(0)
    for trial in range(20):
        lens = rng.integers(1, 60, size=rng.integers(1, 400))
        if trial % 4 == 0:
            lens[rng.integers(0, lens.size)] = 256          # a row that fills a whole tile
        rowptr = np.concatenate([[0], np.cumsum(lens)])
        t = _probe_tiles(rowptr)
        assert t is not None
        assert t[0, 0] == 0 and np.array_equal(t[1:, 0], t[:-1, 0] + t[:-1, 1]) and t[-1, 0] + t[-1, 1] == lens.size
        assert (t[:, 1] >= 1).all() and (t[:, 1] <= 16).all() and (t[:, 3] <= 256).all()
        assert np.array_equal(t[:, 2], rowptr[t[:, 0]]) and np.array_equal(t[:, 3], rowptr[t[:, 0] + t[:, 1]] - rowptr[t[:, 0]])
        # greedy: a tile stops only because the next row would break a limit
        for k in range(t.shape[0] - 1):
            nxt = t[k, 0] + t[k, 1]
            assert t[k, 1] == 16 or t[k, 3] + lens[nxt] > 256
    assert _probe_tiles(np.array([0, 5, 300, 310])) is None     # one row longer than a tile: LDG kernel instead
    assert _probe_tiles(np.array([0])).shape == (0, 4)


def test_region_bucketing_invariants():
    import ctypes as C
    from rdcfes_b200 import lib
    L = lib.load()
    rng = np.random.default_rng(1)
    E, nreg, chunk = 5000, 7, 256
    counted = (rng.random(E) < 0.8).astype(np.uint8)
    region = rng.integers(0, nreg, E).astype(np.int32)
    region[region == 4] = 5                                   # an empty region
    n_counted, n_chunks = C.c_int64(), C.c_int32()
    perm, cptr, rptr = C.c_void_p(), C.c_void_p(), C.c_void_p()
    rc = L.rdc_probe_region_chunks(C.c_int64(E), counted.ctypes.data_as(C.c_void_p), region.ctypes.data_as(C.c_void_p),
                                   C.c_int(nreg), C.c_int(chunk), C.byref(n_counted), C.byref(perm), C.byref(n_chunks),
                                   C.byref(cptr), C.byref(rptr))
    assert rc == 0

    def take(p, n):
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_int32)), shape=(max(n, 1),))[:n].copy()
        L.rdc_free(p)
        return a

    perm, cptr, rptr = take(perm, n_counted.value), take(cptr, n_chunks.value + 1), take(rptr, nreg + 1)
    assert n_counted.value == counted.sum() and np.array_equal(np.sort(perm), np.nonzero(counted)[0])
    assert np.all(np.diff(region[perm]) >= 0)                                    # bucketed by region ...
    for r in range(nreg):
        mine = perm[region[perm] == r]
        assert np.all(np.diff(mine) > 0)                                         # ... element order kept inside
    assert cptr[0] == 0 and cptr[-1] == perm.size and np.all(np.diff(cptr) >= 1) and np.all(np.diff(cptr) <= chunk)
    assert rptr[0] == 0 and rptr[-1] == n_chunks.value and rptr[5] - rptr[4] == 0
    for r in range(nreg):                                                        # no chunk straddles two regions
        for c in range(rptr[r], rptr[r + 1]):
            assert np.all(region[perm[cptr[c]:cptr[c + 1]]] == r)


# ---- solid mechanics, N > 1: which rank adds which penalty terms (host logic of rdc_solid_set_bcs, probed without a device) --
def _solid_bc_worker(rank, world, port, partitioner, et, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import solid_cases as SC
        from oracle import solid as S
        from rdcfes_b200 import solid as G
        from rdcfes_b200 import system as rs
        c = SC.general_case(et, n=5)
        se, sn, sb, bd = c.arrays()
        rows = G.probe_bc_rows(et, c.conn, c.xund, rank, world, partitioner, se, sn)
        owner = rs.probe_partition(et, 3, c.conn, c.xund, rank, world, partitioner)["owner"]
        # every row belongs to a node this rank owns ...
        assert all(owner[n] == rank for n in rows)
        # ... and carries ALL incidences of that node in the boundary sides, in input order: the rank that owns a boundary node adds
        # every penalty term of its row itself (no cross-rank sum), in the same order on any number of ranks
        want = {}
        for k in range(se.shape[0]):
            for j, l in enumerate(S.SIDE_NODES[et][sn[k]]):
                want.setdefault(int(c.conn[se[k], l]), []).append((k, j))
        for n, ent in rows.items():
            assert ent == want[n], (n, ent, want[n])
        # the ranks' rows tile the set of boundary-condition nodes: each exactly once
        cnt = torch.zeros(c.N, dtype=torch.int64)
        cnt[list(rows)] = 1
        dist.all_reduce(cnt)
        expect = torch.zeros(c.N, dtype=torch.int64)
        expect[list(want)] = 1
        assert bool((cnt == expect).all())
        q.put((rank, "ok"))
    except Exception as exc:  # noqa: BLE001
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,partitioner,et", [(2, 0, 4), (2, 1, 8), (3, 0, 4)])
def test_solid_penalty_rows_gloo(world, partitioner, et):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_solid_bc_worker, args=(r, world, port, partitioner, et, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
