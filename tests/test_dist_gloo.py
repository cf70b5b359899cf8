"""world_size-2 gloo tests (CPU) of the host-side logic of the N > 1 path: the strip partition and the
decomposition independence of bench.py's synthetic fields, the 128-byte communicator-id broadcast the
ranks perform before pop_comm_init, and the rank-ordered double-double combine of per-rank partial sums
(the host mirror of sum_ranks_kernel, csrc/pop_reduce.cu) against the exactly rounded sum."""
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def two_sum(a, b):
    s = a + b
    bb = s - a
    return s, (a - (s - bb)) + (b - bb)


def dd_add(x, y):          # csrc/pop_dev.cuh dd_add
    s, e = two_sum(x[0], y[0])
    e += x[1] + y[1]
    hi = s + e
    return hi, e - (hi - s)


def dd_sum(values):
    acc = (0.0, 0.0)
    for v in values:
        acc = dd_add(acc, (float(v), 0.0))
    return acc


def strip_rows(rank, world, ny):   # pop_init: ny_local = ny/P, j0 = rank*ny_local + 1
    nyl = ny // world
    return slice(rank * nyl, (rank + 1) * nyl)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        import bench
        nx, ny, km = 48, 32, 5
        dz = bench.syn.vert_grid("stretched", km)
        kmt = bench.syn.bathymetry(nx, ny, km, 7)
        kmu = bench.syn.kmu_from_kmt(kmt, ew_cyclic=True, ns_type=bench.c.BNDY_TRIPOLE)
        rows = strip_rows(rank, world, ny)
        F = bench.Fields(np, nx, ny, km, dz, kmt, kmu, rows)
        mine = {"T": F.tracer(0, 2, "cur"), "S": F.tracer(1, 3, "old"), "U": F.vel(0, 1, "cur"), "V": F.vel(1, 4, "old"),
                "P": F.psurf("old"), "VDC": F.vdc(1, 2), "VVC": F.vvc(3)}
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        G = bench.Fields(np, nx, ny, km, dz, kmt, kmu, slice(0, ny))
        glob = {"T": G.tracer(0, 2, "cur"), "S": G.tracer(1, 3, "old"), "U": G.vel(0, 1, "cur"), "V": G.vel(1, 4, "old"),
                "P": G.psurf("old"), "VDC": G.vdc(1, 2), "VVC": G.vvc(3)}
        for n in glob:
            assert np.array_equal(np.concatenate([p_[n] for p_ in parts], axis=0), glob[n]), n
        # communicator id: rank 0 creates 128 bytes, everyone ends up with the same bytes
        obj = [bytes(np.random.default_rng(5).integers(0, 256, 128, dtype=np.uint8)) if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        ids = [None] * world
        dist.all_gather_object(ids, obj[0])
        assert len(obj[0]) == 128 and all(i == ids[0] for i in ids)
        # rank-ordered dd combine of the per-rank masked sums
        field = glob["P"] * 1.0e3 + glob["T"] * 1.0e-9
        mask = (kmt > 0).astype(np.float64)
        local = dd_sum((field[rows] * mask[rows]).ravel())
        allp = [None] * world
        dist.all_gather_object(allp, local)
        acc = (0.0, 0.0)
        for r in range(world):
            acc = dd_add(acc, allp[r])
        total = acc[0] + acc[1]
        assert total == math.fsum((field * mask).ravel())
        # every rank computed the same total (no reduction whose order could differ)
        tots = [None] * world
        dist.all_gather_object(tots, total)
        assert all(t == tots[0] for t in tots)
        q.put((rank, "ok"))
    except BaseException as e:  # noqa: BLE001
        q.put((rank, repr(e)))
        raise
    finally:
        dist.destroy_process_group()


def test_two_rank_host_logic_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
