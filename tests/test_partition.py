"""Slab partition (host logic of the multi-GPU path) incl. a world_size-2 gloo run on CPU."""
import os
import socket

import numpy as np
import pytest

from thermalporous_b200 import cases as CS, geo as G
from thermalporous_b200.partition import Slab, slab_range
from thermalporous_b200.physicalparameters import PhysicalParameters


def test_slab_ranges_cover_and_are_disjoint():
    for nl, world in ((85, 1), (85, 2), (85, 4), (85, 8), (7, 3), (16, 16)):
        r = [slab_range(nl, world, k) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == nl
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        prm = PhysicalParameters()
        Slab(G.HomogeneousBoxGeo(2, 2, 2, prm, 1.0, 1.0, 1.0), 4, 3)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    prm = PhysicalParameters()
    prm.S_o = 0.9
    geo = G.SPE10Model3D(6, 8, 10, prm, fields=G.spe10_synthetic(6, 8, 10, seed=3))
    case = CS.WellCase(prm, geo, well_case="default")
    ent = CS.source_entries(case, prm, geo)
    slab = Slab(geo, world, rank)
    loc = slab.localize_sources(ent)
    # every source lands on exactly one rank; field pieces reassemble the global field
    cnt = torch.tensor([float(len(loc))])
    dist.all_reduce(cnt)
    pieces = [None] * world
    dist.all_gather_object(pieces, slab.take(geo.K_z))
    # what rank 0 would broadcast before tpb_comm_init: a 128-byte id
    uid = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    # neighbour planes (what tpb_exchange_static moves over NCCL), here over gloo send/recv
    mine = torch.from_numpy(slab.take(geo.phi))
    lo = torch.zeros(slab.np, dtype=torch.float64)
    if rank == 0:
        dist.send(mine[-slab.np:].contiguous(), dst=1)
    else:
        dist.recv(lo, src=0)
    # the saturation bounds the time loop acts on are global: bench.py's host-side helper reduces them over the slabs
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import NpOps
    sat = np.stack([np.zeros(slab.ncell), np.zeros(slab.ncell), np.linspace(0.2 + rank, 0.5 + rank, slab.ncell)])
    smin, smax = NpOps(dist, None).minmax(sat, 2)
    # the per-step well totals (thermalmodel.py:231-270) are sums over every rank's source cells
    from thermalporous_b200.model import ThermalModel

    class _M:
        pass
    rng = np.random.default_rng(11)
    ug = np.stack([prm.p_ref + rng.uniform(-0.5, 0.5, geo.ncell), rng.uniform(290.0, 330.0, geo.ncell),
                   rng.uniform(0.3, 0.9, geo.ncell)])

    def totals(entries, u, kx, ky, world_):
        m = _M()
        m._entries, m.u, m.params, m.nphase, m.world = entries, torch.from_numpy(np.ascontiguousarray(u)), prm, 2, world_
        cells = np.array([e_[0] for e_ in entries], dtype=np.int64)
        m._src_K = (kx[cells], ky[cells])
        return ThermalModel.well_totals(m)
    t_loc = totals(loc, slab.take(ug), slab.take(geo.K_x), slab.take(geo.K_y), world)
    t_glob = totals(ent, ug, geo.K_x, geo.K_y, 1)
    tot_ok = all((t_glob[k] is None and t_loc[k] is None) or abs(t_loc[k] - t_glob[k]) <= 1e-12 * abs(t_glob[k])
                 for k in t_glob)
    ok = (cnt.item() == len(ent) and np.array_equal(np.concatenate(pieces), geo.K_z) and uid[0] == bytes(range(128))
          and abs(smin - 0.2) < 1e-15 and abs(smax - 1.5) < 1e-15 and tot_ok and t_glob["prod"] is not None
          and slab.local_dims() == (6, 8, 5) and slab.has_lo == (rank == 1) and slab.has_hi == (rank == 0)
          and all(0 <= e[0] < slab.ncell for e in loc)
          and (rank == 0 or np.array_equal(lo.numpy(), geo.phi[slab.c0 - slab.np:slab.c0])))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_partition_over_gloo():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
