"""CPU tests of the host-side slab planning (dml_slab_plan) incl. a world_size-2 gloo run: every rank derives the same
cuts from the same data, the slabs partition the particles, counts are balanced, and the halo sets are mutually consistent
(what rank k sends up is exactly what rank k+1 expects as its lower ghosts)."""
import os
import sys
import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _plan_and_select(rank, world, z, w):
    from din_mol_li_b200 import dml
    cuts = dml.slab_plan(z, world, -1.0, z.max() + 10.0)
    lo, hi = cuts[rank], cuts[rank + 1]
    own = np.flatnonzero((z >= lo) & (z < hi))
    send_hi = own[z[own] >= hi - w] if rank < world - 1 else own[:0]
    send_lo = own[z[own] < lo + w] if rank > 0 else own[:0]
    return cuts, own, send_lo, send_hi


def test_plan_partitions_and_balances():
    rng = np.random.default_rng(3)
    z = rng.uniform(0, 500, 20001)
    parts = [_plan_and_select(r, 4, z, 13.2) for r in range(4)]
    allown = np.concatenate([p[1] for p in parts])
    assert len(allown) == len(z) and len(np.unique(allown)) == len(z)
    counts = [len(p[1]) for p in parts]
    assert max(counts) - min(counts) <= 2
    for r in range(3):          # what r sends up lies within w below the shared face; what r+1 sends down within w above it
        cut = parts[r][0][r + 1]
        assert (z[parts[r][3]] >= cut - 13.2).all() and (z[parts[r][3]] < cut).all()
        assert (z[parts[r + 1][2]] < cut + 13.2).all() and (z[parts[r + 1][2]] >= cut).all()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = np.random.default_rng(11).uniform(0, 300, 5000)
    cuts, own, slo, shi = _plan_and_select(rank, world, z, 13.2)
    import torch
    t = torch.tensor([len(own), len(slo), len(shi)], dtype=torch.int64)
    gathered = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, t)
    q.put((rank, cuts.tolist(), [g.tolist() for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def test_plan_consistent_across_ranks_gloo():
    world, port = 2, 29741
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][1] == out[1][1]                                    # identical cuts on every rank
    g = out[0][2]
    assert g[0][0] + g[1][0] == 5000 and abs(g[0][0] - g[1][0]) <= 2  # partition, balanced
    assert g[0][1] == 0 and g[1][2] == 0 and g[0][2] > 0 and g[1][1] > 0   # open ends have no face, the shared face has both
