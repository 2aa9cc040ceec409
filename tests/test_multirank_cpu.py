"""world_size-2 gloo tests (CPU) of the multi-rank host logic of bench.py: ensemble sharding gives every rank its own
replica and the whole-job figure is SUM(units) / MAX(time) (replicas shard with no data-path collective)."""
import os
import sys
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    ms, units = bench.aggregate_over_ranks(10.0 + 5.0 * rank, 1000.0 * (rank + 1), world)
    seeds = bench.replica_seed(rank)
    q.put((rank, ms, units, seeds))
    dist.barrier()
    dist.destroy_process_group()


def test_aggregate_is_max_time_sum_units_gloo():
    world, port = 2, 29731
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(o[1] == 15.0 and o[2] == 3000.0 for o in out)          # MAX over ranks, SUM over ranks
    assert out[0][3] != out[1][3]                                     # distinct replicas per rank


def test_single_rank_passthrough():
    import bench
    assert bench.aggregate_over_ranks(3.5, 42, 1) == (3.5, 42.0)
