"""CPU, world_size 2 over gloo: the replica sharding logic (which sequence a rank takes, whole-job throughput =
all frames / max-over-ranks time).  The per-frame path itself has no collective."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pf_loader import pfb
    from pfilter_noetic_b200 import shard
    name = shard.sequence_for_rank(rank, world)
    p = pfb.synth.config(name)
    # ranks get different sequences (different seeds -> different scenes)
    scan = pfb.synth.scan(p, 0)
    ms = 100.0 if rank == 0 else 250.0                 # rank 1 is the straggler
    sps, max_ms, total = shard.aggregate_throughput(10, ms)
    out.put((rank, name, int(p.seed), len(scan), sps, max_ms, total))
    dist.barrier()
    dist.destroy_process_group()


def test_replica_sharding_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, n0, s0, l0, sps0, ms0, t0), (r1, n1, s1, l1, sps1, ms1, t1) = res
    assert (n0, n1) == ("cfg5.0", "cfg5.1") and s0 != s1 and l0 > 50000 and l1 > 50000
    assert ms0 == ms1 == 250.0 and t0 == t1 == 20               # max over ranks, sum of frames
    assert abs(sps0 - 20 / 0.25) < 1e-9 and sps0 == sps1


def test_single_rank_passthrough():
    sys.path.insert(0, ROOT)
    from pf_loader import pfb  # noqa: F401
    from pfilter_noetic_b200 import shard
    assert shard.sequence_for_rank(0, 1) == "cfg5.0"          # the same sequence family at every world size
    # configs[4] as a fixed 8-sequence job: every sequence is run exactly once at every world size
    for world in (1, 2, 4, 8):
        owned = [shard.sequences_for_rank(r, world) for r in range(world)]
        assert sorted(sum(owned, [])) == [f"cfg5.{k}" for k in range(8)]
        assert all(len(o) == 8 // world for o in owned)
    sps, ms, total = shard.aggregate_throughput(100, 50.0)
    assert total == 100 and ms == 50.0 and abs(sps - 2000.0) < 1e-9
