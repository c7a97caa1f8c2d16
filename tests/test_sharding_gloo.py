"""Multi-rank host logic on CPU: world_size 2 over gloo (SURVEY.md section 8e -- shards, no data collective)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partitions():
    from ros_vision_b200 import sharding
    for n in (0, 1, 7, 16, 129):
        for g in (1, 2, 4, 8):
            assert sharding.partition_is_valid(n, g)
    assert sharding.frames_for_rank(10, 4, 1) == [1, 5, 9]
    assert sharding.streams_for_rank(8, 8, 3) == [3]
    with pytest.raises(ValueError):
        sharding.frames_for_rank(4, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import pyoracle
    from ros_vision_b200 import sharding, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nframes = 5
    mine = sharding.frames_for_rank(nframes, world, rank)
    cfg = pyoracle.make_config(320, 240, "gray", 2, 0.0)
    local = {}
    for f in mine:  # the per-rank detector is stood in for by the CPU oracle (no GPU in this test)
        sc = synth.make_scene(320, 240, 700 + f, 2, side_range=(50, 80), noise_sigma=2.0, ids=[f, f + 10])
        r = pyoracle.detect(cfg, sc.gray)
        local[f] = [int(i) for i in r.detections["id"]]
    merged = sharding.gather_detections(local)
    slow = sharding.max_over_ranks(0.5 + rank)
    dist.barrier()
    if rank == 0:
        q.put((merged, slow))
    dist.destroy_process_group()


def test_two_rank_gather_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    merged, slow = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(merged) == [0, 1, 2, 3, 4]
    for f, ids in merged.items():
        assert ids == [f, f + 10]
    assert slow == 1.5
