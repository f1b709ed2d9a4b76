"""world_size-2 gloo test of the multi-GPU host logic (frame partition + counter all-gather)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointgnn_b200.utils import sharding


def test_round_robin_partition_is_disjoint_and_complete():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 64):
            owned = [sharding.frames_for_rank(n, r, world) for r in range(world)]
            flat = sorted(i for o in owned for i in o)
            assert flat == list(range(n))
            assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    seeds = {sharding.frame_seed(s, f, r, 8) for r in range(8) for s in range(6) for f in range(8)}
    assert len(seeds) == 8 * 6 * 8          # rank-disjoint synthetic frames


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        mine = sharding.frames_for_rank(13, rank, world)
        counters = {'frames': len(mine), 'device_ms': 100.0 * (rank + 1), 'e2e_ms': 150.0 * (rank + 1),
                    'edges0': 10 * len(mine), 'edges1': 20 * len(mine), 'keypoints': 3 * len(mine)}
        per_rank, summary = sharding.gather_counters(counters)
        if rank == 0:
            out.put((per_rank.tolist(), summary, sharding.throughput(summary)))
    finally:
        dist.destroy_process_group()


def test_counter_all_gather_gloo_world2():
    world = 2
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    per_rank, summary, fps = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(per_rank) == 2
    assert summary['frames'] == 13                      # 7 + 6
    assert summary['device_ms'] == 200.0                # max over ranks
    assert summary['e2e_ms'] == 300.0
    assert summary['edges1'] == 20 * 13
    assert abs(fps - 13 / 0.2) < 1e-9


def test_gather_without_process_group():
    per_rank, summary = sharding.gather_counters({'frames': 4, 'device_ms': 8.0})
    assert per_rank.shape == (1, len(sharding.COUNTER_NAMES))
    assert sharding.throughput(summary) == 500.0
