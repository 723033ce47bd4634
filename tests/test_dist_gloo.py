"""CPU, world_size 2 over gloo: the multi-GPU host logic of the batched mode (robot partitioning and the per-cycle
all-gather of steering commands).  The compute itself needs a GPU and is covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ros_navigation_b200 import dist as D
from ros_navigation_b200.capi import COMMAND_DTYPE


def test_partition_covers_all_robots_once():
    for total in (1, 2, 5, 1024, 16384, 1000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                lo, hi = D.partition(total, r, world)
                assert 0 <= lo <= hi <= total
                seen += list(range(lo, hi))
                for robot in (lo, hi - 1):
                    if lo < hi:
                        assert D.owner_of(robot, total, world) == r
            assert seen == list(range(total))
            sizes = [D.partition(total, r, world)[1] - D.partition(total, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ex = D.CommandExchange(total, torch.device("cpu"))
        for cycle in range(3):
            cmd = np.zeros(ex.n_local, COMMAND_DTYPE)
            robots = np.arange(ex.lo, ex.hi)
            cmd["speed"] = robots * 10 + cycle
            cmd["turnrate"] = -robots
            cmd["picked_angle"] = robots * 0.5
            cmd["flags"] = rank
            ex.local.copy_(torch.from_numpy(cmd.view(np.uint8).reshape(ex.n_local, 16)))
            table = ex.gather().numpy().copy().view(COMMAND_DTYPE).reshape(-1)
            allr = np.arange(total)
            assert np.array_equal(table["speed"], allr * 10 + cycle)
            assert np.array_equal(table["turnrate"], -allr)
            assert np.array_equal(table["picked_angle"], (allr * 0.5).astype(np.float32))
            owners = np.array([D.owner_of(r, total, world) for r in allr])
            assert np.array_equal(table["flags"], owners)
        if ex.even:  # asynchronous, double-buffered form used by the pipelined batched mode
            for cycle in range(4):
                slot = cycle & 1
                ex.wait(slot)
                cmd = np.zeros(ex.n_local, COMMAND_DTYPE)
                cmd["speed"] = np.arange(ex.lo, ex.hi) + 1000 * cycle
                ex.locals[slot].copy_(torch.from_numpy(cmd.view(np.uint8).reshape(ex.n_local, 16)))
                ex.gather_async(slot)
                if cycle >= 1:  # read the previous cycle's table while this one is in flight
                    prev = 1 - slot
                    ex.wait(prev)
                    t = ex.tables[prev].numpy().copy().view(COMMAND_DTYPE).reshape(-1)
                    assert np.array_equal(t["speed"], np.arange(total) + 1000 * (cycle - 1))
            ex.wait()
            # verify(): the check bench.py runs at N > 1 - clean after a real gather, and it notices a damaged row,
            # a stale cycle and two swapped rows on ANY rank
            assert ex.verify(1) == 0 and ex.verify(0) == 0
            if rank == 1:
                ex.tables[1][0, 3] += 1
            assert ex.verify(1) == 1
            if rank == 1:
                ex.tables[1][0, 3] -= 1
                ex.tables[1][[0, 1]] = ex.tables[1][[1, 0]]
            assert ex.verify(1) == (1 if D.owner_of(0, total, world) == D.owner_of(1, total, world) else 2)
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5])
def test_command_allgather_world2_gloo(tmp_path, total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "ok%d" % r)) for r in range(world))
