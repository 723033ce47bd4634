"""GPU (needs >= 2 devices, skipped otherwise): the library's own NCCL binding for the per-cycle command all-gather
(b200nav_fleet_*, used by dist.CommandExchange when it is given a context) against the plain torch.distributed path."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, %r)
from ros_navigation_b200 import capi
from ros_navigation_b200.dist import CommandExchange
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(dev)
with torch.cuda.stream(stream):
    ctx = capi.Context(rank, stream=stream.cuda_stream)
    total = 64 * world
    a = CommandExchange(total, dev, ctx=ctx)      # library NCCL binding
    b = CommandExchange(total, dev)               # torch.distributed
    assert a.fleet is not None and b.fleet is None
    for cycle in range(6):
        slot = cycle & 1
        val = torch.arange(a.n_local * 16, device=dev, dtype=torch.int64).reshape(a.n_local, 16)
        val = ((val * 7 + rank * 31 + cycle * 5) %% 251).to(torch.uint8)
        a.wait(slot); b.wait(slot)
        a.locals[slot].copy_(val); b.locals[slot].copy_(val)
        a.gather_async(slot); b.gather_async(slot)
    a.wait(); b.wait()
    stream.synchronize()
    for slot in (0, 1):
        assert torch.equal(a.tables[slot], b.tables[slot]), slot
        lo, hi = a.lo, a.hi
        assert torch.equal(a.tables[slot][lo:hi], a.locals[slot])
    assert torch.equal(a.gather(), a.tables[0])
    a.close()
    # fused exchange: the VFH+ kernel pushes the commands into every rank's table (NVLink peer mappings)
    import numpy as np
    from ros_navigation_b200 import VFH, DeviceGridMap
    n_local = 24
    rng = np.random.default_rng(100 + rank)
    push = CommandExchange(n_local * world, dev, ctx=ctx, peer_push=True)
    plain = CommandExchange(n_local * world, dev, ctx=ctx)
    assert push.push and not plain.push
    grid = DeviceGridMap(ctx, (6.4, 6.4), 0.05, n_robots=n_local, layers=("master",))
    lay = np.full((grid.cols, grid.rows), np.nan, np.float32)
    for r in range(n_local):
        m = rng.random(lay.shape)
        l2 = lay.copy(); l2[m < 0.5] = 0.0; l2[m > 0.97] = 90.0
        grid.upload("master", l2, robot=r)
    va, vb = VFH(ctx, n_robots=n_local), VFH(ctx, n_robots=n_local)
    for cycle in range(7):
        slot = cycle & 1
        inp = np.zeros(n_local, capi.VFH_INPUT_DTYPE)
        inp["x"], inp["y"] = rng.uniform(-2, 2, n_local), rng.uniform(-2, 2, n_local)
        inp["yaw"], inp["dt"], inp["current_speed"] = rng.uniform(-3, 3, n_local), 0.2, 100
        inp["goal_direction"], inp["goal_distance"], inp["goal_tolerance"] = rng.uniform(0, 180, n_local), 2500.0, 250.0
        d_in = torch.from_numpy(inp.view(np.uint8).reshape(n_local, -1)).to(dev)
        push.vfh_update_push(va, grid, "master", d_in, slot)
        plain.wait(slot)
        vb.update_batched_dev(grid, "master", d_in, plain.locals[slot])
        plain.gather_async(slot)
        push.wait(slot); plain.wait(slot)
        stream.synchronize()
        assert torch.equal(push.tables[slot], plain.tables[slot]), (cycle, rank)
    push.status()
    # Rank skew (ADVICE r1): rank 1 is slow to READ each table (a long spin kernel sits between its wait and its read),
    # rank 0 runs ahead as fast as it can.  Every cycle's decision carries the cycle number (empty map: the picked
    # angle is the goal direction), so a table that a faster writer tore shows rows of a later cycle.
    empty = DeviceGridMap(ctx, (6.4, 6.4), 0.05, n_robots=n_local, layers=("master",))
    vs = VFH(ctx, n_robots=n_local)
    snaps = []
    for cycle in range(12):
        slot = cycle & 1
        inp = np.zeros(n_local, capi.VFH_INPUT_DTYPE)
        inp["dt"], inp["goal_direction"], inp["goal_distance"], inp["goal_tolerance"] = 0.2, 10.0 + cycle, 2500.0, 250.0
        d_in = torch.from_numpy(inp.view(np.uint8).reshape(n_local, -1)).to(dev)
        push.vfh_update_push(vs, empty, "master", d_in, slot)
        push.wait(slot)
        if rank == 1:
            torch.cuda._sleep(int(2e8))          # ~0.1 s on the context's stream before the table is read
        snaps.append(push.tables[slot].clone())  # the read, stream-ordered after the wait (and the sleep)
        push.release(slot)
    stream.synchronize()
    for cycle, snap in enumerate(snaps):
        got = snap.cpu().numpy().view(capi.COMMAND_DTYPE).reshape(-1)["picked_angle"]
        assert np.all(got == np.float32(10.0 + cycle)), (rank, cycle, np.unique(got))
    push.status()
    # b200nav_fleet_cycle_async: inputs from pinned host memory, VFH+ kernel, all-gather and the table back in pinned
    # host memory, enqueue-only (what bench.py's end-to-end loop uses at N > 1) - against the explicit path
    from ros_navigation_b200.capi import check, lib
    vc, vd = VFH(ctx, n_robots=n_local), VFH(ctx, n_robots=n_local)
    cyc = CommandExchange(n_local * world, dev, ctx=ctx)
    h_tab = [torch.zeros(n_local * world, 16, dtype=torch.uint8).pin_memory() for _ in range(2)]
    wants, h_ins = [], []
    for cycle in range(6):
        slot = cycle & 1
        inp = np.zeros(n_local, capi.VFH_INPUT_DTYPE)
        inp["x"], inp["y"] = rng.uniform(-2, 2, n_local), rng.uniform(-2, 2, n_local)
        inp["yaw"], inp["dt"], inp["current_speed"] = rng.uniform(-3, 3, n_local), 0.2, 50 * (cycle & 1)
        inp["goal_direction"], inp["goal_distance"], inp["goal_tolerance"] = rng.uniform(0, 180, n_local), 2500.0, 250.0
        h_in = torch.from_numpy(inp.view(np.uint8).reshape(n_local, -1)).pin_memory()
        h_ins.append(h_in)
        if cycle >= 2:
            check(lib().b200nav_fleet_cycle_wait(cyc.fleet, slot), ctx.h)
            assert torch.equal(h_tab[slot], wants[cycle - 2]), (rank, cycle - 2)
        check(lib().b200nav_fleet_cycle_async(cyc.fleet, vc.h, grid.h, b"master", h_in.data_ptr(), slot,
                                              h_tab[slot].data_ptr()), ctx.h)
        plain.wait(slot)
        vd.update_batched_dev(grid, "master", h_in.to(dev), plain.locals[slot])
        plain.gather_async(slot)
        plain.wait(slot)
        stream.synchronize()
        wants.append(plain.tables[slot].cpu().clone())
    for cycle in (4, 5):
        check(lib().b200nav_fleet_cycle_wait(cyc.fleet, cycle & 1), ctx.h)
        assert torch.equal(h_tab[cycle & 1], wants[cycle]), (rank, cycle)
    cyc.close()
    push.close(); plain.close()
dist.barrier()
dist.destroy_process_group()
print("FLEET_OK", rank)
"""


def test_fleet_allgather_matches_torch_distributed(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611", str(script)],
                         cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.count("FLEET_OK") == 2, out.stdout[-2000:] + out.stderr[-3000:]
