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
