"""GPU box: cost of the per-kernel CUDA events (b200nav_ctx_profile_*) inside bench.py's timed steps.
Runs the C4 device-resident cycle with and without kernel profiling, same L2 flush, per-step events."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
stream = torch.cuda.Stream(device=dev)
with torch.cuda.stream(stream):
    arm = bench.GpuArm("c4", 0, 1024, dev, stream, 1, 1024)
    bench.flush_l2(arm)
    for w in range(50):
        arm.step_dev(w, last=True)
    stream.synchronize()
    for rep in range(3):
        for prof in (False, True):
            arm.ctx.profile_enable(prof)
            ms = bench.timed_steps(torch, stream, arm.step_dev, 8, 40, arm)
            arm.ctx.profile_enable(False)
            print("kernel events %-5s step %.4f ms (median %.4f)" % (prof, float(np.mean(ms)), float(np.median(ms))), flush=True)
