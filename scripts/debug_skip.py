import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from ros_navigation_b200 import capi
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
with torch.cuda.stream(stream):
    arm = bench.GpuArm("c4", 0, 256, dev, stream, 1)
    out = np.zeros(2, np.int64)
    for rep in range(6):
        for c in range(8):
            arm.step_himm_only(c)
        capi.lib().b200nav_himm_debug_tile_stats(arm.grid.h, out.ctypes.data)
        print("rep", rep, "skipped", out[0], "processed", out[1])
    lay = arm.grid.download("laser", robot=3)
    print("nan frac", np.isnan(lay).mean(), "zero frac", (lay == 0).mean(), "marked", (lay > 0).mean())
