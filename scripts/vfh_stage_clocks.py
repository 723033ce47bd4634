"""Debug: per-stage clock cycles of vfh_update_kernel (library built with -DVFH_STAGE_CLOCKS)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from ros_navigation_b200 import capi
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
stream = torch.cuda.Stream(dev)
with torch.cuda.stream(stream):
    arm = bench.GpuArm("c4", 0, 1024, dev, stream, 1)
    for k in range(12):
        arm.step_dev(k)
    stream.synchronize()
    L = capi.lib()
    out = (C.c_ulonglong * 8)()
    L.b200nav_vfh_debug_stage_clocks(out, 1)
    n = 20
    for k in range(n):
        arm.step_dev(12 + k)
    stream.synchronize()
    L.b200nav_vfh_debug_stage_clocks(out, 0)
    names = ["submap info", "window->ranges", "stage M", "stages H,B,K", "mask", "stage S"]
    tot = sum(out[:6])
    for nm, v in zip(names, out[:6]):
        print("%-16s %8.0f cycles/block  %5.1f%%" % (nm, v / (n * 1024), 100.0 * v / tot))
    print("sum %.0f cycles = %.1f us at 1.965 GHz" % (tot / (n * 1024), tot / (n * 1024) / 1965.0))
