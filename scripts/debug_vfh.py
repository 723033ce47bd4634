import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ros_navigation_b200 import VFH, DeviceGridMap, capi
ctx = capi.Context(0)
dg = DeviceGridMap(ctx, (10.0, 10.0), 0.05, layers=("master",))
lay = np.full((200, 200), np.nan, np.float32); lay[100:104, 90:95] = 90.0
dg.upload("master", lay)
v = VFH(ctx)
print("ranges path:", v.Update_VFH(np.full((361, 2), 5000.0), 0, 90.0, 3000.0, 250.0, dt=0.2))
cmd = v.update_from_grid(dg, "master", VFH.make_input(x=0.1, y=0.2, yaw=0.3))
print("grid path:", cmd, (v.ranges()[:, 0] < 5000).sum())
