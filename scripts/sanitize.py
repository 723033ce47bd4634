"""Small mixed workload for compute-sanitizer (memcheck / racecheck): HIMM fan + general + foreign tiles, moves, VFH."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle as O
from ros_navigation_b200 import VFH, DeviceGridMap, capi
from tests.util import lidar_samples, random_samples, assert_layers_equal

ctx = capi.Context(0)
rng = np.random.default_rng(0)
g = O.make_geom(10.0, 6.5, 0.05)
dg = DeviceGridMap(ctx, (10.0, 6.5), 0.05, n_robots=2, layers=("laser",))
dg.alias("master", "laser")
lay = [O.new_layer(g), O.new_layer(g)]
lay[1][:] = (rng.random(lay[1].shape) * 300).astype(np.float32)
dg.upload("laser", lay[1], robot=1)
v = VFH(ctx, n_robots=2)
for it in range(3):
    per = [lidar_samples(rng, g, (0.3, 0.1), 700, 0.2, 4.0, clear_frac=0.1), random_samples(rng, g, 300)]
    off = np.array([0, len(per[0]), len(per[0]) + len(per[1])], np.int32)
    dg.himm_update_batched("laser", np.concatenate(per), off)
    for r in range(2):
        O.himm_update(g, lay[r], per[r])
    inp = np.zeros(2, capi.VFH_INPUT_DTYPE)
    inp["x"], inp["y"], inp["yaw"], inp["dt"] = [0.3, -1.0], [0.1, 0.5], [0.2, 2.0], 0.2
    inp["goal_direction"], inp["goal_distance"], inp["goal_tolerance"] = 90.0, 2000.0, 250.0
    v.update_batched(dg, "master", inp)
for r in range(2):
    assert_layers_equal(dg.download("laser", robot=r), lay[r], "robot %d" % r)
assert dg.layer_format("laser") == "float"

# the same on byte-coded layers (no foreign values): bulk-copy staged tile kernel, coded readers, move
cg = DeviceGridMap(ctx, (10.0, 6.5), 0.05, n_robots=2, layers=("laser",))
cg.alias("master", "laser")
clay = [O.new_layer(g), O.new_layer(g)]
cgeo = [O.make_geom(10.0, 6.5, 0.05), O.make_geom(10.0, 6.5, 0.05)]
for it in range(3):
    per = [lidar_samples(rng, g, (0.3, 0.1), 700, 0.2, 4.0, clear_frac=0.1), random_samples(rng, g, 300)]
    off = np.array([0, len(per[0]), len(per[0]) + len(per[1])], np.int32)
    cg.himm_update_batched("laser", np.concatenate(per), off)
    for r in range(2):
        O.himm_update(cgeo[r], clay[r], per[r])
    v.update_batched(cg, "master", inp)
    cg.move((0.4 * (it + 1), -0.2), robot=0)
    O.move(cgeo[0], [clay[0]], 0.4 * (it + 1), -0.2)
for r in range(2):
    assert_layers_equal(cg.download("laser", robot=r), clay[r], "coded robot %d" % r)
assert cg.layer_format("laser") == "coded"
print("sanitize workload OK")

# scan form (projection fused into the binning kernel, thinned scan with the ifClearEnd quirk) + asynchronous VFH+
sg = DeviceGridMap(ctx, (12.8, 12.8), 0.05, n_robots=3, layers=("laser",))
sg.alias("master", "laser")
info = sg.scan_info(-2.3, 0.0043, 0.1, 6.0, 1080, decimate=True)
for it in range(2):
    poses = rng.uniform(-2, 2, (3, 3))
    ranges = rng.uniform(0.05, 7.0, (3, 1080)).astype(np.float32)
    ranges[rng.random(ranges.shape) < 0.1] = np.inf
    sg.himm_update_scans_batched("laser", info, poses, ranges)
print("sanitize scan-form workload OK")
