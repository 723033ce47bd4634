"""GPU: the C++ drop-in classes (include/b200nav_shim.hpp) driven like the reference node, and the 'next' rows of
SURVEY section 8f that are already built: GridMap::move on the device and toOccupancyGrid."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_layers_equal, lidar_samples

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    from ros_navigation_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def test_cpp_shim_matches_reference(tmp_path):
    """Compiles tests/cpp/shim_check.cpp against the shim header + libb200nav.so and runs it."""
    if not O.have_ref():
        pytest.skip("oracle/_ref missing")
    from ros_navigation_b200 import capi
    O.lib()
    exe = str(tmp_path / "shim_check")
    cmd = ["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests/cpp/shim_check.cpp"),
           "-I" + os.path.join(ROOT, "include"), capi.LIB_PATH, O.LIB_PATH, O.REF_PATH,
           "-Wl,-rpath," + os.path.dirname(capi.LIB_PATH), "-Wl,-rpath," + os.path.dirname(O.LIB_PATH),
           "-Wl,-rpath," + os.path.dirname(O.REF_PATH), "-lpthread"]
    subprocess.run(cmd, check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith("OK"), out.stdout + out.stderr


def test_moving_map_scenario(ctx):
    """mapTest_vfh / mapTest_graph: a 4 m map that follows the robot (MapProvider::loopMoveMap,
    move_control/src/map_provider.cpp:177-188 -> GridMap::move, grid_map_core/src/GridMap.cpp:346-412), interleaved
    with HIMM updates and VFH+ window reads across the circular-buffer seam."""
    from ros_navigation_b200 import VFH, DeviceGridMap
    rng = np.random.default_rng(11)
    g = O.make_geom(4.0, 4.0, 0.05)
    dg = DeviceGridMap(ctx, (4.0, 4.0), 0.05, layers=("master", "laser"))
    laser, master = O.new_layer(g), O.new_layer(g)
    v = VFH(ctx)
    x = y = 0.0
    for step in range(40):
        x += rng.uniform(-0.12, 0.3)
        y += rng.uniform(-0.2, 0.15)
        if step % 2 == 0:
            moved_o = O.move(g, [laser, master], x, y)
            moved_d = dg.move((x, y))
            assert moved_o == moved_d
            (px, py), (s0, s1) = dg.get_geometry()
            assert (px, py, s0, s1) == (g.pos_x, g.pos_y, g.start0, g.start1)
        s = lidar_samples(rng, g, (x, y), 360, 0.2, 2.6, clear_frac=0.05)
        O.himm_update(g, laser, s)
        master[:] = laser
        dg.himm_update("laser", s)
        dg.copy_layer("master", "laser")
        yaw = rng.uniform(-np.pi, np.pi)
        want = O.ranges_from_submap(g, master, x, y, yaw)
        v.update_from_grid(dg, "master", VFH.make_input(x=x, y=y, yaw=yaw))
        assert np.array_equal(v.ranges()[:, 0], want[:, 0]), "step %d" % step
    assert_layers_equal(dg.download("laser"), laser, "laser after moves")
    assert_layers_equal(dg.download("master"), master, "master after moves")
    # large jump: the whole map is dropped
    assert O.move(g, [laser, master], x + 10.0, y - 7.0) == dg.move((x + 10.0, y - 7.0))
    assert_layers_equal(dg.download("laser"), laser, "after jump")
    v.close()
    dg.close()


def test_to_occupancy_grid(ctx):
    """GridMapRosConverter::toOccupancyGrid (grid_map_ros/src/GridMapRosConverter.cpp:251-287) as MapProvider
    publishes it (map_provider.cpp:207-214: layer master, 0..255)."""
    from ros_navigation_b200 import DeviceGridMap
    rng = np.random.default_rng(12)
    for start in [(0, 0), (13, 57)]:
        g = O.make_geom(6.5, 4.0, 0.05, 1.0, -2.0, start)
        dg = DeviceGridMap(ctx, (6.5, 4.0), 0.05, (1.0, -2.0), layers=("master",))
        dg.set_geometry(0, (1.0, -2.0), start)
        layer = O.new_layer(g)
        m = rng.random(layer.shape)
        layer[m < 0.5] = rng.choice(np.arange(0, 190, 10.0), (m < 0.5).sum())
        layer[(m >= 0.5) & (m < 0.55)] = -5.0
        layer[(m >= 0.55) & (m < 0.6)] = 300.0
        dg.upload("master", layer)
        assert np.array_equal(dg.to_occupancy("master", 0.0, 255.0), O.to_occupancy(g, layer, 0.0, 255.0))
        dg.close()


def test_query_blocked_matches_if_blocked(ctx):
    """MapGlobalPlanner::ifBlocked (map_global_planner.h:39-54) as the RRT planner uses it (rrt_planner.cpp:53)."""
    from ros_navigation_b200 import DeviceGridMap
    rng = np.random.default_rng(13)
    for geom, start in [((10.0, 10.0, 0.05, 0.0, 0.0), (0, 0)), ((4.0, 4.0, 0.05, 1.3, -0.7), (17, 63))]:
        g = O.make_geom(*geom, start)
        dg = DeviceGridMap(ctx, geom[:2], geom[2], geom[3:5], layers=("master",))
        dg.set_geometry(0, geom[3:5], start)
        layer = O.new_layer(g)
        m = rng.random(layer.shape)
        layer[m < 0.4] = 0.0
        layer[(m >= 0.4) & (m < 0.405)] = rng.choice([10.0, 90.0, 180.0], ((m >= 0.4) & (m < 0.405)).sum())
        dg.upload("master", layer)
        pts = np.stack([geom[3] + (rng.random(400) - 0.5) * geom[0] * 1.2, geom[4] + (rng.random(400) - 0.5) * geom[1] * 1.2], 1)
        for radius in (0.3, 0.07, 1.0):
            got = dg.query_blocked(pts, radius)
            want = np.array([O.if_blocked(g, layer, p[0], p[1], radius) for p in pts])
            assert np.array_equal(got, want), "radius %g: %d differ" % (radius, (got != want).sum())
            assert want.any() and not want.all()
        dg.close()


def test_error_paths(ctx):
    """Error behaviour of the C ABI: unknown layer, bad robot index, bad arguments -> negative codes, no crash."""
    from ros_navigation_b200 import DeviceGridMap, capi
    dg = DeviceGridMap(ctx, (4.0, 4.0), 0.05, layers=("master",))
    s = np.zeros(4, capi.SAMPLE_DTYPE)
    with pytest.raises(capi.B200NavError) as e:
        dg.himm_update("nope", s)
    assert e.value.code == capi.ENOLAYER
    with pytest.raises(capi.B200NavError) as e:
        dg.himm_update("master", s, robot=3)
    assert e.value.code == capi.EINVAL
    with pytest.raises(capi.B200NavError) as e:
        DeviceGridMap(ctx, (4000.0, 4.0), 0.05)
    assert e.value.code == capi.ERANGE
    dg.close()


def test_two_layer_compose_and_steering_from_it(ctx):
    """the sonar ("range") layer merged into master with the compose the reference carries
    commented out (map_provider.cpp:218-220); the VFH+ window then reads the composed (float-format) master."""
    from ros_navigation_b200 import VFH, DeviceGridMap
    from tests.util import random_samples
    rng = np.random.default_rng(21)
    g = O.make_geom(8.0, 8.0, 0.05)
    dg = DeviceGridMap(ctx, (8.0, 8.0), 0.05, n_robots=2, layers=("master", "laser", "range"))
    laser = [O.new_layer(g), O.new_layer(g)]
    sonar = [O.new_layer(g), O.new_layer(g)]
    v = VFH(ctx, n_robots=2)
    for cycle in range(4):
        per_l = [lidar_samples(rng, g, (0.2 * r, -0.3), 360, 0.3, 3.0) for r in range(2)]
        per_s = [random_samples(rng, g, 25, spread=0.6, clear_frac=0.2, origin=(0.2 * r, -0.3)) for r in range(2)]
        for per, name, store in ((per_l, "laser", laser), (per_s, "range", sonar)):
            off = np.cumsum([0] + [len(p) for p in per]).astype(np.int32)
            dg.himm_update_batched(name, np.concatenate(per), off)
            for r in range(2):
                O.himm_update(g, store[r], per[r])
        dg.compose_master("master", "range", "laser")
    assert dg.layer_format("master") == "float" and dg.layer_format("laser") == "coded"
    for r in range(2):
        want = O.compose_master(sonar[r], laser[r])
        assert_layers_equal(dg.download("master", robot=r), want, "master of robot %d" % r)
        pose = (0.2 * r, -0.3, 0.4)
        v.update_from_grid(dg, "master", VFH.make_input(x=pose[0], y=pose[1], yaw=pose[2]), robot=r)
        assert np.array_equal(v.ranges(robot=r)[:, 0], O.ranges_from_submap(g, want, *pose)[:, 0])
    with pytest.raises(Exception):
        dg.compose_master("laser", "range", "laser")      # destination aliases a source
    v.close()
    dg.close()
