"""GPU: the drop-in class headers (include/move_control/*.h on libb200nav.so) under the reference's OWN node code.

tests/cpp/_build/libnav_dropin.so is the reference's map_provider.cpp and steerer.cpp, compiled unchanged from where
they lie, with "move_control/map_updater.h", "laser_map_updater.h", "range_map_updater.h" and "vfh.h" resolved to the
product's headers: the BOILERPLATE_CODE factory (map_provider.cpp:12-15,262-266) instantiates the product's
LaserMapUpdater / RangeMapUpdater through the reference's constructor signature, Steerer::initVfh (steerer.cpp:46-133)
builds the product's VFH with the reference's 19 arguments.  oracle/_ref/libnav_ref.so is the same harness around the
reference's own classes.  Both are driven through identical scenarios; grids, pseudo-scans, histograms, velocity
commands and the published OccupancyGrid must be identical at every step.
"""
import math

import numpy as np
import pytest

from oracle import navref as N
from ros_navigation_b200 import synth
from tests.util import assert_layers_equal

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (N.have_ref() and N.have_dropin()),
                                 reason="libnav_ref.so / libnav_dropin.so not prebuilt (needs /root/reference at build time)")]


def _scan(world, t, beams, fov, range_max):
    x, y, yaw = [float(v[0]) for v in world.pose(t)]
    r, _ = world.cast(*world.pose(t), beams, fov, range_max)
    return (x, y, yaw), r[0].numpy().astype(np.float32)


def _both(extent, moving):
    return (N.Node(N.REF_PATH, extent, extent, moving=moving, t0=1.0),
            N.Node(N.DROPIN_PATH, extent, extent, moving=moving, t0=1.0))


def _compare_steer(a, b, step):
    for k in ("linear_x", "angular_z", "updated", "plan_ready", "picked_angle", "desired_angle"):
        assert a[k] == b[k], "%s differs at step %d: %r vs %r" % (k, step, a[k], b[k])
    assert np.array_equal(a["ranges"], b["ranges"]), "pseudo-scan differs at step %d" % step
    assert np.array_equal(a["origin_hist"], b["origin_hist"]), "primary histogram differs at step %d" % step
    assert np.array_equal(a["hist"], b["hist"]), "masked histogram differs at step %d" % step


@pytest.mark.parametrize("beams,fov,range_max,extent", [(360, 2 * math.pi, 3.0, 10.0), (1080, 1.5 * math.pi, 12.0, 25.6)])
def test_dropin_node_matches_reference_node(beams, fov, range_max, extent):
    ref, dut = _both(extent, False)
    assert dut.L.navh_is_dropin() == 1 and ref.L.navh_is_dropin() == 0
    world = synth.Worlds(1, extent, 999 + beams)
    amin, ainc = np.float32(-fov / 2), np.float32(fov / beams)
    plan = [[0.0, 0.0], [0.2 * extent, 0.1 * extent], [-0.25 * extent, 0.2 * extent]]
    for n in (ref, dut):
        n.accept_plan(plan)
    emergencies = 0
    for step in range(60):
        t = 1.0 + 0.2 * (step + 1)
        pose, ranges = _scan(world, t, beams, fov, range_max)
        ranges[ranges >= range_max] = np.inf if step % 2 == 0 else np.float32(range_max)
        speed = 0.05 * (step % 5)
        outs = []
        for n in (ref, dut):
            n.set_time(t)
            n.set_frame("base_link", *pose)
            n.set_frame("laser", *pose)
            n.publish_scan(ranges, float(amin), float(ainc), 0.1, range_max)
            n.update_map()
            n.publish_odom(speed)
            outs.append(n.steer())
        assert_layers_equal(dut.layer("laser"), ref.layer("laser"), "laser, step %d" % step)
        assert_layers_equal(dut.layer("master"), ref.layer("master"), "master, step %d" % step)
        _compare_steer(outs[0], outs[1], step)
        emergencies += int(outs[0]["linear_x"] == 0.0)
        ha, hb = ref.hist_msg(), dut.hist_msg()
        assert ha[0] == hb[0] and np.array_equal(ha[1], hb[1]) and np.array_equal(ha[2], hb[2])
    assert np.array_equal(ref.occupancy(), dut.occupancy())
    assert np.nansum(ref.layer("laser")) > 0
    ref.close()
    dut.close()


def test_dropin_moving_map_with_sonars_matches_reference():
    """mapTest_vfh: MapProvider(nh, Length(4,4), ifMoving=true) - the host GridMap is moved behind the updaters' back
    (loopMoveMap), five sonar topics feed the "range" layer, the laser feeds "laser"."""
    ref, dut = _both(4.0, True)
    world = synth.Worlds(1, 10.0, 31337)
    beams, fov, range_max = 360, 2 * math.pi, 3.0
    amin, ainc = np.float32(-fov / 2), np.float32(fov / beams)
    sonars = [("/left_range", "sonar_l", 1.2), ("/right_range", "sonar_r", -1.2), ("/front_range", "sonar_f", 0.0),
              ("/front_left_range", "sonar_fl", 0.6), ("/front_right_range", "sonar_fr", -0.6)]
    rng = np.random.default_rng(4)
    for n in (ref, dut):
        n.accept_plan([[0.0, 0.0], [2.5, 2.0]])
    for step in range(60):
        t = 1.0 + 0.2 * (step + 1)
        pose, ranges = _scan(world, t, beams, fov, range_max)
        son = [(topic, frame, pose[0] + 0.1 * math.cos(pose[2] + d), pose[1] + 0.1 * math.sin(pose[2] + d), pose[2] + d,
                float(np.float32(rng.uniform(0.2, 2.2)))) for topic, frame, d in sonars]
        outs = []
        for n in (ref, dut):
            n.set_time(t)
            n.set_frame("base_link", *pose)
            n.set_frame("laser", *pose)
            if step % 3 == 0:   # from the first step on: outside its map the reference's getSubmap fails and the
                n.move_map()    # empty-map iterator divides by zero - the robot must stay inside
            n.publish_scan(ranges, float(amin), float(ainc), 0.1, range_max)
            for topic, frame, fx, fy, fyaw, rr in son:
                n.set_frame(frame, fx, fy, fyaw)
                n.publish_range(topic, frame, rr, 0.05, 2.0)
            n.update_map()
            n.publish_odom(0.1)
            outs.append(n.steer())
        assert ref.geometry() == dut.geometry()
        for layer in ("laser", "range", "master"):
            assert_layers_equal(dut.layer(layer), ref.layer(layer), "%s, step %d" % (layer, step))
        _compare_steer(outs[0], outs[1], step)
    assert ref.geometry()["start"] != (0, 0), "the map never moved"
    assert np.array_equal(ref.occupancy(), dut.occupancy())
    ref.close()
    dut.close()


def test_dropin_fast_path_from_device_grid_matches_reference_steerer():
    """instead of Steerer::getRangesFromSubmap + Update_VFH on the host copy, the drop-in
    VFH reads the window from the device twin of MapProvider's map (Update_VFH_FromGrid on layer "laser", which the
    laser updater keeps current in HBM) with the goal computed by b200nav_steer_update_goals.  Commands, picked angle
    and both histograms must equal what the reference's own Steerer::update publishes for the same scans and plan."""
    extent, beams, fov, range_max = 10.0, 360, 2 * math.pi, 3.0
    ref, dut = _both(extent, False)
    world = synth.Worlds(1, extent, 2718)
    amin, ainc = np.float32(-fov / 2), np.float32(fov / beams)
    plan = np.array([[0.0, 0.0], [1.5, 1.0], [-2.0, 1.5], [0.5, -2.5]])
    ref.accept_plan(plan)
    idx, compared = 1, 0
    for step in range(80):
        t = 1.0 + 0.2 * (step + 1)
        pose, ranges = _scan(world, t, beams, fov, range_max)
        speed = 0.04 * (step % 6)
        for n in (ref, dut):
            n.set_time(t)
            n.set_frame("base_link", *pose)
            n.set_frame("laser", *pose)
            n.publish_scan(ranges, float(amin), float(ainc), 0.1, range_max)
            n.update_map()
            n.publish_odom(speed)
        want = ref.steer()
        rc, idx, got = dut.steer_from_grid(plan, idx, speed, "laser")
        assert rc in (0, 1)
        assert got["plan_ready"] == want["plan_ready"], step
        if not want["plan_ready"]:
            break
        for k in ("linear_x", "angular_z", "picked_angle", "desired_angle"):
            assert got[k] == want[k], "%s differs at step %d: %r vs %r" % (k, step, got[k], want[k])
        assert np.array_equal(got["origin_hist"], want["origin_hist"]) and np.array_equal(got["hist"], want["hist"]), step
        compared += 1
    assert compared >= 40
    ref.close()
    dut.close()
