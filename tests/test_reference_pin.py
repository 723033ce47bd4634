"""Pins the oracle restatement (oracle/himm_oracle.cpp) against the REFERENCE ITSELF.

oracle/_ref/libnav_ref.so is the reference's own move_control (MapProvider, LaserMapUpdater, RangeMapUpdater, Steerer,
VFH) and grid_map_core sources, compiled unchanged where they lie under /root/reference against stand-in headers for
the middleware this image lacks (oracle/standin: Eigen, roscpp, tf, laser_geometry ...; oracle/Makefile).  Every test
here drives that library and the restatement with the same inputs and demands identical bits:

  map_updater.h:38-71 (lineOnMap / clearCell / markCell) + LineIterator.cpp     -> oracle_himm_update
  laser_map_updater.cpp:38-144 (intake: rate limit, thinning, ifClearEnd quirk) -> oracle_project_scan (+ spec sincos)
  range_map_updater.cpp:38-76 (sonar rays)                                      -> samples built by the test
  map_provider.cpp:190-223 (update + compose), :177-188 (move)                  -> oracle_himm_update / oracle_move
  steerer.cpp:147-191 (getRangesFromSubmap), :221-270 (update)                  -> oracle_ranges_from_submap,
                                                                                   oracle_goal_from_pose, RefVFH
  map_global_planner.h:39-54 with CircleIterator.cpp                            -> oracle_if_blocked
"""
import math

import numpy as np
import pytest

from oracle import navref as N
from oracle import oracle as O
from ros_navigation_b200 import synth
from tests.util import assert_layers_equal, lidar_samples, random_samples

pytestmark = pytest.mark.skipif(not N.have_ref(), reason="oracle/_ref/libnav_ref.so not built (needs /root/reference)")


@pytest.mark.parametrize("lx,ly,res,pos", [(10.0, 10.0, 0.05, (0.0, 0.0)),       # C1
                                            (7.3, 5.9, 0.05, (0.3, -0.2)),       # odd sizes, off-centre
                                            (102.4, 102.4, 0.05, (0.0, 0.0)),    # C2: 2048 x 2048
                                            (40.96, 40.96, 0.02, (1.0, 2.0))])   # C3 resolution, 2048 x 2048
def test_himm_core_matches_reference(lx, ly, res, pos):
    rng = np.random.default_rng(int(lx * 100))
    g = O.make_geom(lx, ly, res, *pos)
    core = N.Core(lx, ly, res, pos)
    assert (core.rows, core.cols) == (g.rows, g.cols)
    layer = O.new_layer(g)
    origin = np.array(pos)
    for it in range(12):
        if it % 4 == 3:   # scattered rays, every start inside the map (a ray that misses the map is UB in the reference)
            s = random_samples(rng, g, 300, spread=0.98, clear_frac=0.3)
            s["ex"] = pos[0] + (rng.random(300) - 0.5) * lx * 1.6   # ends may lie outside
            s["ey"] = pos[1] + (rng.random(300) - 0.5) * ly * 1.6
        else:
            origin = origin + rng.uniform(-0.4, 0.4, 2)
            s = lidar_samples(rng, g, origin, 720, 0.1, 0.7 * max(lx, ly), clear_frac=0.15)
        b1, b2 = np.zeros(4), np.zeros(4)
        O.himm_update(g, layer, s, b1)
        core.update(s, b2)
        assert np.array_equal(b1, b2)
    assert_layers_equal(core.layer(), layer, "reference lineOnMap vs oracle")
    assert np.nansum(layer) > 0


def test_line_iterator_matches_reference():
    """The real grid_map::LineIterator against oracle_line_cells on random segments (both ends anywhere, start side
    always reaching the map) - beyond the three vendored known-answer cases of tests/test_oracle_kat.py."""
    rng = np.random.default_rng(11)
    for (lx, ly, res, pos) in [(8.0, 5.0, 1.0, (0.0, 0.0)), (10.0, 10.0, 0.05, (0.5, -0.25)), (3.0, 7.0, 0.1, (0, 0))]:
        g = O.make_geom(lx, ly, res, *pos)
        core = N.Core(lx, ly, res, pos)
        for _ in range(400):
            sx = pos[0] + (rng.random() - 0.5) * lx * 0.99
            sy = pos[1] + (rng.random() - 0.5) * ly * 0.99
            ex = pos[0] + (rng.random() - 0.5) * lx * 2.5
            ey = pos[1] + (rng.random() - 0.5) * ly * 2.5
            if rng.random() < 0.3:   # outside -> inside as well
                sx, sy, ex, ey = ex, ey, sx, sy
            want = core.line(sx, sy, ex, ey)
            got = O.line_cells(g, sx, sy, ex, ey)
            assert np.array_equal(got, want), (sx, sy, ex, ey)


def test_move_and_wrapped_lines_match_reference():
    """GridMap::move (GridMap.cpp:346-412) then rays across the seam: the reference walks wrapped buffer indices
    without unwrapping (SURVEY H4 d); the oracle must do the same."""
    rng = np.random.default_rng(3)
    lx = ly = 4.0
    g = O.make_geom(lx, ly, 0.05)
    core = N.Core(lx, ly, 0.05)
    layer = O.new_layer(g)
    centre = np.zeros(2)
    for it in range(25):
        centre = centre + rng.uniform(-0.35, 0.35, 2)
        moved_ref = core.move(*centre)
        moved = O.move(g, [layer], *centre)
        assert moved == moved_ref
        start, posn = core.start_index()
        assert start == (g.start0, g.start1) and posn == (g.pos_x, g.pos_y)
        s = lidar_samples(rng, g, centre + rng.uniform(-0.1, 0.1, 2), 360, 0.1, 3.0, clear_frac=0.1)
        O.himm_update(g, layer, s)
        core.update(s)
        assert_layers_equal(core.layer(), layer, "after move %d" % it)


def test_if_blocked_matches_reference_circle_iterator():
    rng = np.random.default_rng(8)
    g = O.make_geom(10.0, 10.0, 0.05)
    core = N.Core(10.0, 10.0, 0.05, layer="master")
    layer = O.new_layer(g)
    for _ in range(6):
        O.himm_update(g, layer, lidar_samples(rng, g, rng.uniform(-2, 2, 2), 360, 0.5, 4.0))
    core.set_layer(layer)
    hits = 0
    for _ in range(500):
        x, y = rng.uniform(-4.6, 4.6, 2)
        want = core.blocked(x, y, 0.3)
        assert O.if_blocked(g, layer, x, y, 0.3) == want
        hits += want
    assert 0 < hits < 500


def _scan(world, t, beams, fov, range_max):
    x, y, yaw = [float(v[0]) for v in world.pose(t)]
    r, _ = world.cast(*world.pose(t), beams, fov, range_max)
    return (x, y, yaw), r[0].numpy().astype(np.float32)


@pytest.mark.parametrize("beams,fov,range_max,extent", [(360, 2 * math.pi, 3.0, 10.0),      # C1: not thinned
                                                         (1080, 1.5 * math.pi, 6.0, 25.6)])  # C4: thinned to ~1 deg
def test_node_pipeline_matches_reference(beams, fov, range_max, extent):
    """Whole reference node (scan callback -> MapProvider::updateMap -> Steerer::update) against the oracle pipeline,
    step by step: laser / master layers, pseudo-scan, both histograms, velocity command."""
    node = N.Node(N.REF_PATH, extent, extent, moving=False, t0=1.0)
    g = O.make_geom(extent, extent, 0.05)
    layer = O.new_layer(g)
    world = synth.Worlds(1, extent, 4242 + beams)
    vfh = O.RefVFH()
    amin, ainc = np.float32(-fov / 2), np.float32(fov / beams)
    goal = (0.25 * extent, -0.2 * extent)
    node.accept_plan([[0.0, 0.0], goal])
    flagged = 0
    for step in range(30):
        t = 1.0 + 0.2 * (step + 1)
        pose, ranges = _scan(world, t, beams, fov, range_max)
        # what a real driver reports for "nothing hit": +inf on even steps, range_max itself on odd ones
        ranges[ranges >= range_max] = np.inf if step % 2 == 0 else np.float32(range_max)
        node.set_time(t)
        node.set_frame("base_link", *pose)
        node.set_frame("laser", *pose)
        node.publish_scan(ranges, float(amin), float(ainc), 0.1, range_max)
        node.publish_scan(np.ones_like(ranges), float(amin), float(ainc), 0.1, range_max)   # same instant: rate-limited away
        node.update_map()
        s = O.project_scan(amin, ainc, 0.1, range_max, ranges, pose, decimate=True)
        flagged += int(s["clear_end"].sum())
        O.himm_update(g, layer, s)
        assert_layers_equal(node.layer("laser"), layer, "laser, step %d" % step)
        assert_layers_equal(node.layer("master"), layer, "master, step %d" % step)

        node.publish_odom(0.1)
        out = node.steer()
        assert out["updated"]
        rp = node.robot_pose()
        want = O.ranges_from_submap(g, layer, rp[0], rp[1], rp[2])
        assert np.array_equal(out["ranges"], want[:, 0]), "pseudo-scan, step %d" % step
        gdir, gdist = O.goal_from_pose(rp[0], rp[1], rp[2], *goal)
        cs, ct = vfh.update(want, int(0.1 * 1000.0), gdir, gdist, 250.0, 0.2)
        st = vfh.state()
        assert np.array_equal(st["origin_hist"], out["origin_hist"])
        assert np.array_equal(st["hist"], out["hist"])
        assert out["picked_angle"] == np.float32(st["picked"])
        assert out["linear_x"] == float(np.float32(cs)) / 1000.0
        assert out["angular_z"] == ct * math.pi / 180.0
    if beams > 400:
        assert flagged > 0, "the thinned-scan ifClearEnd quirk was never exercised"
    node.close()


def test_sonar_rays_and_moving_map_match_reference():
    """mapTest_vfh shape: 4 m map that follows the robot (MapProvider(nh, Length(4,4), true)), laser + five sonars."""
    node = N.Node(N.REF_PATH, 4.0, 4.0, moving=True, t0=1.0)
    g = O.make_geom(4.0, 4.0, 0.05)
    laser, sonar, master = O.new_layer(g), O.new_layer(g), O.new_layer(g)
    world = synth.Worlds(1, 10.0, 77)
    beams, fov, range_max = 360, 2 * math.pi, 3.0
    amin, ainc = np.float32(-fov / 2), np.float32(fov / beams)
    sonars = [("/left_range", "sonar_l", 1.2), ("/right_range", "sonar_r", -1.2), ("/front_range", "sonar_f", 0.0)]
    rng = np.random.default_rng(2)
    for step in range(40):
        t = 1.0 + 0.2 * (step + 1)
        pose, ranges = _scan(world, t, beams, fov, range_max)
        node.set_time(t)
        node.set_frame("base_link", *pose)
        node.set_frame("laser", *pose)
        if step % 3 == 2:   # loopMoveMap fires between updates
            assert node.move_map() == int(O.move(g, [laser, sonar, master], pose[0], pose[1]))
            geo = node.geometry()
            assert geo["start"] == (g.start0, g.start1) and geo["pos"] == (g.pos_x, g.pos_y)
        node.publish_scan(ranges, float(amin), float(ainc), 0.1, range_max)
        sonar_samples = []
        for topic, frame, dyaw in sonars:
            fx, fy, fyaw = pose[0] + 0.1 * math.cos(pose[2] + dyaw), pose[1] + 0.1 * math.sin(pose[2] + dyaw), pose[2] + dyaw
            node.set_frame(frame, fx, fy, fyaw)
            rr = np.float32(rng.uniform(0.2, 2.2))
            max_r = np.float32(2.0)
            node.publish_range(topic, frame, float(rr), 0.05, float(max_r))
            s_, c_ = O.sincos(fyaw)
            # tf stand-in: p_map = (c*px - s*py) + x0, (s*px + c*py) + y0 with py = 0
            ex, ey = (c_ * float(rr) - s_ * 0.0) + fx, (s_ * float(rr) + c_ * 0.0) + fy
            sonar_samples.append((fx, fy, ex, ey, 0 if rr < max_r else 1))
        node.update_map()
        O.himm_update(g, laser, O.project_scan(amin, ainc, 0.1, range_max, ranges, pose, decimate=True))
        ss = np.array(sonar_samples)
        O.himm_update(g, sonar, O.make_samples(ss[:, 0], ss[:, 1], ss[:, 2], ss[:, 3], ss[:, 4].astype(np.int32)))
        master[...] = laser   # map_provider.cpp:221
        assert_layers_equal(node.layer("laser"), laser, "laser, step %d" % step)
        assert_layers_equal(node.layer("range"), sonar, "range, step %d" % step)
        assert_layers_equal(node.layer("master"), master, "master, step %d" % step)
    occ = node.occupancy()
    assert np.array_equal(occ, O.to_occupancy(g, master))
    node.close()


def test_steer_goal_glue_matches_reference_steerer():
    """b200nav_steer_update_goals (the product's host-side restatement of Steerer::update's caller glue,
    steerer.cpp:228-256: waypoint advance, desiredDist, desiredAngle, odometry speed) against the reference's own
    Steerer::update on a multi-waypoint plan, until the plan is exhausted."""
    from ros_navigation_b200 import capi
    L = capi.lib()
    rng = np.random.default_rng(5)
    node = N.Node(N.REF_PATH, 10.0, 10.0, False, t0=1.0)
    plan = np.array([[0, 0], [0.5, 0.2], [0.6, 0.25], [2.0, 1.0], [2.1, 1.1], [1.0, 2.0]], float)
    node.accept_plan(plan)
    idx, offs = np.array([1], np.int32), np.array([0, len(plan)], np.int32)
    inp, done = np.zeros(1, capi.VFH_INPUT_DTYPE), np.zeros(1, np.uint8)
    pos, leg, compared = np.zeros(2), 1, 0
    for step in range(2000):
        if leg < len(plan):   # walk towards the current waypoint with some wobble
            d = plan[leg] - pos
            pos = pos + 0.02 * d / max(np.hypot(*d), 1e-9) + rng.uniform(-0.004, 0.004, 2)
            if np.hypot(*(plan[leg] - pos)) < 0.2:
                leg += 1
        yaw, v = rng.uniform(-3.1, 3.1), rng.uniform(0, 0.3)
        node.set_time(1.0 + 0.2 * (step + 1))
        node.set_frame("base_link", pos[0], pos[1], yaw)
        node.publish_odom(v)
        out, rp = node.steer(), node.robot_pose()
        poses, speed = np.array([[rp[0], rp[1], rp[2]]]), np.array([v])
        assert L.b200nav_steer_update_goals(1, poses.ctypes.data, plan.ctypes.data, offs.ctypes.data, idx.ctypes.data,
                                            250.0, speed.ctypes.data, inp.ctypes.data, done.ctypes.data) == 0
        assert bool(done[0]) == (not out["plan_ready"]), step
        if done[0]:
            break
        assert np.float32(inp["goal_direction"][0]) == np.float32(out["desired_angle"]), step
        assert int(inp["current_speed"][0]) == int(v * 1000.0)
        compared += 1
    assert done[0] == 1 and compared > 100 and idx[0] == len(plan)
    node.close()


def test_himm_core_random_geometries_and_boundary_endpoints():
    """Property test of the same pin: geometries the fixed cases do not reach (resolutions that do not divide the
    lengths, far-away map centres, thin maps) and end points placed EXACTLY on cell borders (multiples of the
    resolution from the map corner), where the two association orders of the index arithmetic and the truncation
    toward zero decide the cell."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    @settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(st.floats(0.6, 9.0), st.floats(0.6, 9.0), st.sampled_from([0.05, 0.03, 0.07, 0.1, 0.125, 0.02]),
           st.floats(-2000.0, 2000.0), st.floats(-2000.0, 2000.0), st.integers(0, 2 ** 31 - 1))
    def check(lx, ly, res, px, py, seed):
        rng = np.random.default_rng(seed)
        g = O.make_geom(lx, ly, res, px, py)
        core = N.Core(lx, ly, res, (px, py))
        assert (core.rows, core.cols) == (g.rows, g.cols)
        layer = O.new_layer(g)
        Lx, Ly = g.rows * res, g.cols * res
        for it in range(4):
            n = 120
            s = random_samples(rng, g, n, spread=0.97, clear_frac=0.25)     # starts inside the map
            # a third of the end points exactly on cell borders (also on the map's outer border), a third outside
            kx, ky = rng.integers(0, g.rows + 1, n), rng.integers(0, g.cols + 1, n)
            on_border = rng.random(n) < 0.34
            s["ex"] = np.where(on_border, (px + Lx / 2) - kx * res, s["ex"])
            s["ey"] = np.where(on_border, (py + Ly / 2) - ky * res, s["ey"])
            outside = rng.random(n) < 0.3
            s["ex"] = np.where(outside & ~on_border, px + (rng.random(n) - 0.5) * Lx * 2.2, s["ex"])
            s["ey"] = np.where(outside & ~on_border, py + (rng.random(n) - 0.5) * Ly * 2.2, s["ey"])
            b1, b2 = np.zeros(4), np.zeros(4)
            O.himm_update(g, layer, s, b1)
            core.update(s, b2)
            assert np.array_equal(b1, b2)
        assert_layers_equal(core.layer(), layer, "lx=%r ly=%r res=%r pos=(%r, %r) seed=%d" % (lx, ly, res, px, py, seed))
        core.close()

    check()
