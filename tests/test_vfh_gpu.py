"""VFH+ parity: CUDA path (through the C ABI) vs the REFERENCE move_control::VFH (oracle/_ref) and vs the committed
golden vectors generated from it.

Bars (BASELINE.json north_star): primary polar histogram within 1e-5 relative (this implementation sums in the
reference's order, so the test demands bit equality and reports the relative error otherwise); binary / masked
histograms, picked direction, speed and turn rate exact.
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import lidar_samples

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PRIMARY_RTOL = 1e-5  # north_star tolerance for the primary histogram


@pytest.fixture(scope="module")
def ctx():
    from ros_navigation_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def params_from_dict(p):
    from ros_navigation_b200 import VfhParams
    kw = {k: (int(v) if k in ("window_diameter", "sector_angle", "max_speed", "max_speed_narrow_opening",
                              "max_speed_wide_opening", "max_acceleration", "min_turnrate", "max_turnrate_0ms",
                              "max_turnrate_1ms") else float(v)) for k, v in p.items()}
    return VfhParams(**kw)


def check_primary(got, want, what):
    if np.array_equal(got, want):
        return
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-30)
    assert rel.max() <= PRIMARY_RTOL, "%s: primary histogram rel err %g" % (what, rel.max())


@pytest.mark.parametrize("case", ["w30_default", "w30_allidx", "w33", "w30_fast", "w60_fixed_safety"])
def test_update_ranges_vs_golden(ctx, case):
    """Update_VFH on host pseudo-scans vs golden vectors produced by the reference class."""
    from ros_navigation_b200 import VFH
    z = np.load(os.path.join(GOLD, "vfh_golden.npz"))
    g = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(case + "/")}
    p = dict(zip(O.VFH_PARAM_ORDER, g["params"]))
    v = VFH(ctx, params_from_dict(p))
    for s in range(len(g["dt"])):
        cs, ct = v.Update_VFH(g["ranges"][s], int(g["speed"][s]), float(g["gdir"][s]), float(g["gdist"][s]),
                              float(g["tol"][s]), dt=float(g["dt"][s]))
        st = v.state()
        what = "%s step %d" % (case, s)
        check_primary(st["origin_hist"], g["origin_hist"][s], what)
        assert np.array_equal(st["last_binary"], g["last_binary"][s]), what + " binary"
        assert np.array_equal(st["hist"], g["hist"][s]), what + " masked"
        assert st["picked"] == g["picked"][s], what + " picked"
        assert (cs, ct) == (g["out_speed"][s], g["out_turn"][s]), what + " command"
        assert st["blocked_radius"] == g["blocked"][s], what
    v.close()


@pytest.mark.parametrize("kw", [dict(), dict(window_diameter=33), dict(window_diameter=129, cell_size=20.0),
                                dict(sector_angle=2, window_diameter=41, safety_dist_1ms=200.0)])
def test_tables_match_reference_init(ctx, kw):
    if not O.have_ref():
        pytest.skip("oracle/_ref missing")
    from ros_navigation_b200 import VFH
    ref = O.RefVFH(**kw)
    p = dict(ref.params)
    v = VFH(ctx, params_from_dict(p))
    assert v.hist_size == ref.hist_size and v.num_tables == ref.num_tables
    rd, rs, rb = ref.cell_tables()
    for t in range(ref.num_tables):
        d, s, b, m, mtr = v.tables(t)
        assert np.array_equal(d.view(np.uint32), rd.view(np.uint32))
        assert np.array_equal(s.view(np.uint32), rs.view(np.uint32))
        assert np.array_equal(b.view(np.uint32), rb.view(np.uint32))
        rm, ok = ref.sector_masks(t)
        assert ok and np.array_equal(m, rm)
    assert np.array_equal(mtr, ref.min_turning_radius())
    v.close()


def run_live(ctx, steps, seed, **kw):
    """Random pseudo-scan sequence through both the reference class and the CUDA path."""
    from ros_navigation_b200 import VFH
    from tests.golden.make_golden import scan_sequence
    rng = np.random.default_rng(seed)
    ref = O.RefVFH(**kw)
    v = VFH(ctx, params_from_dict(ref.params))
    ranges = scan_sequence(rng, steps, even_only=bool(seed % 2))
    for s in range(steps):
        speed = int(rng.integers(-10, ref.params["max_speed"] + 1))
        gdir = float(np.float32(rng.uniform(0, 360))) if rng.random() < 0.6 else 90.0
        gdist = float(np.float32(rng.uniform(50, 5000)))
        dt = float(rng.choice([0.2, 0.1, 0.35, 0.0]))
        rcs, rct = ref.update(ranges[s], speed, gdir, gdist, 250.0, dt)
        cs, ct = v.Update_VFH(ranges[s], speed, gdir, gdist, 250.0, dt=dt)
        a, b = ref.state(), v.state()
        what = "seed %d step %d" % (seed, s)
        check_primary(b["origin_hist"], a["origin_hist"], what)
        assert np.array_equal(b["last_binary"], a["last_binary"]), what
        assert np.array_equal(b["hist"], a["hist"]), what
        assert b["picked"] == a["picked"] and b["last_picked"] == a["last_picked"], what
        assert (cs, ct) == (rcs, rct), what
        assert b["last_chosen_speed"] == a["last_chosen_speed"], what
    v.close()


@pytest.mark.parametrize("seed,kw", [(101, dict()), (102, dict()), (103, dict(window_diameter=33)),
                                     (104, dict(window_diameter=129, cell_size=20.0)),
                                     (105, dict(weight_desired_dir=5.0, weight_current_dir=3.0, max_speed=400)),
                                     (106, dict(sector_angle=10)), (107, dict(sector_angle=3, window_diameter=21)),
                                     (108, dict(safety_dist_1ms=10.0)),            # one sector table
                                     (109, dict(max_turnrate_1ms=10, max_speed=1000, robot_radius=250.0))])
def test_update_ranges_vs_live_reference(ctx, seed, kw):
    if not O.have_ref():
        pytest.skip("oracle/_ref missing")
    run_live(ctx, 80, seed, **kw)


def test_flags_and_current_max_speed(ctx):
    """Emergency stop, hemmed-in and cant-turn paths are exercised and flagged; SetCurrentMaxSpeed (vfh.cpp:144-166)
    rebuilds the turning-radius table like the reference."""
    if not O.have_ref():
        pytest.skip("oracle/_ref missing")
    from ros_navigation_b200 import VFH, capi
    ref = O.RefVFH()
    v = VFH(ctx, params_from_dict(ref.params))
    seen = 0
    rng = np.random.default_rng(3)
    free = np.full((361, 2), 5000.0)
    wall_close = free.copy()
    wall_close[:, 0] = 120.0           # everything 12 cm away: inside the safety distance -> emergency
    wall_mid = free.copy()
    wall_mid[:, 0] = 900.0             # blocked all around but outside the safety distance -> hemmed in
    for step in range(60):
        if step == 20:
            ref.set_current_max_speed(120)
            v.SetCurrentMaxSpeed(120)
        r = [free, wall_close, wall_mid, free][step % 4] if step < 40 else free
        gdist = float(np.float32(rng.uniform(50, 400))) if step >= 40 else 3000.0
        gdir = float(np.float32(rng.uniform(0, 180)))
        speed = int(rng.integers(0, 200))
        rcs, rct = ref.update(r, speed, gdir, gdist, 250.0, 0.2)
        inp = VFH.make_input(dt=0.2, speed=speed, goal_dir=gdir, goal_dist=gdist, tol=250.0)
        out = np.zeros(1, capi.COMMAND_DTYPE)
        capi.check(capi.lib().b200nav_vfh_update_ranges(v.h, 0, np.ascontiguousarray(r).ctypes.data, inp.ctypes.data,
                                                         out.ctypes.data), ctx.h)
        assert (int(out["speed"][0]), int(out["turnrate"][0])) == (rcs, rct), "step %d" % step
        assert float(out["picked_angle"][0]) == ref.state()["picked"]
        seen |= int(out["flags"][0])
        if step % 4 == 1 and step < 40:
            assert int(out["flags"][0]) & capi.CMD_EMERGENCY
            assert np.array_equal(v.state()["origin_hist"], np.ones(72, np.float32))
        if step % 4 == 2 and step < 40:
            assert int(out["flags"][0]) & capi.CMD_HEMMED_IN
    assert seen & capi.CMD_EMERGENCY and seen & capi.CMD_HEMMED_IN and seen & capi.CMD_CANT_TURN
    assert np.array_equal(v.tables(0)[4][:121], ref.min_turning_radius())
    v.close()


def make_world_layer(rng, g, density=0.03):
    layer = O.new_layer(g)
    occ = rng.random(layer.shape)
    layer[occ < 0.15] = 0.0
    m = (occ >= 0.15) & (occ < 0.15 + density)
    layer[m] = rng.choice([10.0, 30.0, 90.0, 180.0, 3.0, 3.5], m.sum())
    return layer


@pytest.mark.parametrize("geom,start,tma", [
    ((10.0, 10.0, 0.05, 0.0, 0.0), (0, 0), True),
    ((10.0, 10.0, 0.05, 0.0, 0.0), (0, 0), False),
    ((4.0, 4.0, 0.05, 1.3, -0.7), (17, 63), True),    # circular buffer: windows wrap -> coalesced-load path
    ((6.5, 3.5, 0.05, -2.0, 5.0), (0, 0), True),      # rows % 4 != 0 -> no TMA
])
def test_ranges_from_grid_vs_oracle_and_golden(ctx, geom, start, tma):
    """Steerer::getRangesFromSubmap on the device (stage R) vs the oracle restatement: exact."""
    from ros_navigation_b200 import VFH, DeviceGridMap
    rng = np.random.default_rng(7)
    g = O.make_geom(*geom, start)
    dg = DeviceGridMap(ctx, geom[:2], geom[2], geom[3:5], layers=("master",))
    dg.set_geometry(0, geom[3:5], start)
    layer = make_world_layer(rng, g)
    dg.upload("master", layer)
    v = VFH(ctx)
    v.disable_tma(not tma)
    for k in range(40):
        x = geom[3] + (rng.random() - 0.5) * geom[0] * 1.1
        y = geom[4] + (rng.random() - 0.5) * geom[1] * 1.1
        yaw = rng.uniform(-np.pi, np.pi)
        want = O.ranges_from_submap(g, layer, x, y, yaw)
        v.update_from_grid(dg, "master", VFH.make_input(x=x, y=y, yaw=yaw))
        got = v.ranges()
        assert np.array_equal(got[:, 0], want[:, 0]), "pose %d (%g,%g,%g): %d entries differ" % (
            k, x, y, yaw, (got[:, 0] != want[:, 0]).sum())
    # committed golden poses for the same geometry family
    z = np.load(os.path.join(GOLD, "ranges_golden.npz"))
    for i in range(3):
        gg = z["%d/geom" % i]
        if tuple(gg[:5]) != tuple(geom) or tuple(gg[5:].astype(int)) != tuple(start):
            continue
        dg.upload("master", z["%d/layer" % i])
        for p, want in zip(z["%d/poses" % i], z["%d/ranges" % i]):
            v.update_from_grid(dg, "master", VFH.make_input(x=p[0], y=p[1], yaw=p[2]))
            assert np.array_equal(v.ranges()[:, 0], want)
    v.close()
    dg.close()


def test_full_pipeline_single_robot_vs_reference(ctx):
    """mapTest_vfh scenario (BASELINE config 1): scans -> HIMM grid -> master -> pseudo-scan -> VFH+ command, every
    stage compared with the oracle / reference class."""
    if not O.have_ref():
        pytest.skip("oracle/_ref missing")
    import torch
    from ros_navigation_b200 import VFH, DeviceGridMap, LaserMapUpdater, synth
    for window in (30, 33):
        cfg = synth.CONFIGS["c1"]
        w = synth.Worlds(1, cfg["extent"], synth.config_seed("c1"))
        g = O.make_geom(cfg["extent"], cfg["extent"], cfg["res"])
        laser, master = O.new_layer(g), O.new_layer(g)
        dg = DeviceGridMap(ctx, (cfg["extent"],) * 2, cfg["res"], layers=("master",))
        upd = LaserMapUpdater(dg, "laser")
        ref = O.RefVFH(window_diameter=window)
        v = VFH(ctx, params_from_dict(ref.params))
        speed = 0
        for step in range(150):
            t = step * 0.2
            x, y, yaw = w.pose(t)
            rng_, ang = w.cast(x, y, yaw, cfg["beams"], cfg["fov"], cfg["range_max"])
            s8, off = synth.samples_from_scan(x, y, yaw, rng_, ang, cfg["range_max"])
            s = synth.samples_to_numpy(s8)
            # MapProvider::updateMap (map_provider.cpp:190-205)
            bb_o, bb_d = np.zeros(4), np.zeros(4)
            O.himm_update(g, laser, s, bb_o)
            master[:] = laser
            upd.push_samples(s)
            upd.updateMap(bb_d)
            dg.copy_layer("master", "laser")
            assert np.array_equal(bb_o, bb_d)
            # Steerer::update (steerer.cpp:221-270)
            inp = synth.vfh_inputs_to_numpy(synth.vfh_inputs(w, t, 0.2, speed))
            px, py, pyaw = float(inp["x"][0]), float(inp["y"][0]), float(inp["yaw"][0])
            want_r = O.ranges_from_submap(g, master, px, py, pyaw)
            rcs, rct = ref.update(want_r, speed, float(inp["goal_direction"][0]), float(inp["goal_distance"][0]),
                                  250.0, 0.2)
            cmd = v.update_from_grid(dg, "master", inp)
            assert np.array_equal(v.ranges()[:, 0], want_r[:, 0]), "step %d ranges" % step
            a, b = ref.state(), v.state()
            check_primary(b["origin_hist"], a["origin_hist"], "step %d" % step)
            assert np.array_equal(b["hist"], a["hist"]) and np.array_equal(b["last_binary"], a["last_binary"])
            assert (int(cmd["speed"]), int(cmd["turnrate"])) == (rcs, rct), "step %d command" % step
            assert float(cmd["picked_angle"]) == a["picked"]
            speed = rcs
        from tests.util import assert_layers_equal
        assert_layers_equal(dg.download("master"), master, "master layer")
        v.close()
        dg.close()


def test_batched_robots_vs_reference(ctx):
    """Batched mode: every robot has its own grid, pose stream and VFH state."""
    if not O.have_ref():
        pytest.skip("oracle/_ref missing")
    from ros_navigation_b200 import VFH, DeviceGridMap, synth
    n = 24
    rng = np.random.default_rng(9)
    geom = (12.8, 12.8, 0.05, 0.0, 0.0)
    g = O.make_geom(*geom)
    dg = DeviceGridMap(ctx, geom[:2], geom[2], n_robots=n, layers=("laser",))
    dg.alias("master", "laser")
    layers = []
    for r in range(n):
        lay = make_world_layer(rng, g, density=0.004 if r % 3 else 0.05)
        layers.append(lay)
        dg.upload("laser", lay, robot=r)
    refs = [O.RefVFH() for _ in range(n)]
    v = VFH(ctx, n_robots=n)
    from ros_navigation_b200.capi import VFH_INPUT_DTYPE
    speeds = np.zeros(n, np.int32)
    for step in range(25):
        inp = np.zeros(n, VFH_INPUT_DTYPE)
        inp["x"] = (rng.random(n) - 0.5) * 11
        inp["y"] = (rng.random(n) - 0.5) * 11
        inp["yaw"] = rng.uniform(-np.pi, np.pi, n)
        inp["dt"] = 0.2
        inp["current_speed"] = speeds
        inp["goal_direction"] = rng.uniform(0, 360, n).astype(np.float32)
        inp["goal_distance"] = rng.uniform(100, 4000, n).astype(np.float32)
        inp["goal_tolerance"] = 250.0
        out = v.update_batched(dg, "master", inp)
        for r in range(n):
            want_r = O.ranges_from_submap(g, layers[r], inp["x"][r], inp["y"][r], inp["yaw"][r])
            rcs, rct = refs[r].update(want_r, int(speeds[r]), float(inp["goal_direction"][r]),
                                      float(inp["goal_distance"][r]), 250.0, 0.2)
            assert (int(out["speed"][r]), int(out["turnrate"][r])) == (rcs, rct), "step %d robot %d" % (step, r)
            assert float(out["picked_angle"][r]) == refs[r].state()["picked"]
            if r in (0, 5, 23):
                a, b = refs[r].state(), v.state(r)
                check_primary(b["origin_hist"], a["origin_hist"], "step %d robot %d" % (step, r))
                assert np.array_equal(b["hist"], a["hist"])
        speeds = out["speed"].copy()
    v.close()
    dg.close()
