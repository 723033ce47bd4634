"""GPU parity at the BASELINE configurations the round-1 suite left open:

  C3  8192 x 8192 @ 2 cm, 4096-beam scans, VFH+ with a 129 x 129 window read from the grid (2.58 m submap):
      layer bit-exact, pseudo-scan exact vs oracle_ranges_from_submap (steerer.cpp:147-191 at that submap size),
      commands / histograms vs the live reference VFH class built with window_diameter = 129, cell_size = 20.
  C5  16 384 robots x 256 x 256 at full size, 12 closed-loop cycles: spot robots bit-exact (layer, pseudo-scan,
      commands) against oracle + reference VFH.
  C4-shaped fleet against the reference's OWN node code: every robot is one MapProvider + LaserMapUpdater + Steerer
      + VFH of oracle/_ref/libnav_ref.so (navh_fleet_cycle_samples, the CPU arm bench.py times) - velocity commands of
      the batched device path must equal the Twists the reference nodes publish.
"""
import math

import numpy as np
import pytest

from oracle import navref as N
from oracle import oracle as O
from tests.util import assert_layers_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ros_navigation_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def test_c3_pipeline_large_window(ctx):
    import torch
    from ros_navigation_b200 import VFH, DeviceGridMap, VfhParams, capi, synth
    if not O.have_ref():
        pytest.skip("oracle/_ref missing")
    cfg = synth.CONFIGS["c3"]
    w = synth.Worlds(1, cfg["extent"], synth.config_seed("c3"))
    g = O.make_geom(cfg["extent"], cfg["extent"], cfg["res"])
    dg = DeviceGridMap(ctx, (cfg["extent"],) * 2, cfg["res"], layers=("laser",))
    dg.alias("master", "laser")
    layer = O.new_layer(g)
    v = VFH(ctx, VfhParams(window_diameter=cfg["window"], cell_size=cfg["cell"], submap_length=cfg["submap"]))
    ref = O.RefVFH(window_diameter=cfg["window"], cell_size=cfg["cell"])
    tabs, rtabs = v.tables(0), ref.cell_tables()
    for k in range(3):
        assert np.array_equal(tabs[k], rtabs[k], equal_nan=True), "W=129 table %d" % k
    speed, occupied_windows, blocked_bins = 0, 0, 0
    for step in range(8):
        t = step * 1.5
        x, y, yaw = w.pose(t)
        rng_, ang = w.cast(x, y, yaw, cfg["beams"], cfg["fov"], cfg["range_max"])
        # something inside the 2.58 m window: a block of beams ends 0.4 - 1.2 m from the robot (left / ahead / right)
        lo = (step * 577) % (cfg["beams"] - 400)
        rng_[0, lo:lo + 300] = torch.linspace(0.4, 1.2, 300, dtype=rng_.dtype)
        s8, off = synth.samples_from_scan(x, y, yaw, rng_, ang, cfg["range_max"], keep_max=(step % 3 == 1))
        s = synth.samples_to_numpy(s8)
        O.himm_update(g, layer, s)
        dg.himm_update("laser", s)
        inp = synth.vfh_inputs_to_numpy(synth.vfh_inputs(w, t, 0.1, speed))
        cmd = v.update_from_grid(dg, "master", inp)
        want_r = O.ranges_from_submap(g, layer, float(inp["x"][0]), float(inp["y"][0]), float(inp["yaw"][0]),
                                      cfg["submap"])
        got_r = v.ranges()
        assert np.array_equal(got_r[:, 0], want_r[:, 0]), "pseudo-scan, step %d" % step
        occupied_windows += int((want_r[:, 0] < 5000.0).any())
        cs, ct = ref.update(want_r, speed, float(inp["goal_direction"][0]), float(inp["goal_distance"][0]), 250.0, 0.1)
        st, mine = ref.state(), v.state()
        assert np.array_equal(mine["origin_hist"], st["origin_hist"]), "primary histogram, step %d" % step
        assert np.array_equal(mine["hist"], st["hist"]), "masked histogram, step %d" % step
        assert (int(cmd["speed"]), int(cmd["turnrate"])) == (cs, ct), "command, step %d" % step
        assert np.float32(cmd["picked_angle"]) == np.float32(st["picked"])
        blocked_bins += int((st["hist"] > 0).sum())
        speed = cs
    assert_layers_equal(dg.download("laser"), layer, "c3 layer after 8 scans")
    assert occupied_windows >= 6 and blocked_bins > 0, (occupied_windows, blocked_bins)
    v.close()
    dg.close()


def test_c5_full_size_closed_loop(ctx):
    import torch
    from ros_navigation_b200 import VFH, DeviceGridMap, VfhParams, capi, synth
    cfg = synth.CONFIGS["c5"]
    n = cfg["robots"]
    dev = torch.device("cuda:0")
    w = synth.Worlds(n, cfg["extent"], synth.config_seed("c5"), device=dev)
    dg = DeviceGridMap(ctx, (cfg["extent"],) * 2, cfg["res"], n_robots=n, layers=("laser",))
    dg.alias("master", "laser")
    v = VFH(ctx, VfhParams(window_diameter=cfg["window"], cell_size=cfg["cell"], submap_length=cfg["submap"]),
            n_robots=n)
    g = O.make_geom(cfg["extent"], cfg["extent"], cfg["res"])
    spots = [0, 1, 4095, 8192, 12345, n - 1]
    layers = {r: O.new_layer(g) for r in spots}
    have_ref = O.have_ref()
    refs = {r: O.RefVFH() for r in spots} if have_ref else {}
    speeds = torch.zeros(n, dtype=torch.int32, device=dev)
    cmd_dev = torch.zeros(n, 16, dtype=torch.uint8, device=dev)
    checked = 0
    for step in range(12):
        t = 0.2 * step
        x, y, yaw = w.pose(t)
        rng_, ang = w.cast(x, y, yaw, cfg["beams"], cfg["fov"], cfg["range_max"])
        org, xy, clr, off = synth.cloud_from_scan(x, y, yaw, rng_, ang, cfg["range_max"], keep_max=(step % 4 == 2))
        inp_dev = synth.vfh_inputs(w, t, 0.2, speeds)
        torch.cuda.synchronize()
        dg.himm_update_cloud_batched_dev("laser", org, xy, clr, off, int(off[-1]), cfg["beams"])
        v.update_batched_dev(dg, "master", inp_dev, cmd_dev)
        ctx.synchronize()
        out = cmd_dev.cpu().numpy().view(capi.COMMAND_DTYPE).reshape(-1)
        inp = synth.vfh_inputs_to_numpy(inp_dev)
        offs = off.cpu().numpy()
        for r in spots:
            sl = slice(int(offs[r]), int(offs[r + 1]))
            cnt = sl.stop - sl.start
            xy_r, clr_r = xy[sl].cpu(), clr[sl].cpu()
            O.himm_update(g, layers[r], O.make_samples(np.full(cnt, float(x[r])), np.full(cnt, float(y[r])),
                                                       xy_r[:, 0].double().numpy(), xy_r[:, 1].double().numpy(),
                                                       clr_r.numpy()))
            if have_ref:
                want_r = O.ranges_from_submap(g, layers[r], inp["x"][r], inp["y"][r], inp["yaw"][r])
                assert np.array_equal(v.ranges(robot=r)[:, 0], want_r[:, 0]), "pseudo-scan, step %d robot %d" % (step, r)
                cs, ct = refs[r].update(want_r, int(inp["current_speed"][r]), float(inp["goal_direction"][r]),
                                        float(inp["goal_distance"][r]), 250.0, 0.2)
                assert (int(out["speed"][r]), int(out["turnrate"][r])) == (cs, ct), "step %d robot %d" % (step, r)
                checked += 1
        speeds = torch.from_numpy(out["speed"].astype(np.int32)).to(dev)
    for r in spots:
        assert_layers_equal(dg.download("laser", robot=r), layers[r], "c5 robot %d after 12 cycles" % r)
    assert (not have_ref) or checked == 12 * len(spots)
    v.close()
    dg.close()


@pytest.mark.skipif(not N.have_ref(), reason="oracle/_ref/libnav_ref.so not prebuilt")
def test_fleet_matches_reference_nodes(ctx):
    """The batched device path against a fleet of the reference's own nodes (the CPU arm of bench.py)."""
    import torch
    from ros_navigation_b200 import VFH, DeviceGridMap, capi, synth
    cfg = synth.CONFIGS["c4"]
    n, cycles = 24, 25
    w = synth.Worlds(n, cfg["extent"], 20260117)
    nodes = [N.Node(N.REF_PATH, cfg["extent"], cfg["extent"], False, t0=1.0) for _ in range(n)]
    dg = DeviceGridMap(ctx, (cfg["extent"],) * 2, cfg["res"], n_robots=n, layers=("laser",))
    dg.alias("master", "laser")
    v = VFH(ctx, n_robots=n)
    speeds = np.zeros(n, np.int32)
    moving = 0
    for c in range(cycles):
        t_sim = 0.2 * c
        x, y, yaw = w.pose(t_sim)
        rng_, ang = w.cast(x, y, yaw, cfg["beams"], cfg["fov"], cfg["range_max"])
        s8, off = synth.samples_from_scan(x, y, yaw, rng_, ang, cfg["range_max"], keep_max=(c % 5 == 4))
        samples, offs = synth.samples_to_numpy(s8), off.numpy()
        gx, gy, _ = w.pose(t_sim + 15.0)
        goals = np.stack([gx.numpy(), gy.numpy()], 1)
        poses = np.stack([x.numpy(), y.numpy(), yaw.numpy()], 1)
        # the reference reads its pose back through tf (quaternion -> yaw): feed the device path the same numbers
        inp = np.zeros(n, capi.VFH_INPUT_DTYPE)
        for r in range(n):
            nodes[r].set_frame("base_link", *poses[r])
            rp = nodes[r].robot_pose()
            gdir, gdist = O.goal_from_pose(rp[0], rp[1], rp[2], goals[r, 0], goals[r, 1])
            inp[r] = (rp[0], rp[1], rp[2], 0.2, speeds[r], gdir, gdist, 250.0)
        want = N.fleet_cycle_samples(nodes, 1.0 + 0.2 * (c + 1), poses, samples, offs, goals, speeds / 1000.0, threads=4)
        dg.himm_update_batched("laser", samples, offs)
        out = v.update_batched(dg, "master", inp)
        for r in range(n):
            assert want[r, 0] == float(np.float32(out["speed"][r])) / 1000.0, "speed, cycle %d robot %d" % (c, r)
            assert want[r, 1] == int(out["turnrate"][r]) * math.pi / 180.0, "turn rate, cycle %d robot %d" % (c, r)
        moving += int((out["speed"] > 0).sum())
        speeds = out["speed"].astype(np.int32).copy()
    for r in range(0, n, 5):
        assert_layers_equal(dg.download("laser", robot=r), nodes[r].layer("laser"), "robot %d" % r)
    assert moving > cycles * n // 4
    for nd in nodes:
        nd.close()
    v.close()
    dg.close()
