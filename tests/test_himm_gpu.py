"""HIMM parity: CUDA path (through the C ABI) vs the CPU oracle, bit-exact on identical sample sequences.

Reference behaviour under test: LaserMapUpdater::updateMap (move_control/src/laser_map_updater.cpp:7-21),
MapUpdater::lineOnMap/clearCell/markCell (move_control/include/move_control/map_updater.h:38-71),
grid_map::LineIterator (grid_map_core/src/iterators/LineIterator.cpp).
"""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_layers_equal, lidar_samples, random_samples

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from ros_navigation_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def make_pair(ctx, lx, ly, res, pos=(0.0, 0.0), start=(0, 0), n_robots=1, layers=("laser",)):
    from ros_navigation_b200 import DeviceGridMap
    g = O.make_geom(lx, ly, res, pos[0], pos[1], start)
    dg = DeviceGridMap(ctx, (lx, ly), res, pos, n_robots=n_robots, layers=layers)
    assert (dg.rows, dg.cols) == (g.rows, g.cols)
    if tuple(start) != (0, 0):
        for r in range(n_robots):
            dg.set_geometry(r, pos, start)
    return g, dg


@pytest.mark.parametrize("lx,ly,res,pos,start", [
    (10.0, 10.0, 0.05, (0.0, 0.0), (0, 0)),       # C1 grid: 200 x 200
    (4.0, 4.0, 0.05, (1.3, -0.7), (17, 63)),      # mapTest_vfh moving map, circular buffer shifted
    (6.5, 3.5, 0.05, (-2.0, 5.0), (0, 0)),        # 130 x 70: not a multiple of the tile, rows % 4 != 0
    (12.85, 15.0, 0.05, (0.0, 0.0), (0, 0)),      # 257 x 300
    (8.0, 5.0, 1.0, (0.0, 0.0), (0, 0)),          # the grid_map test geometry
])
def test_random_samples_bit_exact(ctx, lx, ly, res, pos, start):
    rng = np.random.default_rng(1)
    g, dg = make_pair(ctx, lx, ly, res, pos, start)
    layer = O.new_layer(g)
    bb_o = np.zeros(4)
    bb_d = np.zeros(4)
    for batch in range(6):
        n = [1, 7, 360, 1080, 33, 500][batch]
        s = random_samples(rng, g, n) if batch % 2 == 0 else \
            lidar_samples(rng, g, (pos[0] + 0.3, pos[1] - 0.2), n, 0.2, min(lx, ly) * 0.7, clear_frac=0.1)
        O.himm_update(g, layer, s, bb_o)
        dg.himm_update("laser", s, bbox=bb_d)
        assert_layers_equal(dg.download("laser"), layer, "batch %d" % batch)
        assert np.array_equal(bb_o, bb_d)
    vals = layer[~np.isnan(layer)]
    assert set(np.unique(vals)).issubset(set(np.arange(0, 190, 10.0)))
    dg.close()


def test_order_dependence_same_cell(ctx):
    """Many beams ending in / crossing the same few cells: the saturating clear/mark sequence must be replayed in
    sample order.  Includes repeats of the same beam and alternating clear_end."""
    rng = np.random.default_rng(2)
    g, dg = make_pair(ctx, 10.0, 10.0, 0.05)
    layer = O.new_layer(g)
    for rep in range(8):
        n = 400
        # a tight bundle of beams from one origin to nearly the same end point, plus a few crossing beams
        ex = 1.0 + 0.12 * rng.random(n)
        ey = 2.0 + 0.12 * rng.random(n)
        ce = (rng.random(n) < 0.35).astype(np.int32)
        s = O.make_samples(np.full(n, -2.0), np.full(n, -1.0), ex.astype(np.float32), ey.astype(np.float32), ce)
        cross = O.make_samples(ex[:50] - 1.0, ey[:50] + 1.0, ex[:50] + 0.5, ey[:50] - 0.5, np.zeros(50, np.int32))
        allS = np.concatenate([s[:200], cross, s[200:]])
        O.himm_update(g, layer, allS)
        dg.himm_update("laser", allS)
        assert_layers_equal(dg.download("laser"), layer, "rep %d" % rep)
    dg.close()


def test_chunking_more_beams_than_list_capacity(ctx):
    rng = np.random.default_rng(3)
    g, dg = make_pair(ctx, 10.0, 10.0, 0.05)
    layer = O.new_layer(g)
    s = np.concatenate([lidar_samples(rng, g, (0.5, 0.5), 3000, 0.1, 4.0), random_samples(rng, g, 3500)])
    O.himm_update(g, layer, s)
    dg.himm_update("laser", s)
    assert_layers_equal(dg.download("laser"), layer, "7000 samples in one batch")
    dg.close()


def test_very_long_single_batch_is_split():
    """A single-robot batch whose binning scratch would exceed the budget is applied as consecutive launches.  Run in
    a child process with a 1 MiB budget (2048 x 2048 grid: 256 KiB of masks per 2048 samples -> 8192 samples per
    launch) so that 20000 samples need three launches."""
    import subprocess
    import sys
    code = """
import numpy as np
from oracle import oracle as O
from ros_navigation_b200 import DeviceGridMap, capi
from tests.util import assert_layers_equal, lidar_samples, random_samples
ctx = capi.Context(0)
rng = np.random.default_rng(33)
g = O.make_geom(102.4, 102.4, 0.05)
dg = DeviceGridMap(ctx, (102.4, 102.4), 0.05, layers=("laser",))
layer = O.new_layer(g)
s = np.concatenate([lidar_samples(rng, g, (1.0, -2.0), 17000, 0.5, 20.0, clear_frac=0.1), random_samples(rng, g, 3000, spread=0.3)])
O.himm_update(g, layer, s)
dg.himm_update("laser", s)
assert_layers_equal(dg.download("laser"), layer, "20000 samples")
print("SPLIT_OK")
"""
    import os
    env = dict(os.environ, B200NAV_MASK_BUDGET_MB="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert "SPLIT_OK" in out.stdout, out.stdout + out.stderr


def test_several_scans_and_sonar_rays_in_one_update(ctx):
    """One update may carry several messages (different origins) - LaserMapUpdater keeps buffering between updates -
    and the sonar updaters push one ray per message into layer "range" (range_map_updater.cpp:38-76)."""
    rng = np.random.default_rng(9)
    g, dg = make_pair(ctx, 10.0, 10.0, 0.05, layers=("laser", "range"))
    laser, sonar = O.new_layer(g), O.new_layer(g)
    for cycle in range(5):
        parts = [lidar_samples(rng, g, (rng.uniform(-2, 2), rng.uniform(-2, 2)), n, 0.2, 4.0, clear_frac=0.05)
                 for n in (360, 17, 360, 45)]
        s = np.concatenate(parts)          # batches of 32 straddle the scans: fan and general schedules interleave
        O.himm_update(g, laser, s)
        dg.himm_update("laser", s)
        rays = random_samples(rng, g, 5 * 4, spread=0.9, clear_frac=0.3)   # 5 sonars, a few messages each
        O.himm_update(g, sonar, rays)
        dg.himm_update("range", rays)
    assert_layers_equal(dg.download("laser"), laser, "laser")
    assert_layers_equal(dg.download("range"), sonar, "range")
    dg.close()


def test_edge_cases(ctx):
    g, dg = make_pair(ctx, 8.0, 5.0, 1.0)
    layer = O.new_layer(g)
    dg.himm_update("laser", np.zeros(0, O.SAMPLE_DTYPE))            # empty batch
    assert_layers_equal(dg.download("laser"), layer, "empty")
    inf, nan = np.inf, np.nan
    s = O.make_samples(
        [0.0, -8.0, 0.0, 0.0, 2.2, -7.0, 0.0, nan, 3.9],
        [0.0, 8.0, 0.0, 0.0, 1.2, -9.0, 0.0, 0.0, 2.4],
        [9.0, 8.0, 0.0, inf, 2.2, 8.0, nan, 1.0, 30.0],
        [6.0, 8.0, 0.0, 1.0, 1.2, 8.0, 0.0, 1.0, 2.4],
        [0, 0, 0, 0, 1, 0, 0, 0, 0])
    # end outside | no intersection | start==end | inf end | start==end clear | both outside | NaN end | NaN start |
    # start inside, end far outside along a row
    O.himm_update(g, layer, s)
    dg.himm_update("laser", s)
    assert_layers_equal(dg.download("laser"), layer, "edge cases")
    dg.close()


def test_uploaded_foreign_values(ctx):
    """Layers uploaded from the host may hold values outside the HIMM set; clear/mark must still match."""
    rng = np.random.default_rng(4)
    g, dg = make_pair(ctx, 6.0, 6.0, 0.05)
    layer = (rng.random((g.cols, g.rows)) * 400 - 100).astype(np.float32)
    layer[rng.random(layer.shape) < 0.2] = np.nan
    layer[rng.random(layer.shape) < 0.1] = 155.5
    dg.upload("laser", layer)
    for _ in range(3):
        s = lidar_samples(rng, g, (0.1, 0.2), 720, 0.3, 2.9, clear_frac=0.05)
        O.himm_update(g, layer, s)
        dg.himm_update("laser", s)
    assert_layers_equal(dg.download("laser"), layer, "foreign values")
    dg.close()


def test_batched_ragged_robots(ctx):
    rng = np.random.default_rng(5)
    n_robots = 7
    g, dg = make_pair(ctx, 12.8, 12.8, 0.05, n_robots=n_robots)
    layers = [O.new_layer(g) for _ in range(n_robots)]
    for cycle in range(4):
        per = []
        for r in range(n_robots):
            n = [0, 1, 1080, 37, 2500, 360, 5][(r + cycle) % 7]
            per.append(lidar_samples(rng, g, (rng.random() * 4 - 2, rng.random() * 4 - 2), n, 0.2, 6.0,
                                     fov=1.5 * np.pi, clear_frac=0.05) if n else np.zeros(0, O.SAMPLE_DTYPE))
        offsets = np.zeros(n_robots + 1, np.int32)
        offsets[1:] = np.cumsum([len(p) for p in per])
        allS = np.concatenate(per)
        bb_d = np.zeros((n_robots, 4))
        dg.himm_update_batched("laser", allS, offsets, bbox=bb_d)
        for r in range(n_robots):
            bb_o = np.zeros(4)
            O.himm_update(g, layers[r], per[r], bb_o)
            assert_layers_equal(dg.download("laser", robot=r), layers[r], "cycle %d robot %d" % (cycle, r))
            assert np.array_equal(bb_o, bb_d[r])
    dg.close()


def test_free_tile_shortcut_sees_external_writes(ctx):
    """The tile kernel skips tiles whose beams only clear columns known to be all 0.  Every other writer of the layer
    (upload, clear, move, copy) must invalidate that knowledge: interleave them with clear-only scans."""
    rng = np.random.default_rng(8)
    g, dg = make_pair(ctx, 12.8, 12.8, 0.05, layers=("laser", "master"))
    layer = O.new_layer(g)
    origin = (0.3, -0.2)

    def scans(n, clear_frac):
        for _ in range(n):
            s = lidar_samples(rng, g, origin, 720, 3.0, 5.5, clear_frac=clear_frac)
            O.himm_update(g, layer, s)
            dg.himm_update("laser", s)

    scans(25, 0.0)                      # marks at the ends, everything in between becomes 0
    scans(3, 1.0)                       # clear-only: large all-free regions are now known and skipped
    assert_layers_equal(dg.download("laser"), layer, "after warm-up")
    # (1) upload: plant obstacles and unknown cells inside the free region
    ys, xs = np.nonzero(layer == 0.0)
    pick = rng.choice(len(ys), 400, replace=False)
    layer[ys[pick[:200]], xs[pick[:200]]] = 90.0
    layer[ys[pick[200:]], xs[pick[200:]]] = np.nan
    dg.upload("laser", layer)
    scans(2, 1.0)
    assert_layers_equal(dg.download("laser"), layer, "after upload")
    # (2) copy from another layer
    other = layer.copy()
    other[other == 0.0] = 40.0
    dg.upload("master", other)
    dg.copy_layer("laser", "master")
    layer[:] = other
    scans(2, 1.0)
    assert_layers_equal(dg.download("laser"), layer, "after copy_layer")
    # (3) move: strips become NaN
    scans(6, 1.0)
    assert O.move(g, [layer], 1.0, -0.6) == dg.move((1.0, -0.6))
    scans(2, 1.0)
    assert_layers_equal(dg.download("laser"), layer, "after move")
    # (4) clear
    dg.clear("laser")
    layer[:] = np.nan
    scans(2, 0.5)
    assert_layers_equal(dg.download("laser"), layer, "after clear")
    dg.close()


def test_cloud_form_matches_sample_form(ctx):
    """b200nav_himm_update_cloud_batched (origin per robot + float32 points) == the RangeSample form, bit for bit."""
    rng = np.random.default_rng(6)
    n_robots = 5
    g, dg = make_pair(ctx, 12.8, 12.8, 0.05, n_robots=n_robots)
    layers = [O.new_layer(g) for _ in range(n_robots)]
    for cycle in range(3):
        counts = [(0, 1080, 7, 333, 2100)[(r + cycle) % 5] for r in range(n_robots)]
        offsets = np.zeros(n_robots + 1, np.int32)
        offsets[1:] = np.cumsum(counts)
        origins = rng.uniform(-3, 3, (n_robots, 2))
        xy = rng.uniform(-8, 8, (offsets[-1], 2)).astype(np.float32)
        clear = (rng.random(offsets[-1]) < 0.1).astype(np.uint8)
        bb_d = np.zeros((n_robots, 4))
        dg.himm_update_cloud_batched("laser", origins, xy, clear if cycle != 1 else None, offsets, bbox=bb_d)
        for r in range(n_robots):
            sl = slice(offsets[r], offsets[r + 1])
            n = counts[r]
            s = O.make_samples(np.full(n, origins[r, 0]), np.full(n, origins[r, 1]), xy[sl, 0].astype(np.float64),
                               xy[sl, 1].astype(np.float64), clear[sl] if cycle != 1 else np.zeros(n, np.int32))
            bb_o = np.zeros(4)
            O.himm_update(g, layers[r], s, bb_o)
            assert_layers_equal(dg.download("laser", robot=r), layers[r], "cycle %d robot %d" % (cycle, r))
            assert np.array_equal(bb_o, bb_d[r])
    dg.close()


def test_cloud_form_pipelined_copy_groups(ctx):
    """Large host-side cloud batches are copied in groups of robots on a second stream, pipelined with the prep kernel
    (>= 65536 points and >= 8 robots): same result, ragged groups included."""
    rng = np.random.default_rng(16)
    n_robots = 100
    g, dg = make_pair(ctx, 12.8, 12.8, 0.05, n_robots=n_robots)
    layers = [O.new_layer(g) for _ in range(n_robots)]
    for cycle in range(2):
        counts = [(1080, 0, 1500, 700, 1080, 3)[(r + cycle) % 6] for r in range(n_robots)]
        offsets = np.zeros(n_robots + 1, np.int32)
        offsets[1:] = np.cumsum(counts)
        assert offsets[-1] >= 65536
        origins = rng.uniform(-3, 3, (n_robots, 2))
        ang = rng.uniform(-np.pi, np.pi, offsets[-1])
        rad = rng.uniform(0.2, 6.0, offsets[-1])
        own = np.repeat(np.arange(n_robots), counts)
        xy = np.stack([origins[own, 0] + rad * np.cos(ang), origins[own, 1] + rad * np.sin(ang)], 1).astype(np.float32)
        clear = (rng.random(offsets[-1]) < 0.05).astype(np.uint8)
        dg.himm_update_cloud_batched("laser", origins, xy, clear, offsets)
        for r in range(0, n_robots, 3):
            sl = slice(offsets[r], offsets[r + 1])
            s = O.make_samples(np.full(counts[r], origins[r, 0]), np.full(counts[r], origins[r, 1]),
                               xy[sl, 0].astype(np.float64), xy[sl, 1].astype(np.float64), clear[sl])
            O.himm_update(g, layers[r], s)
            assert_layers_equal(dg.download("laser", robot=r), layers[r], "cycle %d robot %d" % (cycle, r))
    dg.close()


def test_async_cycles_overlap_and_match_synchronous_calls(ctx):
    """b200nav_himm_update_cloud_batched_async + b200nav_vfh_update_batched_async from pinned host buffers, several
    cycles in flight (the cloud copy of cycle i+1 overlaps the kernels of cycle i): grids and commands equal those of
    the synchronous calls on a second grid."""
    import torch
    from ros_navigation_b200 import VFH, DeviceGridMap, capi
    rng = np.random.default_rng(17)
    n_robots, beams, cycles = 96, 720, 6
    grids = [DeviceGridMap(ctx, (12.8, 12.8), 0.05, n_robots=n_robots, layers=("laser",)) for _ in range(2)]
    vfhs = [VFH(ctx, n_robots=n_robots) for _ in range(2)]
    data, tickets = [], []
    for c in range(cycles):
        origins = rng.uniform(-3, 3, (n_robots, 2))
        offsets = np.arange(n_robots + 1, dtype=np.int32) * beams
        assert offsets[-1] >= 65536
        ang = np.tile(np.linspace(-np.pi, np.pi, beams, endpoint=False), n_robots)
        rad = rng.uniform(0.3, 5.0, offsets[-1])
        own = np.repeat(np.arange(n_robots), beams)
        xy = np.stack([origins[own, 0] + rad * np.cos(ang), origins[own, 1] + rad * np.sin(ang)], 1).astype(np.float32)
        clear = (rng.random(offsets[-1]) < 0.05).astype(np.uint8)
        inp = np.zeros(n_robots, capi.VFH_INPUT_DTYPE)
        inp["x"], inp["y"], inp["yaw"], inp["dt"] = origins[:, 0], origins[:, 1], rng.uniform(-3, 3, n_robots), 0.2
        inp["current_speed"], inp["goal_direction"], inp["goal_distance"], inp["goal_tolerance"] = 100, 90.0, 3000.0, 250.0
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        data.append(dict(origins=pin(origins), xy=pin(xy), clear=pin(clear), offsets=pin(offsets),
                         inp=pin(inp.view(np.uint8).reshape(n_robots, -1)), inp_np=inp,
                         out=torch.zeros(n_robots, 16, dtype=torch.uint8).pin_memory()))
    for c, d in enumerate(data):  # enqueue everything without waiting
        grids[0].himm_update_cloud_batched_async("laser", d["origins"], d["xy"], d["clear"], d["offsets"])
        vfhs[0].update_batched_async(grids[0], "laser", d["inp"], d["out"])
        tickets.append(ctx.fence())
        if c == 2:   # stream-ordered writers / state readers in the middle of the pipeline (they join the side stream)
            grids[0].move((0.5, -0.25), robot=2)
            mid_state = vfhs[0].state(robot=3)
    want = []
    for c, d in enumerate(data):
        grids[1].himm_update_cloud_batched("laser", d["origins"].numpy(), d["xy"].numpy(), d["clear"].numpy(),
                                           d["offsets"].numpy())
        want.append(vfhs[1].update_batched(grids[1], "laser", d["inp_np"]))
        if c == 2:
            grids[1].move((0.5, -0.25), robot=2)
            ref_state = vfhs[1].state(robot=3)
            for k in ref_state:
                assert np.array_equal(np.asarray(mid_state[k]), np.asarray(ref_state[k])), k
    for c in (3, 0, 5):
        ctx.wait(tickets[c])
        got = data[c]["out"].numpy().view(capi.COMMAND_DTYPE).reshape(-1)
        assert np.array_equal(got, want[c]), "commands of cycle %d" % c
    ctx.synchronize()
    for c in range(cycles):
        assert np.array_equal(data[c]["out"].numpy().view(capi.COMMAND_DTYPE).reshape(-1), want[c])
    for r in range(0, n_robots, 7):
        assert_layers_equal(grids[0].download("laser", robot=r), grids[1].download("laser", robot=r), "robot %d" % r)
    with pytest.raises(capi.B200NavError):
        ctx.wait(10 ** 6)
    for x in vfhs + grids:
        x.close()


def test_long_steady_state_fleet_sequence(ctx):
    """150 cycles of the bench's own synthetic workload (rooms with discs and boxes, 1080-beam noisy scans along
    Lissajous paths) for a small fleet: walls saturate, free space is re-cleared every cycle, marks and clears keep
    meeting in the same cells - the steady state the throughput numbers are measured in.  Grids bit-exact against the
    oracle half way and at the end; the closed loop (HIMM -> window -> VFH+) equals oracle + reference VFH throughout."""
    import torch
    from ros_navigation_b200 import VFH, DeviceGridMap, synth
    n, cycles, beams, rmax = 6, 150, 1080, 6.0
    W = synth.Worlds(n, 12.8, seed=4242)
    g = O.make_geom(12.8, 12.8, 0.05)
    dg = DeviceGridMap(ctx, (12.8, 12.8), 0.05, n_robots=n, layers=("laser",))
    dg.alias("master", "laser")
    layers = [O.new_layer(g) for _ in range(n)]
    have_ref = O.have_ref()
    refs = [O.RefVFH() for _ in range(n)] if have_ref else None
    v = VFH(ctx, n_robots=n)
    speeds = np.zeros(n, np.int32)
    for c in range(cycles):
        t = 0.2 * c
        x, y, yaw = W.pose(t)
        r, ang = W.cast(x, y, yaw, beams, 1.5 * np.pi, rmax)
        org, xy, clr, off = synth.cloud_from_scan(x, y, yaw, r, ang, rmax, keep_max=(c % 7 == 3))
        dg.himm_update_cloud_batched("laser", org.numpy(), xy.numpy(), clr.numpy(), off.numpy())
        offs = off.numpy()
        for k in range(n):
            sl = slice(offs[k], offs[k + 1])
            cnt = offs[k + 1] - offs[k]
            O.himm_update(g, layers[k], O.make_samples(np.full(cnt, float(x[k])), np.full(cnt, float(y[k])),
                                                       xy[sl, 0].double().numpy(), xy[sl, 1].double().numpy(),
                                                       clr[sl].numpy()))
        inp = synth.vfh_inputs_to_numpy(synth.vfh_inputs(W, t, 0.2, torch.from_numpy(speeds)))
        out = v.update_batched(dg, "master", inp)
        if have_ref and c % 5 == 0:
            for k in range(n):
                want_r = O.ranges_from_submap(g, layers[k], inp["x"][k], inp["y"][k], inp["yaw"][k])
                rcs, rct = refs[k].update(want_r, int(speeds[k]), float(inp["goal_direction"][k]),
                                          float(inp["goal_distance"][k]), 250.0, 0.2)
                assert (int(out["speed"][k]), int(out["turnrate"][k])) == (rcs, rct), "cycle %d robot %d" % (c, k)
        elif have_ref:
            for k in range(n):   # keep the reference's state (hysteresis, last picked angle, speed) in step
                refs[k].update(O.ranges_from_submap(g, layers[k], inp["x"][k], inp["y"][k], inp["yaw"][k]),
                               int(speeds[k]), float(inp["goal_direction"][k]), float(inp["goal_distance"][k]), 250.0, 0.2)
        speeds = out["speed"].astype(np.int32).copy()
        if c in (cycles // 2, cycles - 1):
            for k in range(n):
                assert_layers_equal(dg.download("laser", robot=k), layers[k], "cycle %d robot %d" % (c, k))
    sat = sum(int((np.nan_to_num(l) >= 150).sum()) for l in layers)
    assert sat > 200, "the sequence should reach saturated wall cells (%d)" % sat
    v.close()
    dg.close()


@pytest.mark.parametrize("n_robots,beams", [(1, 360), (1, 1080), (5, 720)])
def test_uneven_beam_lengths_saturating_walls(ctx, n_robots, beams):
    """Scans whose neighbouring beams differ wildly in length (a comb of near and far returns), replayed until the
    walls saturate: consecutive 32-beam batches of a tile then cover very different step ranges, which is what the
    multi-warp kernel's wavefront pipeline has to keep in order (a short batch must not let its successor overtake
    an earlier, longer one)."""
    rng = np.random.default_rng(beams + n_robots)
    g = O.make_geom(10.0, 10.0, 0.05)
    from ros_navigation_b200 import DeviceGridMap
    dg = DeviceGridMap(ctx, (10.0, 10.0), 0.05, n_robots=n_robots, layers=("laser",))
    layers = [O.new_layer(g) for _ in range(n_robots)]
    centre = rng.uniform(-1.0, 1.0, (n_robots, 2))
    comb = np.where(np.arange(beams) // rng.integers(3, 40) % 2 == 0, 0.35, 2.9)   # near / far teeth
    for cycle in range(90):
        per = []
        for r in range(n_robots):
            origin = centre[r] + 0.02 * cycle * np.array([np.cos(0.1 * cycle + r), np.sin(0.07 * cycle)])
            th = -np.pi + 2 * np.pi * np.arange(beams) / beams
            rad = comb * (1.0 + 0.02 * rng.standard_normal(beams))
            if cycle % 9 == 4:
                rad = np.roll(rad, int(rng.integers(1, 60)))
            ex = (origin[0] + rad * np.cos(th)).astype(np.float32).astype(np.float64)
            ey = (origin[1] + rad * np.sin(th)).astype(np.float32).astype(np.float64)
            ce = (rng.random(beams) < 0.03).astype(np.int32)
            per.append(O.make_samples(np.full(beams, origin[0]), np.full(beams, origin[1]), ex, ey, ce))
        off = np.concatenate([[0], np.cumsum([len(p) for p in per])]).astype(np.int32)
        dg.himm_update_batched("laser", np.concatenate(per), off)
        for r in range(n_robots):
            O.himm_update(g, layers[r], per[r])
        if cycle % 15 == 14:
            for r in range(n_robots):
                assert_layers_equal(dg.download("laser", robot=r), layers[r], "cycle %d robot %d" % (cycle, r))
    assert max(float(np.nanmax(l)) for l in layers) >= 170.0
    dg.close()


def test_c2_sized_grid_scan_sequence(ctx):
    """BASELINE config 2 geometry: 2048 x 2048 @ 5 cm, 1080-beam / 270 deg scans up to 30 m."""
    import torch
    from ros_navigation_b200 import synth
    cfg = synth.CONFIGS["c2"]
    w = synth.Worlds(1, cfg["extent"], synth.config_seed("c2"))
    g, dg = make_pair(ctx, cfg["extent"], cfg["extent"], cfg["res"])
    layer = O.new_layer(g)
    for step in range(12):
        x, y, yaw = w.pose(step * 1.0)
        rng_, ang = w.cast(x, y, yaw, cfg["beams"], cfg["fov"], cfg["range_max"])
        s8, off = synth.samples_from_scan(x, y, yaw, rng_, ang, cfg["range_max"], keep_max=(step % 3 == 0))
        s = synth.samples_to_numpy(s8)
        O.himm_update(g, layer, s)
        dg.himm_update("laser", s)
    assert_layers_equal(dg.download("laser"), layer, "c2 sequence")
    dg.close()


def test_c3_sized_grid_few_scans(ctx):
    """BASELINE config 3 geometry: 8192 x 8192 @ 2 cm, 4096 beams up to 60 m (256 MiB layer)."""
    from ros_navigation_b200 import synth
    cfg = synth.CONFIGS["c3"]
    w = synth.Worlds(1, cfg["extent"], synth.config_seed("c3"))
    g, dg = make_pair(ctx, cfg["extent"], cfg["extent"], cfg["res"])
    layer = O.new_layer(g)
    for step in range(3):
        x, y, yaw = w.pose(step * 2.0)
        rng_, ang = w.cast(x, y, yaw, cfg["beams"], cfg["fov"], cfg["range_max"])
        s8, off = synth.samples_from_scan(x, y, yaw, rng_, ang, cfg["range_max"])
        s = synth.samples_to_numpy(s8)
        O.himm_update(g, layer, s)
        dg.himm_update("laser", s)
    assert_layers_equal(dg.download("laser"), layer, "c3 sequence")
    dg.close()


def test_grid_with_more_tiles_than_the_binning_bitmap(ctx):
    """9024 x 8320 cells = 141 x 130 = 18 330 tiles per robot: more than the binning kernel's per-CTA tile bitmap
    (16 384), so every first-touch probe goes straight to global memory; also not a whole number of tiles in x."""
    rng = np.random.default_rng(77)
    g, dg = make_pair(ctx, 180.48, 166.4, 0.02)
    layer = O.new_layer(g)
    for step, origin in enumerate([(10.0, -20.0), (-60.0, 55.0), (85.0, 80.0)]):
        s = lidar_samples(rng, g, origin, 2000, 0.5, 70.0, clear_frac=0.1)
        O.himm_update(g, layer, s)
        dg.himm_update("laser", s)
    assert_layers_equal(dg.download("laser"), layer, "large-grid sequence")
    assert np.nansum(layer) > 0
    dg.close()


def test_c4_full_size_properties(ctx):
    """BASELINE config 4 at full size (1024 robots x 512 x 512, device-resident samples): spot robots are compared
    with the oracle bit-for-bit; all robots satisfy the size-independent HIMM invariants (value set {NaN,0..180}
    and: re-applying the same scan 18 times drives every ray cell that is not an end cell to 0)."""
    import torch
    from ros_navigation_b200 import DeviceGridMap, synth
    cfg = synth.CONFIGS["c4"]
    n = cfg["robots"]
    dev = torch.device("cuda:0")
    w = synth.Worlds(n, cfg["extent"], synth.config_seed("c4"), device=dev)
    dg = DeviceGridMap(ctx, (cfg["extent"], cfg["extent"]), cfg["res"], n_robots=n, layers=("laser",))
    g = O.make_geom(cfg["extent"], cfg["extent"], cfg["res"])
    spots = [0, 1, 511, 1023]
    layers = {r: O.new_layer(g) for r in spots}
    for step in range(3):
        x, y, yaw = w.pose(step * 0.2)
        rng_, ang = w.cast(x, y, yaw, cfg["beams"], cfg["fov"], cfg["range_max"])
        s8, off = synth.samples_from_scan(x, y, yaw, rng_, ang, cfg["range_max"])
        torch.cuda.synchronize()
        dg.himm_update_batched_dev("laser", s8, off, int(off[-1]), cfg["beams"])
        ctx.synchronize()
        sn, offn = synth.samples_to_numpy(s8), off.cpu().numpy()
        for r in spots:
            O.himm_update(g, layers[r], sn[offn[r]:offn[r + 1]])
    for r in spots:
        assert_layers_equal(dg.download("laser", robot=r), layers[r], "c4 robot %d" % r)
    # invariants over all robots, checked on the device layer through torch (plumbing only)
    for _ in range(18):
        dg.himm_update_batched_dev("laser", s8, off, int(off[-1]), cfg["beams"])
    ctx.synchronize()
    for r in spots:
        for _ in range(18):
            O.himm_update(g, layers[r], sn[offn[r]:offn[r + 1]])
        assert_layers_equal(dg.download("laser", robot=r), layers[r], "c4 robot %d after 18 repeats" % r)
    sample_robots = list(range(0, n, 97))
    for r in sample_robots:
        lay = dg.download("laser", robot=r)
        vals = np.unique(lay[~np.isnan(lay)])
        assert set(vals).issubset(set(np.arange(0, 190, 10.0))), (r, vals)
    dg.close()


def test_free_block_summaries_never_claim_a_non_free_block(ctx):
    """The tile kernel drops beam batches that only re-clear blocks its per-tile summary calls free.  The summary must
    therefore never claim a block that holds anything but 0.0 - checked against the downloaded layer after a long
    sequence with moves, uploads and clears in between (every other writer of the layer must reset it)."""
    from ros_navigation_b200 import DeviceGridMap, synth
    from ros_navigation_b200.capi import lib
    n, beams, rmax = 40, 1080, 6.0     # > 32 robots: the one-warp kernel (the one that keeps block summaries)
    W = synth.Worlds(n, 12.8, seed=99)
    dg = DeviceGridMap(ctx, (12.8, 12.8), 0.05, n_robots=n, layers=("laser",))
    nt = ((dg.rows + 63) // 64) * ((dg.cols + 63) // 64)
    tiles_r = (dg.rows + 63) // 64

    def check(tag):
        claimed = 0
        for robot in (0, 7, n - 1):
            lay = dg.download("laser", robot=robot)            # [col][row]
            out = np.zeros(nt, np.uint64)
            assert lib().b200nav_himm_debug_free_summary(dg.h, b"laser", robot, out.ctypes.data, nt) == nt
            for t in range(nt):
                tc, tr = t // tiles_r, t % tiles_r
                w64 = int(out[t])
                for bit in range(64):
                    if (w64 >> bit) & 1:
                        bc, br = bit // 8, bit % 8
                        blk = lay[tc * 64 + bc * 8:tc * 64 + bc * 8 + 8, tr * 64 + br * 8:tr * 64 + br * 8 + 8]
                        assert blk.size == 0 or np.all(blk == 0.0), (tag, robot, t, bc, br)
                        claimed += 1
        return claimed

    total = 0
    spots = (0, 7, n - 1)     # the same robots are also followed by the oracle: batches dropped as "only re-clears
    geoms = {r: O.make_geom(12.8, 12.8, 0.05) for r in spots}   # free blocks" must not change a single cell
    layers = {r: O.new_layer(geoms[r]) for r in spots}
    for c in range(40):
        x, y, yaw = W.pose(0.2 * c)
        r, ang = W.cast(x, y, yaw, beams, 1.5 * np.pi, rmax)
        org, xy, clr, off = synth.cloud_from_scan(x, y, yaw, r, ang, rmax)
        dg.himm_update_cloud_batched("laser", org.numpy(), xy.numpy(), clr.numpy(), off.numpy())
        offs = off.numpy()
        for k in spots:
            sl = slice(offs[k], offs[k + 1])
            cnt = offs[k + 1] - offs[k]
            O.himm_update(geoms[k], layers[k], O.make_samples(np.full(cnt, float(x[k])), np.full(cnt, float(y[k])),
                                                              xy[sl, 0].double().numpy(), xy[sl, 1].double().numpy(),
                                                              clr[sl].numpy()))
        if c == 15:
            dg.move((0.4, -0.3), robot=7)
            O.move(geoms[7], [layers[7]], 0.4, -0.3)
        if c == 22:
            lay = dg.download("laser", robot=0)
            lay[100:140, 90:130] = 50.0
            dg.upload("laser", lay, robot=0)
            layers[0][100:140, 90:130] = 50.0
        if c == 30:
            dg.clear("laser")
            for k in spots:
                layers[k][...] = np.nan
        if c % 8 == 7:
            total += check("cycle %d" % c)
            for k in spots:
                assert_layers_equal(dg.download("laser", robot=k), layers[k], "cycle %d robot %d" % (c, k))
    # FLOAT layers only ever record whole tiles, and so does the multi-warp kernel when it is forced on a fleet
    if dg.layer_format("laser") == "coded" and os.environ.get("B200NAV_MW_HEAVY") != "2":
        assert total > 100, "the summaries never recorded a free block"
    dg.close()


def test_random_geometries_and_boundary_endpoints(ctx):
    """The GPU side of tests/test_reference_pin.py::test_himm_core_random_geometries_and_boundary_endpoints:
    resolutions that do not divide the lengths, far-away map centres, end points exactly on cell borders."""
    rng = np.random.default_rng(20261017)
    for case in range(16):
        lx, ly = rng.uniform(0.6, 9.0, 2)
        res = float(rng.choice([0.05, 0.03, 0.07, 0.1, 0.125, 0.02]))
        px, py = rng.uniform(-2000.0, 2000.0, 2)
        g, dg = make_pair(ctx, float(lx), float(ly), res, (float(px), float(py)))
        layer = O.new_layer(g)
        Lx, Ly = g.rows * res, g.cols * res
        for it in range(4):
            n = 160
            s = random_samples(rng, g, n, spread=0.97, clear_frac=0.25)
            kx, ky = rng.integers(0, g.rows + 1, n), rng.integers(0, g.cols + 1, n)
            on_border = rng.random(n) < 0.34
            s["ex"] = np.where(on_border, (px + Lx / 2) - kx * res, s["ex"])
            s["ey"] = np.where(on_border, (py + Ly / 2) - ky * res, s["ey"])
            outside = rng.random(n) < 0.3
            s["ex"] = np.where(outside & ~on_border, px + (rng.random(n) - 0.5) * Lx * 2.2, s["ex"])
            s["ey"] = np.where(outside & ~on_border, py + (rng.random(n) - 0.5) * Ly * 2.2, s["ey"])
            b1, b2 = np.zeros(4), np.zeros(4)
            O.himm_update(g, layer, s, b1)
            dg.himm_update("laser", s, bbox=b2)
            assert np.array_equal(b1, b2), "bbox, case %d" % case
        assert_layers_equal(dg.download("laser"), layer, "case %d: %r x %r @ %r at (%r, %r)" % (case, lx, ly, res, px, py))
        dg.close()
