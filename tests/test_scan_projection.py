"""LaserScan -> RangeSample projection (spec in ros_navigation_b200/csrc/scan_project.h).
CPU part: the oracle restatement against plain numpy, the simplifyLaserScan selection against hand-derived cases and
against the C ABI's host helper.  GPU part: the scan form of the HIMM update (projection fused into the binning
kernel) against oracle projection + oracle HIMM, bit for bit."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_layers_equal


def test_sincos_is_accurate_and_exactly_periodic_in_quadrants():
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.linspace(-7, 7, 4001), rng.uniform(-3e4, 3e4, 4000), [0.0, np.pi / 2, -np.pi / 4]])
    err = 0.0
    for x in xs:
        s, c = O.sincos(x)
        err = max(err, abs(s - np.sin(x)), abs(c - np.cos(x)))
    assert err < 4e-16
    assert O.sincos(0.0) == (0.0, 1.0)


def test_scan_select_follows_simplifyLaserScan():
    """laser_map_updater.cpp:114-144: ranges[0], then every index at which the float accumulator reaches 0.017; the
    projection then uses the LAST accumulated value as angle_increment; scans with increment >= 0.017 are untouched
    (:74)."""
    sel, used = O.scan_select(np.float32(0.01), 10)
    assert list(sel) == [0, 1, 3, 5, 7, 9] and used == np.float32(np.float32(0.01) + np.float32(0.01))
    sel, used = O.scan_select(np.float32(0.0175), 5)          # >= 0.017: kept as is
    assert list(sel) == [0, 1, 2, 3, 4] and used == np.float32(0.0175)
    sel, used = O.scan_select(np.float32(1.5 * np.pi / 1080), 1080)   # the 270 degree / 1080 beam lidar of C2, C4, C5
    assert len(sel) == 271 and list(sel[:4]) == [0, 3, 7, 11]
    acc = np.float32(0)
    for _ in range(4):
        acc = np.float32(acc + np.float32(1.5 * np.pi / 1080))
    assert used == acc
    sel, used = O.scan_select(np.float32(0.01), 10, decimate=False)
    assert list(sel) == list(range(10)) and used == np.float32(0.01)
    assert len(O.scan_select(np.float32(0.001), 0)[0]) == 0


def test_c_abi_scan_select_equals_oracle():
    from ros_navigation_b200 import capi
    L = capi.lib()
    for inc, n, dec in [(0.01, 10, 1), (1.5 * np.pi / 1080, 1080, 1), (2 * np.pi / 360, 360, 1), (0.004, 4096, 0),
                        (0.0169999, 33, 1), (2 * np.pi / 4096, 4096, 1)]:
        info = np.zeros(1, capi.SCAN_INFO_DTYPE)
        info["angle_increment"], info["n_ranges"], info["decimate"] = inc, n, dec
        sel = np.zeros(n + 1, np.int32)
        used = C.c_float()
        k = L.b200nav_scan_select(info.ctypes.data, sel.ctypes.data, n + 1, C.byref(used))
        want_sel, want_used = O.scan_select(np.float32(inc), n, bool(dec))
        assert k == len(want_sel) and np.array_equal(sel[:k], want_sel) and used.value == want_used
        assert L.b200nav_scan_select(info.ctypes.data, None, 0, None) == k   # count only


def test_projection_against_numpy_and_drop_rules():
    rng = np.random.default_rng(1)
    n = 360
    inc = np.float32(2 * np.pi / n)
    ranges = rng.uniform(0.3, 5.0, n).astype(np.float32)
    ranges[[3, 50, 51, 200, 359]] = [np.inf, np.nan, 6.0, 0.05, 5.99]   # 6.0 == range_max and 0.05 < range_min drop
    pose = (1.25, -0.75, 0.6)
    s = O.project_scan(-np.pi, inc, 0.1, 6.0, ranges, pose, decimate=True)   # 1 degree spacing: nothing thinned
    keep = np.array([(r >= np.float32(0.1)) and (r < np.float32(6.0)) for r in ranges])
    assert len(s) == keep.sum() == n - 4
    j = np.arange(n)[keep]
    ang = np.float64(np.float32(-np.pi)) + j.astype(np.float64) * np.float64(inc)
    lx = (ranges[keep].astype(np.float64) * np.cos(ang)).astype(np.float32)
    ly = (ranges[keep].astype(np.float64) * np.sin(ang)).astype(np.float32)
    X = ((np.cos(pose[2]) * lx - np.sin(pose[2]) * ly) + pose[0]).astype(np.float32)
    Y = ((np.sin(pose[2]) * lx + np.cos(pose[2]) * ly) + pose[1]).astype(np.float32)
    # libm and the specified polynomial may differ in the last bit of sin / cos: allow one float32 ulp
    assert np.all(np.abs(s["ex"] - X) <= np.spacing(np.abs(X).astype(np.float32)))
    assert np.all(np.abs(s["ey"] - Y) <= np.spacing(np.abs(Y).astype(np.float32)))
    assert np.all(s["sx"] == pose[0]) and np.all(s["sy"] == pose[1]) and np.all(s["clear_end"] == 0)
    assert np.all(s["ex"] == s["ex"].astype(np.float32)) and np.all(s["ey"] == s["ey"].astype(np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("beams,fov,decimate", [(1080, 1.5 * np.pi, True), (1080, 1.5 * np.pi, False),
                                                 (360, 2 * np.pi, True), (4096, 2 * np.pi, True)])
def test_scan_form_update_matches_oracle(beams, fov, decimate):
    """b200nav_himm_update_scans_batched[_dev]: projection inside the binning kernel == oracle projection + oracle
    HIMM, grids bit-exact; dropped readings (inf / NaN / out of range) do nothing."""
    import torch
    from ros_navigation_b200 import DeviceGridMap, capi
    ctx = capi.Context(0)
    rng = np.random.default_rng(beams + int(decimate))
    n_robots = 5
    g = O.make_geom(12.8, 12.8, 0.05)
    dg = DeviceGridMap(ctx, (12.8, 12.8), 0.05, n_robots=n_robots, layers=("laser",))
    layers = [O.new_layer(g) for _ in range(n_robots)]
    angle_min, inc = np.float32(-fov / 2), np.float32(fov / beams)
    info = dg.scan_info(angle_min, inc, 0.2, 6.0, beams, decimate)
    for cycle in range(4):
        poses = np.stack([rng.uniform(-3, 3, n_robots), rng.uniform(-3, 3, n_robots), rng.uniform(-4, 4, n_robots)], 1)
        poses[0] = [7.5, 0.0, np.pi]          # sensor outside the map: rays are clipped in (general schedule)
        ranges = rng.uniform(0.1, 7.0, (n_robots, beams)).astype(np.float32)   # some below min, some beyond max
        ranges[:, ::97] = np.inf
        ranges[1, 5] = np.nan
        if cycle == 2:   # enqueue-only form from pinned host buffers
            hp, hr = torch.from_numpy(poses).pin_memory(), torch.from_numpy(ranges).pin_memory()
            dg.himm_update_scans_batched_async("laser", info, hp, hr)
            ctx.synchronize()
        elif cycle % 2 == 0:
            dg.himm_update_scans_batched("laser", info, poses, ranges)
        else:
            dg.himm_update_scans_batched_dev("laser", info, torch.from_numpy(poses).cuda(), torch.from_numpy(ranges).cuda())
            ctx.synchronize()
        for r in range(n_robots):
            O.himm_update(g, layers[r], O.project_scan(angle_min, inc, 0.2, 6.0, ranges[r], poses[r], decimate))
    for r in range(n_robots):
        assert_layers_equal(dg.download("laser", robot=r), layers[r], "robot %d" % r)
    dg.close()
    ctx.close()
