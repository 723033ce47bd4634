"""GPU: the two device formats of a layer (ros_navigation_b200/csrc/cells.cuh).  Layers start byte-coded (one byte per
cell in 64 x 64 tile records); values outside {NaN, 0, 10, ..., 180} or a request for the raw float pointer move a
layer to the reference's float matrix.  Every API result must be the same in both formats."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_layers_equal, lidar_samples, random_samples

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    from ros_navigation_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


def himm_values(rng, shape):
    """Random cells from the closed HIMM set, all codes present."""
    vals = np.concatenate([[np.nan], np.arange(0, 190, 10)]).astype(np.float32)
    out = vals[rng.integers(0, len(vals), size=shape)]
    out.flat[:len(vals)] = vals
    return np.ascontiguousarray(out)


@pytest.mark.parametrize("lx,ly", [(10.0, 10.0), (6.5, 3.5), (3.2, 12.85), (0.05, 0.05)])
def test_coded_round_trip_and_readers(ctx, lx, ly):
    """upload -> download is the identity on HIMM-set data for any grid size (also not a multiple of the 64-cell tile),
    and the coded readers (occupancy export, blocked query) agree with the oracle."""
    from ros_navigation_b200 import DeviceGridMap
    rng = np.random.default_rng(5)
    g = O.make_geom(lx, ly, 0.05)
    dg = DeviceGridMap(ctx, (lx, ly), 0.05, n_robots=3, layers=("master",))
    # layers are [cols][rows] C-contiguous arrays == column-major rows x cols (oracle.new_layer)
    data = [himm_values(rng, (g.cols, g.rows)) if g.rows * g.cols >= 20 else np.full((g.cols, g.rows), 30, np.float32)
            for _ in range(3)]
    for r in (2, 0, 1):
        dg.upload("master", data[r], robot=r)
    assert dg.layer_format("master") == "coded"
    for r in range(3):
        assert_layers_equal(dg.download("master", robot=r), data[r], "robot %d" % r)
        assert np.array_equal(dg.to_occupancy("master", 0.0, 180.0, robot=r), O.to_occupancy(g, data[r], 0.0, 180.0))
    pts = rng.uniform(-0.6, 0.6, size=(64, 2)) * [lx, ly]
    got = dg.query_blocked(pts, radius=0.3, robot=1)
    want = np.array([O.if_blocked(g, data[1], float(x), float(y), 0.3) for x, y in pts])
    assert np.array_equal(np.asarray(got, bool), want)
    dg.close()


def test_foreign_upload_switches_layer_to_float(ctx):
    """A value outside the HIMM set cannot be coded: the whole layer (all robots) moves to the float format, keeps
    every other robot's content, and later HIMM updates stay bit-exact."""
    from ros_navigation_b200 import DeviceGridMap
    rng = np.random.default_rng(6)
    g = O.make_geom(8.0, 6.0, 0.05)
    dg = DeviceGridMap(ctx, (8.0, 6.0), 0.05, n_robots=2, layers=("laser",))
    dg.alias("master", "laser")
    lay = [O.new_layer(g), O.new_layer(g)]
    s0 = lidar_samples(rng, g, (0.5, 0.5), 500, 0.2, 3.5)
    off = np.array([0, len(s0), len(s0)], np.int32)
    dg.himm_update_batched("laser", s0, off)
    O.himm_update(g, lay[0], s0)
    assert dg.layer_format("laser") == "coded"
    lay[1][:] = (rng.random(lay[1].shape) * 250 - 20).astype(np.float32)   # arbitrary floats, some negative
    lay[1][3, 4] = -0.0
    dg.upload("laser", lay[1], robot=1)
    assert dg.layer_format("laser") == "float" and dg.layer_format("master") == "float"   # the alias follows
    for it in range(3):
        per = [lidar_samples(rng, g, (0.5, 0.5), 400, 0.2, 3.5, clear_frac=0.1), random_samples(rng, g, 300)]
        off = np.array([0, len(per[0]), len(per[0]) + len(per[1])], np.int32)
        dg.himm_update_batched("master", np.concatenate(per), off)
        for r in range(2):
            O.himm_update(g, lay[r], per[r])
    for r in range(2):
        got = dg.download("laser", robot=r)
        assert_layers_equal(got, lay[r], "robot %d" % r)
    assert np.signbit(dg.download("laser", robot=1)[3, 4]) == np.signbit(lay[1][3, 4])
    dg.close()


def test_devptr_and_copy_between_formats(ctx):
    """The raw device pointer is the reference's float matrix (conversion on demand); copy_layer gives the destination
    the source's format; clear works in both."""
    import torch
    from ros_navigation_b200 import DeviceGridMap
    rng = np.random.default_rng(7)
    g = O.make_geom(5.0, 7.0, 0.05)
    dg = DeviceGridMap(ctx, (5.0, 7.0), 0.05, layers=("laser", "copy", "other"))
    layer = O.new_layer(g)
    s = lidar_samples(rng, g, (0.0, 0.0), 720, 0.2, 3.0)
    dg.himm_update("laser", s)
    O.himm_update(g, layer, s)
    dg.copy_layer("copy", "laser")                      # coded -> coded
    assert dg.layer_format("copy") == "coded"
    assert_layers_equal(dg.download("copy"), layer, "coded copy")
    p = dg.layer_devptr("laser")                       # converts
    assert p and dg.layer_format("laser") == "float"
    assert_layers_equal(dg.download("laser"), layer, "after conversion")
    dg.copy_layer("other", "laser")                     # float -> coded destination becomes float
    assert dg.layer_format("other") == "float"
    assert_layers_equal(dg.download("other"), layer, "float copy")
    dg.copy_layer("laser", "copy")                      # coded -> float destination becomes coded again
    assert dg.layer_format("laser") == "coded"
    s2 = lidar_samples(rng, g, (0.4, -0.3), 720, 0.2, 3.0)
    for name in ("laser", "other"):                     # one coded, one float: same result
        dg.himm_update(name, s2)
    O.himm_update(g, layer, s2)
    assert_layers_equal(dg.download("laser"), layer, "coded after copies")
    assert_layers_equal(dg.download("other"), layer, "float after copies")
    dg.clear()
    for name in ("laser", "copy", "other"):
        assert np.isnan(dg.download(name)).all()
    dg.close()


def test_float_layer_mode_runs_the_same_suites():
    """B200NAV_FLOAT_LAYERS=1 creates every layer in the float format: the HIMM and VFH parity suites must pass there
    too (the format is chosen at layer creation, so this runs in a child process)."""
    env = dict(os.environ, B200NAV_FLOAT_LAYERS="1")
    out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_himm_gpu.py", "tests/test_vfh_gpu.py", "-m", "gpu",
                          "-x", "-q", "-k", "not very_long and not full_size", "-p", "no:cacheprovider"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=400)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]


@pytest.mark.parametrize("mode", ["0", "2"])
def test_both_tile_kernels_run_the_himm_suite(mode):
    """By default fleets of up to 32 robots are walked by himm_tile_coded_mw_kernel (CTAs of four warps share a tile
    as a wavefront pipeline over the beam batches) and larger ones by the one-warp kernel.  B200NAV_MW_HEAVY=0 / =2
    force one or the other for every fleet size: same bits either way."""
    env = dict(os.environ, B200NAV_MW_HEAVY=mode)
    out = subprocess.run([sys.executable, "-m", "pytest", "tests/test_himm_gpu.py", "-m", "gpu", "-x", "-q", "-k",
                          "random or order or chunking or several or edge or batched or cloud_form_matches or long_steady "
                          "or uneven or c2_sized",
                          "-p", "no:cacheprovider"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]


def test_red_calibration_is_a_plain_measurement(ctx):
    """b200nav_ctx_calibrate_red (bench.py's roofline_red): a positive, finite rate; rejects a useless buffer; a buffer
    that fits L2 is not slower than one that does not."""
    from ros_navigation_b200 import capi
    small = ctx.calibrate_red(32 << 20)
    big = ctx.calibrate_red(1 << 30)
    assert np.isfinite(small) and np.isfinite(big) and small > 1e9 and big > 1e9
    assert small > 0.8 * big
    with pytest.raises(capi.B200NavError):
        ctx.calibrate_red(16)
