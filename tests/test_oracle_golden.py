"""CPU: the oracle reproduces the committed golden vectors (tests/golden/*.npz, see make_golden.py)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import assert_layers_equal

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_himm_oracle_reproduces_golden():
    z = np.load(os.path.join(GOLD, "himm_golden.npz"))
    for name in sorted({k.split("/")[0] for k in z.files}):
        lx, ly, res, px, py, s0, s1 = z[name + "/geom"]
        g = O.make_geom(lx, ly, res, px, py, (int(s0), int(s1)))
        samples = z[name + "/samples"].view(O.SAMPLE_DTYPE)
        layer = O.new_layer(g)
        pos = 0
        for n in z[name + "/batch"]:
            O.himm_update(g, layer, samples[pos:pos + n])
            pos += n
        assert_layers_equal(layer, z[name + "/layer"], name)


def test_ranges_oracle_reproduces_golden():
    z = np.load(os.path.join(GOLD, "ranges_golden.npz"))
    for i in range(3):
        gg = z["%d/geom" % i]
        g = O.make_geom(*gg[:5], (int(gg[5]), int(gg[6])))
        layer = z["%d/layer" % i]
        for p, want in zip(z["%d/poses" % i], z["%d/ranges" % i]):
            got = O.ranges_from_submap(g, layer, *p)
            assert np.array_equal(got[:, 0], want)


def test_reference_vfh_reproduces_golden():
    if not O.have_ref():
        pytest.skip("oracle/_ref not built")
    z = np.load(os.path.join(GOLD, "vfh_golden.npz"))
    for case in sorted({k.split("/")[0] for k in z.files}):
        g = {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(case + "/")}
        p = dict(zip(O.VFH_PARAM_ORDER, g["params"]))
        v = O.RefVFH(**p)
        for s in range(len(g["dt"])):
            cs, ct = v.update(g["ranges"][s], int(g["speed"][s]), float(g["gdir"][s]), float(g["gdist"][s]),
                              float(g["tol"][s]), float(g["dt"][s]))
            st = v.state()
            assert (cs, ct) == (g["out_speed"][s], g["out_turn"][s])
            assert np.array_equal(st["origin_hist"], g["origin_hist"][s])
            assert st["picked"] == g["picked"][s]


def test_submap_window_sizes():
    """getSubmapInformation (GridMapMath.cpp:246-296): a 1.5 m window is 30 or 31 cells at 5 cm and is clipped
    at the map border."""
    g = O.make_geom(10.0, 10.0, 0.05)
    sizes = set()
    rng = np.random.default_rng(0)
    for _ in range(200):
        info = O.submap_info(g, rng.uniform(-4, 4), rng.uniform(-4, 4), 1.5, 1.5)
        sizes.add(info["size"])
    assert sizes.issubset({(30, 30), (30, 31), (31, 30), (31, 31)}) and (31, 31) in sizes
    info = O.submap_info(g, 4.9, -4.9, 1.5, 1.5)
    assert info["size"][0] < 30 and info["size"][1] < 30
    assert O.submap_info(g, 0.0, 0.0, 1.5, 1.5)["size"] == (31, 31)
