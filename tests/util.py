"""Shared helpers for the parity tests."""
import numpy as np

from oracle import oracle as O


def canon(layer):
    """float32 bit patterns with every NaN canonicalised."""
    a = np.ascontiguousarray(layer, dtype=np.float32).copy()
    bits = a.view(np.uint32)
    bits[np.isnan(a)] = 0x7FC00000
    return bits


def assert_layers_equal(got, want, what=""):
    g, w = canon(got), canon(want)
    if not np.array_equal(g, w):
        bad = np.argwhere(g != w)
        c, r = bad[0]
        raise AssertionError("%s: %d cells differ, first at (row %d, col %d): got %r want %r" %
                             (what, len(bad), r, c, got[c, r], want[c, r]))


def random_samples(rng, g, n, spread=1.4, clear_frac=0.2, origin=None):
    """n samples around the map (some ends/starts outside), optionally sharing one origin like a lidar."""
    lx, ly = g.len_x * spread, g.len_y * spread
    if origin is None:
        sx = g.pos_x + (rng.random(n) - 0.5) * lx
        sy = g.pos_y + (rng.random(n) - 0.5) * ly
    else:
        sx = np.full(n, origin[0])
        sy = np.full(n, origin[1])
    ex = g.pos_x + (rng.random(n) - 0.5) * lx
    ey = g.pos_y + (rng.random(n) - 0.5) * ly
    # laser_geometry delivers float32 end points
    ex = ex.astype(np.float32).astype(np.float64)
    ey = ey.astype(np.float32).astype(np.float64)
    ce = (rng.random(n) < clear_frac).astype(np.int32)
    return O.make_samples(sx, sy, ex, ey, ce)


def lidar_samples(rng, g, origin, n, rmin, rmax, fov=2 * np.pi, clear_frac=0.0):
    th = -fov / 2 + fov * np.arange(n) / n
    r = rmin + (rmax - rmin) * rng.random(n)
    ex = (origin[0] + r * np.cos(th)).astype(np.float32).astype(np.float64)
    ey = (origin[1] + r * np.sin(th)).astype(np.float32).astype(np.float64)
    ce = (rng.random(n) < clear_frac).astype(np.int32)
    return O.make_samples(np.full(n, origin[0]), np.full(n, origin[1]), ex, ey, ce)
