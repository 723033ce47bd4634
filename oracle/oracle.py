"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE -- never imported by the product package).

Two libraries:
  oracle/_build/liboracle.so   restatement of the HIMM / grid_map / pseudo-scan path (himm_oracle.cpp)
  oracle/_ref/libvfh_ref.so    the unmodified reference move_control::VFH (vfh_ref_wrap.cpp + reference vfh.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
REF_PATH = os.path.join(HERE, "_ref", "libvfh_ref.so")


def build(verbose=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    out = subprocess.run(["make", "-C", HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


class Geom(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("res", C.c_double), ("len_x", C.c_double),
                ("len_y", C.c_double), ("pos_x", C.c_double), ("pos_y", C.c_double), ("start0", C.c_int),
                ("start1", C.c_int)]


SAMPLE_DTYPE = np.dtype([("sx", "<f8"), ("sy", "<f8"), ("ex", "<f8"), ("ey", "<f8"), ("clear_end", "<i4"),
                         ("pad", "<i4")])
assert SAMPLE_DTYPE.itemsize == 40

_lib = None
_ref = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        gp = C.POINTER(Geom)
        L.oracle_geom_init.argtypes = [gp] + [C.c_double] * 5
        L.oracle_is_inside.argtypes = [gp, C.c_double, C.c_double]
        L.oracle_index_from_position.argtypes = [gp, C.c_double, C.c_double, _ip, _ip]
        L.oracle_position_from_index.argtypes = [gp, C.c_int, C.c_int, _dp, _dp]
        L.oracle_index_shift_from_position_shift.argtypes = [C.c_double, C.c_double, C.c_double, _ip, _ip]
        L.oracle_line_cells.argtypes = [gp] + [C.c_double] * 4 + [C.c_void_p, C.c_int]
        L.oracle_himm_update.argtypes = [gp, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_himm_update.restype = C.c_longlong
        L.oracle_himm_update_as_written.argtypes = [gp, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.oracle_himm_update_as_written.restype = C.c_longlong
        L.oracle_himm_count.argtypes = [gp, C.c_void_p, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.oracle_submap_info.argtypes = [gp] + [C.c_double] * 4 + [_ip] * 4 + [_dp] * 4
        L.oracle_get_submap.argtypes = [gp, C.c_void_p] + [C.c_double] * 4 + [C.c_void_p, C.c_int, _ip, _ip]
        L.oracle_ranges_from_submap.argtypes = [gp, C.c_void_p] + [C.c_double] * 4 + [C.c_void_p]
        L.oracle_move.argtypes = [gp, C.POINTER(C.c_void_p), C.c_int, C.c_double, C.c_double]
        L.oracle_to_occupancy.argtypes = [gp, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
        L.oracle_goal_from_pose.argtypes = [C.c_double] * 5 + [_fp, _fp]
        L.oracle_if_blocked.argtypes = [gp, C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.oracle_compose_master.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong]
        L.oracle_sincos.argtypes = [C.c_double, _dp, _dp]
        L.oracle_scan_select.argtypes = [C.c_float, C.c_int, C.c_int, C.c_void_p, _fp]
        L.oracle_project_scan.argtypes = [C.c_float] * 4 + [C.c_void_p, C.c_int, C.c_int, C.c_void_p] + [C.c_double] * 3 + [C.c_void_p]
        _lib = L
    return _lib


def have_ref():
    return os.path.exists(REF_PATH)


def ref():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_PATH):
            build()
        if not os.path.exists(REF_PATH):
            raise RuntimeError("oracle/_ref/libvfh_ref.so missing and /root/reference not present to build it")
        R = C.CDLL(REF_PATH)
        R.vfhref_create.argtypes = [_dp]
        R.vfhref_create.restype = C.c_void_p
        R.vfhref_destroy.argtypes = [C.c_void_p]
        R.vfhref_set_time.argtypes = [C.c_double]
        R.vfhref_get_time.restype = C.c_double
        R.vfhref_set_current_max_speed.argtypes = [C.c_void_p, C.c_int]
        R.vfhref_update.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, _ip, _ip]
        R.vfhref_hist_size.argtypes = [C.c_void_p]
        R.vfhref_window.argtypes = [C.c_void_p]
        R.vfhref_num_tables.argtypes = [C.c_void_p]
        R.vfhref_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        R.vfhref_get_cell_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        R.vfhref_get_sector_masks.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        R.vfhref_get_min_turning_radius.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        R.vfhref_get_cell_mag.argtypes = [C.c_void_p, C.c_void_p]
        _ref = R
    return _ref


# ----------------------------------------------------------------------------------------------------------------
# Thin Python conveniences over the C oracle
# ----------------------------------------------------------------------------------------------------------------

def make_geom(len_x, len_y, res, pos_x=0.0, pos_y=0.0, start=(0, 0)):
    g = Geom()
    lib().oracle_geom_init(C.byref(g), len_x, len_y, res, pos_x, pos_y)
    g.start0, g.start1 = int(start[0]), int(start[1])
    return g


def new_layer(g):
    """Column-major rows x cols float layer, NaN-initialised (GridMap.cpp:51-70,624-629)."""
    return np.full((g.cols, g.rows), np.nan, dtype=np.float32)  # [col][row] == column-major (row, col)


def index_from_position(g, x, y):
    r, c = C.c_int(), C.c_int()
    ok = lib().oracle_index_from_position(C.byref(g), x, y, C.byref(r), C.byref(c))
    return (r.value, c.value) if ok else None


def position_from_index(g, r, c):
    x, y = C.c_double(), C.c_double()
    ok = lib().oracle_position_from_index(C.byref(g), r, c, C.byref(x), C.byref(y))
    return (x.value, y.value) if ok else None


def is_inside(g, x, y):
    return bool(lib().oracle_is_inside(C.byref(g), x, y))


def line_cells(g, sx, sy, ex, ey, cap=1 << 16):
    buf = np.zeros((cap, 2), dtype=np.int32)
    n = lib().oracle_line_cells(C.byref(g), sx, sy, ex, ey, buf.ctypes.data, cap)
    return buf[:min(n, cap)].copy()


def make_samples(sx, sy, ex, ey, clear_end=None):
    n = len(ex)
    s = np.zeros(n, dtype=SAMPLE_DTYPE)
    s["sx"], s["sy"], s["ex"], s["ey"] = sx, sy, ex, ey
    if clear_end is not None:
        s["clear_end"] = clear_end
    return s


def himm_update(g, layer, samples, bbox=None, as_written=False):
    """Apply samples in order to `layer` ([col][row] float32, modified in place). Returns #visits."""
    assert layer.dtype == np.float32 and layer.flags.c_contiguous and layer.shape == (g.cols, g.rows)
    samples = np.ascontiguousarray(samples, dtype=SAMPLE_DTYPE)
    if bbox is None:
        bbox = np.zeros(4, dtype=np.float64)
    fn = lib().oracle_himm_update_as_written if as_written else lib().oracle_himm_update
    return fn(C.byref(g), layer.ctypes.data, samples.ctypes.data, len(samples), bbox.ctypes.data)


def himm_count(g, samples):
    samples = np.ascontiguousarray(samples, dtype=SAMPLE_DTYPE)
    v, m = C.c_longlong(), C.c_longlong()
    lib().oracle_himm_count(C.byref(g), samples.ctypes.data, len(samples), C.byref(v), C.byref(m))
    return v.value, m.value


def submap_info(g, cx, cy, lx, ly):
    tr, tc, sr, sc = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    px, py, sx, sy = C.c_double(), C.c_double(), C.c_double(), C.c_double()
    ok = lib().oracle_submap_info(C.byref(g), cx, cy, lx, ly, C.byref(tr), C.byref(tc), C.byref(sr), C.byref(sc),
                                  C.byref(px), C.byref(py), C.byref(sx), C.byref(sy))
    if not ok:
        return None
    return dict(tl=(tr.value, tc.value), size=(sr.value, sc.value), pos=(px.value, py.value),
                length=(sx.value, sy.value))


def get_submap(g, layer, cx, cy, lx, ly):
    out = np.zeros(g.rows * g.cols, dtype=np.float32)
    sr, sc = C.c_int(), C.c_int()
    ok = lib().oracle_get_submap(C.byref(g), layer.ctypes.data, cx, cy, lx, ly, out.ctypes.data, out.size,
                                 C.byref(sr), C.byref(sc))
    if not ok:
        return None
    return out[:sr.value * sc.value].reshape(sc.value, sr.value).copy()


def ranges_from_submap(g, master, rx, ry, yaw, submap_len=1.5):
    ranges = np.zeros((361, 2), dtype=np.float64)
    lib().oracle_ranges_from_submap(C.byref(g), master.ctypes.data, rx, ry, yaw, submap_len, ranges.ctypes.data)
    return ranges


def compose_master(range_layer, laser_layer):
    """master = range (+) laser (the commented-out compose, map_provider.cpp:218-220)."""
    a = np.ascontiguousarray(range_layer, dtype=np.float32)
    b = np.ascontiguousarray(laser_layer, dtype=np.float32)
    out = np.empty_like(a)
    lib().oracle_compose_master(a.ctypes.data, b.ctypes.data, out.ctypes.data, a.size)
    return out


def sincos(x):
    s, c = C.c_double(), C.c_double()
    lib().oracle_sincos(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def scan_select(angle_increment, n_ranges, decimate=True):
    """simplifyLaserScan: (indices of the projected readings, angle increment the projection uses)."""
    sel = np.zeros(n_ranges + 1, np.int32)
    used = C.c_float()
    n = lib().oracle_scan_select(float(angle_increment), int(n_ranges), int(bool(decimate)), sel.ctypes.data,
                                 C.byref(used))
    return sel[:n].copy(), used.value


def project_scan(angle_min, angle_increment, range_min, range_max, ranges, pose, decimate=True):
    """RangeSamples of one LaserScan taken at sensor pose (x, y, yaw) in the map frame (spec: scan_project.h)."""
    ranges = np.ascontiguousarray(ranges, dtype=np.float32)
    sel, used = scan_select(np.float32(angle_increment), len(ranges), decimate)
    out = np.zeros(len(sel), dtype=SAMPLE_DTYPE)
    n = lib().oracle_project_scan(float(np.float32(angle_min)), used, float(np.float32(range_min)),
                                  float(np.float32(range_max)), sel.ctypes.data, len(sel),
                                  int(bool(decimate) and float(np.float32(angle_increment)) < 0.017), ranges.ctypes.data,
                                  float(pose[0]), float(pose[1]), float(pose[2]), out.ctypes.data)
    return out[:n].copy()


def move(g, layers, x, y):
    arr = (C.c_void_p * len(layers))(*[l.ctypes.data for l in layers])
    return bool(lib().oracle_move(C.byref(g), arr, len(layers), x, y))


def to_occupancy(g, layer, data_min=0.0, data_max=255.0):
    out = np.zeros(g.rows * g.cols, dtype=np.int8)
    lib().oracle_to_occupancy(C.byref(g), layer.ctypes.data, data_min, data_max, out.ctypes.data)
    return out


def if_blocked(g, master, x, y, radius=0.3):
    return bool(lib().oracle_if_blocked(C.byref(g), master.ctypes.data, x, y, radius))


def goal_from_pose(rx, ry, yaw, tx, ty):
    a, d = C.c_float(), C.c_float()
    lib().oracle_goal_from_pose(rx, ry, yaw, tx, ty, C.byref(a), C.byref(d))
    return a.value, d.value


# Steerer::initVfh defaults (move_control/src/steerer.cpp:69-121), in VFH constructor order + robot_radius.
VFH_DEFAULTS = dict(cell_size=100.0, window_diameter=30, sector_angle=5, safety_dist_0ms=10.0, safety_dist_1ms=50.0,
                    max_speed=200, max_speed_narrow_opening=200, max_speed_wide_opening=300, max_acceleration=200,
                    min_turnrate=40, max_turnrate_0ms=40, max_turnrate_1ms=40, min_turn_radius_safety_factor=1.0,
                    free_space_cutoff_0ms=2000000.0, obs_cutoff_0ms=4000000.0, free_space_cutoff_1ms=2000000.0,
                    obs_cutoff_1ms=4000000.0, weight_desired_dir=10.0, weight_current_dir=1.0, robot_radius=178.0)
VFH_PARAM_ORDER = list(VFH_DEFAULTS.keys())


class RefVFH:
    """The reference move_control::VFH behind a deterministic clock (vfh.h:182-361)."""

    def __init__(self, **kw):
        p = dict(VFH_DEFAULTS)
        p.update(kw)
        self.params = p
        arr = (C.c_double * 20)(*[float(p[k]) for k in VFH_PARAM_ORDER])
        self._R = ref()
        self._R.vfhref_set_time(1000.0)
        self._h = self._R.vfhref_create(arr)
        self._t = 1000.0
        self.hist_size = self._R.vfhref_hist_size(self._h)
        self.window = self._R.vfhref_window(self._h)
        self.num_tables = self._R.vfhref_num_tables(self._h)

    def __del__(self):
        try:
            self._R.vfhref_destroy(self._h)
        except Exception:
            pass

    def set_current_max_speed(self, s):
        self._R.vfhref_set_current_max_speed(self._h, int(s))

    def update(self, ranges, speed, goal_dir, goal_dist, tol, dt):
        """dt = seconds since the previous update (or since Init for the first one)."""
        ranges = np.ascontiguousarray(ranges, dtype=np.float64).reshape(361, 2)
        self._t = round(self._t + dt, 6)
        self._R.vfhref_set_time(self._t)
        cs, ct = C.c_int(), C.c_int()
        self._R.vfhref_update(self._h, ranges.ctypes.data, int(speed), float(goal_dir), float(goal_dist), float(tol),
                              C.byref(cs), C.byref(ct))
        return cs.value, ct.value

    def state(self):
        n = self.hist_size
        oh = np.zeros(n, np.float32)
        h = np.zeros(n, np.float32)
        lb = np.zeros(n, np.float32)
        f = np.zeros(4, np.float32)
        i = np.zeros(3, np.int32)
        self._R.vfhref_get_state(self._h, oh.ctypes.data, h.ctypes.data, lb.ctypes.data, f.ctypes.data, i.ctypes.data)
        return dict(origin_hist=oh, hist=h, last_binary=lb, picked=float(f[0]), last_picked=float(f[1]),
                    desired=float(f[2]), blocked_radius=float(f[3]), last_chosen_speed=int(i[0]),
                    max_speed_for_picked=int(i[1]), current_max_speed=int(i[2]))

    def cell_tables(self):
        W = self.window
        d = np.zeros((W, W), np.float32)
        s = np.zeros((W, W), np.float32)
        b = np.zeros((W, W), np.float32)
        self._R.vfhref_get_cell_tables(self._h, d.ctypes.data, s.ctypes.data, b.ctypes.data)
        return d, s, b

    def sector_masks(self, table):
        W = self.window
        nw = (self.hist_size + 31) // 32
        m = np.zeros((W, W, nw), np.uint32)
        ok = self._R.vfhref_get_sector_masks(self._h, table, m.ctypes.data, nw)
        return m, bool(ok)

    def min_turning_radius(self):
        n = self.state()["current_max_speed"] + 1
        out = np.zeros(n, np.int32)
        self._R.vfhref_get_min_turning_radius(self._h, out.ctypes.data, n)
        return out
