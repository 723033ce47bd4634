"""ctypes face of oracle/nav_harness.cpp (TEST INFRASTRUCTURE -- never imported by the product package).

Two builds of the same harness export the same navh_* functions:
  oracle/_ref/libnav_ref.so          the reference's own MapProvider / Steerer / updaters / VFH / grid_map_core,
                                     compiled where they lie under /root/reference (oracle/Makefile)
  tests/cpp/_build/libnav_dropin.so  the reference's map_provider.cpp / steerer.cpp / grid_map_core with the product's
                                     drop-in updater and VFH headers (include/move_control/) on libb200nav.so
                                     (tests/cpp/Makefile)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(HERE, "_ref", "libnav_ref.so")
DROPIN_PATH = os.path.join(os.path.dirname(HERE), "tests", "cpp", "_build", "libnav_dropin.so")


class SteerOut(C.Structure):
    _fields_ = [("linear_x", C.c_double), ("angular_z", C.c_double), ("updated", C.c_int32),
                ("plan_ready", C.c_int32), ("picked_angle", C.c_float), ("desired_angle", C.c_float),
                ("ranges", C.c_double * 361), ("hist", C.c_float * 72), ("origin_hist", C.c_float * 72)]


_libs = {}


def have_ref():
    return os.path.exists(REF_PATH)


def have_dropin():
    return os.path.exists(DROPIN_PATH)


def load(path):
    if path in _libs:
        return _libs[path]
    L = C.CDLL(path)
    vp, cp, dp, fp, ip = C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.navh_create.argtypes = [C.c_double, C.c_double, C.c_int, C.POINTER(cp), dp, C.c_int, C.c_double]
    L.navh_create.restype = vp
    L.navh_destroy.argtypes = [vp]
    L.navh_set_time.argtypes = [vp, C.c_double]
    L.navh_set_frame.argtypes = [vp, cp, C.c_double, C.c_double, C.c_double]
    L.navh_publish_scan.argtypes = [vp, cp, cp, C.c_float, C.c_float, C.c_float, C.c_float, vp, C.c_int]
    L.navh_publish_range.argtypes = [vp, cp, cp, C.c_float, C.c_float, C.c_float]
    L.navh_publish_odom.argtypes = [vp, C.c_double]
    L.navh_update_map.argtypes = [vp]
    L.navh_move_map.argtypes = [vp]
    L.navh_publish_map.argtypes = [vp, vp, C.c_int]
    L.navh_geometry.argtypes = [vp, vp, vp]
    L.navh_get_layer.argtypes = [vp, cp, vp, C.c_int]
    L.navh_accept_plan.argtypes = [vp, vp, C.c_int]
    L.navh_robot_pose.argtypes = [vp, vp]
    L.navh_steer.argtypes = [vp, C.POINTER(SteerOut)]
    L.navh_last_hist_msg.argtypes = [vp, vp, vp]
    L.navh_fleet_cycle.argtypes = [vp, C.c_int, C.c_double, vp, vp, C.c_int, C.c_float, C.c_float, C.c_float,
                                   C.c_float, vp, vp, vp, C.c_int]
    if L.navh_is_dropin():
        L.navh_steer_from_grid.argtypes = [vp, vp, C.c_int, ip, C.c_double, cp, C.POINTER(SteerOut)]
    if not L.navh_is_dropin():
        L.navh_fleet_cycle_samples.argtypes = [vp, C.c_int, C.c_double, vp, vp, vp, vp, vp, vp, C.c_int]
        L.navh_core_create.argtypes = [C.c_double] * 5 + [cp]
        L.navh_core_create.restype = vp
        L.navh_core_destroy.argtypes = [vp]
        L.navh_core_size.argtypes = [vp, ip, ip]
        L.navh_core_update.argtypes = [vp, vp, C.c_int, vp]
        L.navh_core_move.argtypes = [vp, C.c_double, C.c_double]
        L.navh_core_start_index.argtypes = [vp, vp, vp]
        L.navh_core_get_layer.argtypes = [vp, cp, vp, C.c_int]
        L.navh_core_set_layer.argtypes = [vp, cp, vp]
        L.navh_core_line.argtypes = [vp] + [C.c_double] * 4 + [vp, C.c_int]
        L.navh_core_blocked.argtypes = [vp, cp, C.c_double, C.c_double, C.c_double]
    _libs[path] = L
    return L


class Node:
    """One reference node: MapProvider(nh, Length(len_x, len_y), moving) + Steerer (nav_only_vfh_node.cpp:40-41)."""

    def __init__(self, path, len_x, len_y, moving=False, params=None, t0=1.0):
        self.L = load(path)
        params = params or {}
        keys = (C.c_char_p * max(1, len(params)))(*[k.encode() for k in params])
        vals = (C.c_double * max(1, len(params)))(*[float(v) for v in params.values()])
        self.h = self.L.navh_create(len_x, len_y, int(moving), keys, vals, len(params), t0)
        if not self.h:
            raise RuntimeError("navh_create failed (" + path + ")")
        geo = np.zeros(4, np.int32)
        self.L.navh_geometry(self.h, geo.ctypes.data, np.zeros(5).ctypes.data)
        self.rows, self.cols = int(geo[0]), int(geo[1])

    def close(self):
        if self.h:
            self.L.navh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_time(self, t):
        self.L.navh_set_time(self.h, float(t))

    def set_frame(self, frame, x, y, yaw):
        self.L.navh_set_frame(self.h, frame.encode(), float(x), float(y), float(yaw))

    def publish_scan(self, ranges, angle_min, angle_increment, range_min, range_max, topic="/laser_scan",
                     frame="laser"):
        r = np.ascontiguousarray(ranges, dtype=np.float32)
        self.L.navh_publish_scan(self.h, topic.encode(), frame.encode(), angle_min, angle_increment, range_min,
                                 range_max, r.ctypes.data, len(r))

    def publish_range(self, topic, frame, rng, min_range, max_range):
        self.L.navh_publish_range(self.h, topic.encode(), frame.encode(), rng, min_range, max_range)

    def publish_odom(self, vx):
        self.L.navh_publish_odom(self.h, float(vx))

    def update_map(self):
        self.L.navh_update_map(self.h)

    def move_map(self):
        return self.L.navh_move_map(self.h)

    def geometry(self):
        geo = np.zeros(4, np.int32)
        dbl = np.zeros(5, np.float64)
        self.L.navh_geometry(self.h, geo.ctypes.data, dbl.ctypes.data)
        return dict(rows=int(geo[0]), cols=int(geo[1]), start=(int(geo[2]), int(geo[3])), pos=(dbl[0], dbl[1]),
                    length=(dbl[2], dbl[3]), res=dbl[4])

    def layer(self, name):
        out = np.zeros((self.cols, self.rows), np.float32)
        n = self.L.navh_get_layer(self.h, name.encode(), out.ctypes.data, out.size)
        if n != out.size:
            raise RuntimeError("no layer " + name)
        return out

    def occupancy(self):
        out = np.zeros(self.rows * self.cols, np.int8)
        n = self.L.navh_publish_map(self.h, out.ctypes.data, out.size)
        assert n == out.size
        return out

    def accept_plan(self, xy):
        a = np.ascontiguousarray(xy, dtype=np.float64).reshape(-1, 2)
        self.L.navh_accept_plan(self.h, a.ctypes.data, len(a))

    def robot_pose(self):
        p = np.zeros(3, np.float64)
        if self.L.navh_robot_pose(self.h, p.ctypes.data) != 0:
            return None
        return p

    def steer(self):
        o = SteerOut()
        self.L.navh_steer(self.h, C.byref(o))
        return dict(linear_x=o.linear_x, angular_z=o.angular_z, updated=bool(o.updated), plan_ready=bool(o.plan_ready),
                    picked_angle=o.picked_angle, desired_angle=o.desired_angle,
                    ranges=np.frombuffer(o.ranges, dtype=np.float64).copy(),
                    hist=np.frombuffer(o.hist, dtype=np.float32).copy(),
                    origin_hist=np.frombuffer(o.origin_hist, dtype=np.float32).copy())

    def steer_from_grid(self, plan, plan_index, odom_speed, layer="laser"):
        """Drop-in build only: goal glue + VFH::Update_VFH_FromGrid on the device twin. Returns (rc, index, out)."""
        plan = np.ascontiguousarray(plan, np.float64).reshape(-1, 2)
        idx = C.c_int(int(plan_index))
        o = SteerOut()
        rc = self.L.navh_steer_from_grid(self.h, plan.ctypes.data, len(plan), C.byref(idx), float(odom_speed),
                                         layer.encode(), C.byref(o))
        return rc, idx.value, dict(linear_x=o.linear_x, angular_z=o.angular_z, updated=bool(o.updated),
                                   plan_ready=bool(o.plan_ready), picked_angle=o.picked_angle,
                                   desired_angle=o.desired_angle,
                                   hist=np.frombuffer(o.hist, dtype=np.float32).copy(),
                                   origin_hist=np.frombuffer(o.origin_hist, dtype=np.float32).copy())

    def hist_msg(self):
        a = np.zeros(36, np.uint16)
        b = np.zeros(36, np.uint16)
        n = self.L.navh_last_hist_msg(self.h, a.ctypes.data, b.ctypes.data)
        return n, a, b


def fleet_cycle(nodes, t, poses, ranges, scan, goals=None, speeds=None, threads=1):
    """One native call: every node takes one scan and one decision (navh_fleet_cycle). Returns commands [n, 2]."""
    L = nodes[0].L
    n = len(nodes)
    hs = (C.c_void_p * n)(*[nd.h for nd in nodes])
    poses = np.ascontiguousarray(poses, np.float64)
    ranges = np.ascontiguousarray(ranges, np.float32)
    out = np.zeros((n, 2), np.float64)
    g = np.ascontiguousarray(goals, np.float64) if goals is not None else None
    s = np.ascontiguousarray(speeds, np.float64) if speeds is not None else None
    L.navh_fleet_cycle(hs, n, float(t), poses.ctypes.data, ranges.ctypes.data, ranges.shape[1], scan["angle_min"],
                       scan["angle_increment"], scan["range_min"], scan["range_max"],
                       g.ctypes.data if g is not None else None, s.ctypes.data if s is not None else None,
                       out.ctypes.data, int(threads))
    return out


def fleet_cycle_samples(nodes, t, poses, samples, offsets, goals=None, speeds=None, threads=1):
    """navh_fleet_cycle_samples: every reference node applies its RangeSamples (oracle.SAMPLE_DTYPE records) and takes
    one steering decision, robots block-partitioned over `threads` std::threads inside ONE native call."""
    L = nodes[0].L
    n = len(nodes)
    hs = (C.c_void_p * n)(*[nd.h for nd in nodes])
    poses = np.ascontiguousarray(poses, np.float64)
    samples = np.ascontiguousarray(samples)
    assert samples.dtype.itemsize == 40
    offsets = np.ascontiguousarray(offsets, np.int32)
    out = np.zeros((n, 2), np.float64)
    g = np.ascontiguousarray(goals, np.float64) if goals is not None else None
    s = np.ascontiguousarray(speeds, np.float64) if speeds is not None else None
    L.navh_fleet_cycle_samples(hs, n, float(t), poses.ctypes.data, samples.ctypes.data, offsets.ctypes.data,
                               g.ctypes.data if g is not None else None, s.ctypes.data if s is not None else None,
                               out.ctypes.data, int(threads))
    return out


class Core:
    """The reference's MapUpdater::lineOnMap on a grid_map::GridMap of any geometry (libnav_ref.so only)."""

    def __init__(self, len_x, len_y, res, pos=(0.0, 0.0), layer="laser"):
        self.L = load(REF_PATH)
        self.layer_name = layer
        self.h = self.L.navh_core_create(len_x, len_y, res, pos[0], pos[1], layer.encode())
        r, c = C.c_int(), C.c_int()
        self.L.navh_core_size(self.h, C.byref(r), C.byref(c))
        self.rows, self.cols = r.value, c.value

    def close(self):
        if self.h:
            self.L.navh_core_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update(self, samples, bbox=None):
        """samples: structured array with sx, sy, ex, ey, clear_end (oracle.SAMPLE_DTYPE)."""
        s5 = np.stack([samples["sx"], samples["sy"], samples["ex"], samples["ey"],
                       samples["clear_end"].astype(np.float64)], axis=1)
        s5 = np.ascontiguousarray(s5, np.float64)
        self.L.navh_core_update(self.h, s5.ctypes.data, len(s5), bbox.ctypes.data if bbox is not None else None)

    def move(self, x, y):
        return bool(self.L.navh_core_move(self.h, x, y))

    def start_index(self):
        s = np.zeros(2, np.int32)
        p = np.zeros(2, np.float64)
        self.L.navh_core_start_index(self.h, s.ctypes.data, p.ctypes.data)
        return (int(s[0]), int(s[1])), (float(p[0]), float(p[1]))

    def layer(self, name=None):
        out = np.zeros((self.cols, self.rows), np.float32)
        n = self.L.navh_core_get_layer(self.h, (name or self.layer_name).encode(), out.ctypes.data, out.size)
        assert n == out.size
        return out

    def set_layer(self, data, name=None):
        d = np.ascontiguousarray(data, np.float32)
        assert d.shape == (self.cols, self.rows)
        self.L.navh_core_set_layer(self.h, (name or self.layer_name).encode(), d.ctypes.data)

    def line(self, sx, sy, ex, ey, cap=1 << 16):
        buf = np.zeros((cap, 2), np.int32)
        n = self.L.navh_core_line(self.h, sx, sy, ex, ey, buf.ctypes.data, cap)
        return buf[:min(n, cap)].copy()

    def blocked(self, x, y, radius=0.3, name=None):
        return bool(self.L.navh_core_blocked(self.h, (name or self.layer_name).encode(), x, y, radius))
