"""Host-side mirror of move_control::VFH (move_control/include/move_control/vfh.h:182-361) over the C ABI."""
import ctypes as C
import time

import numpy as np

from . import capi
from .capi import COMMAND_DTYPE, VFH_INPUT_DTYPE, VfhParamsC, check, lib, ptr


def VfhParams(**kw):
    """Steerer::initVfh defaults (move_control/src/steerer.cpp:69-121), overridable by keyword."""
    p = VfhParamsC()
    lib().b200nav_vfh_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


class VFH:
    """n_robots independent VFH+ controllers sharing one parameter set.

    Single-robot use mirrors the reference class: Update_VFH(ranges, speed, goal_dir, goal_dist, tol) returns
    (chosen_speed, chosen_turnrate); Hist / OriginHist are refreshed after each update (vfh.h:243-244).
    """

    def __init__(self, ctx, params=None, n_robots=1):
        self.ctx = ctx
        self.params = params if params is not None else VfhParams()
        h = C.c_void_p()
        check(lib().b200nav_vfh_create(ctx.h, C.byref(self.params), int(n_robots), C.byref(h)), ctx.h)
        self.h = h
        self.n_robots = n_robots
        self.hist_size = lib().b200nav_vfh_hist_size(h)
        self.num_tables = lib().b200nav_vfh_num_tables(h)
        self.window = self.params.window_diameter
        self.Hist = np.zeros(self.hist_size, np.float32)
        self.OriginHist = np.zeros(self.hist_size, np.float32)
        self._last_time = time.time()
        self._picked = 90.0

    # -- reference getters/setters ----------------------------------------------------------------------------
    def getHistSize(self):
        return self.hist_size

    def getSectorAngle(self):
        return self.params.sector_angle

    def GetPickedAngle(self):
        return self._picked

    def GetMaxTurnrate(self, speed):
        return lib().b200nav_vfh_get_max_turnrate(self.h, int(speed))

    def SetCurrentMaxSpeed(self, s):
        check(lib().b200nav_vfh_set_current_max_speed(self.h, int(s)), self.ctx.h)

    def disable_tma(self, disable=True):
        check(lib().b200nav_vfh_debug_disable_tma(self.h, int(disable)), self.ctx.h)

    @staticmethod
    def make_input(x=0.0, y=0.0, yaw=0.0, dt=0.2, speed=0, goal_dir=90.0, goal_dist=3000.0, tol=250.0):
        a = np.zeros(1, VFH_INPUT_DTYPE)
        a["x"], a["y"], a["yaw"], a["dt"] = x, y, yaw, dt
        a["current_speed"], a["goal_direction"], a["goal_distance"], a["goal_tolerance"] = speed, goal_dir, goal_dist, tol
        return a

    # -- updates ---------------------------------------------------------------------------------------------------
    def Update_VFH(self, ranges, current_speed, goal_direction, goal_distance, goal_distance_tolerance, dt=None,
                   robot=0):
        """VFH::Update_VFH (vfh.cpp:480-605).  dt defaults to the wall-clock time since the previous call, like the
        reference's gettimeofday bookkeeping; pass it explicitly for reproducible runs."""
        if dt is None:
            now = time.time()
            dt, self._last_time = now - self._last_time, now
        ranges = np.ascontiguousarray(ranges, dtype=np.float64).reshape(361, 2)
        inp = self.make_input(dt=dt, speed=current_speed, goal_dir=goal_direction, goal_dist=goal_distance,
                              tol=goal_distance_tolerance)
        out = np.zeros(1, COMMAND_DTYPE)
        check(lib().b200nav_vfh_update_ranges(self.h, robot, ranges.ctypes.data, inp.ctypes.data, out.ctypes.data),
              self.ctx.h)
        self._after(robot, out)
        return int(out["speed"][0]), int(out["turnrate"][0])

    def update_from_grid(self, grid, layer, inp, robot=0):
        """Steerer::getRangesFromSubmap + Update_VFH reading the device grid (steerer.cpp:260-263)."""
        inp = np.ascontiguousarray(inp, dtype=VFH_INPUT_DTYPE)
        out = np.zeros(1, COMMAND_DTYPE)
        check(lib().b200nav_vfh_update_grid(self.h, grid.h, layer.encode(), robot, inp.ctypes.data, out.ctypes.data),
              self.ctx.h)
        self._after(robot, out)
        return out[0]

    def update_batched(self, grid, layer, inputs):
        inputs = np.ascontiguousarray(inputs, dtype=VFH_INPUT_DTYPE)
        assert len(inputs) == self.n_robots
        out = np.zeros(self.n_robots, COMMAND_DTYPE)
        check(lib().b200nav_vfh_update_batched(self.h, grid.h, layer.encode(), inputs.ctypes.data, out.ctypes.data),
              self.ctx.h)
        return out

    def update_batched_async(self, grid, layer, host_inputs, host_out):
        """Enqueue-only form of update_batched: host buffers (pinned) in and out, valid until the context caught up."""
        check(lib().b200nav_vfh_update_batched_async(self.h, grid.h, layer.encode(), ptr(host_inputs), ptr(host_out)),
              self.ctx.h)

    def update_batched_dev(self, grid, layer, dev_inputs, dev_out):
        check(lib().b200nav_vfh_update_batched_dev(self.h, grid.h, layer.encode(), ptr(dev_inputs), ptr(dev_out)),
              self.ctx.h)

    def _after(self, robot, out):
        self._picked = float(out["picked_angle"][0])
        check(lib().b200nav_vfh_read_state(self.h, robot, self.OriginHist.ctypes.data, self.Hist.ctypes.data, None,
                                           None, None), self.ctx.h)

    # -- state / tables for parity checks ----------------------------------------------------------------------
    def state(self, robot=0):
        n = self.hist_size
        oh, h, lb = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        f, i = np.zeros(4, np.float32), np.zeros(2, np.int32)
        check(lib().b200nav_vfh_read_state(self.h, robot, oh.ctypes.data, h.ctypes.data, lb.ctypes.data,
                                           f.ctypes.data, i.ctypes.data), self.ctx.h)
        return dict(origin_hist=oh, hist=h, last_binary=lb, picked=float(f[0]), last_picked=float(f[1]),
                    desired=float(f[2]), blocked_radius=float(f[3]), last_chosen_speed=int(i[0]),
                    max_speed_for_picked=int(i[1]))

    def ranges(self, robot=0):
        r = np.zeros((361, 2), np.float64)
        check(lib().b200nav_vfh_read_ranges(self.h, robot, r.ctypes.data), self.ctx.h)
        return r

    def tables(self, table=0):
        W, nw = self.window, (self.hist_size + 31) // 32
        d, s, b = (np.zeros((W, W), np.float32) for _ in range(3))
        m = np.zeros((W, W, nw), np.uint32)
        mtr = np.zeros(self.params.max_speed + 1, np.int32)
        check(lib().b200nav_vfh_get_tables(self.h, table, d.ctypes.data, s.ctypes.data, b.ctypes.data, m.ctypes.data,
                                           mtr.ctypes.data), self.ctx.h)
        return d, s, b, m, mtr

    def close(self):
        if self.h:
            lib().b200nav_vfh_destroy(self.h)
            self.h = None
