"""Host-side mirrors of the reference's mapping classes, forwarding to the C ABI.

  DeviceGridMap    ~ grid_map::GridMap as MapProvider uses it (move_control/src/map_provider.cpp:17-41,145-149,
                     190-223): named float layers, geometry, move(), getSubmap's consumers read the device layer.
  LaserMapUpdater  ~ move_control::LaserMapUpdater (move_control/include/move_control/laser_map_updater.h:13-43,
                     move_control/src/laser_map_updater.cpp:7-21): buffers RangeSamples, updateMap() drains them
                     in order into its layer and returns the touched bounding box.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import SAMPLE_DTYPE, check, lib, ptr


class DeviceGridMap:
    def __init__(self, ctx, length, resolution, position=(0.0, 0.0), n_robots=1, layers=("master",)):
        self.ctx = ctx
        h = C.c_void_p()
        check(lib().b200nav_grid_create(ctx.h, float(length[0]), float(length[1]), float(resolution),
                                        float(position[0]), float(position[1]), int(n_robots), C.byref(h)), ctx.h)
        self.h = h
        r, c, n = C.c_int(), C.c_int(), C.c_int()
        lib().b200nav_grid_size(h, C.byref(r), C.byref(c), C.byref(n))
        self.rows, self.cols, self.n_robots = r.value, c.value, n.value
        self.resolution = float(resolution)
        for name in layers:
            self.add(name)

    # -- grid_map::GridMap surface -----------------------------------------------------------------------------
    def add(self, layer):
        check(lib().b200nav_grid_add_layer(self.h, layer.encode()), self.ctx.h)

    def alias(self, alias, target):
        check(lib().b200nav_grid_alias_layer(self.h, alias.encode(), target.encode()), self.ctx.h)

    def copy_layer(self, dst, src):
        """map_[dst] = map_[src] (MapProvider::composeMasterMapFromLayerdMap, map_provider.cpp:221)."""
        check(lib().b200nav_grid_copy_layer(self.h, dst.encode(), src.encode()), self.ctx.h)

    def compose_master(self, dst="master", range_layer="range", laser_layer="laser"):
        """dst = range (+) laser as in the commented-out compose of map_provider.cpp:218-220."""
        check(lib().b200nav_grid_compose_master(self.h, dst.encode(), range_layer.encode(), laser_layer.encode()),
              self.ctx.h)

    def clear(self, layer=None):
        check(lib().b200nav_grid_clear(self.h, layer.encode() if layer else None), self.ctx.h)

    def upload(self, layer, data, robot=0):
        """data: [cols][rows] float32 (== column-major rows x cols)."""
        data = np.ascontiguousarray(data, dtype=np.float32)
        assert data.size == self.rows * self.cols
        check(lib().b200nav_grid_upload(self.h, robot, layer.encode(), data.ctypes.data), self.ctx.h)

    def download(self, layer, robot=0):
        out = np.empty((self.cols, self.rows), dtype=np.float32)
        check(lib().b200nav_grid_download(self.h, robot, layer.encode(), out.ctypes.data), self.ctx.h)
        return out

    def set_geometry(self, robot, position, start_index=(0, 0)):
        check(lib().b200nav_grid_set_geometry(self.h, robot, float(position[0]), float(position[1]),
                                              int(start_index[0]), int(start_index[1])), self.ctx.h)

    def get_geometry(self, robot=0):
        x, y, s0, s1 = C.c_double(), C.c_double(), C.c_int(), C.c_int()
        check(lib().b200nav_grid_get_geometry(self.h, robot, C.byref(x), C.byref(y), C.byref(s0), C.byref(s1)),
              self.ctx.h)
        return (x.value, y.value), (s0.value, s1.value)

    def move(self, position, robot=0):
        moved = C.c_int()
        check(lib().b200nav_grid_move(self.h, robot, float(position[0]), float(position[1]), C.byref(moved)),
              self.ctx.h)
        return bool(moved.value)

    def to_occupancy(self, layer="master", data_min=0.0, data_max=255.0, robot=0):
        out = np.empty(self.rows * self.cols, dtype=np.int8)
        check(lib().b200nav_grid_to_occupancy(self.h, robot, layer.encode(), data_min, data_max, out.ctypes.data),
              self.ctx.h)
        return out

    def query_blocked(self, points, radius=0.3, layer="master", robot=0):
        """MapGlobalPlanner::ifBlocked for an array of points [n,2] -> bool[n] (map_global_planner.h:39-54)."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)
        out = np.zeros(len(pts), np.uint8)
        check(lib().b200nav_grid_query_blocked(self.h, robot, layer.encode(), pts.ctypes.data, len(pts), float(radius),
                                               out.ctypes.data), self.ctx.h)
        return out.astype(bool)

    def layer_written(self, layer, robot=-1):
        check(lib().b200nav_grid_layer_written(self.h, layer.encode(), int(robot)), self.ctx.h)

    def layer_format(self, layer):
        """"coded" (one byte per cell, the default) or "float" (the reference's float matrix)."""
        rc = lib().b200nav_grid_layer_format(self.h, layer.encode())
        if rc < 0:
            check(rc, self.ctx.h)
        return "coded" if rc == 1 else "float"

    def layer_devptr(self, layer):
        return lib().b200nav_grid_layer_devptr(self.h, layer.encode())

    # -- HIMM -----------------------------------------------------------------------------------------------------
    def himm_update(self, layer, samples, robot=0, bbox=None):
        samples = np.ascontiguousarray(samples, dtype=SAMPLE_DTYPE)
        check(lib().b200nav_himm_update(self.h, robot, layer.encode(), samples.ctypes.data, len(samples), ptr(bbox)),
              self.ctx.h)

    def himm_update_batched(self, layer, samples, offsets, bbox=None):
        samples = np.ascontiguousarray(samples, dtype=SAMPLE_DTYPE)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        assert len(offsets) == self.n_robots + 1
        check(lib().b200nav_himm_update_batched(self.h, layer.encode(), samples.ctypes.data, offsets.ctypes.data,
                                                ptr(bbox)), self.ctx.h)

    def himm_update_batched_dev(self, layer, dev_samples, dev_offsets, total, max_per_robot):
        check(lib().b200nav_himm_update_batched_dev(self.h, layer.encode(), ptr(dev_samples), ptr(dev_offsets),
                                                    int(total), int(max_per_robot)), self.ctx.h)

    def himm_update_cloud_batched(self, layer, origins, xy, clear_end, offsets, bbox=None):
        """Compact form: origins [n_robots,2] f64, xy [total,2] f32, clear_end [total] u8 or None, offsets."""
        origins = np.ascontiguousarray(origins, dtype=np.float64)
        xy = np.ascontiguousarray(xy, dtype=np.float32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        ce = None if clear_end is None else np.ascontiguousarray(clear_end, dtype=np.uint8)
        check(lib().b200nav_himm_update_cloud_batched(self.h, layer.encode(), origins.ctypes.data, xy.ctypes.data,
                                                      None if ce is None else ce.ctypes.data, offsets.ctypes.data,
                                                      ptr(bbox)), self.ctx.h)

    @staticmethod
    def scan_info(angle_min, angle_increment, range_min, range_max, n_ranges, decimate=True):
        from .capi import SCAN_INFO_DTYPE
        info = np.zeros(1, SCAN_INFO_DTYPE)
        info["angle_min"], info["angle_increment"] = angle_min, angle_increment
        info["range_min"], info["range_max"] = range_min, range_max
        info["n_ranges"], info["decimate"] = n_ranges, int(bool(decimate))
        return info

    def himm_update_scans_batched(self, layer, info, poses, ranges):
        """Scan form (b200nav_himm_update_scans_batched): poses [n_robots,3] f64 sensor x, y, yaw; ranges
        [n_robots, n_ranges] f32; the projection runs inside the binning kernel."""
        poses = np.ascontiguousarray(poses, dtype=np.float64)
        ranges = np.ascontiguousarray(ranges, dtype=np.float32)
        assert poses.shape == (self.n_robots, 3) and ranges.shape == (self.n_robots, int(info["n_ranges"][0]))
        check(lib().b200nav_himm_update_scans_batched(self.h, layer.encode(), info.ctypes.data, poses.ctypes.data,
                                                      ranges.ctypes.data), self.ctx.h)

    def himm_update_scans_batched_async(self, layer, info, host_poses, host_ranges):
        """Enqueue-only form: pinned host buffers that stay valid until the context caught up."""
        check(lib().b200nav_himm_update_scans_batched_async(self.h, layer.encode(), info.ctypes.data, ptr(host_poses),
                                                            ptr(host_ranges)), self.ctx.h)

    def himm_update_scans_batched_dev(self, layer, info, dev_poses, dev_ranges):
        check(lib().b200nav_himm_update_scans_batched_dev(self.h, layer.encode(), info.ctypes.data, ptr(dev_poses),
                                                          ptr(dev_ranges)), self.ctx.h)

    def himm_update_cloud_batched_async(self, layer, origins, xy, clear_end, offsets):
        """Enqueue-only form: the arguments must be host buffers (pinned torch tensors / numpy arrays) that stay valid
        and unchanged until the context has caught up (Context.wait / synchronize)."""
        check(lib().b200nav_himm_update_cloud_batched_async(self.h, layer.encode(), ptr(origins), ptr(xy),
                                                            ptr(clear_end), ptr(offsets)), self.ctx.h)

    def himm_update_cloud_batched_dev(self, layer, dev_origins, dev_xy, dev_clear_end, dev_offsets, total,
                                      max_per_robot):
        check(lib().b200nav_himm_update_cloud_batched_dev(self.h, layer.encode(), ptr(dev_origins), ptr(dev_xy),
                                                          ptr(dev_clear_end), ptr(dev_offsets), int(total),
                                                          int(max_per_robot)), self.ctx.h)

    def himm_last_stats(self):
        """(cell visits, marks, beams) of the last update (roofline accounting)."""
        out = np.zeros(3, np.int64)
        check(lib().b200nav_himm_last_stats(self.h, out.ctypes.data), self.ctx.h)
        return int(out[0]), int(out[1]), int(out[2])

    def close(self):
        if self.h:
            lib().b200nav_grid_destroy(self.h)
            self.h = None


class LaserMapUpdater:
    """Same surface as move_control::LaserMapUpdater minus the ROS intake: samples are pushed by the caller
    (bufferIncomingMsg's output, laser_map_updater.cpp:53-70) and updateMap() drains them in order."""

    def __init__(self, grid, sensor_type="laser", robot=0):
        self.grid = grid
        self.type_name = sensor_type
        self.robot = robot
        grid.add(sensor_type)  # MapUpdater ctor: add the layer if it does not exist (map_updater.h:10-14)
        self._buffer = []

    def getTypeName(self):
        return self.type_name

    def push_samples(self, samples):
        self._buffer.append(np.ascontiguousarray(samples, dtype=SAMPLE_DTYPE))

    def updateMap(self, bbox):
        """bbox: np.float64[4] = minX,minY,maxX,maxY in/out (laser_map_updater.cpp:7-21)."""
        if not self._buffer:
            return
        samples = np.concatenate(self._buffer)
        self._buffer = []
        self.grid.himm_update(self.type_name, samples, robot=self.robot, bbox=bbox)
