"""ctypes bindings of include/b200nav.h.  Fails loudly if the CUDA library is missing."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libb200nav.so")

OK, EINVAL, ECUDA, ENOMEM, ENOLAYER, ERANGE, ENODEVICE = 0, -1, -2, -3, -4, -5, -6
CMD_EMERGENCY, CMD_HEMMED_IN, CMD_CANT_TURN, CMD_NO_SUBMAP = 1, 2, 4, 8

SAMPLE_DTYPE = np.dtype([("sx", "<f8"), ("sy", "<f8"), ("ex", "<f8"), ("ey", "<f8"), ("clear_end", "<i4"),
                         ("reserved", "<i4")])
VFH_INPUT_DTYPE = np.dtype([("x", "<f8"), ("y", "<f8"), ("yaw", "<f8"), ("dt", "<f8"), ("current_speed", "<i4"),
                            ("goal_direction", "<f4"), ("goal_distance", "<f4"), ("goal_tolerance", "<f4")])
SCAN_INFO_DTYPE = np.dtype([("angle_min", "<f4"), ("angle_increment", "<f4"), ("range_min", "<f4"),
                            ("range_max", "<f4"), ("n_ranges", "<i4"), ("decimate", "<i4")])
COMMAND_DTYPE = np.dtype([("speed", "<i4"), ("turnrate", "<i4"), ("picked_angle", "<f4"), ("flags", "<u4")])
assert SAMPLE_DTYPE.itemsize == 40 and VFH_INPUT_DTYPE.itemsize == 48 and COMMAND_DTYPE.itemsize == 16


class VfhParamsC(C.Structure):
    _fields_ = [("cell_size", C.c_double), ("window_diameter", C.c_int32), ("sector_angle", C.c_int32),
                ("safety_dist_0ms", C.c_double), ("safety_dist_1ms", C.c_double), ("max_speed", C.c_int32),
                ("max_speed_narrow_opening", C.c_int32), ("max_speed_wide_opening", C.c_int32),
                ("max_acceleration", C.c_int32), ("min_turnrate", C.c_int32), ("max_turnrate_0ms", C.c_int32),
                ("max_turnrate_1ms", C.c_int32), ("reserved0", C.c_int32),
                ("min_turn_radius_safety_factor", C.c_double), ("free_space_cutoff_0ms", C.c_double),
                ("obs_cutoff_0ms", C.c_double), ("free_space_cutoff_1ms", C.c_double), ("obs_cutoff_1ms", C.c_double),
                ("weight_desired_dir", C.c_double), ("weight_current_dir", C.c_double), ("robot_radius", C.c_double),
                ("submap_length", C.c_double), ("occupied_threshold", C.c_double)]


class B200NavError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200nav error %d: %s" % (code, msg))
        self.code = code


# Every symbol include/b200nav.h declares (checked by tests/test_capi_symbols.py against the header text).
_SIGNATURES = {
    "b200nav_ctx_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200nav_ctx_destroy": (C.c_int, [C.c_void_p]),
    "b200nav_ctx_synchronize": (C.c_int, [C.c_void_p]),
    "b200nav_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "b200nav_last_error": (C.c_char_p, [C.c_void_p]),
    "b200nav_ctx_launch_count": (C.c_int64, [C.c_void_p]),
    "b200nav_ctx_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "b200nav_ctx_profile_read": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "b200nav_ctx_profile_select": (C.c_int, [C.c_void_p, C.c_char_p]),
    "b200nav_himm_last_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200nav_himm_update_cloud_batched": (C.c_int, [C.c_void_p, C.c_char_p] + [C.c_void_p] * 5),
    "b200nav_himm_update_cloud_batched_dev": (C.c_int, [C.c_void_p, C.c_char_p] + [C.c_void_p] * 4 + [C.c_int, C.c_int]),
    "b200nav_grid_create": (C.c_int, [C.c_void_p] + [C.c_double] * 5 + [C.c_int, C.POINTER(C.c_void_p)]),
    "b200nav_grid_destroy": (C.c_int, [C.c_void_p]),
    "b200nav_grid_size": (C.c_int, [C.c_void_p] + [C.POINTER(C.c_int)] * 3),
    "b200nav_grid_add_layer": (C.c_int, [C.c_void_p, C.c_char_p]),
    "b200nav_grid_alias_layer": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "b200nav_grid_copy_layer": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "b200nav_grid_clear": (C.c_int, [C.c_void_p, C.c_char_p]),
    "b200nav_grid_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]),
    "b200nav_grid_download": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p]),
    "b200nav_grid_set_geometry": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]),
    "b200nav_grid_get_geometry": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                            C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b200nav_grid_move": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int)]),
    "b200nav_grid_to_occupancy": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_float, C.c_float, C.c_void_p]),
    "b200nav_grid_query_blocked": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p]),
    "b200nav_grid_layer_written": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "b200nav_grid_layer_devptr": (C.c_void_p, [C.c_void_p, C.c_char_p]),
    "b200nav_ctx_flush_l2": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t]),
    "b200nav_ctx_calibrate_red": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_double)]),
    "b200nav_ctx_fence": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "b200nav_ctx_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "b200nav_himm_update_cloud_batched_async": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                          C.c_void_p]),
    "b200nav_vfh_update_batched_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]),
    "b200nav_scan_select": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_float)]),
    "b200nav_himm_update_scans_batched": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200nav_himm_update_scans_batched_async": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200nav_himm_update_scans_batched_dev": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200nav_fleet_unique_id": (C.c_int, [C.c_void_p]),
    "b200nav_fleet_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "b200nav_fleet_gather_async": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]),
    "b200nav_fleet_push_region": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "b200nav_fleet_push_connect": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200nav_fleet_table": (C.c_void_p, [C.c_void_p, C.c_int]),
    "b200nav_fleet_release": (C.c_int, [C.c_void_p, C.c_int]),
    "b200nav_fleet_cycle_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int,
                                            C.c_void_p]),
    "b200nav_fleet_cycle_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "b200nav_steer_update_goals": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200nav_vfh_update_batched_dev_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]),
    "b200nav_fleet_status": (C.c_int, [C.c_void_p]),
    "b200nav_fleet_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "b200nav_fleet_destroy": (C.c_int, [C.c_void_p]),
    "b200nav_grid_compose_master": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]),
    "b200nav_grid_has_layer": (C.c_int, [C.c_void_p, C.c_char_p]),
    "b200nav_grid_layer_format": (C.c_int, [C.c_void_p, C.c_char_p]),
    "b200nav_himm_update": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p]),
    "b200nav_himm_update_batched": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200nav_himm_update_batched_dev": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "b200nav_vfh_default_params": (None, [C.POINTER(VfhParamsC)]),
    "b200nav_vfh_create": (C.c_int, [C.c_void_p, C.POINTER(VfhParamsC), C.c_int, C.POINTER(C.c_void_p)]),
    "b200nav_vfh_destroy": (C.c_int, [C.c_void_p]),
    "b200nav_vfh_set_current_max_speed": (C.c_int, [C.c_void_p, C.c_int]),
    "b200nav_vfh_hist_size": (C.c_int, [C.c_void_p]),
    "b200nav_vfh_num_tables": (C.c_int, [C.c_void_p]),
    "b200nav_vfh_get_max_turnrate": (C.c_int, [C.c_void_p, C.c_int]),
    "b200nav_vfh_update_ranges": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200nav_vfh_update_grid": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_void_p]),
    "b200nav_vfh_update_batched": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]),
    "b200nav_vfh_update_batched_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]),
    "b200nav_vfh_read_state": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 5),
    "b200nav_vfh_read_ranges": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "b200nav_vfh_get_tables": (C.c_int, [C.c_void_p, C.c_int] + [C.c_void_p] * 5),
}
_DEBUG_SIGNATURES = {
    "b200nav_vfh_debug_disable_tma": (C.c_int, [C.c_void_p, C.c_int]),
    "b200nav_himm_debug_tile_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200nav_himm_debug_batch_stats": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200nav_himm_debug_free_summary": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_void_p, C.c_int]),
}

_lib = None


def build(verbose=False):
    """Compile csrc/libb200nav.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("libb200nav.so build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def lib():
    """Load the CUDA library.  Raises (no fallback) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback for this path)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in list(_SIGNATURES.items()) + list(_DEBUG_SIGNATURES.items()):
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, ctx=None):
    if rc != OK:
        msg = lib().b200nav_last_error(ctx)
        raise B200NavError(rc, msg.decode() if msg else "")


def ptr(a):
    """Address of a numpy array / torch tensor / int (device or host)."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError(type(a))


class Context:
    """b200nav_ctx: one CUDA device + stream."""

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        rc = lib().b200nav_ctx_create(int(device), stream, C.byref(h))
        if rc != OK:
            msg = lib().b200nav_last_error(None)
            raise B200NavError(rc, msg.decode() if msg else "")
        self.h = h
        self.device = device

    def synchronize(self):
        check(lib().b200nav_ctx_synchronize(self.h), self.h)

    def flush_l2(self, write_bytes, read_bytes=0):
        check(lib().b200nav_ctx_flush_l2(self.h, int(write_bytes), int(read_bytes)), self.h)

    def calibrate_red(self, buffer_bytes):
        """Unordered 4-byte reductions per second at random words of a scratch buffer (measurement aid)."""
        v = C.c_double()
        check(lib().b200nav_ctx_calibrate_red(self.h, int(buffer_bytes), C.byref(v)), self.h)
        return v.value

    def fence(self):
        """Mark the current end of the stream; returns a ticket for wait()."""
        t = C.c_int()
        check(lib().b200nav_ctx_fence(self.h, C.byref(t)), self.h)
        return t.value

    def wait(self, ticket):
        """Block until everything enqueued before fence() -> ticket has completed."""
        check(lib().b200nav_ctx_wait(self.h, int(ticket)), self.h)

    @property
    def stream(self):
        return lib().b200nav_ctx_stream(self.h)

    @property
    def launches(self):
        return int(lib().b200nav_ctx_launch_count(self.h))

    def profile_enable(self, on=True):
        check(lib().b200nav_ctx_profile_enable(self.h, int(on)), self.h)

    def profile_select(self, names=None):
        """Time only the named kernels (iterable of names; None: all) while profiling is enabled."""
        arg = ",".join(names).encode() if names else None
        check(lib().b200nav_ctx_profile_select(self.h, arg), self.h)

    def profile_read(self, name):
        """(total device ms, timed launches) of kernel `name` since profile_enable(True)."""
        ms, n = C.c_double(), C.c_int64()
        check(lib().b200nav_ctx_profile_read(self.h, name.encode(), C.byref(ms), C.byref(n)), self.h)
        return ms.value, n.value

    def close(self):
        if self.h:
            lib().b200nav_ctx_destroy(self.h)
            self.h = None
