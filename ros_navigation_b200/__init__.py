"""ros_navigation_b200 -- B200-native HIMM mapping + VFH+ steering (drop-in for the hot path of
jmloveyj/ros_navigation: move_control::MapUpdater / LaserMapUpdater / VFH + Steerer::getRangesFromSubmap).

The product is the C-ABI library `csrc/libb200nav.so` (include/b200nav.h).  This package is the Python face used
by tests and bench.py: ctypes bindings (`capi`) and thin host-side mirrors of the reference classes (`mapping`,
`vfh`).  There is no CPU fallback: importing works without a GPU (so the symbol table can be checked), but creating
a context without a CUDA device raises.
"""
from . import capi  # noqa: F401
from .mapping import DeviceGridMap, LaserMapUpdater  # noqa: F401
from .vfh import VFH, VfhParams  # noqa: F401

__all__ = ["capi", "DeviceGridMap", "LaserMapUpdater", "VFH", "VfhParams"]
