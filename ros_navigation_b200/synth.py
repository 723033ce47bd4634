"""Seeded synthetic worlds, trajectories and laser scans (workload synthesis for tests and bench.py).

Not part of the hot path: it only produces the inputs the reference's intake side would deliver
(MapUpdater::RangeSample buffers, move_control/include/move_control/map_updater.h:28-32, and the pose/goal scalars of
Steerer::update, move_control/src/steerer.cpp:221-263).  Plain torch, runs on CPU or GPU.

World: a rectangular room covering 90 % of the map extent, centred on the map, plus
K = ceil(area/25 m^2) random discs (r in [0.15, 0.6] m) and K boxes (side in [0.3, 1.5] m).  Rays are cast
analytically in fp64; range noise N(0, 0.01 m) as in custom_description/description/lidar.urdf.xacro:47-51.
Scan -> sample projection mirrors laser_geometry as used by LaserMapUpdater::bufferIncomingMsg
(move_control/src/laser_map_updater.cpp:38-73): end points are float32 x/y in the map frame, readings with
r >= range_max are dropped (or, with keep_max=True, kept with clear_end=1).
Trajectory: Lissajous x = A sin(2 pi t/T1), y = B sin(2 pi t/T2 + pi/3), A = B = 30 % of the extent, yaw = heading.
"""
import math

import numpy as np
import torch

SAMPLE_BYTES = 40


class Worlds:
    """n independent worlds with identical extent."""

    def __init__(self, n, extent, seed, device="cpu"):
        self.n, self.extent, self.device = n, float(extent), torch.device(device)
        g = torch.Generator(device="cpu")
        g.manual_seed(int(seed))
        k = max(1, math.ceil(extent * extent / 25.0))
        self.k = k
        half = 0.45 * extent
        self.room = half  # room = [-half, half]^2
        u = lambda *s: torch.rand(*s, generator=g, dtype=torch.float64)
        self.disc_c = ((u(n, k, 2) - 0.5) * 2 * half * 0.95).to(self.device)
        self.disc_r = (0.15 + 0.45 * u(n, k)).to(self.device)
        self.box_c = ((u(n, k, 2) - 0.5) * 2 * half * 0.95).to(self.device)
        self.box_h = (0.15 + 0.6 * u(n, k, 2)).to(self.device)
        # per-robot trajectory parameters
        self.T1 = (60.0 + 60.0 * u(n)).to(self.device)
        self.T2 = (45.0 + 60.0 * u(n)).to(self.device)
        self.phase = (2 * math.pi * u(n)).to(self.device)
        self.noise_gen = torch.Generator(device=self.device)
        self.noise_gen.manual_seed(int(seed) + 7919)

    def pose(self, t):
        """Robot pose at time t (seconds): x, y, yaw tensors [n] (fp64)."""
        A = 0.30 * self.extent
        w1, w2 = 2 * math.pi / self.T1, 2 * math.pi / self.T2
        a1, a2 = w1 * t + self.phase, w2 * t + self.phase + math.pi / 3
        x, y = A * torch.sin(a1), A * torch.sin(a2)
        vx, vy = A * w1 * torch.cos(a1), A * w2 * torch.cos(a2)
        return x, y, torch.atan2(vy, vx)

    def goal(self, t, lookahead_s=15.0):
        """Goal = the trajectory point lookahead_s ahead; returns (goal_dir deg, goal_dist mm) as Steerer::update
        computes them (steerer.cpp:232-256), float32."""
        x, y, yaw = self.pose(t)
        gx, gy, _ = self.pose(t + lookahead_s)
        dx = ((gx - x) * 1000.0).float()
        dy = ((gy - y) * 1000.0).float()
        dist = torch.hypot(dx, dy)
        a = torch.atan2(dy.double(), dx.double()) - yaw + math.pi / 2
        two_pi = 2 * math.pi
        npos = torch.fmod(torch.fmod(a, two_pi) + two_pi, two_pi)
        return (npos * 180.0 / math.pi).float(), dist

    def cast(self, x, y, yaw, n_beams, fov, range_max, noise_std=0.01, chunk=256):
        """Ranges [n, n_beams] (fp64) of a lidar at pose (x, y, yaw); beam angles span [-fov/2, fov/2)."""
        ang = (-fov / 2 + fov * torch.arange(n_beams, dtype=torch.float64, device=self.device) / n_beams)
        out = torch.empty(self.n, n_beams, dtype=torch.float64, device=self.device)
        for s in range(0, self.n, chunk):
            e = min(self.n, s + chunk)
            th = yaw[s:e, None] + ang[None, :]
            dx, dy = torch.cos(th), torch.sin(th)
            ox, oy = x[s:e, None], y[s:e, None]
            inf = torch.full_like(dx, float("inf"))
            # room walls (from inside): exit distance of the box [-room, room]^2
            tx = torch.where(dx > 0, (self.room - ox) / dx, torch.where(dx < 0, (-self.room - ox) / dx, inf))
            ty = torch.where(dy > 0, (self.room - oy) / dy, torch.where(dy < 0, (-self.room - oy) / dy, inf))
            t = torch.minimum(tx, ty)
            # discs
            cx, cy, r = self.disc_c[s:e, None, :, 0], self.disc_c[s:e, None, :, 1], self.disc_r[s:e, None, :]
            fx, fy = ox[..., None] - cx, oy[..., None] - cy
            b = fx * dx[..., None] + fy * dy[..., None]
            c = fx * fx + fy * fy - r * r
            disc = b * b - c
            td = -b - torch.sqrt(torch.clamp(disc, min=0))
            td = torch.where((disc > 0) & (td > 0), td, torch.full_like(td, float("inf")))
            t = torch.minimum(t, td.min(dim=-1).values)
            # boxes (slab method)
            bx, by = self.box_c[s:e, None, :, 0], self.box_c[s:e, None, :, 1]
            hx, hy = self.box_h[s:e, None, :, 0], self.box_h[s:e, None, :, 1]
            idx = 1.0 / torch.where(dx == 0, torch.full_like(dx, 1e-300), dx)[..., None]
            idy = 1.0 / torch.where(dy == 0, torch.full_like(dy, 1e-300), dy)[..., None]
            t1x, t2x = (bx - hx - ox[..., None]) * idx, (bx + hx - ox[..., None]) * idx
            t1y, t2y = (by - hy - oy[..., None]) * idy, (by + hy - oy[..., None]) * idy
            tn = torch.maximum(torch.minimum(t1x, t2x), torch.minimum(t1y, t2y))
            tf = torch.minimum(torch.maximum(t1x, t2x), torch.maximum(t1y, t2y))
            tb = torch.where((tn <= tf) & (tn > 0), tn, torch.full_like(tn, float("inf")))
            t = torch.minimum(t, tb.min(dim=-1).values)
            if noise_std > 0:
                t = t + noise_std * torch.randn(t.shape, generator=self.noise_gen, dtype=torch.float64,
                                                device=self.device)
            out[s:e] = torch.clamp(t, min=0.05)
        return out, ang


def samples_from_scan(x, y, yaw, ranges, ang, range_max, keep_max=False):
    """Project scans to RangeSample records.

    Returns (samples uint8 [total, 40] laid out as b200nav_sample, offsets int32 [n+1]); beam order is preserved
    per robot."""
    n, nb = ranges.shape
    dev = ranges.device
    hit = ranges < range_max
    r = torch.clamp(ranges, max=range_max)
    th = yaw[:, None] + ang[None, :]
    ex = (x[:, None] + r * torch.cos(th)).float().double()  # float32 cloud points (PointCloud2 x/y)
    ey = (y[:, None] + r * torch.sin(th)).float().double()
    keep = torch.ones_like(hit) if keep_max else hit
    counts = keep.sum(dim=1)
    offsets = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    offsets[1:] = torch.cumsum(counts, 0).int()
    sx = x[:, None].expand(n, nb)[keep]
    sy = y[:, None].expand(n, nb)[keep]
    total = int(sx.numel())
    buf = torch.empty(total, 5, dtype=torch.float64, device=dev)
    buf[:, 0], buf[:, 1], buf[:, 2], buf[:, 3] = sx, sy, ex[keep], ey[keep]
    ints = torch.zeros(total, 2, dtype=torch.int32, device=dev)
    ints[:, 0] = (~hit[keep]).int()
    buf[:, 4] = ints.view(torch.float64).reshape(total)
    return buf.view(torch.uint8).reshape(total, SAMPLE_BYTES), offsets


def cloud_from_scan(x, y, yaw, ranges, ang, range_max, keep_max=False):
    """Same projection as samples_from_scan in the compact cloud form of b200nav_himm_update_cloud_batched:
    (origins f64 [n,2], xy f32 [total,2], clear_end u8 [total], offsets i32 [n+1])."""
    n, nb = ranges.shape
    dev = ranges.device
    hit = ranges < range_max
    r = torch.clamp(ranges, max=range_max)
    th = yaw[:, None] + ang[None, :]
    ex = (x[:, None] + r * torch.cos(th)).float()
    ey = (y[:, None] + r * torch.sin(th)).float()
    keep = torch.ones_like(hit) if keep_max else hit
    offsets = torch.zeros(n + 1, dtype=torch.int32, device=dev)
    offsets[1:] = torch.cumsum(keep.sum(dim=1), 0).int()
    xy = torch.stack([ex[keep], ey[keep]], dim=1).contiguous()
    clear = (~hit[keep]).to(torch.uint8).contiguous()
    origins = torch.stack([x, y], dim=1).contiguous()
    return origins, xy, clear, offsets


def samples_to_numpy(samples_u8):
    from .capi import SAMPLE_DTYPE
    a = samples_u8.detach().cpu().contiguous().numpy()
    return a.view(SAMPLE_DTYPE).reshape(-1)


def vfh_inputs(worlds, t, dt, current_speed):
    """b200nav_vfh_input records (uint8 [n, 48]) for time t."""
    x, y, yaw = worlds.pose(t)
    gdir, gdist = worlds.goal(t)
    n, dev = worlds.n, worlds.device
    buf = torch.empty(n, 6, dtype=torch.float64, device=dev)
    buf[:, 0], buf[:, 1], buf[:, 2], buf[:, 3] = x, y, yaw, float(dt)
    tail = torch.empty(n, 4, dtype=torch.int32, device=dev)
    tail[:, 0] = current_speed if torch.is_tensor(current_speed) else int(current_speed)
    tail[:, 1] = gdir.view(torch.int32)
    tail[:, 2] = gdist.view(torch.int32)
    tail[:, 3] = torch.full((n,), 250.0, dtype=torch.float32, device=dev).view(torch.int32)
    buf[:, 4:6] = tail.view(torch.float64).reshape(n, 2)
    return buf.view(torch.uint8).reshape(n, 48)


def vfh_inputs_to_numpy(inp_u8):
    from .capi import VFH_INPUT_DTYPE
    return inp_u8.detach().cpu().contiguous().numpy().view(VFH_INPUT_DTYPE).reshape(-1)


# The five BASELINE.json configurations.
CONFIGS = {
    "c1": dict(extent=10.0, res=0.05, beams=360, fov=2 * math.pi, range_max=3.0, robots=1, window=30, cell=100.0,
               submap=1.5, rate=5.0),
    "c2": dict(extent=102.4, res=0.05, beams=1080, fov=1.5 * math.pi, range_max=30.0, robots=1, window=30, cell=100.0,
               submap=1.5, rate=40.0),
    "c3": dict(extent=163.84, res=0.02, beams=4096, fov=2 * math.pi, range_max=60.0, robots=1, window=129, cell=20.0,
               submap=2.58, rate=10.0),
    "c4": dict(extent=25.6, res=0.05, beams=1080, fov=1.5 * math.pi, range_max=12.0, robots=1024, window=30,
               cell=100.0, submap=1.5, rate=5.0),
    "c5": dict(extent=12.8, res=0.05, beams=1080, fov=1.5 * math.pi, range_max=6.0, robots=16384, window=30,
               cell=100.0, submap=1.5, rate=5.0),
}


def config_seed(name, rank=0):
    idx = sorted(CONFIGS).index(name) + 1
    return 0x5EED0000 + idx * 65536 + rank
