#!/usr/bin/env python
"""bench.py -- laser scans/sec fused into the HIMM grid + VFH+ steering decisions/sec (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload c4] [--impl reference]

A step = one cycle of the batched perception-to-steering path: every robot fuses ONE laser scan into its HIMM grid
(b200nav_himm_update_batched*) and takes ONE VFH+ decision from the window of that grid
(b200nav_vfh_update_batched*); with N > 1 ranks the per-robot 16-byte steering commands are all-gathered (NCCL).
Default workload = BASELINE config 4 ("batched 1024 independent robots, each a 512x512 grid and 1080-beam scans"):
robots are block-partitioned over the ranks with no data-path collective.  --scaling weak (default) gives every GPU
the configuration's 1024 robots (N x 1024 in total); --scaling strong splits the 1024 robots over the N GPUs.  Inputs are synthetic (ros_navigation_b200/synth.py), pre-staged in HBM for `value` and in pinned host
memory for `e2e`.  L2 is flushed between timed steps; every step is timed with CUDA events on the launching stream.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "laser scans/sec fused into HIMM grid + VFH+ steering decisions/sec"
UNIT = "scans/s (each scan = 1 HIMM grid update + 1 VFH+ decision)"
N_CYCLES = 8          # distinct pre-generated cycles replayed round-robin (inputs > L2 at N=1)
L2_FLUSH_BYTES = 256 << 20
L2_FLUSH_LIGHT_BYTES = 160 << 20   # > 126 MB of L2


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------------------
# workload construction (shared by both arms)
# ----------------------------------------------------------------------------------------------------------------
def local_robot_range(total, rank, world):
    from ros_navigation_b200.dist import partition
    return partition(total, rank, world)


class Cycles:
    """Pre-generated scan cycles for the robots [lo, hi) of a config."""

    def __init__(self, cfg_name, lo, hi, device, n_cycles=N_CYCLES, want_samples=False):
        import torch
        from ros_navigation_b200 import synth
        self.cfg = cfg = synth.CONFIGS[cfg_name]
        self.n = hi - lo
        self.device = device
        seed = synth.config_seed(cfg_name, rank=lo)
        self.worlds = synth.Worlds(self.n, cfg["extent"], seed, device=device)
        dt = 1.0 / cfg["rate"]
        self.dt = dt
        self.samples, self.offsets, self.inputs, self.totals = [], [], [], []
        self.origins, self.xy, self.clear = [], [], []
        self.poses, self.ranges32 = [], []   # scan form (end-to-end path): sensor pose + raw float32 ranges
        for c in range(n_cycles):
            t = c * dt * 5  # spread the replayed poses a little so cycles are distinct
            x, y, yaw = self.worlds.pose(t)
            r, ang = self.worlds.cast(x, y, yaw, cfg["beams"], cfg["fov"], cfg["range_max"])
            if want_samples:  # 40-byte RangeSamples (CPU arm)
                s8, off = synth.samples_from_scan(x, y, yaw, r, ang, cfg["range_max"])
                self.samples.append(s8.contiguous())
            else:             # compact cloud form (GPU arm): origin per robot + float32 end points
                org, xy, clr, off = synth.cloud_from_scan(x, y, yaw, r, ang, cfg["range_max"])
                self.poses.append(torch.stack([x, y, yaw], dim=1).contiguous())
                self.ranges32.append(r.float().contiguous())
                self.origins.append(org)
                self.xy.append(xy)
                self.clear.append(clr)
            self.offsets.append(off.contiguous())
            self.totals.append(int(off[-1]))
            self.inputs.append(synth.vfh_inputs(self.worlds, t, dt, 150).contiguous())
        if device.type == "cuda":
            torch.cuda.synchronize(device)


# ----------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, load_sm = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                if len(parts) > 9 and float(parts[9]) > 0:
                    load_sm.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            use = sorted(load_sm) if load_sm else sorted(sm)
            out.update(sm_mhz=use[len(use) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       samples_under_load=len(load_sm),
                       window="warm-up + timed steps + end-to-end steps (GPU busy throughout)")
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU algorithm on the host cores (oracle restatement + reference vfh.cpp)
# ----------------------------------------------------------------------------------------------------------------
class CpuArm:
    """The reference's CPU implementation of the path on the host cores, for a bounded sample of robots.

    kind "reference" (configs whose geometry the reference's MapProvider / Steerer hard-code: 5 cm cells, 30-cell VFH
    window, 1.5 m submap = c1, c2, c4, c5): oracle/_ref/libnav_ref.so = the reference's own move_control and
    grid_map_core sources compiled in place.  Every robot is one reference node (MapProvider + LaserMapUpdater +
    Steerer + VFH); a cycle is ONE native call (navh_fleet_cycle_samples) that block-partitions the robots over
    std::threads - no Python in the loop: the scan's RangeSamples go into the LaserMapUpdater's buffer, then
    MapProvider::updateMap (per-cell string-keyed layer lookup, master := laser copy) and Steerer::update (submap copy
    of all layers, pseudo-scan, VFH::Update_VFH) run as written.
    kind "port" (c3: 2 cm cells / 129-cell window cannot be configured in the reference's node classes, or when
    oracle/_ref is missing): oracle restatement of HIMM + pseudo-scan and the reference vfh.cpp, one robot per core."""

    def __init__(self, cfg_name, n_robots, n_cycles=4):
        import numpy as np
        import torch
        from oracle import navref as N
        from oracle import oracle as O
        from ros_navigation_b200 import synth
        self.O, self.N, self.np = O, N, np
        self.n = n_robots
        cyc = Cycles(cfg_name, 0, n_robots, torch.device("cpu"), n_cycles=n_cycles, want_samples=True)
        self.cfg = cfg = cyc.cfg
        self.cores = max(1, min(os.cpu_count() or 1, n_robots))
        self.samples = [synth.samples_to_numpy(s) for s in cyc.samples]
        self.offsets = [o.numpy() for o in cyc.offsets]
        self.inputs = [synth.vfh_inputs_to_numpy(i) for i in cyc.inputs]
        self.n_cycles = n_cycles
        self.submap = cfg["submap"]
        node_shaped = (cfg["res"] == 0.05 and cfg["window"] == 30 and cfg["cell"] == 100.0 and cfg["submap"] == 1.5)
        self.cycle_no = 0
        if N.have_ref() and node_shaped and not os.environ.get("B200NAV_CPU_PORT"):
            self.kind = "reference"
            self.parts = ("reference move_control + grid_map_core sources compiled in place (oracle/_ref/libnav_ref.so): "
                          "MapProvider::updateMap + Steerer::update per robot, std::thread pool inside one native call")
            self.nodes = [N.Node(N.REF_PATH, cfg["extent"], cfg["extent"], False, t0=1.0) for _ in range(n_robots)]
            self.poses = [np.stack([i["x"], i["y"], i["yaw"]], 1) for i in self.inputs]
            self.speeds = [i["current_speed"].astype(np.float64) / 1000.0 for i in self.inputs]
            self.goals = []
            for c in range(n_cycles):   # the waypoint behind synth.vfh_inputs' goal_direction / goal_distance
                gx, gy, _ = cyc.worlds.pose(c * cyc.dt * 5 + 15.0)
                self.goals.append(np.stack([gx.numpy(), gy.numpy()], 1))
            return
        self.kind = "port"
        self.parts = ("HIMM + pseudo-scan: oracle restatement; VFH+: " +
                      ("unmodified reference vfh.cpp" if O.have_ref() else "not run (oracle/_ref missing)"))
        self.geom = O.make_geom(cfg["extent"], cfg["extent"], cfg["res"])
        self.layers = [O.new_layer(self.geom) for _ in range(n_robots)]
        self.vfh = None
        if O.have_ref():
            self.vfh = [O.RefVFH(window_diameter=cfg["window"], cell_size=cfg["cell"]) for _ in range(n_robots)]

    def _robot_cycle(self, r, c):
        O = self.O
        off = self.offsets[c]
        O.himm_update(self.geom, self.layers[r], self.samples[c][off[r]:off[r + 1]])
        inp = self.inputs[c][r]
        rng = O.ranges_from_submap(self.geom, self.layers[r], float(inp["x"]), float(inp["y"]), float(inp["yaw"]),
                                   self.submap)
        if self.vfh is not None:
            self.vfh[r].update(rng, int(inp["current_speed"]), float(inp["goal_direction"]),
                               float(inp["goal_distance"]), float(inp["goal_tolerance"]), float(inp["dt"]))

    def run_cycle(self, c):
        """One cycle over all sample robots on all cores; returns seconds."""
        c = c % self.n_cycles
        if self.kind == "reference":
            self.cycle_no += 1
            t0 = time.perf_counter()
            self.last_cmd = self.N.fleet_cycle_samples(self.nodes, 1.0 + 0.2 * self.cycle_no, self.poses[c],
                                                       self.samples[c], self.offsets[c], self.goals[c], self.speeds[c],
                                                       threads=self.cores)
            return time.perf_counter() - t0
        chunks = [list(range(i, self.n, self.cores)) for i in range(self.cores)]

        def work(rs):
            for r in rs:
                self._robot_cycle(r, c)

        t0 = time.perf_counter()
        ths = [threading.Thread(target=work, args=(rs,)) for rs in chunks if rs]
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        return time.perf_counter() - t0


def cpu_sample_size(cfg_name):
    """Robots in the CPU arm's bounded sample.  Single-robot configurations (c1-c3) are one robot on ONE core, as the
    reference runs them; batched configurations use a robot sample spread over all host cores (at least 8 robots per
    core so that the block partition is balanced)."""
    cores = os.cpu_count() or 1
    return {"c1": 1, "c2": 1, "c3": 1, "c4": max(256, 8 * cores), "c5": max(512, 16 * cores)}.get(cfg_name, 64)


def run_reference_arm(args, rank, world):
    """bench.py --impl reference: the CPU implementation on the host cores; rank 0 only."""
    if rank != 0:
        return
    arm = CpuArm(args.workload, cpu_sample_size(args.workload))
    for w in range(args.warmup):
        arm.run_cycle(w)
    secs = 0.0
    for k in range(args.steps):
        secs += arm.run_cycle(args.warmup + k)
    value = arm.n * args.steps / secs
    sample = "%d robots x %d cycles of workload %s per run; %s" % (arm.n, args.steps, args.workload, arm.parts)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * secs / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32 cells / f64 geometry", "data": "synthetic",
        "config": workload_config(args.workload, world, (args.robots or arm.cfg["robots"]) * (world if args.scaling == "weak" else 1)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(name, world, robots_total):
    from ros_navigation_b200 import synth
    cfg = synth.CONFIGS[name]
    rows = int(round(cfg["extent"] / cfg["res"]))
    return {
        "workload": "%s: %d robots x %dx%d grid @%g m, %d-beam scans <= %g m, VFH+ 72 sectors window %d" % (
            name, robots_total, rows, rows, cfg["res"], cfg["beams"], cfg["range_max"], cfg["window"]),
        "robots": robots_total, "grid": [rows, rows], "beams": cfg["beams"], "parallelism": "robots/%d" % world,
        "l2": "flushed between timed steps (256 MiB write + 256 MiB read); %d distinct cycles replayed" % N_CYCLES,
    }


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
class GpuArm:
    def __init__(self, cfg_name, lo, hi, device, stream, world, robots_total=None):
        robots_total = robots_total or (hi - lo)
        import torch
        from ros_navigation_b200 import VFH, DeviceGridMap, VfhParams, capi
        self.torch = torch
        self.world = world
        self.cyc = Cycles(cfg_name, lo, hi, device)
        cfg = self.cfg = self.cyc.cfg
        self.n = hi - lo
        self.ctx = capi.Context(device.index, stream=stream.cuda_stream)
        self.grid = DeviceGridMap(self.ctx, (cfg["extent"],) * 2, cfg["res"], n_robots=self.n, layers=("laser",))
        self.grid.alias("master", "laser")  # map_["master"] = map_["laser"] without the copy
        self.vfh = VFH(self.ctx, VfhParams(window_diameter=cfg["window"], cell_size=cfg["cell"],
                                           submap_length=cfg["submap"]), n_robots=self.n)
        from ros_navigation_b200.dist import CommandExchange
        self.exchange = (CommandExchange(robots_total, device, ctx=None if os.environ.get("B200NAV_TORCH_EXCHANGE") else self.ctx,
                                         peer_push=bool(os.environ.get("B200NAV_PEER_PUSH")))
                         if (world > 1 and not os.environ.get("B200NAV_BENCH_NO_EXCHANGE")) else None)
        self.cmd = self.exchange.local if self.exchange else torch.zeros(self.n, 16, dtype=torch.uint8, device=device)
        self.gathered = self.exchange.table if self.exchange else None
        # pinned host copies for the end-to-end path
        self.h_origins = [o.cpu().pin_memory() for o in self.cyc.origins]
        self.h_xy = [o.cpu().pin_memory() for o in self.cyc.xy]
        self.h_clear = [o.cpu().pin_memory() for o in self.cyc.clear]
        self.h_offsets = [o.cpu().pin_memory() for o in self.cyc.offsets]
        self.h_inputs = [i.cpu().pin_memory() for i in self.cyc.inputs]
        self.h_poses = [o.cpu().pin_memory() for o in self.cyc.poses]
        self.h_ranges = [o.cpu().pin_memory() for o in self.cyc.ranges32]
        self.scan_info = self.grid.scan_info(-cfg["fov"] / 2, cfg["fov"] / cfg["beams"], 0.0, cfg["range_max"],
                                             cfg["beams"], decimate=False)
        self.d_inputs_e2e = torch.zeros_like(self.cyc.inputs[0])
        self.h_cmd = torch.zeros((robots_total if world > 1 else self.n), 16, dtype=torch.uint8).pin_memory()
        self.h_cmds = [self.h_cmd, torch.zeros_like(self.h_cmd).pin_memory()]

    def all_gather(self):
        if self.exchange is not None:
            self.exchange.gather()

    def step_dev(self, i, last=False):
        """One device-resident cycle.  With N > 1 the all-gather of this cycle's commands is started asynchronously
        (double-buffered) and overlaps the next cycle's kernels; the last timed step waits for everything."""
        c = i % N_CYCLES
        self.grid.himm_update_cloud_batched_dev("laser", self.cyc.origins[c], self.cyc.xy[c], self.cyc.clear[c],
                                                self.cyc.offsets[c], self.cyc.totals[c], self.cfg["beams"])
        if self.exchange is None:
            self.vfh.update_batched_dev(self.grid, "master", self.cyc.inputs[c], self.cmd)
            return
        slot = i & 1
        if self.exchange.push:  # fused: the VFH+ kernel stores the commands into every rank's table itself
            self.exchange.wait(1 - slot)      # the previous cycle's table is complete (a consumer would read it here)
            self.exchange.release(1 - slot)   # ... and has been read: peers may write that slot again
            self.exchange.vfh_update_push(self.vfh, self.grid, "master", self.cyc.inputs[c], slot)
        else:
            self.exchange.wait(slot)  # the gather that last read this buffer has finished
            self.vfh.update_batched_dev(self.grid, "master", self.cyc.inputs[c], self.exchange.locals[slot])
            self.exchange.gather_async(slot)
        if last:
            self.exchange.wait()

    def step_himm_only(self, c):
        self.grid.himm_update_cloud_batched_dev("laser", self.cyc.origins[c], self.cyc.xy[c], self.cyc.clear[c],
                                                self.cyc.offsets[c], self.cyc.totals[c], self.cfg["beams"])

    def step_e2e(self, i):
        """Host buffers in, host result out: H2D of samples/offsets/VFH inputs and D2H of the commands inside."""
        c = i % N_CYCLES
        from ros_navigation_b200.capi import check, lib
        check(lib().b200nav_himm_update_cloud_batched(self.grid.h, b"laser", self.h_origins[c].data_ptr(),
                                                      self.h_xy[c].data_ptr(), self.h_clear[c].data_ptr(),
                                                      self.h_offsets[c].data_ptr(), None), self.ctx.h)
        self.d_inputs_e2e.copy_(self.h_inputs[c], non_blocking=True)
        if self.exchange is not None:
            self.exchange.wait()
        self.vfh.update_batched_dev(self.grid, "master", self.d_inputs_e2e, self.cmd)
        self.all_gather()
        src = self.gathered if self.gathered is not None else self.cmd
        self.h_cmd.copy_(src, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()

    def run_e2e_pipelined(self, first, n, flush=True, depth=2, scans=True):
        """n end-to-end cycles through the asynchronous C-ABI calls, at most `depth` cycles in flight: every cycle
        copies its raw scans (float32 ranges + sensor poses; scans=False: the projected cloud) and VFH inputs from
        pinned host memory, runs the three kernels and copies
        the commands (N > 1: the all-gathered table) back to pinned host memory.  The cloud copy of cycle i+1 overlaps
        the tile / VFH+ kernels of cycle i.  The L2 flush runs in-stream between cycles (inside the caller's timed
        region)."""
        tickets = []
        busy = 0.0   # host time spent enqueueing (everything except the waits for tickets)
        for k in range(n):
            t_k = time.perf_counter()
            i = first + k
            c, slot = i % N_CYCLES, k & 1
            if flush:
                flush_l2_light(self)
            if scans:
                self.grid.himm_update_scans_batched_async("laser", self.scan_info, self.h_poses[c], self.h_ranges[c])
            else:
                self.grid.himm_update_cloud_batched_async("laser", self.h_origins[c], self.h_xy[c], self.h_clear[c],
                                                          self.h_offsets[c])
            if self.exchange is None:
                self.vfh.update_batched_async(self.grid, "master", self.h_inputs[c], self.h_cmds[slot])
            elif self.exchange.fleet is not None and not self.exchange.push:
                # N > 1 without torch in the loop: inputs up, VFH+ kernel, NCCL all-gather and the copy of the whole
                # fleet's table to pinned host memory are ONE enqueue-only library call on side streams
                from ros_navigation_b200.capi import check, lib
                if k >= 2:   # blocks until the slot's previous cycle is on the host: a wait, not enqueue work
                    t_w = time.perf_counter()
                    check(lib().b200nav_fleet_cycle_wait(self.exchange.fleet, slot), self.ctx.h)
                    busy -= time.perf_counter() - t_w
                check(lib().b200nav_fleet_cycle_async(self.exchange.fleet, self.vfh.h, self.grid.h, b"master",
                                                      self.h_inputs[c].data_ptr(), slot, self.h_cmds[slot].data_ptr()),
                      self.ctx.h)
            else:
                self.d_inputs_e2e.copy_(self.h_inputs[c], non_blocking=True)
                if self.exchange.push:
                    self.exchange.vfh_update_push(self.vfh, self.grid, "master", self.d_inputs_e2e, slot)
                else:
                    self.exchange.wait(slot)
                    self.vfh.update_batched_dev(self.grid, "master", self.d_inputs_e2e, self.exchange.locals[slot])
                    self.exchange.gather_async(slot)
                self.exchange.wait(slot)  # stream-side wait: the copy below follows the exchange
                self.h_cmds[slot].copy_(self.exchange.tables[slot], non_blocking=True)
                self.exchange.release(slot)  # peer push: the table has been read, peers may write the slot again
            tickets.append(self.ctx.fence())
            busy += time.perf_counter() - t_k
            if k >= depth:
                self.ctx.wait(tickets[k - depth])
        self.e2e_enqueue_s = busy
        self.ctx.synchronize()
        if self.exchange is not None and self.exchange.fleet is not None and not self.exchange.push:
            from ros_navigation_b200.capi import check, lib
            for slot in (0, 1):
                check(lib().b200nav_fleet_cycle_wait(self.exchange.fleet, slot), self.ctx.h)

    def e2e_bytes(self, c):
        h2d = self.h_poses[c].numel() * 8 + self.h_ranges[c].numel() * 4 + self.h_inputs[c].numel()
        return h2d, self.h_cmd.numel()

    def tile_stats(self):
        """(tile work items dropped by the known-free shortcut, items walked) since the previous call."""
        import numpy as np
        from ros_navigation_b200.capi import lib
        out = np.zeros(2, np.int64)
        lib().b200nav_himm_debug_tile_stats(self.grid.h, out.ctypes.data)
        b = np.zeros(2, np.int64)
        lib().b200nav_himm_debug_batch_stats(self.grid.h, b.ctypes.data)
        self.last_batch_stats = (int(b[0]), int(b[1]))   # 32-beam batches set up / dropped as "only re-clears free blocks"
        return int(out[0]), int(out[1])

    def algorithmic_bytes(self, c):
        """8 B per cell visit + 8 B per mark + 36 B per beam, for cycle c (this rank)."""
        self.step_himm_only(c)
        visits, marks, beams = self.grid.himm_last_stats()
        return 8 * visits + 8 * marks + 36 * beams, visits, marks, beams


def tile_phase(ctx):
    """Device time of the tile walk of an update: the one-warp-per-tile kernel (fleets) or the multi-warp kernel
    (fleets of up to 32 robots: every tile on an 8-warp CTA); an update launches exactly one of them.
    Returns (summed ms, number of updates, parts)."""
    one_ms, one_n = ctx.profile_read("himm_tile")
    mw_ms, mw_n = ctx.profile_read("himm_tile_mw")
    parts = {}
    if one_n:
        parts["himm_tile_coded_kernel"] = one_ms / one_n
    if mw_n:
        parts["himm_tile_coded_mw_kernel"] = mw_ms / mw_n
    return one_ms + mw_ms, max(one_n, mw_n), parts


def flush_l2_light(arm):
    """In-stream flush for the pipelined end-to-end loop: write a buffer larger than L2 (160 MiB > 126 MB).  The
    write-back of these lines happens inside the timed region like everything else there."""
    arm.ctx.flush_l2(L2_FLUSH_LIGHT_BYTES, 0)


def flush_l2(arm):
    """Write a 256 MiB buffer (evicts everything), then read another 256 MiB one so that the dirty lines of the
    write are themselves written back BEFORE the timed step starts (their write-back is not our kernels' traffic)."""
    arm.ctx.flush_l2(L2_FLUSH_BYTES, L2_FLUSH_BYTES)


def timed_steps(torch, stream, step_fn, first, n, arm):
    """n steps, each bracketed by CUDA events on `stream`, L2 flushed before every step. Returns ms list."""
    evs = []
    t0 = time.perf_counter()
    for k in range(n):
        flush_l2(arm)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        step_fn(first + k, last=(k == n - 1))
        b.record(stream)
        evs.append((a, b))
    arm.host_enqueue_ms_per_step = (time.perf_counter() - t0) * 1000.0 / max(n, 1)   # host side of the loop alone
    stream.synchronize()
    return [a.elapsed_time(b) for a, b in evs]


def run_gpu_arm(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from ros_navigation_b200 import synth
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    cfg = synth.CONFIGS[args.workload]
    per_config = args.robots or cfg["robots"]
    # weak scaling (default): every GPU carries the configuration's full robot count; strong: the configuration's
    # robots are split over the GPUs ("1024 robots sharded across 1/2/4/8 B200").
    robots_total = per_config * world if args.scaling == "weak" else per_config
    lo, hi = local_robot_range(robots_total, rank, world)
    stream = torch.cuda.Stream(device)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else \
        (6650.0, "fallback (B200_PROFILING.md)")

    with torch.cuda.stream(stream):
        arm = GpuArm(args.workload, lo, hi, device, stream, world, robots_total)
        # algorithmic bytes per cycle (also warms every cycle once)
        alg = [arm.algorithmic_bytes(c) for c in range(N_CYCLES)]
        flush_l2(arm)  # allocates the flush scratch outside the timed loops
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
            time.sleep(0.3)  # let nvidia-smi come up; sampling then covers warm-up, timed and end-to-end steps
        for w in range(max(args.warmup, 50)):
            arm.step_dev(w, last=True)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        # Inside the timed steps only the dominant kernel (the tile walk) is bracketed by events: every timed launch
        # costs two event records that keep the next launch from overlapping the previous kernel's tail (all three
        # kernels timed: +0.016 ms per cycle, scripts/event_overhead_probe.py).  The other kernels' durations come
        # from a second pass over the same steps below.
        arm.ctx.profile_select(("himm_tile", "himm_tile_mw"))
        arm.ctx.profile_enable(True)
        launches0 = arm.ctx.launches
        arm.tile_stats()   # reset the skipped / processed counters
        t_wall0 = time.perf_counter()
        ms = timed_steps(torch, stream, arm.step_dev, args.warmup, args.steps, arm)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t_wall0
        launches = arm.ctx.launches - launches0
        tiles_skipped, tiles_processed = arm.tile_stats()
        main_batches = arm.last_batch_stats
        tile_ms, tile_n, tile_parts = tile_phase(arm.ctx)
        arm.ctx.profile_enable(False)
        # ---- kernel pass (not part of `value`): the same steps with every kernel timed ----
        arm.ctx.profile_select(None)
        arm.ctx.profile_enable(True)
        ms_all_timed = timed_steps(torch, stream, arm.step_dev, args.warmup, args.steps, arm)
        torch.cuda.synchronize()
        prep_ms, prep_n = arm.ctx.profile_read("himm_prep")
        vfh_ms, vfh_n = arm.ctx.profile_read("vfh_update")
        tile2_ms, tile2_n, _ = tile_phase(arm.ctx)
        arm.ctx.profile_enable(False)
        # ---- outside the timed region: proof that the exchange delivered every rank's rows to every rank ----
        exchange_bad = None
        if arm.exchange is not None:
            exchange_bad = 0
            for k in range(2):   # one more cycle per slot, then checksum every rank's block on every rank
                arm.step_dev(args.warmup + args.steps + k, last=True)
                exchange_bad += arm.exchange.verify((args.warmup + args.steps + k) & 1)
        # ---- end-to-end: host buffers through the C ABI ----
        arm.run_e2e_pipelined(0, max(3, args.warmup // 2))
        # cost of the in-stream L2 flushes alone (reported next to the raw number, never subtracted from it)
        stream.synchronize()
        t0 = time.perf_counter()
        for k in range(args.steps):
            flush_l2_light(arm)
        stream.synchronize()
        flush_s = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        stream.synchronize()
        t0 = time.perf_counter()
        arm.run_e2e_pipelined(args.warmup, args.steps)
        e2e_s = time.perf_counter() - t0
        e2e_enqueue_ms = getattr(arm, "e2e_enqueue_s", 0.0) * 1000.0 / max(args.steps, 1)
        # the same loop without the explicit flush (reported beside the headline, never instead of it): a cycle's own
        # working set - touched tile records in and out, beam segments, masks, scans - already exceeds L2
        if world > 1:
            dist.barrier()
        stream.synchronize()
        t0 = time.perf_counter()
        arm.run_e2e_pipelined(args.warmup, args.steps, flush=False)
        e2e_noflush_s = time.perf_counter() - t0
        clk = clocks.stop() if rank == 0 else None

    # first-pass number (all ranks take part): layers cleared to NaN, the first N_CYCLES cycles timed
    cold = None
    if not args.no_extra:
        try:
            cold = cold_grid_numbers(torch, stream, arm, robots_total, alg, hbm_peak)
        except Exception as e:
            cold = {"error": repr(e)}
    # N > 1: the same run also shards the configurations' FIXED robot counts over the ranks (strong scaling):
    # BASELINE config 4 (1024 robots / N per GPU) and config 5 (16 384 robots / N per GPU, with the all-gather)
    sharded = None
    if world > 1 and not args.no_extra and args.scaling == "weak":
        sharded = {}
        for name in ("c4", "c5"):
            try:
                sharded[name] = sharded_numbers(torch, dist, device, name, rank, world)
            except Exception as e:
                sharded[name] = {"error": repr(e)}
    total_ms = float(sum(ms))
    t = torch.tensor([total_ms, e2e_s * 1000.0, e2e_noflush_s * 1000.0], dtype=torch.float64, device=device)
    by_rank = None
    if world > 1:
        mine = torch.tensor([total_ms / args.steps, tile_ms / max(tile_n, 1),
                             getattr(arm, "host_enqueue_ms_per_step", 0.0) or 0.0, e2e_enqueue_ms,
                             e2e_s * 1000.0 / args.steps], dtype=torch.float64, device=device)
        every = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(every, mine)
        by_rank = {"ms_per_step": [round(float(e[0]), 4) for e in every],
                   "tile_kernel_ms": [round(float(e[1]), 4) for e in every],
                   "host_enqueue_ms_per_step": [round(float(e[2]), 4) for e in every],
                   "e2e_host_enqueue_ms_per_step": [round(float(e[3]), 4) for e in every],
                   "e2e_ms_per_step": [round(float(e[4]), 4) for e in every]}
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, e2e_noflush_ms = float(t[0]), float(t[1]), float(t[2])
    value = robots_total * args.steps / (total_ms / 1000.0)
    e2e_value = robots_total * args.steps / (e2e_ms / 1000.0)

    if rank != 0:
        return
    # roofline of the dominant kernel (himm_tile_kernel), this rank's launches
    used = [alg[(args.warmup + k) % N_CYCLES] for k in range(args.steps)]
    alg_bytes = float(np.mean([u[0] for u in used]))
    tile_avg_ms = tile_ms / max(tile_n, 1)
    achieved = alg_bytes / (tile_avg_ms / 1000.0) / 1e9
    h2d, d2h = arm.e2e_bytes(0)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload)
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "u8 cell codes (f32 at the API) / f64 geometry", "data": "synthetic",
        "config": workload_config(args.workload, world, robots_total),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_ms / args.steps,
                "how": "wall clock over all timed cycles through the asynchronous C-ABI calls, 2 cycles in flight: pinned "
                       "host buffers in (raw float32 scans + sensor poses, projected on the device; VFH inputs), commands back to pinned host memory; L2 is flushed in-stream before every "
                       "cycle (160 MiB write) and that flush is INSIDE this time",
                "l2_flush_ms_per_step": flush_s * 1000.0 / args.steps,
                "without_explicit_flush": {
                    "value": robots_total * args.steps / (e2e_noflush_ms / 1000.0),
                    "ms_per_step": e2e_noflush_ms / args.steps,
                    "note": "same loop, no flush kernel between cycles: a cycle's own working set (about 100 MB of "
                            "touched tile records read and written, 35 MB of beam segments, masks, 4.5 MB of scans) "
                            "exceeds the 126 MB L2; reported for reference, the headline keeps the flush"}},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "tile walk = " + " + ".join(sorted(tile_parts)), "achieved": achieved,
                     "peak": hbm_peak, "kernel_parts_ms": tile_parts,
                     "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                     "avg_launch_ms": tile_avg_ms, "launches_timed": int(tile_n),
                     "visits_per_launch": float(np.mean([u[1] for u in used])),
                     "beam_batches_per_launch": {"set_up": main_batches[0] / max(args.steps, 1),
                                                 "dropped_only_free_blocks": main_batches[1] / max(args.steps, 1),
                                                 "note": "a 32-beam batch of a tile whose segments all lie in 8x8-cell "
                                                         "blocks known to be free (and not marked in this update) only "
                                                         "re-clears free cells: it is dropped before the walk; its visits "
                                                         "still count in algorithmic_bytes"},
                     "tile_items_per_launch": {"walked": tiles_processed / max(args.steps, 1),
                                               "dropped_known_free": tiles_skipped / max(args.steps, 1),
                                               "note": "steady state (free space already 0): a touched tile whose cells are "
                                                       "all 0 and that receives no mark cannot change and is dropped before any "
                                                       "copy; its visits are still counted in algorithmic_bytes (the reference "
                                                       "performs them) - see cold_grid for the first-pass number"}},
        "kernel_ms_per_step": {"himm_prep": prep_ms / max(prep_n, 1), "himm_tile": tile_avg_ms,
                               "vfh_update": vfh_ms / max(vfh_n, 1),
                               "how": "himm_tile: events around the tile kernel inside the timed steps (the roofline's "
                                      "launch time); himm_prep / vfh_update: a second pass over the same steps with "
                                      "every kernel timed (not part of `value`: the extra event records cost "
                                      "launch overlap)",
                               "all_kernels_timed_pass": {"ms_per_step": sum(ms_all_timed) / max(len(ms_all_timed), 1),
                                                          "himm_tile": tile2_ms / max(tile2_n, 1)}},
        "wall_s_timed_region": wall, "by_rank": by_rank,
        "host_enqueue_ms_per_step": getattr(arm, "host_enqueue_ms_per_step", None),
        "e2e_host_enqueue_ms_per_step": e2e_enqueue_ms, "cpu_affinity": args.cpu_affinity,
        "exchange_verified": (None if exchange_bad is None else
                              {"ranks": world, "mismatching_blocks": exchange_bad,
                               "how": "after the timed region: 2 more cycles, every rank checksums each rank's block of "
                                      "the table it received against the producer's own checksum (all-gathered)"}),
        "exchange": (None if arm.exchange is None else
                     ("peer push fused into the VFH+ kernel (NVLink P2P stores)" if arm.exchange.push else
                      ("NCCL all-gather, library binding" if arm.exchange.fleet is not None else
                       "torch.distributed all-gather"))),
    }
    # The resource that actually binds the tile kernel is warp-instruction issue (ncu: DRAM ~12 %, issue ~70 %): second
    # roofline = warp instructions per launch (ncu smsp__inst_executed.sum of the same workload, profiles/issue.json)
    # / this run's launch time, against 148 SMs x 4 schedulers x 1 instruction per clock at the sampled SM clock.
    try:
        issue = json.load(open(os.path.join(ROOT, "profiles", "issue.json"))).get(args.workload)
    except Exception:
        issue = None
    if issue and clk and clk.get("sm_mhz"):
        peak_issue = 148 * 4 * clk["sm_mhz"] * 1e6
        ach = issue["warp_instructions_per_launch"] / (tile_avg_ms / 1e3)
        line["roofline_issue"] = {"bound": "issue", "kernel": "himm_tile_coded_kernel", "achieved": ach / 1e9,
                                  "peak": peak_issue / 1e9, "unit": "G warp-instructions/s", "frac": ach / peak_issue,
                                  "warp_instructions_per_launch": issue["warp_instructions_per_launch"],
                                  "warp_instructions_per_visit": issue["warp_instructions_per_launch"] /
                                  max(float(np.mean([u[1] for u in used])), 1.0),
                                  "source": issue.get("source")}
    # SURVEY 8d's second denominator: cell visits per second against the rate at which this GPU retires unordered
    # 4-byte reductions at random addresses (measured in this run, outside the timed region) - the ceiling of a
    # one-atomic-per-visit design.  The tile kernel uses no atomics per visit (every cell has one owner); the number
    # says what that ordering-exact design costs or gains against the unordered one.
    try:
        red_l2 = arm.ctx.calibrate_red(32 << 20)
        red_hbm = arm.ctx.calibrate_red(4 << 30)
        visits_per_s = float(np.mean([u[1] for u in used])) / (tile_avg_ms / 1e3)
        line["roofline_red"] = {"bound": "unordered L2 reductions (not used by this path)", "unit": "G/s",
                                "visits_per_s": visits_per_s / 1e9,
                                "red_peak_32MiB_buffer": red_l2 / 1e9, "red_peak_4GiB_buffer": red_hbm / 1e9,
                                "visits_over_red_peak_32MiB": visits_per_s / red_l2,
                                "visits_over_red_peak_4GiB": visits_per_s / red_hbm,
                                "how": "b200nav_ctx_calibrate_red: RED.ADD.u32 at xorshift-random words, "
                                       "148 x 8 CTAs x 256 threads x 256 reductions, CUDA events"}
    except Exception as e:  # noqa: BLE001 - a measurement aid must not cost the bench line
        line["roofline_red"] = {"error": str(e)[:200]}
    # the two smaller kernels against the same HBM peak (SURVEY section 8d accounting; both are latency bound)
    beams = float(np.mean([u[3] for u in used]))
    n_sub = int(math.ceil(arm.cfg["submap"] / arm.cfg["res"])) + 1
    prep_avg, vfh_avg = prep_ms / max(prep_n, 1), vfh_ms / max(vfh_n, 1)
    prep_bytes = beams * (9 + 24)               # cloud point in (8 B xy + 1 B flag) + 24 B BeamSeg out
    vfh_bytes = arm.n * (4 * n_sub * n_sub + 2888 + 3 * 288 + 16)
    line["roofline_other"] = {
        "himm_prep_kernel": {"bytes_per_launch": prep_bytes, "achieved": prep_bytes / (prep_avg / 1e3) / 1e9,
                             "frac": prep_bytes / (prep_avg / 1e3) / 1e9 / hbm_peak, "unit": "GB/s",
                             "note": "instruction / latency bound (fp64 clipping, warp-aggregated RED.OR binning)"},
        "vfh_update_kernel": {"bytes_per_launch": vfh_bytes, "achieved": vfh_bytes / (vfh_avg / 1e3) / 1e9,
                              "frac": vfh_bytes / (vfh_avg / 1e3) / 1e9 / hbm_peak, "unit": "GB/s",
                              "note": "latency bound: one 128-thread CTA per robot, %.2f waves" % (arm.n / (148.0 * 8))},
    }
    # ---- CPU baseline on a bounded sample (rank 0, N == 1 only) ----
    if world == 1 and not args.no_cpu:
        try:
            os.sched_setaffinity(0, args.full_affinity)   # the CPU arm uses all host cores
            cpu = CpuArm(args.workload, cpu_sample_size(args.workload))
            t_cal = cpu.run_cycle(0)
            n_cyc = int(max(2, min(4000, 20.0 / max(t_cal, 1e-3))))   # the first (cold) cycle overestimates: lands at 10-15 s
            secs = sum(cpu.run_cycle(1 + k) for k in range(n_cyc))
            line["cpu_baseline"] = {"value": cpu.n * n_cyc / secs, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind,
                                    "sample": ("%d robots x %d cycles of workload %s (%.1f s); " + cpu.parts) % (
                                        cpu.n, n_cyc, args.workload, secs)}
            if cpu.kind == "reference":
                # beside the reference as written (per-cell string-keyed layer lookup, master copy, submap
                # copy of all layers) also the restatement with those hoisted: oracle HIMM + pseudo-scan + reference
                # vfh.cpp, one Python thread per core around native calls (a shorter sample: ~4 s)
                del cpu
                os.environ["B200NAV_CPU_PORT"] = "1"
                try:
                    port = CpuArm(args.workload, cpu_sample_size(args.workload))
                    t_cal = port.run_cycle(0)
                    n_cyc = int(max(2, min(500, 4.0 / max(t_cal, 1e-3))))
                    secs = sum(port.run_cycle(1 + k) for k in range(n_cyc))
                    line["cpu_baseline"]["hoisted_port"] = {
                        "value": port.n * n_cyc / secs, "cores": port.cores,
                        "sample": ("%d robots x %d cycles (%.1f s); " + port.parts) % (port.n, n_cyc, secs)}
                finally:
                    del os.environ["B200NAV_CPU_PORT"]
        except Exception as e:  # the baseline must never take the GPU number down
            line.setdefault("cpu_baseline", {})["error"] = repr(e)
        pin_to_gpu_numa_node(local_rank)   # back next to the GPU for the remaining GPU legs
    if cold is not None:
        line["cold_grid"] = cold
    if sharded is not None:
        line["strong_scaling"] = sharded
    if world == 1 and args.workload == "c4" and not args.no_extra:
        del arm
        torch.cuda.empty_cache()
        line["other_workloads"] = {}
        for key, fn in (("c1", single_robot_numbers), ("c2", single_robot_numbers), ("c3", single_robot_numbers),
                        ("c5", batched_numbers), ("c4_float_layers", float_layer_numbers)):
            try:
                line["other_workloads"][key] = fn(device, key)
            except Exception as e:
                line["other_workloads"][key] = {"error": repr(e)}
    print(json.dumps(line), flush=True)


def cold_grid_numbers(torch, stream, arm, robots_total, alg, hbm_peak):
    """The same workload on a map that knows nothing yet: every layer cleared to NaN, then the first N_CYCLES cycles
    (nothing is known free, every touched tile is walked, every end cell is marked for the first time)."""
    import numpy as np
    with torch.cuda.stream(stream):
        if arm.exchange is not None:
            arm.exchange.wait()
        arm.grid.clear("laser")
        stream.synchronize()
        if arm.world > 1:   # ranks reach this point at different times; a step's events include the exchange
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()
        arm.ctx.profile_enable(True)
        arm.tile_stats()
        ms = timed_steps(torch, stream, arm.step_dev, 0, N_CYCLES, arm)
        skipped, walked = arm.tile_stats()
        tile_ms, tile_n, _ = tile_phase(arm.ctx)
        arm.ctx.profile_enable(False)
    t = torch.tensor([float(sum(ms))], dtype=torch.float64, device=arm.cyc.device)
    if arm.world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t[0])
    alg_bytes = float(np.mean([a[0] for a in alg]))
    tile_avg = tile_ms / max(tile_n, 1)
    return {"value": robots_total * N_CYCLES / (total_ms / 1000.0), "unit": UNIT, "ms_per_step": total_ms / N_CYCLES,
            "steps": N_CYCLES, "himm_tile_ms": tile_avg,
            "roofline_frac": alg_bytes / (tile_avg / 1e3) / 1e9 / hbm_peak,
            "tile_items_per_launch": {"walked": walked / N_CYCLES, "dropped_known_free": skipped / N_CYCLES},
            "beam_batches_per_launch": {"set_up": arm.last_batch_stats[0] / N_CYCLES,
                                        "dropped_only_free_blocks": arm.last_batch_stats[1] / N_CYCLES},
            "how": "layer cleared to NaN, cycles 0..%d timed with CUDA events, L2 flushed before each" % (N_CYCLES - 1)}


def sharded_numbers(torch, dist, device, name, rank, world, steps=12, warm=40):
    """A configuration's own robot count split over the ranks (every rank calls this): device-timed like the main
    line (events per step, L2 flushed, barrier on both sides, max over ranks), exchange included and verified."""
    from ros_navigation_b200 import synth
    robots_total = synth.CONFIGS[name]["robots"]
    lo, hi = local_robot_range(robots_total, rank, world)
    stream = torch.cuda.Stream(device)
    with torch.cuda.stream(stream):
        arm = GpuArm(name, lo, hi, device, stream, world, robots_total)
        flush_l2(arm)   # allocates this context's flush scratch outside the timed loop (else the ranks skew)
        for w in range(warm):
            arm.step_dev(w, last=True)
        stream.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        ms = timed_steps(torch, stream, arm.step_dev, warm, steps, arm)
        torch.cuda.synchronize()
        dist.barrier()
        bad = 0
        if arm.exchange is not None:
            for k in range(2):
                arm.step_dev(warm + steps + k, last=True)
                bad += arm.exchange.verify((warm + steps + k) & 1)
    t = torch.tensor([float(sum(ms))], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = {"workload": workload_config(name, world, robots_total)["workload"], "robots_per_gpu": hi - lo,
           "value": robots_total * steps / (float(t[0]) / 1000.0), "unit": UNIT, "ms_per_step": float(t[0]) / steps,
           "scaling": "strong", "exchange_mismatching_blocks": bad}
    if arm.exchange is not None:
        arm.exchange.close()
    del arm
    torch.cuda.empty_cache()
    return out


def float_layer_numbers(device, name, steps=12):
    """C4 on FLOAT layers (the reference's own column-major float matrices in HBM) instead of the byte-coded tile
    records: what the path costs when the layer layout north_star names is kept on the device."""
    out = batched_numbers(device, "c4", steps=steps, warm=60, float_layers=True)
    out["layers"] = "float [robot][col][row] (Eigen::MatrixXf layout), 4 B per cell"
    out["roofline"]["kernel"] = "himm_tile_kernel (FloatView)"
    return out


def single_robot_numbers(device, name, scans=300):
    """The single-robot configurations (C1 200x200 / 360 beams, C2 2048x2048 / 1080 beams, C3 8192x8192 @ 2 cm /
    4096 beams / 129-cell VFH window): scans replayed back to back (latency bound: one robot cannot fill the GPU)."""
    import torch
    stream = torch.cuda.Stream(device)
    with torch.cuda.stream(stream):
        arm = GpuArm(name, 0, 1, device, stream, 1)
        for w in range(10):
            arm.step_dev(w)
        stream.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for k in range(scans):
            arm.step_dev(k)
        b.record(stream)
        stream.synchronize()
        dev_ms = a.elapsed_time(b)
        for w in range(5):
            arm.step_e2e(w)
        t0 = time.perf_counter()
        for k in range(scans):
            arm.step_e2e(k)
        e2e_s = time.perf_counter() - t0
    arm.ctx.profile_enable(True)
    with torch.cuda.stream(stream):
        for k in range(50):
            arm.step_dev(k)
    kms = {n: arm.ctx.profile_read(n)[0] / 50.0 for n in ("himm_prep", "himm_tile", "himm_tile_mw", "vfh_update")}  # ms per scan
    arm.ctx.profile_enable(False)
    return {"workload": workload_config(name, 1, 1)["workload"], "value": scans / (dev_ms / 1000.0),
            "e2e": scans / e2e_s, "unit": UNIT, "ms_per_scan": dev_ms / scans, "kernel_ms": kms,
            "l2": "no flush: consecutive scans of one robot hit the same tile records"}


def batched_numbers(device, name, steps=12, warm=12, float_layers=False):
    """A second batched configuration (BASELINE config 5: 16384 robots x 256x256) on this GPU: device-resident value
    and the tile kernel's roofline fraction, measured like the main line (events per step, L2 flushed)."""
    import numpy as np
    import torch
    from ros_navigation_b200 import synth
    stream = torch.cuda.Stream(device)
    robots = synth.CONFIGS[name]["robots"]
    with torch.cuda.stream(stream):
        arm = GpuArm(name, 0, robots, device, stream, 1)
        if float_layers:   # asking for the raw device pointer moves the layer to the float layout for good
            arm.grid.layer_devptr("laser")
            assert arm.grid.layer_format("laser") != "coded"
        alg = [arm.algorithmic_bytes(c) for c in range(N_CYCLES)]
        for w in range(warm):
            arm.step_dev(w)
        stream.synchronize()
        arm.ctx.profile_enable(True)
        arm.tile_stats()
        ms = timed_steps(torch, stream, arm.step_dev, warm, steps, arm)
        skipped, walked = arm.tile_stats()
        tile_ms, tile_n, tile_parts = tile_phase(arm.ctx)
        kms = {n: arm.ctx.profile_read(n)[0] / max(arm.ctx.profile_read(n)[1], 1) for n in ("himm_prep", "vfh_update")}
        kms["himm_tile"] = tile_ms / max(tile_n, 1)
        kms.update(tile_parts)
        arm.ctx.profile_enable(False)
    peak = 6543.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", peak)
    except Exception:
        pass
    used = float(np.mean([alg[(warm + k) % N_CYCLES][0] for k in range(steps)]))
    achieved = used / (tile_ms / max(tile_n, 1) / 1000.0) / 1e9
    out = {"workload": workload_config(name, 1, robots)["workload"], "value": robots * steps / (sum(ms) / 1000.0),
           "unit": UNIT, "ms_per_step": sum(ms) / steps, "kernel_ms_per_step": kms,
           "tile_items_per_launch": {"walked": walked / steps, "dropped_known_free": skipped / steps},
           "roofline": {"kernel": "himm_tile_coded_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "avg_launch_ms": tile_ms / max(tile_n, 1)}}
    # the VFH+ kernel against the issue peak where it is throughput bound (many waves): ncu instruction count of the
    # same workload (profiles/issue.json) / this run's launch time
    try:
        vi = json.load(open(os.path.join(ROOT, "profiles", "issue.json"))).get(name + "_vfh")
        mhz = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("sm_max_mhz", 1965.0)
    except Exception:
        vi, mhz = None, 1965.0
    if vi and not float_layers:
        ach = vi["warp_instructions_per_launch"] / (kms["vfh_update"] / 1e3)
        out["vfh_roofline_issue"] = {"kernel": "vfh_update_kernel", "achieved": ach / 1e9, "peak": 148 * 4 * mhz * 1e6 / 1e9,
                                     "unit": "G warp-instructions/s", "frac": ach / (148 * 4 * mhz * 1e6),
                                     "decisions_per_launch": robots, "us_per_launch": kms["vfh_update"] * 1e3,
                                     "source": vi.get("source")}
    del arm
    torch.cuda.empty_cache()
    return out


def pin_to_gpu_numa_node(local_rank):
    """Run this rank (and allocate its pinned buffers) on the CPUs of the NUMA node its GPU hangs off: with one
    process per GPU the end-to-end feed otherwise crosses the socket interconnect for half of the GPUs."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        cpus = set()
        for part in open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return "%d cpus of the GPU's NUMA node (%s)" % (len(cpus), bus)
    except Exception as e:
        return "unchanged (%s)" % type(e).__name__


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--robots", type=int, default=0, help="override the config's robot count")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: robots-per-GPU fixed (N x config robots in total); strong: config robots split over N")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip other_workloads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    # the GPU legs run on the CPUs next to the GPU (pinned buffers, enqueue thread); the CPU baseline leg gets every
    # host core back (run_gpu_arm restores args.full_affinity before it starts its threads)
    args.full_affinity = os.sched_getaffinity(0)
    args.cpu_affinity = pin_to_gpu_numa_node(local_rank)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the b200nav path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_gpu_arm(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
