"""Where does the end-to-end cycle time go?  H2D bandwidth, host enqueue cost, pipelined cycles with / without flush."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
stream = torch.cuda.Stream(dev)
with torch.cuda.stream(stream):
    arm = bench.GpuArm("c4", 0, 1024, dev, stream, 1)
    n = 40
    arm.run_e2e_pipelined(0, 8)
    # (a) raw H2D of the cloud
    d = torch.empty_like(arm.cyc.xy[0])
    stream.synchronize()
    t0 = time.perf_counter()
    for k in range(n):
        d.copy_(arm.h_xy[k % 8][: d.shape[0]] if arm.h_xy[k % 8].shape[0] >= d.shape[0] else arm.h_xy[0], non_blocking=True)
    stream.synchronize()
    dt = (time.perf_counter() - t0) / n
    print("H2D xy %.1f MB: %.3f ms -> %.1f GB/s" % (d.numel() * 4 / 1e6, dt * 1e3, d.numel() * 4 / dt / 1e9))
    for flush in (True, False):
        for depth in (1, 2, 4):
            stream.synchronize()
            t0 = time.perf_counter()
            arm.run_e2e_pipelined(8, n, flush=flush, depth=depth)
            print("pipelined flush=%s depth=%d: %.3f ms/cycle" % (flush, depth, (time.perf_counter() - t0) / n * 1e3))
    # host enqueue cost: time the enqueue loop alone (GPU far behind? then wait)
    stream.synchronize()
    t0 = time.perf_counter()
    for k in range(n):
        c = k % 8
        arm.grid.himm_update_cloud_batched_async("laser", arm.h_origins[c], arm.h_xy[c], arm.h_clear[c], arm.h_offsets[c])
        arm.vfh.update_batched_async(arm.grid, "master", arm.h_inputs[c], arm.h_cmds[k & 1])
    t1 = time.perf_counter()
    stream.synchronize()
    t2 = time.perf_counter()
    print("enqueue only: %.3f ms/cycle host, %.3f ms/cycle until done" % ((t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
    stream.synchronize()
    t0 = time.perf_counter()
    for k in range(n):
        arm.step_dev(k)
    stream.synchronize()
    print("device-resident back to back (no flush): %.3f ms/cycle" % ((time.perf_counter() - t0) / n * 1e3))
