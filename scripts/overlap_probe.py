"""How much does running the cycle as G independent robot groups on G streams buy (prep of one group overlapping the
tile walk of another)?  G contexts (each with its own stream) x 1024/G robots of workload c4, against one context."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
K = 40
for G in (1, 2, 4, 8):
    arms = []
    per = 1024 // G
    for g in range(G):
        st = torch.cuda.Stream(dev)
        with torch.cuda.stream(st):
            arms.append((bench.GpuArm("c4", g * per, (g + 1) * per, dev, st, 1), st))
    for w in range(60):
        for arm, st in arms:
            with torch.cuda.stream(st):
                arm.step_dev(w)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        t0 = time.perf_counter()
        for k in range(K):
            for arm, st in arms:
                with torch.cuda.stream(st):
                    arm.step_dev(k)
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / K)
    print("groups %d: %.4f ms per cycle of 1024 robots (no L2 flush, wall clock)" % (G, best * 1e3), flush=True)
    del arms
    torch.cuda.empty_cache()
