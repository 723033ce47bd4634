#!/bin/bash
# Runs on the GPU box: VFH+ kernel time on C4 (1024 decisions per launch) and C5 (16 384) for every variants/*.so.
cp ros_navigation_b200/csrc/libb200nav.so /tmp/default.so
for v in variants/*.so; do
  cp $v ros_navigation_b200/csrc/libb200nav.so
  python - "$v" <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
for name in ("c4", "c5"):
    o = bench.batched_numbers(dev, name, steps=12, warm=12)
    k = o["kernel_ms_per_step"]
    print(sys.argv[1], name, "ms/step %.4f" % o["ms_per_step"], "vfh %.4f prep %.4f tile %.4f" % (k["vfh_update"], k["himm_prep"], k["himm_tile"]), flush=True)
PY
done
cp /tmp/default.so ros_navigation_b200/csrc/libb200nav.so
