#!/bin/bash
# Runs on the GPU box: single-robot configurations (C1, C2, C3) for every variants/*.so (see build_variants.sh).
cp ros_navigation_b200/csrc/libb200nav.so /tmp/default.so
for v in variants/*.so; do
  cp $v ros_navigation_b200/csrc/libb200nav.so
  python - "$v" <<'PY'
import sys, torch
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
for name in ("c1", "c2", "c3"):
    o = bench.single_robot_numbers(dev, name)
    print(sys.argv[1], name, "ms/scan %.4f" % o["ms_per_scan"], {k: round(v, 4) for k, v in o["kernel_ms"].items()}, flush=True)
PY
done
cp /tmp/default.so ros_navigation_b200/csrc/libb200nav.so
