#!/bin/bash
# Runs on the GPU box: racecheck of scripts/sanitize.py for every variants/*.so, multi-warp pipeline kernel (default for
# the small fleets of sanitize.py) and one-warp kernel (B200NAV_MW_HEAVY=0).
cp ros_navigation_b200/csrc/libb200nav.so /tmp/default.so
for v in variants/*.so; do
  cp $v ros_navigation_b200/csrc/libb200nav.so
  for mw in 1 0; do
    B200NAV_MW_HEAVY=$mw timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitize.py > /tmp/rc.full 2>&1
    echo "$v MW_HEAVY=$mw: $(grep -E 'RACECHECK SUMMARY' /tmp/rc.full) $(grep -c 'OK' /tmp/rc.full) workloads OK; $(grep -E 'Error|Traceback' /tmp/rc.full | head -2 | tr '\n' ' ' | cut -c1-200)"
  done
done
cp /tmp/default.so ros_navigation_b200/csrc/libb200nav.so
