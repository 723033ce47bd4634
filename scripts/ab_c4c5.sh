#!/bin/bash
# Runs on the GPU box: C4 and C5 kernel times for every variants/*.so (see build_variants.sh).
cp ros_navigation_b200/csrc/libb200nav.so /tmp/default.so
for v in variants/*.so; do
  cp $v ros_navigation_b200/csrc/libb200nav.so
  for wl in c4 c5; do
    timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('$v $wl', 'value %.0f' % d['value'], 'step %.4f' % d['ms_per_step'], {a: round(b,4) for a,b in k.items() if a in ('himm_prep','himm_tile','vfh_update')})"
  done
done
cp /tmp/default.so ros_navigation_b200/csrc/libb200nav.so
