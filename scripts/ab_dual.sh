#!/bin/bash
# Runs on the GPU box: variants/<name>.so with B200NAV_TILE_CTAS_PER_SM taken from the name's trailing number (if any).
cp ros_navigation_b200/csrc/libb200nav.so /tmp/default.so
for v in variants/*.so; do
  cp $v ros_navigation_b200/csrc/libb200nav.so
  n=$(basename $v .so | grep -o '[0-9]*$')
  export B200NAV_TILE_CTAS_PER_SM=${n:-28}
  echo "$v (CTAs/SM $B200NAV_TILE_CTAS_PER_SM): $(timeout 600 python -m pytest tests/test_himm_gpu.py tests/test_layer_formats_gpu.py -m gpu -x -q 2>&1 | tail -1)"
  for rep in 1 2; do
    timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('$v', 'value %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'], 'prep %.4f tile %.4f vfh %.4f' % (k['himm_prep'], k['himm_tile'], k['vfh_update']))"
  done
done
cp /tmp/default.so ros_navigation_b200/csrc/libb200nav.so
