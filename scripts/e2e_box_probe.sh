#!/bin/bash
# End-to-end number of ONE rank on whatever box this runs on (the 8-GPU hosts feed a single GPU more slowly than the
# 1-GPU hosts): three repetitions, then the same under taskset to the first 8 cores.
for i in 1 2 3; do python bench.py --steps 40 --warmup 8 --no-cpu --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('plain    value %.0f ms %.4f e2e %.0f (%.4f ms) enqueue %.4f' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['e2e']['ms_per_step'], l['e2e_host_enqueue_ms_per_step']))"; done
for i in 1 2; do taskset -c 0-7 python bench.py --steps 40 --warmup 8 --no-cpu --no-extra 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print('taskset  value %.0f ms %.4f e2e %.0f (%.4f ms) enqueue %.4f' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['e2e']['ms_per_step'], l['e2e_host_enqueue_ms_per_step']))"; done
python scripts/e2e_probe.py 2>&1 | tail -12
