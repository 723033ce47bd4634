#!/bin/bash
# A/B on the GPU box: heavy (origin) tiles on multi-warp CTAs (default) against one warp per tile (B200NAV_MW_HEAVY=0)
for mw in 0 1 2; do
  B200NAV_MW_HEAVY=$mw python bench.py --steps 40 --warmup 8 --no-cpu 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('MW=$mw value %.0f ms %.4f e2e %.0f (%.4f ms) kernels %s' % (l['value'], l['ms_per_step'], l['e2e']['value'], l['e2e']['ms_per_step'], {k: round(v,4) for k,v in l["kernel_ms_per_step"].items() if isinstance(v, float)}))
for k,v in (l.get('other_workloads') or {}).items():
    print('   ',k, round(v.get('value',0)), v.get('ms_per_scan', v.get('ms_per_step')), v.get('kernel_ms', v.get('kernel_ms_per_step')))
print('    cold', l.get('cold_grid',{}).get('ms_per_step'))
"
done
