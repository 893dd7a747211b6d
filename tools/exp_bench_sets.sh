#!/bin/bash
# bench step with different (r x C) sets routed to the mma.sync kernel; prints local-correlation ms per scale
for set in "" "4x32" "4x32,6x64,7x64" "4x32,6x64,7x64,2x16"; do
  GFB_MMA_AUTO="$set" python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); r=d['roofline']
print('set=[$set]', 'ms/step %.3f'%d['ms_per_step'], 'frac %.3f'%r['frac'], ' '.join('%s=%.3f'%(k.replace('pass','p').replace('_scale','s'),v['ms_per_step']) for k,v in r['by_scale'].items()))
"
done
