mkdir -p gpurun_out
for i in 1 2 3; do python bench.py --steps 5 --warmup 3 --no-cpu --only cfr 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)['cfr']; print('cfr us/iter', round(d['us_per_iteration'],2), 'many ms', round(d['many_deals']['ms'],3))
"; done > gpurun_out/cfr_repeat.txt 2>&1
