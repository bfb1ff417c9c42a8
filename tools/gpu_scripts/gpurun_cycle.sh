python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for e in 0 1; do
if [ $e = 1 ]; then export MR_NO_CLUSTER_CULL=1; echo "no cluster cull"; fi
python tools/stage_probe.py sphere 0 | grep flags
python tools/stage_probe.py bench 0 | grep flags
python tools/stage_probe.py cloud 0 | grep flags
python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('fps', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'warm', round(d['warm_l2_pipelined']['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'stages', {k: round(v*1000,1) for k,v in d['stage_ms'].items()})
"
done
