python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/stage_probe.py sphere 0 | grep flags
python tools/stage_probe.py bench 0 | grep flags
python tools/stage_probe.py cloud 0 | grep flags
for n in 4 6; do echo "ctas/sm $n"; MR_RASTER_CTAS_PER_SM=$n python tools/stage_probe.py sphere 0 | grep flags;  MR_RASTER_CTAS_PER_SM=$n python tools/stage_probe.py cloud 0 | grep flags; done
python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('fps', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'warm', round(d['warm_l2_pipelined']['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'stages', {k: round(v*1000,1) for k,v in d['stage_ms'].items()})
"
