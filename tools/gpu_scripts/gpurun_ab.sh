# A/B of library variants: VARIANTS="base libminirender_b200" SCENES="sphere bench cloud" [TESTS=1] [BENCH=1]
if [ "${TESTS:-1}" = 1 ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -3; fi
for v in ${VARIANTS:-base libminirender_b200}; do
  echo "== $v"; for sc in ${SCENES:-sphere bench cloud}; do MINIRENDER_B200_LIB=$PWD/minirender_b200/lib/$v.so python tools/stage_probe.py $sc 0 | grep flags; done
done
if [ "${BENCH:-1}" = 1 ]; then python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print('fps', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'stages', {k: round(v*1000,1) for k,v in d['stage_ms'].items()})"; fi
