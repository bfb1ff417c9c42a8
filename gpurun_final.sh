python -m pytest tests -m gpu -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
compute-sanitizer --tool racecheck --print-limit 3 python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "golden" 2>&1 | tail -4
python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_r1.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r1.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline','cpu_baseline','clocks','gpu_launches','warm_l2_pipelined')})"
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
