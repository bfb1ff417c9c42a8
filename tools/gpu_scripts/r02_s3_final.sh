# session 3, final: GPU tests, smoke, default bench line, reference arm
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/s3_final_gputests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py 2>&1 | tail -1 > gpurun_out/s3_final_bench.json
python -c "
import json; d=json.load(open('gpurun_out/s3_final_bench.json')); print({k:d.get(k) for k in ('value','ms_per_step','steps','e2e','roofline','cpu_baseline','clocks','gpu_launches')})"
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
