python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:'k_raster|k_setup|k_vertex' -s 6 -c 3 -o gpurun_out/prof_s3_final python tools/prof_run.py sphere 4 > gpurun_out/prof.log 2>&1
tail -1 gpurun_out/prof.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_s3_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c k_raster gpurun_out/launches_s3_final.csv
python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_s3_final.json; python -c "
import json; d=json.load(open('gpurun_out/bench_s3_final.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','e2e_blocking','e2e_rgb8','roofline','frame_roofline','cpu_baseline','clocks','gpu_launches','warm_l2_pipelined','stage_ms')})"
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
python bench.py --workload turntable2m --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
python tools/stage_probe.py benchtex 0 | grep flags
python tools/stage_probe.py sphere2m 0 | grep flags
