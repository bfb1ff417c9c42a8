ncu --set full --clock-control none --import-source on -k regex:'k_raster|k_setup|k_vertex' -s 6 -c 3 -o gpurun_out/prof_r1_final python tools/prof_run.py sphere 4 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c k_raster gpurun_out/launches_final.csv
