# ncu --set full capture of the three frame kernels (source-level) + launch list of a short bench run
ncu --set full --clock-control none --import-source on -k regex:'k_raster|k_setup|k_vertex' -s 6 -c 3 -o gpurun_out/prof_${TAG:-cur} python tools/prof_run.py ${SCENE:-sphere} 4 > gpurun_out/prof.log 2>&1
tail -1 gpurun_out/prof.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG:-cur}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -E "k_vertex|k_setup|k_raster" gpurun_out/launches_${TAG:-cur}.csv | tail -6 | cut -c1-200
