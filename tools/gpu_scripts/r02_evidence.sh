# round-2 evidence: GPU tests, smoke, bench line, launch list, --set full of the two frame kernels (source level),
# frame DRAM traffic without cache flushes (application replay), sanitizer logs. TAG names the outputs.
T=${TAG:-r02}
python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/gputests_$T.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 200 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_$T.json; cut -c1-300 gpurun_out/bench_$T.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -cE "k_geom|k_raster" gpurun_out/launches_$T.csv
ncu --set full --clock-control none --import-source on -k regex:'k_raster|k_geom' -s 8 -c 2 -o gpurun_out/prof_$T python tools/prof_run.py sphere 6 > gpurun_out/prof.log 2>&1
tail -1 gpurun_out/prof.log
ncu --replay-mode application --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_read_lookup_hit.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_red.sum -k regex:'k_raster|k_geom' -s 20 -c 6 --csv --log-file gpurun_out/frame_dram_$T.csv python tools/prof_run.py sphere 16 > /dev/null 2>&1
grep -c dram__bytes gpurun_out/frame_dram_$T.csv
if [ "${SANITIZE:-1}" = 1 ]; then
compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "golden or strips or immediate or empty or batch" > gpurun_out/sanitizer_memcheck_$T.txt 2>&1; tail -4 gpurun_out/sanitizer_memcheck_$T.txt
compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "golden" > gpurun_out/sanitizer_racecheck_$T.txt 2>&1; tail -4 gpurun_out/sanitizer_racecheck_$T.txt
fi
