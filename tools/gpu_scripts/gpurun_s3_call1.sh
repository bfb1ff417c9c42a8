python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_s3a.json; cut -c1-300 gpurun_out/bench_s3a.json
ncu --set full --clock-control none --import-source on -k regex:'k_raster|k_setup|k_vertex' -s 6 -c 3 -o gpurun_out/prof_s3a python tools/prof_run.py sphere 4 > gpurun_out/prof.log 2>&1
tail -1 gpurun_out/prof.log; ls -la gpurun_out
python tools/stage_probe.py sphere 0 | grep flags
python tools/stage_probe.py bench 0 | grep flags
python tools/stage_probe.py cloud 0 | grep flags
