ncu --set full --clock-control none --import-source on -k regex:'k_raster|k_setup|k_vertex' -s 6 -c 3 -o gpurun_out/prof_cur python tools/prof_run.py ${1:-sphere} 4 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
