ncu --set full --clock-control none --import-source on -k regex:'k_raster|k_setup' -s 4 -c 2 -o gpurun_out/prof_r1_v6 python tools/prof_run.py sphere 4 > gpurun_out/prof.log 2>&1
tail -3 gpurun_out/prof.log
