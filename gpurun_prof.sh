set -x
ncu --set full --clock-control none --import-source on -k regex:'k_raster|k_setup' -s 4 -c 2 -o gpurun_out/prof_r1_v1 python tools_prof_run.py sphere 4 > gpurun_out/prof.log 2>&1
tail -5 gpurun_out/prof.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v1.csv python tools_prof_run.py sphere 3 > /dev/null 2>&1
tail -12 gpurun_out/launches_v1.csv
