ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_run.py sphere 6 > /dev/null 2>&1
tail -8 gpurun_out/launches.csv | awk -F'","' '{print $5, $(NF)}'
