# Does k_raster find k_geom's records in L2? DRAM bytes per kernel without cache flushes, (a) one-pass kernel replay
# with two metrics, (b) application replay (no save / restore of device memory between passes).
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:'k_raster|k_geom' -s 20 -c 6 --csv --log-file gpurun_out/l2res_a.csv python tools/prof_run.py sphere 16 > /dev/null 2>&1
ncu --replay-mode application --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_read_lookup_hit.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_red.sum -k regex:'k_raster|k_geom' -s 20 -c 6 --csv --log-file gpurun_out/l2res_b.csv python tools/prof_run.py sphere 16 > /dev/null 2>&1
for f in a b; do echo "== $f"; grep -E "k_raster|k_geom" gpurun_out/l2res_$f.csv | awk -F'","' '{print $1, substr($5,1,30), $(NF-2), $NF}' | tr -d '"'; done
