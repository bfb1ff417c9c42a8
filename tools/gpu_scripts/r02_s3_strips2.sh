# session 3: strips with the second calibration stage (stores going to rank 0), N GPUs
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload strips4k --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/s3_strips${N}_remote_cal.json
python - <<PY
import json
d=json.load(open('gpurun_out/s3_strips${N}_remote_cal.json'))
d=d.get('strips4k', d)
for k in ('ms_per_step','speedup_vs_single_gpu_frame','single_gpu_frame_ms','assembled_frame_identical_to_single_gpu','strip_device_ms_per_rank','strip_rows_per_rank','nvlink_bytes_per_frame'):
    print(k, d.get(k))
for c in d.get('strip_balancing', []): print(c)
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --workload strips4k --local-calibration --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/s3_strips${N}_local_cal.json
python - <<PY
import json
d=json.load(open('gpurun_out/s3_strips${N}_local_cal.json'))
d=d.get('strips4k', d)
for k in ('ms_per_step','speedup_vs_single_gpu_frame','assembled_frame_identical_to_single_gpu','strip_device_ms_per_rank','strip_rows_per_rank'):
    print(k, d.get(k))
PY
