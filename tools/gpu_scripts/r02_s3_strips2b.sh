N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload strips4k --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/s3_strips${N}_b.json
python - <<PY
import json
d=json.load(open('gpurun_out/s3_strips${N}_b.json'))
d=d.get('strips4k', d)
for k in ('host_enqueue_ms_per_frame_per_rank','ms_per_step','ms_per_step_device_rank0','speedup_vs_single_gpu_frame','single_gpu_frame_ms','assembled_frame_identical_to_single_gpu','strip_device_ms_per_rank','strip_rows_per_rank'):
    print(k, d.get(k))
PY
