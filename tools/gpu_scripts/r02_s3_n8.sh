# session 3: default bench line on 8 GPUs (view batch + strips4k with the second calibration stage)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/s3_bench_n8.json
python - <<PY
import json
d=json.load(open('gpurun_out/s3_bench_n8.json'))
print({k: d.get(k) for k in ('value','ms_per_step','n_gpus')}, json.dumps(d.get('e2e'))[:200])
d=d.get('strips4k', d)
for k in ("host_enqueue_ms_per_frame_per_rank","ms_per_step","speedup_vs_single_gpu_frame","single_gpu_frame_ms","assembled_frame_identical_to_single_gpu","strip_device_ms_per_rank","strip_geom_ms_per_rank","strip_raster_ms_per_rank",'strip_rows_per_rank','nvlink_bytes_per_frame'):
    print(k, d.get(k))
for c in d.get('strip_balancing', []): print(c)
PY
