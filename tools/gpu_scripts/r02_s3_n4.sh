# session 3: default bench line on 4 GPUs
timeout 42 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 4 --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/s3_bench_n4.json
python - <<PY
import json
d=json.load(open('gpurun_out/s3_bench_n4.json'))
print({k: d.get(k) for k in ('value','ms_per_step','n_gpus')}, json.dumps(d.get('e2e'))[:120])
d=d.get('strips4k', d)
for k in ('ms_per_step','speedup_vs_single_gpu_frame','assembled_frame_identical_to_single_gpu','strip_device_ms_per_rank','strip_rows_per_rank'):
    print(k, d.get(k))
PY
