# session 3: 2 GPUs, final code: the NCCL / peer-store strip test and the default bench line for N = 2
nvidia-smi -L | head -2
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/s3_multi2_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/s3_bench_n2.json
cut -c1-900 gpurun_out/s3_bench_n2.json
