nvidia-smi -L | wc -l
for n in 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $n --workload turntable2m --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --workload strips4k --gather peer --steps 20 --warmup 3 2>&1 | tail -1 | cut -c1-260
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 4 --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
