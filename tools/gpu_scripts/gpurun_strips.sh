python bench.py --workload strips4k --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-700
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --workload strips4k --gather peer --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --workload strips4k --gather nccl --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-300
