# session 3, call B: many-mesh test, host time per frame of the 10 000-mesh scene after the second host pass
python -m pytest tests/test_gpu_abi.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/s3b_gputests.txt
python tools/host_overhead.py cloud 1500 2>&1 | tee gpurun_out/s3b_host_cloud.txt
MINIRENDER_B200_HOST_THREADS=1 python tools/host_overhead.py cloud 300 2>&1 | tee gpurun_out/s3b_host_cloud_1thread.txt
MINIRENDER_B200_HOST_THREADS=4 python tools/host_overhead.py cloud 1000 2>&1 | tee gpurun_out/s3b_host_cloud_4threads.txt
MINIRENDER_B200_HOST_THREADS=16 python tools/host_overhead.py cloud 1000 2>&1 | tee gpurun_out/s3b_host_cloud_16threads.txt
