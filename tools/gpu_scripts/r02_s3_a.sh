# session 3, call A: GPU tests on the host-pool build, host time per frame of the 10 000-mesh scene
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/s3a_gputests.txt
nproc
python tools/host_overhead.py cloud 1000 2>&1 | tee gpurun_out/s3a_host_cloud.txt
MINIRENDER_B200_HOST_THREADS=1 python tools/host_overhead.py cloud 300 2>&1 | tee gpurun_out/s3a_host_cloud_1thread.txt
python tools/host_overhead.py sphere 2000 2>&1 | tee gpurun_out/s3a_host_sphere.txt
