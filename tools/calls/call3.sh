set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_router.py tests/test_partition.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/c3_tests.log
SHN_TEST_MULTI_GPU=1 timeout 600 python -m pytest tests/test_host_binary.py -m gpu -x -q -k two_gpus 2>&1 | tail -25 > gpurun_out/c3_hostbin.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --partitioned only --part-workload sift10m --part-ef 64 > gpurun_out/c3_part2.json 2> gpurun_out/c3_part2.log
tail -5 gpurun_out/c3_part2.log
