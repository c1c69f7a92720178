set -x
timeout 900 python -m pytest tests/test_router.py tests/test_partition.py tests/test_host_binary.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/c8_tests.log
cat gpurun_out/c8_tests.log
for cfg in "8 8" "8 16" "0 8" "8 30"; do
  set -- $cfg
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --partitioned only --part-workload sift10m --part-ef 64 --cache-ratio $1 --halo-ratio $2 > gpurun_out/c8_part2_c$1_h$2.json 2> gpurun_out/c8_part2_c$1_h$2.log
  grep -v "^\*\|OMP\|^$" gpurun_out/c8_part2_c$1_h$2.log | tail -4
  cat gpurun_out/c8_part2_c$1_h$2.json
done
