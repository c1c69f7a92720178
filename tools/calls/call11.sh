set -x
timeout 900 python -m pytest tests/test_router.py tests/test_partition.py tests/test_host_binary.py -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/part_overhead.py 10000000 1000000 64 > gpurun_out/c11_overhead.log 2>&1
cat gpurun_out/c11_overhead.log | cut -c1-300
for cfg in "8 0" "8 8" "8 16"; do
  set -- $cfg
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --partitioned only --part-workload sift10m --part-ef 64 --cache-ratio $1 --halo-ratio $2 2> gpurun_out/c11_part2_c$1_h$2.log | grep -v NCCL > gpurun_out/c11_part2_c$1_h$2.json
  grep -i "error\|Traceback" gpurun_out/c11_part2_c$1_h$2.log | head -5
  python - <<P
import json
d=json.load(open("gpurun_out/c11_part2_c$1_h$2.json"))["partitioned"]
print({k:d[k] for k in ("value","efficiency_vs_whole_index_replicas","rows_remote","rows_halo","identical_to_whole_index","step_ms_rank0")})
P
done
