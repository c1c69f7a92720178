set -x
for cfg in "8 0" "8 8" "8 16"; do
  set -- $cfg
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --partitioned only --part-workload sift10m --part-ef 64 --cache-ratio $1 --halo-ratio $2 2> gpurun_out/c12_part2_c$1_h$2.log | grep -v "^NCCL version" > gpurun_out/c12_part2_c$1_h$2.json
  python - <<P
import json
d=json.load(open("gpurun_out/c12_part2_c$1_h$2.json"))["partitioned"]
print({k:d[k] for k in ("value","efficiency_vs_whole_index_replicas","rows_remote","rows_halo","identical_to_whole_index","step_ms_rank0")})
P
done
