set -x
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --partitioned only --part-workload sift10m --part-ef 64 2> gpurun_out/c13_part8.log | grep -v "^NCCL version" > gpurun_out/c13_part8.json
grep -i "error\|Traceback" gpurun_out/c13_part8.log | head -5
python - <<P
import json
d=json.load(open("gpurun_out/c13_part8.json"))["partitioned"]
print({k:d[k] for k in ("value","efficiency_vs_whole_index_replicas","rows_hot","rows_local","rows_remote","rows_halo","identical_to_whole_index","step_ms_rank0","sent_rank0","received_rank0","part_sizes","block_s")})
P
