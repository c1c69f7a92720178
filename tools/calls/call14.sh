set -x
for h in 8 16; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 --partitioned only --part-workload sift10m --part-ef 64 --halo-ratio $h 2> gpurun_out/c14_part8_h$h.log | grep -v "^NCCL version" > gpurun_out/c14_part8_h$h.json
grep -i "error\|Traceback" gpurun_out/c14_part8_h$h.log | head -5
python - <<P
import json
d=json.load(open("gpurun_out/c14_part8_h$h.json"))["partitioned"]
print({k:d[k] for k in ("value","efficiency_vs_whole_index_replicas","rows_hot","rows_local","rows_remote","rows_halo","identical_to_whole_index","step_ms_rank0","block_s")})
P
done
