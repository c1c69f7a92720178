set -x
date
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/c28_bench2.log | grep -v "^NCCL version" > gpurun_out/c28_bench2.json
date
grep -i "error\|Traceback" gpurun_out/c28_bench2.log | head -5
python - <<P
import json
d=json.load(open("gpurun_out/c28_bench2.json"))
print({k:d[k] for k in ("value","ms_per_step","e2e")})
p=d["partitioned"]
print({k:p[k] for k in p if k not in ("design","workload","sweep_whole_index","clocks")})
P
