set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/route_debug.py 2000000 200000 visits > gpurun_out/c7_dbg_vis.log 2>&1
grep -v "^\*\|OMP\|^$" gpurun_out/c7_dbg_vis.log | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --partitioned only --part-workload sift10m --part-ef 64 > gpurun_out/c7_part2.json 2> gpurun_out/c7_part2.log
grep -v "^\*\|OMP\|^$" gpurun_out/c7_part2.log | tail -12
cat gpurun_out/c7_part2.json
