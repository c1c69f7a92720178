set -x
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/route_debug.py 2000000 200000 > gpurun_out/c6_dbg_novis.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/route_debug.py 2000000 200000 visits > gpurun_out/c6_dbg_vis.log 2>&1
grep -v "^\*\|OMP\|^$" gpurun_out/c6_dbg_novis.log | tail -20
grep -v "^\*\|OMP\|^$" gpurun_out/c6_dbg_vis.log | tail -20
