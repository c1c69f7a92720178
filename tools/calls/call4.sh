set -x
timeout 1200 python -m pytest tests/test_search_parity.py tests/test_partition.py tests/test_router.py tests/test_build.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/c4_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/route_debug.py 2000000 200000 > gpurun_out/c4_dbg_novis.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/route_debug.py 2000000 200000 visits > gpurun_out/c4_dbg_vis.log 2>&1
timeout 600 python tools/var_perf.py 10000000 128 1000000 16,32,64,128,256 0 > gpurun_out/c4_varperf.log 2>&1
