set -x
SHN_SKIP_C1=1 timeout 1500 python -m pytest tests/test_search_parity.py tests/test_partition.py tests/test_router.py tests/test_build.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/c16_tests.log
SHN_SKIP_C1=1 true
cat gpurun_out/c16_tests.log
timeout 1200 python tools/var_perf2.py 10000000 128 1000000 16,64,128,256 nola,b200 > gpurun_out/c16_ab.log 2>&1
cat gpurun_out/c16_ab.log
