set -x
SHN_SKIP_C1=1 timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/c17_tests.log
cat gpurun_out/c17_tests.log
timeout 1200 python tools/var_perf2.py 10000000 128 1000000 16,32,64,128,256 b200 > gpurun_out/c17_ab.log 2>&1
cat gpurun_out/c17_ab.log
