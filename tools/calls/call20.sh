set -x
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/c20_tests.log
cat gpurun_out/c20_tests.log
timeout 1200 python tools/var_perf2.py 10000000 128 1000000 16,64,128,256 b200,p3b4,p4b4,p3b5,p2b6 > gpurun_out/c20_ab.log 2>&1
cat gpurun_out/c20_ab.log
