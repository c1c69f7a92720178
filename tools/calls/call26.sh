set -x
timeout 900 python tools/var_perf2.py 10000000 128 1000000 16,64,128,256 nopf,b200,bulk > gpurun_out/c26_ab.log 2>&1; cat gpurun_out/c26_ab.log
