set -x
timeout 1200 python tools/var_perf2.py 10000000 128 1000000 16,64,128,256 o45,a45,a46,a84,b200 > gpurun_out/c5_ab.log 2>&1
