set -x
nvidia-smi -L
SHN_TEST_SEQUENTIAL_BUILD=1 timeout 600 python -m pytest tests/test_build.py -q -m gpu -k sequential 2>&1 | tail -15 > gpurun_out/c1_seq.log
timeout 600 python tools/var_perf.py 10000000 128 500000 64,128,256 0,2048,4096,8192 > gpurun_out/c1_varperf.log 2>&1
timeout 300 python tools/one_search.py 10000000 200000 256 > gpurun_out/c1_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:search_kernel -c 1 -o gpurun_out/r2_ef256_base python tools/one_search.py 10000000 200000 256 > gpurun_out/c1_ncu.log 2>&1
ls -la gpurun_out
