set -x
timeout 300 python tools/one_search.py 10000000 1000000 64 > gpurun_out/c2_plain.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:search_kernel -c 1 -o gpurun_out/r2_ef64_base python tools/one_search.py 10000000 1000000 64 > gpurun_out/c2_ncu.log 2>&1
