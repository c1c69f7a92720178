set -x
SHN_SKIP_C1=1 timeout 900 python -m pytest tests/test_search_parity.py tests/test_full_size.py -m gpu -x -q 2>&1 | tail -4
SHN_WIDE_FROM_EF=1 SHN_SKIP_C1=1 timeout 900 python -m pytest tests/test_search_parity.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python tools/wide_probe.py 10000000 128 1000000 72,80,96,100,112,128,160,200 > gpurun_out/c21_wide128.log 2>&1; cat gpurun_out/c21_wide128.log
timeout 900 python tools/wide_probe.py 20000000 96 1000000 64,100,128 > gpurun_out/c21_wide96.log 2>&1; cat gpurun_out/c21_wide96.log
