set -x
SHN_VIS_COMPACT=1 SHN_SKIP_C1=1 timeout 600 python -m pytest tests/test_search_parity.py tests/test_full_size.py tests/test_partition.py tests/test_router.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python tools/env_probe.py SHN_VIS_COMPACT 0,1 10000000 128 1000000 16,32,64,128,256 > gpurun_out/c29_compact.log 2>&1; cat gpurun_out/c29_compact.log
