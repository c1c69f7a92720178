set -x
SHN_SKIP_C1=1 timeout 1500 python -m pytest tests/test_search_parity.py tests/test_full_size.py tests/test_partition.py tests/test_router.py tests/test_build.py -m gpu -x -q 2>&1 | tail -4
timeout 900 python tools/var_perf.py 10000000 128 1000000 16,32,64,128,256 0 > gpurun_out/c22_perf.log 2>&1; cat gpurun_out/c22_perf.log
timeout 900 python bench.py --cpu-seconds 1 --steps 3 2> gpurun_out/c22_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['roofline']['frac']); [print(s) for s in d['sweep']]"
