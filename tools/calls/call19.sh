set -x
SHN_SKIP_C1=1 timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/c19_tests.log
cat gpurun_out/c19_tests.log
timeout 600 python tools/bf_perf.py 20000000 10000 96 > gpurun_out/c19_bf_d96.log 2>&1; cat gpurun_out/c19_bf_d96.log
timeout 600 python tools/bf_perf.py 10000000 10000 128 > gpurun_out/c19_bf_d128.log 2>&1; cat gpurun_out/c19_bf_d128.log
timeout 600 python tools/bf_perf.py 5000000 10000 200 > gpurun_out/c19_bf_d200.log 2>&1; cat gpurun_out/c19_bf_d200.log
SHN_SEARCH_CHUNKS=1 timeout 900 python bench.py --cpu-seconds 1 > gpurun_out/c19_bench_chunks1.json 2> gpurun_out/c19_bench_chunks1.log; echo "rc=$?"
timeout 900 python bench.py --cpu-seconds 1 > gpurun_out/c19_bench_chunks4.json 2> gpurun_out/c19_bench_chunks4.log; echo "rc=$?"
python - <<P
import json
for f in ("c19_bench_chunks1","c19_bench_chunks4"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"])
P
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > gpurun_out/r2_ncu_launches.log 2>&1; echo "rc=$?"
grep -c "search_kernel" gpurun_out/r2_bench_launches.csv
