set -x
date
timeout 900 python bench.py > gpurun_out/r2_bench_sift10m.json 2> gpurun_out/r2_bench_sift10m.log; echo "rc=$?"
date
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_sift10m_reference.json 2> gpurun_out/r2_bench_sift10m_reference.log; echo "rc=$?"
date
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 --cpu-seconds 1 > gpurun_out/r2_ncu_launches.log 2>&1; echo "rc=$?"
date
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:search_kernel -c 1 -f -o gpurun_out/r2_search_ef64 python tools/one_search.py 10000000 1000000 64 > gpurun_out/r2_ncu_ef64.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:search_kernel -c 1 -f -o gpurun_out/r2_search_ef256 python tools/one_search.py 10000000 200000 256 > gpurun_out/r2_ncu_ef256.log 2>&1; echo "rc=$?"
date
timeout 600 python bench.py --workload sift1m --cpu-seconds 5 > gpurun_out/r2_bench_sift1m.json 2> gpurun_out/r2_bench_sift1m.log; echo "rc=$?"
timeout 600 python bench.py --workload gist1m --cpu-seconds 5 > gpurun_out/r2_bench_gist1m.json 2> gpurun_out/r2_bench_gist1m.log; echo "rc=$?"
timeout 900 python bench.py --workload t2i10m --ef 250 --zipf 1.0 --cpu-seconds 5 > gpurun_out/r2_bench_t2i10m_ef250_zipf1.json 2> gpurun_out/r2_bench_t2i10m.log; echo "rc=$?"
date
timeout 1200 python -m pytest tests/test_search_parity.py -m gpu -x -q -s -k c1 2>&1 | tail -8 > gpurun_out/r2_c1.log; cat gpurun_out/r2_c1.log
date
