set -x
date
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/c25_tests.log; cat gpurun_out/c25_tests.log
date
timeout 900 python bench.py > gpurun_out/r2_final_bench_sift10m.json 2> gpurun_out/r2_final_bench_sift10m.log; echo "rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/r2_final_bench_sift10m_reference.json 2> gpurun_out/r2_final_bench_sift10m_reference.log; echo "rc=$?"
timeout 900 python bench.py --workload t2i10m --ef 250 --zipf 1.0 --cpu-seconds 5 > gpurun_out/r2_final_bench_t2i10m_ef250_zipf1.json 2> gpurun_out/r2_final_bench_t2i10m.log; echo "rc=$?"
date
python - <<P
import json
for f in ("r2_final_bench_sift10m","r2_final_bench_sift10m_reference","r2_final_bench_t2i10m_ef250_zipf1"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"], d["e2e"]["value"], d.get("roofline",{}).get("frac"), d.get("parity_at_scale"), d["config"]["ef"], d["config"]["recall_at_10"], d.get("clocks"))
P
