set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --cpu-seconds 3 > gpurun_out/c30_bench.json 2> gpurun_out/c30_bench.log; echo "rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/c30_bench.json")); print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity_at_scale"], d["clocks"])
for s in d["sweep"]: print(s)
P
