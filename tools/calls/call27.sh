set -x
python -c "import __graft_entry__ as g; g.smoke()"
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --cpu-seconds 5 > gpurun_out/c27_bench.json 2> gpurun_out/c27_bench.log; echo "rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/c27_bench.json")); print(d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["parity_at_scale"], d["clocks"])
for s in d["sweep"]: print(s)
P
