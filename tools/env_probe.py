"""Scratch: QPS at several ef with an environment variable set to each of the given values (one process per value).
    python tools/env_probe.py SHN_VIS_COMPACT 0,1 10000000 128 1000000 16,64,128,256"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
var, values, n, dim, nq, efs = sys.argv[1:7]
for v in values.split(","):
    env = dict(os.environ, **{var: v})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "var_perf.py"), n, dim, nq, efs, "0"], env=env, capture_output=True, text=True)
    print(f"{var}={v}", r.stdout.strip() or r.stderr[-500:], flush=True)
