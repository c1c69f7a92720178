"""Scratch: QPS at several ef for the narrow / wide kernel configuration (SHN_WIDE_FROM_EF), one process per setting."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n, dim, nq, efs = sys.argv[1:5]
for w in ("1000000", "1"):
    env = dict(os.environ, SHN_WIDE_FROM_EF=w)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "var_perf.py"), n, dim, nq, efs, "0"], env=env, capture_output=True, text=True)
    print("wide_from_ef=" + w, r.stdout.strip() or r.stderr[-500:], flush=True)
