"""Scratch: QPS of several library variants (libshn_<v>.so) on ONE n x dim GPU-built index shape; one process per variant."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n, dim, nq, efs = sys.argv[1:5]
for v in sys.argv[5].split(","):
    env = dict(os.environ)
    env["SHN_LIB"] = os.path.join(ROOT, "dm-hnsw-reference_b200", f"libshn_{v}.so")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "var_perf.py"), n, dim, nq, efs, "0"], env=env, capture_output=True, text=True)
    print(r.stdout.strip() or r.stderr[-500:], flush=True)
