"""dm-hnsw-reference_b200 — B200-native HNSW search engine behind the SHINE compute-node contract.

The product is libshn_b200.so (hand-written sm_100a CUDA + a C ABI, include/shn.h) and the host binary
host/shine_b200; this Python package is only the ctypes binding the tests and bench.py drive it through.
The directory name is not an importable identifier: load it with `load_package()` from __graft_entry__.py or
put this directory on sys.path and `import shn`.
"""
from . import parallel, shn  # noqa: F401
from .shn import (Index, Router, ShnError, bruteforce_topk, bruteforce_topk_device, build_library, device_view, draw_levels, library_path, route_queries,
                  repartition_dumps, set_build_option)  # noqa: F401
