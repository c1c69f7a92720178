set -x
timeout 300 python -m pytest tests/test_partition.py tests/test_router.py -m gpu -x -q 2>&1 | tail -3
