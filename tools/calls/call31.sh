set -x
python -c "import __graft_entry__ as g; g.smoke()"
SHN_SKIP_C1=1 timeout 600 python -m pytest tests/test_search_parity.py tests/test_abi.py tests/test_host_binary.py -m gpu -x -q 2>&1 | tail -3
SHN_SEARCH_CHUNKS=4 python - <<P
import sys, numpy as np
sys.path[:0]=["tests","oracle","."]
import __graft_entry__ as g, datagen
pkg=g.load_package()
base,q=datagen.base_and_queries(5000,5,16)
with pkg.Index.build(base,8,40) as ix:
    a=ix.search(q,3,16)[0]
    import os; os.environ["SHN_SEARCH_CHUNKS"]="1"
    b=ix.search(q,3,16)[0]
print("tiny batch, forced 4 chunks == 1 chunk:", bool((a==b).all()))
P
