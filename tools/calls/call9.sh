set -x
timeout 600 python tools/part_overhead.py 10000000 1000000 64 > gpurun_out/c9_vmm.log 2>&1
SHN_SHARE_PLAIN=1 timeout 600 python tools/part_overhead.py 10000000 1000000 64 > gpurun_out/c9_plain.log 2>&1
tail -3 gpurun_out/c9_vmm.log gpurun_out/c9_plain.log
