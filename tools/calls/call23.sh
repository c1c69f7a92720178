set -x
timeout 600 python tools/var_perf.py 10000000 128 1000000 16,64,100,128,256 0 > gpurun_out/c23_perf.log 2>&1; cat gpurun_out/c23_perf.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_search_parity.py -m gpu -x -q -k "golden" > gpurun_out/c23_memcheck_search.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c23_memcheck_search.log
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_router.py -m gpu -x -q -k "halo and 2-10" > gpurun_out/c23_memcheck_router.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c23_memcheck_router.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_bruteforce.py -m gpu -x -q -k "certificate or (tensor_core_path_is_exact and 1000-5)" > gpurun_out/c23_memcheck_bf.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/c23_memcheck_bf.log
