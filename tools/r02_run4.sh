#!/bin/bash
# round 2, GPU call 4: chase the illegal access seen in test_both_objects_posed[front64] (subset run), tolerance tests
mkdir -p gpurun_out/r02_d
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "contacts or exhaustive or edge or host_api or both_objects" > gpurun_out/r02_d/pytest_subset.log 2>&1; echo "subset rc=$?"; tail -2 gpurun_out/r02_d/pytest_subset.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_objects" > gpurun_out/r02_d/sanitizer_both.log 2>&1; echo "sanitizer rc=$?"
grep -E "Invalid|Error|at 0x|by thread|Address|fclgpu|kernel|passed|failed" gpurun_out/r02_d/sanitizer_both.log | head -40
timeout 600 python -m pytest tests/test_zz_gpu_tolerance.py tests/test_fcl_shim.py -m gpu -x -q > gpurun_out/r02_d/pytest_tol.log 2>&1; echo "tolerance rc=$?"; tail -15 gpurun_out/r02_d/pytest_tol.log
