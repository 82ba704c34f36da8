#!/bin/bash
# Last GPU call of round 1 (short budget): new mesh <-> sphere distance parity first, then its timing, a memcheck
# pass over the new kernel, the default bench line at HEAD, and as much of the whole GPU suite as the time allows.
set -u
O=gpurun_out/r01_last
mkdir -p $O
timeout 300 python -m pytest tests/test_zz_gpu_mesh_sphere_distance.py -x -q > $O/pytest_sphere_distance.log 2>&1; echo "sphere tests rc=$?" | tee -a $O/rc.txt
timeout 200 python tools/mesh_sphere_timing.py > $O/mesh_sphere_timing.log 2>&1; echo "timing rc=$?" | tee -a $O/rc.txt
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_mesh_sphere_distance.py -x -q -k "tiny or known" > $O/sanitizer_sphere_distance.log 2>&1; echo "memcheck rc=$?" | tee -a $O/rc.txt
timeout 300 python bench.py > $O/bench_distance.json 2> $O/bench_distance.err; echo "bench rc=$?" | tee -a $O/rc.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $O/rc.txt
tail -n 3 $O/pytest_sphere_distance.log $O/mesh_sphere_timing.log $O/sanitizer_sphere_distance.log $O/bench_distance.json $O/pytest_gpu.log
