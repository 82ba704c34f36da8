#!/bin/bash
# Evidence run for the mesh <-> sphere distance (GPU box; what profiles/r01_*sphere* came from):
# whole GPU suite, randomised GPU-vs-oracle stress, kernel-variant timing, memcheck over the sphere-distance tests,
# one `ncu --set full` capture of the default kernel at the bench's batch size, the bench line.
# Outputs under gpurun_out/sphere/ ; summaries are copied into profiles/ afterwards.
set -u
O=gpurun_out/sphere
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $O/rc.txt
timeout 100 python tests/stress/stress_parity.py 45 991 > $O/stress.log 2>&1; echo "stress rc=$?" | tee -a $O/rc.txt
timeout 100 python tools/mesh_sphere_timing.py > $O/mesh_sphere_timing.log 2>&1; echo "timing rc=$?" | tee -a $O/rc.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_mesh_sphere_distance.py -x -q -k "tiny or known" > $O/sanitizer_sphere_distance.log 2>&1; echo "memcheck rc=$?" | tee -a $O/rc.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"distance_mesh_sphere" -c 1 -f -o $O/full_sphere_distance \
    python tools/profile_run.py --workload sphere_distance --poses 1000000 --traversal 3 --launches 1 > $O/full_sphere_distance.log 2>&1; echo "ncu rc=$?" | tee -a $O/rc.txt
python tools/ncu_summary.py $O/full_sphere_distance.ncu-rep > $O/full_sphere_distance.summary.txt 2>&1
python tools/ncu_by_function.py $O/full_sphere_distance.ncu-rep >> $O/full_sphere_distance.summary.txt 2>&1
timeout 100 python bench.py --workload sphere_distance > $O/bench_sphere_distance.json 2> $O/bench_sphere_distance.err; echo "bench sphere rc=$?" | tee -a $O/rc.txt
tail -n 3 $O/pytest_gpu.log $O/stress.log $O/sanitizer_sphere_distance.log
grep distance $O/mesh_sphere_timing.log
head -n 48 $O/full_sphere_distance.summary.txt
cut -c1-400 $O/bench_sphere_distance.json
