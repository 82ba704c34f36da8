#!/bin/bash
# Round-end evidence at HEAD: whole GPU suite, mesh <-> sphere timing, ncu --set full of the default sphere-distance
# kernel, memcheck over the sphere-distance tests, the three bench lines.
set -u
O=gpurun_out/r01_c
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $O/rc.txt
timeout 100 python tools/mesh_sphere_timing.py > $O/mesh_sphere_timing.log 2>&1; echo "timing rc=$?" | tee -a $O/rc.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"distance_mesh_sphere" -c 1 -f -o $O/full_sphere_distance \
    python tools/profile_run.py --workload sphere_distance --poses 1000000 --traversal 3 --launches 1 > $O/full_sphere_distance.log 2>&1; echo "ncu rc=$?" | tee -a $O/rc.txt
python tools/ncu_summary.py $O/full_sphere_distance.ncu-rep > $O/full_sphere_distance.summary.txt 2>&1
python tools/ncu_by_function.py $O/full_sphere_distance.ncu-rep >> $O/full_sphere_distance.summary.txt 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_zz_gpu_mesh_sphere_distance.py -x -q -k "tiny or known" > $O/sanitizer_sphere_distance.log 2>&1; echo "memcheck rc=$?" | tee -a $O/rc.txt
timeout 200 python bench.py > $O/bench_distance.json 2> $O/bench_distance.err; echo "bench distance rc=$?" | tee -a $O/rc.txt
timeout 200 python bench.py --workload collide > $O/bench_collide.json 2> $O/bench_collide.err; echo "bench collide rc=$?" | tee -a $O/rc.txt
timeout 200 python bench.py --workload contacts > $O/bench_contacts.json 2> $O/bench_contacts.err; echo "bench contacts rc=$?" | tee -a $O/rc.txt
tail -n 3 $O/pytest_gpu.log $O/sanitizer_sphere_distance.log
grep distance $O/mesh_sphere_timing.log
head -n 45 $O/full_sphere_distance.summary.txt
cut -c1-330 $O/bench_*.json
