#!/bin/bash
set -u
O=gpurun_out/r01_d
mkdir -p $O
timeout 300 python -m pytest tests/test_zz_gpu_mesh_sphere_distance.py -x -q > $O/pytest_sphere_distance.log 2>&1; echo "sphere tests rc=$?" | tee -a $O/rc.txt
timeout 100 python tools/mesh_sphere_timing.py > $O/mesh_sphere_timing.log 2>&1; echo "timing rc=$?" | tee -a $O/rc.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/rc.txt
timeout 200 python bench.py --workload sphere_distance > $O/bench_sphere_distance.json 2> $O/bench_sphere_distance.err; echo "bench sphere rc=$?" | tee -a $O/rc.txt
tail -n 4 $O/pytest_sphere_distance.log $O/smoke.log
grep distance $O/mesh_sphere_timing.log
cat $O/bench_sphere_distance.json; tail -n 5 $O/bench_sphere_distance.err
