#!/bin/bash
set -u
O=gpurun_out/r01_g
mkdir -p $O
timeout 60 python tools/mesh_sphere_timing.py > $O/mesh_sphere_timing.log 2>&1; echo "timing rc=$?" | tee -a $O/rc.txt
timeout 100 python -m pytest tests/test_zz_gpu_mesh_sphere_distance.py -x -q > $O/pytest_sphere_distance.log 2>&1; echo "sphere tests rc=$?" | tee -a $O/rc.txt
grep distance $O/mesh_sphere_timing.log; tail -n 3 $O/pytest_sphere_distance.log
