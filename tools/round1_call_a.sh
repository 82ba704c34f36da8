#!/bin/bash
# whole GPU suite, randomised GPU-vs-oracle stress (new sphere distance included), ncu capture of the new kernel
set -u
O=gpurun_out/r01_a
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "gpu suite rc=$?" | tee -a $O/rc.txt
timeout 200 python tests/stress/stress_parity.py 100 777 > $O/stress.log 2>&1; echo "stress rc=$?" | tee -a $O/rc.txt
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"distance_mesh_sphere_kernel" -c 1 -f -o $O/full_sphere_distance \
    python tools/profile_run.py --workload sphere_distance --poses 1000000 --traversal 3 --launches 1 > $O/full_sphere_distance.log 2>&1; echo "ncu rc=$?" | tee -a $O/rc.txt
python tools/ncu_summary.py $O/full_sphere_distance.ncu-rep > $O/full_sphere_distance.summary.txt 2>&1
python tools/ncu_by_function.py $O/full_sphere_distance.ncu-rep >> $O/full_sphere_distance.summary.txt 2>&1
tail -n 4 $O/pytest_gpu.log $O/stress.log $O/full_sphere_distance.log
head -n 40 $O/full_sphere_distance.summary.txt
