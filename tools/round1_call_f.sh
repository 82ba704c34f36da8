#!/bin/bash
# last short check: the sphere bench line through bench.py, the plain-C ABI client and the C++ shim on a device
set -u
O=gpurun_out/r01_f
mkdir -p $O
timeout 100 python bench.py --workload sphere_distance > $O/bench_sphere_distance.json 2> $O/bench_sphere_distance.err; echo "bench sphere rc=$?" | tee -a $O/rc.txt
timeout 100 python -m pytest tests/test_c_abi.py tests/test_fcl_shim.py -x -q > $O/pytest_abi_shim.log 2>&1; echo "abi+shim rc=$?" | tee -a $O/rc.txt
tail -n 3 $O/pytest_abi_shim.log; tail -n 3 $O/bench_sphere_distance.err; cut -c1-700 $O/bench_sphere_distance.json
