#!/bin/bash
# round 2, GPU call 3: A/B of the ordered-front contact kernel (rolled / inline / occupancy variants), parity of the default first
mkdir -p gpurun_out/r02_c
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "contacts or exhaustive or edge or host_api or both_objects" > gpurun_out/r02_c/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_c/pytest.log
export TRAV=3
bash tools/ab_bench.sh "default def rin rall rall3 unr rin5 rsat" "contacts"
