#!/bin/bash
# round 2, GPU call 38: randomised differential run and racecheck at HEAD (front kernel with dynamic stack, ordered kernel with
# pre-expanded entries)
O=gpurun_out/r02_ao
mkdir -p $O
timeout 400 python tests/stress/stress_parity.py 150 20261018 > $O/stress_parity.log 2>&1; echo "stress rc=$?"; tail -2 $O/stress_parity.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 $O/sanitizer_racecheck.log
