#!/bin/bash
# round 2, GPU call 23: locality order of the batch for the lane-per-query collide kernel -- parity + A/B
O=gpurun_out/r02_z
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_broadphase.py -m gpu -x -q -k "cfg1 or cfg3 or contacts or exhaustive or binary or both_objects or edge or synthetic or full_size or overloads or pinned or tiny or broadphase" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest.log
for v in 16384 0 16384 0; do
  timeout 300 python bench.py --steps 10 --warmup 3 --workload collide --no-cpu-baseline --no-e2e --opt order_queries=$v 2> $O/ab_$v.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('order_queries=$v value %.4g q/s  step_ms %.3f kernel_ms %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms']))
except Exception as e: print('$v FAILED', e)"
done
for v in 16384 0; do
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed --clock-control none -k regex:"collide_pooled|order_" -c 4 \
    python tools/profile_run.py --workload collide --poses 1000000 --traversal 3 --launches 1 --opt order_queries=$v 2>&1 | grep -E "^  [a-z_:A-Z<>]+.*\(|gpu__time|ratio|hit_rate|issue_active|lsu_wave" | sed "s/^/order=$v /"
done
