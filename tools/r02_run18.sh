#!/bin/bash
# round 2, GPU call 18: L2 discard of the contact staging A/B + contact parity + DRAM traffic
O=gpurun_out/r02_r
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "contacts or exhaustive or binary_mode or pinned or overflow or synthetic or tiny or single_query" > $O/pytest_contacts.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_contacts.log
for v in 1 0 1 0; do
  timeout 300 python bench.py --steps 5 --warmup 3 --workload contacts --no-cpu-baseline --no-e2e --opt contact_discard=$v 2> $O/ab_$v.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('discard=$v value %.4g q/s  kernel_ms %.3f' % (d['value'], d['roofline']['kernel_ms']))
except Exception as e: print('$v FAILED', e)"
done
for v in 1 0; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_write.sum,dram__bytes_read.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct --clock-control none -k regex:"collide_ordered" -c 1 \
    python tools/profile_run.py --workload contacts --poses 1000000 --traversal 3 --launches 1 --opt contact_discard=$v 2>&1 | grep -E "gpu__time|dram__|l1tex|lts__" | sed "s/^/discard=$v /"
done
