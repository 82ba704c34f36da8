#!/bin/bash
# round 2, GPU call 28: ncu of the counts-only front kernel on cfg4 (and on env/rob at 1M)
O=gpurun_out/r02_ae
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"collide_front_kernel" -s 30 -c 1 -f -o $O/full_cfg4 \
    python bench.py --workload cfg4 --poses 250000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/full_cfg4.log 2>&1
python tools/ncu_summary.py $O/full_cfg4.ncu-rep > $O/full_cfg4.summary.txt 2>&1
python tools/ncu_by_function.py $O/full_cfg4.ncu-rep >> $O/full_cfg4.summary.txt 2>&1
head -48 $O/full_cfg4.summary.txt
