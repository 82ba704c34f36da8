#!/bin/bash
# Round-end evidence run (GPU box): bench lines for the three workloads, the ncu launch list of the default
# bench command, and one `ncu --set full` capture per dominant kernel at the bench's batch size (1M poses).
# Outputs under gpurun_out/final/ ; summaries are copied into profiles/ afterwards.
set -u
O=gpurun_out/final
mkdir -p $O
python bench.py > $O/bench_distance.json 2> $O/bench_distance.err
python bench.py --workload collide > $O/bench_collide.json 2> $O/bench_collide.err
python bench.py --workload contacts > $O/bench_contacts.json 2> $O/bench_contacts.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_distance.json 2> $O/bench_reference_distance.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_distance_bench.csv \
    python bench.py --steps 2 --warmup 1 > $O/launches_bench.log 2>&1
for w in distance collide contacts; do
  ncu --set full --clock-control none --import-source on -k regex:"distance_warp_kernel|collide_pooled_kernel|collide_deferred_kernel" \
      -c 1 -f -o $O/full_$w python tools/profile_run.py --workload $w --poses 1000000 --traversal 3 --launches 1 > $O/full_$w.log 2>&1
  python tools/ncu_summary.py $O/full_$w.ncu-rep > $O/full_$w.summary.txt 2>&1
  python tools/ncu_by_function.py $O/full_$w.ncu-rep >> $O/full_$w.summary.txt 2>&1
done
tail -n 3 $O/bench_*.json
