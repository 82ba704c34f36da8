#!/bin/bash
# round 2, evidence run at HEAD (one GPU): default bench line, reference arm, ncu launch list of the default bench command, one
# `ncu --set full` capture per dominant kernel at the bench's batch size, GPU suite, compute-sanitizer over every kernel family.
set -u
O=gpurun_out/r02_final
mkdir -p $O
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-big > $O/launches_bench.log 2>&1; echo "launch list rc=$?"
for w in distance collide contacts; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"distance_warp_kernel|collide_pooled_kernel|collide_ordered_kernel" \
      -c 1 -f -o $O/full_$w python tools/profile_run.py --workload $w --poses 1000000 --traversal 3 --launches 1 > $O/full_$w.log 2>&1
  python tools/ncu_summary.py $O/full_$w.ncu-rep > $O/full_$w.summary.txt 2>&1
  python tools/ncu_by_function.py $O/full_$w.ncu-rep >> $O/full_$w.summary.txt 2>&1
  grep -E "::|gpu__time_duration|dram__bytes" $O/full_$w.summary.txt | head -4
done
# the counts-only front kernel on cfg4 and the big-model distance instantiation on cfg5 (reduced batches, like the default bench line)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"collide_front_kernel" -s 30 -c 1 -f -o $O/full_cfg4 \
    python bench.py --workload cfg4 --poses 250000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/full_cfg4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"distance_warp_kernel" -s 2 -c 1 -f -o $O/full_cfg5_distance \
    python bench.py --workload cfg5 --poses 100000 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/full_cfg5_distance.log 2>&1
for w in cfg4 cfg5_distance; do
  python tools/ncu_summary.py $O/full_$w.ncu-rep > $O/full_$w.summary.txt 2>&1
  python tools/ncu_by_function.py $O/full_$w.ncu-rep >> $O/full_$w.summary.txt 2>&1
  grep -E "::|gpu__time_duration|dram__bytes" $O/full_$w.summary.txt | head -4
done
rm -f $O/full_collide.ncu-rep $O/full_cfg4.ncu-rep $O/full_cfg5_distance.ncu-rep  # gpurun_out is limited to 64 MiB; the summaries stay
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_memcheck.log
