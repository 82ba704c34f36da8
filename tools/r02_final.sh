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
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_memcheck.log
