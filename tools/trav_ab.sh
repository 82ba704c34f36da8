#!/bin/bash
# A/B over the `traversal` option on the default library: tools/trav_ab.sh "3 4" "collide contacts"
for t in $1; do for w in $2; do
  timeout 300 python bench.py --steps 3 --warmup 3 --workload $w --traversal $t --no-cpu-baseline --no-e2e > gpurun_out/trav_${t}_$w.json 2> gpurun_out/trav_${t}_$w.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/trav_${t}_$w.json")); print("traversal %s %-9s value %.4g q/s  kernel_ms %.3f" % ("$t", "$w", d["value"], d["roofline"]["kernel_ms"]))
except Exception as e:
    print("traversal $t $w FAILED", e); print(open("gpurun_out/trav_${t}_$w.err").read()[-500:])
PY
done; done
