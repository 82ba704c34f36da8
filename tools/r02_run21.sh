#!/bin/bash
# round 2, GPU call 21: compact contact records -- parity + end-to-end rates
O=gpurun_out/r02_w
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fcl_shim.py tests/test_c_abi.py -m gpu -x -q -k "compact or contacts or shim or abi or c_client" > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.log
timeout 600 python bench.py --workload contacts --steps 6 --no-cpu-baseline > $O/bench_contacts.json 2> $O/bench_contacts.err; echo "bench rc=$?"; tail -2 $O/bench_contacts.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_w/bench_contacts.json"))
e=d["e2e"]
print("kernel value %.4g  e2e full %.4g  ids %.4g  f32 %.4g" % (d["value"], e["value"], e["compact_ids"]["value"], e["compact_f32"]["value"]))
PY
