#!/bin/bash
# round 2, GPU call 41 (2 GPUs): where the extra per-step time of the short cfg5 / cfg1 steps at N > 1 comes from
O=gpurun_out/r02_aq
mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) bench.py --gpus 2 "${@}"; }
show() { python -c "
import json,sys
d=json.load(open('$1'))
print('%-22s' % '$2', {k: round(v['ms_per_step'],3) for k,v in (d.get('workloads') or {}).items() if v} or round(d['ms_per_step'],3))"; }
run --workload cfg5 --poses 100000 --steps 5 --no-cpu-baseline --no-e2e > $O/cfg5_events.json 2> $O/cfg5_events.err; show $O/cfg5_events.json "cfg5 per-step events"
FCLGPU_BENCH_TIMING=loop run --workload cfg5 --poses 100000 --steps 5 --no-cpu-baseline --no-e2e > $O/cfg5_loop.json 2> $O/cfg5_loop.err; show $O/cfg5_loop.json "cfg5 loop"
FCLGPU_BENCH_TIMING=loop FCLGPU_BENCH_NO_GATHER=1 run --workload cfg5 --poses 100000 --steps 5 --no-cpu-baseline --no-e2e > $O/cfg5_loop_nogather.json 2> $O/cfg5_loop_nogather.err; show $O/cfg5_loop_nogather.json "cfg5 loop, no gather"
run --workload all --steps 5 --no-cpu-baseline > $O/all.json 2> $O/all.err; show $O/all.json "all per-step events"
python -c "
import json
d=json.load(open('$O/all.json')); print('all: value %.4g e2e %.4g ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
