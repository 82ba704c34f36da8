#!/bin/bash
# copies the outputs of tools/r02_final.sh (gpurun_out/r02_final) into profiles/ and refreshes the derived files
O=gpurun_out/r02_final
cp $O/bench_default.json profiles/r02_bench_default.json
cp $O/bench_reference.json profiles/r02_bench_reference.json
cp $O/launches_bench.csv profiles/r02_launches_bench.csv
cp $O/full_distance.summary.txt profiles/r02_ncu_distance.txt
cp $O/full_contacts.summary.txt profiles/r02_ncu_contacts.txt
cp $O/full_collide.summary.txt profiles/r02_ncu_collide.txt
cp $O/full_cfg4.summary.txt profiles/r02_ncu_cfg4_front.txt
cp $O/full_cfg5_distance.summary.txt profiles/r02_ncu_cfg5_distance.txt
cp $O/pytest_gpu.log profiles/r02_pytest_gpu.log
cp $O/sanitizer_memcheck.log profiles/r02_sanitizer_memcheck.log
python - <<'PY'
import json, re, csv, collections
def dram(path):
    t = open(path).read()
    r = float(re.search(r"dram__bytes_read.sum = ([0-9.]+)", t).group(1)); w = float(re.search(r"dram__bytes_write.sum = ([0-9.]+)", t).group(1))
    # ncu_summary prints Mbyte for reads; writes are Mbyte or Gbyte (values < 10 with a kernel that writes GBs are Gbyte)
    return r, w
p = 'profiles/r02_traffic.json'; t = json.load(open(p))
r, w = dram('profiles/r02_ncu_distance.txt'); t["distance"]["dram_bytes_per_launch"] = int((r + w) * 1e6)
r, w = dram('profiles/r02_ncu_collide.txt'); t["collide"]["dram_bytes_per_launch"] = int((r + w) * 1e6)
r, w = dram('profiles/r02_ncu_contacts.txt'); t["contacts"]["dram_bytes_per_launch"] = int(r * 1e6 + (w * 1e9 if w < 50 else w * 1e6))
json.dump(t, open(p, 'w'), indent=1)
rows = list(csv.reader(open('profiles/r02_launches_bench.csv', errors='ignore')))
hdr = None; agg = collections.Counter(); cnt = collections.Counter()
for r_ in rows:
    if 'Kernel Name' in r_: hdr = r_; continue
    if hdr and len(r_) == len(hdr):
        d = dict(zip(hdr, r_))
        try: v = float(d['Metric Value'].replace(',', ''))
        except ValueError: continue
        unit = d.get('Metric Unit', '')
        ms = v / 1e6 if unit in ('ns', 'nsecond') else (v / 1e3 if unit in ('us', 'usecond') else (v if unit in ('ms', 'msecond') else v * 1e3))
        k = d['Kernel Name'].split('(')[0][:70]
        agg[k] += ms; cnt[k] += 1
tot = sum(agg.values())
with open('profiles/r02_launches_bench.txt', 'w') as f:
    f.write("ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-big` (gpu__time_duration.sum, --clock-control none;\n"
            "profiles/r02_launches_bench.csv), aggregated per kernel.  Every launch of the run is listed: warm-up and timed steps, the\n"
            "kernel-only timing passes, the end-to-end passes (the host API runs distance in 131072-pose chunks, so its launches are 1/8 of a\n"
            "step each), ONE traversal-0 launch per workload that measures the reference traversal's n_bv / n_leaf for the roofline\n"
            "(distance_thread_kernel<1>, collide_thread_kernel<1>: not timed, not part of a step) and the three micro-benchmarks.\n"
            "Only the repo's own kernels run, plus torch fill kernels (the L2 flush between steps and result-buffer initialisation).\n\n")
    for k, v in agg.most_common(20): f.write("%-72s launches %4d  total %9.3f ms  share %.3f\n" % (k, cnt[k], v, v / tot))
d = json.load(open('profiles/r02_bench_default.json'))
print("headline value %.4g e2e %.4g frac %.3f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"]))
for k, w_ in d["workloads"].items():
    rr = w_.get("roofline") or {}
    print("%-9s value %.4g e2e %s ms %.3f bound %s frac %s" % (k, w_["value"], (w_.get("e2e") or {}).get("value"), w_["ms_per_step"], rr.get("bound"), rr.get("frac")))
PY
