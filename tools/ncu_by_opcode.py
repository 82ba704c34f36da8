"""Aggregate the SASS source page of an .ncu-rep by opcode: stall samples, executed warp instructions.
    python tools/ncu_by_opcode.py rep.ncu-rep [top]"""
import csv, io, subprocess, sys, collections
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
si, ii, ai, sp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Address Space"), hdr.index("# Samples")
agg = collections.defaultdict(lambda: [0, 0])
ts = ti = 0
for r in rows[2:]:
    if len(r) <= max(si, ii, sp):
        continue
    toks = r[si].split()
    if not toks:
        continue
    op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
    op = op.split(".")[0] + ("." + r[ai] if r[ai] not in ("-", "") else "")
    s, i = int(r[sp] or 0), int(r[ii] or 0)
    agg[op][0] += s
    agg[op][1] += i
    ts += s
    ti += i
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
print("%-22s %9s %9s" % ("opcode", "samples%", "instr%"))
for op, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s %9.1f %9.1f" % (op, 100.0 * s / max(ts, 1), 100.0 * i / max(ti, 1)))
