"""Top source lines (CUDA-C view) of an .ncu-rep by stall samples.
    python tools/ncu_by_line.py rep.ncu-rep [file-substring] [top]"""
import csv, io, subprocess, sys
path = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur, hdr, rows, tot = None, None, [], 0.0
for row in csv.reader(io.StringIO(txt)):
    if not row:
        continue
    if row[0] == "File Path":
        cur = row[1]; continue
    if row[0] == "Line No":
        hdr = row; continue
    if hdr is None:
        continue
    try:
        line = int(row[0])
    except ValueError:
        continue
    d = dict(zip(hdr, row))
    def num(k):
        try: return float(d.get(k) or 0)
        except ValueError: return 0.0
    s, i, t = num("# Samples"), num("Instructions Executed"), num("Thread Instructions Executed")
    tot += s
    if filt in cur:
        rows.append((s, i, t, cur.split("/")[-1], line, d.get("Source", "")[:90]))
rows.sort(reverse=True)
for s, i, t, f, line, src in rows[:top]:
    print("%5.2f%% lanes %4.1f  %s:%d  %s" % (100 * s / max(tot, 1), t / max(i, 1), f, line, src.strip()))
