"""Attribute ncu source-page samples / executed instructions to source functions.
    python tools/ncu_by_function.py gpurun_out/prof.ncu-rep"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def func_table(path):
    """(start_line, name) for every function-looking definition in a source file."""
    out = []
    try:
        src = open(path).read().split("\n")
    except OSError:
        return out
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:FD|__device__|__global__|static|inline|__forceinline__|\s)*[\w:<>\*&\s]+?\b(\w+)\s*\([^;]*$")
    for i, line in enumerate(src, 1):
        if line.startswith((" ", "\t", "//", "#", "}")) or "(" not in line:
            continue
        m = pat.match(line)
        if m and m.group(1) not in ("if", "for", "while", "switch", "return"):
            out.append((i, m.group(1)))
    return out


def main():
    path = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    cur_file, hdr = None, None
    agg = defaultdict(lambda: [0.0, 0.0, 0.0])  # samples, warp instr, thread instr
    tables = {}
    for row in csv.reader(io.StringIO(txt)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1]
            tables.setdefault(cur_file, func_table(cur_file))
            continue
        if row[0] == "Line No":
            hdr = row
            continue
        if hdr is None or row[0] in ("Function Name", "Kernel Name") or row[0] == "":
            continue
        try:
            line = int(row[0])
        except ValueError:
            continue
        d = dict(zip(hdr, row))
        name = "?"
        for start, fn in tables.get(cur_file, []):
            if start <= line:
                name = fn
        key = (cur_file.split("/")[-1], name)
        def num(k):
            try:
                return float(d.get(k) or 0)
            except ValueError:
                return 0.0

        agg[key][0] += num("# Samples")
        agg[key][1] += num("Instructions Executed")
        agg[key][2] += num("Thread Instructions Executed")
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print(f"{'file':22s} {'function':28s} {'samples%':>9s} {'instr%':>8s} {'lanes/instr':>11s}")
    for (f, n), v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if v[0] / ts < 0.003 and v[1] / ti < 0.003:
            continue
        print(f"{f:22s} {n:28s} {100 * v[0] / ts:9.1f} {100 * v[1] / ti:8.1f} {v[2] / max(v[1], 1):11.1f}")


if __name__ == "__main__":
    main()
