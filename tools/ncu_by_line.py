"""Per-source-line instruction / stall-sample shares of one kernel: joins `nvdisasm -g` line info of the built cubin
with the SASS page of an ncu report (same build).  usage: ncu_by_line.py <report.ncu-rep> <mangled-kernel-substring>"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, sub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "multi-robot-fabrics_b200", "libmrf_b200.so")], cwd=d,
                       capture_output=True)
        cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=d, capture_output=True, text=True).stdout.split("\n")
    ins, cur, on = [], None, False
    for l in dis:
        if l.startswith("\t.section\t.text."):
            on = sub in l
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            ins.append((m.group(2), cur))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    if len(data) != len(ins):
        sys.exit(f"build differs from the report: {len(ins)} SASS instructions here, {len(data)} in the report")
    n_by, s_by = collections.Counter(), collections.Counter()
    src = {}
    for (sass, loc), r in zip(ins, data):
        n_by[loc] += int(r[ix["Instructions Executed"]] or 0)
        s_by[loc] += int(r[ix["# Samples"]] or 0)
    tn, ts = sum(n_by.values()), sum(s_by.values())
    files = {}
    print(f"total warp instructions {tn}, samples {ts}")
    print("| file:line | inst % | samples % | source |\n|---|---|---|---|")
    for loc, n in n_by.most_common(top):
        f, ln = loc
        if f not in files:
            p = os.path.join(ROOT, "multi-robot-fabrics_b200", "csrc", f)
            files[f] = open(p).read().split("\n") if os.path.exists(p) else []
        text = files[f][ln - 1].strip()[:110] if ln - 1 < len(files[f]) else ""
        print(f"| {f}:{ln} | {100 * n / tn:.2f} | {100 * s_by[loc] / ts:.2f} | `{text}` |")


if __name__ == "__main__":
    main()
