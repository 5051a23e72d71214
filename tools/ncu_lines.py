"""Developer tool: per-source-line instruction / stall-sample shares of a kernel in an .ncu-rep (needs -lineinfo).
    python tools/ncu_lines.py report.ncu-rep [top]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, data = None, None, {}
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    iI, iW = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    key = (cur.split("/")[-1], ln)
    old = data.get(key, (0, 0, ""))
    data[key] = (old[0] + int(r[iI]), old[1] + int(r[iW]), r[1].strip()[:100])
tot = sum(v[0] for v in data.values())
tots = sum(v[1] for v in data.values())
print("total warp instructions", tot, "stall samples", tots)
for k, v in sorted(data.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / max(tots, 1) * 100:5.1f}% stall  {k[0]}:{k[1]}: {v[2]}")
