"""Stall samples of an .ncu-rep grouped in consecutive blocks of SASS instructions, with the landmark
instructions of each block (helps attributing time to kernel phases without the GUI).
python scripts/ncu_phase.py report.ncu-rep [block]"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 150
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# several kernels in one report: sections start with a "Kernel Name" row; pick the one matching argv[3] (default: last)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if starts:
    want = sys.argv[3] if len(sys.argv) > 3 else None
    pick = starts[-1]
    if want:
        pick = next((i for i in starts if want in rows[i][1]), pick)
    end = next((i for i in starts if i > pick), len(rows))
    print("kernel:", rows[pick][1][:100])
    rows = rows[pick:end]
hi = next(i for i, r in enumerate(rows) if "Source" in r)
h = rows[hi]
c_src = h.index("Source")
c_samp = h.index("# Samples") if "# Samples" in h else h.index("Warp Stall Sampling (All Samples)")
c_exec = next((h.index(n) for n in ("# Instructions Executed", "Instructions Executed") if n in h), None)
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[c_samp] or 0), int(r[c_exec] or 0) if c_exec is not None else 0, r[c_src]))
    except Exception:
        pass
tot = sum(d[0] for d in data) or 1
texec = sum(d[1] for d in data) or 1
LAND = re.compile(r"MUFU\.EX2|MUFU\.LG2|MUFU\.SIN|LDTM|STTM|UTCHMMA|BAR\.SYNC|PHASECHK|UTCBAR|DADD|DFMA|DMUL|IMAD\.WIDE\.U32|SHFL|STG|LDG|ATOMS|STL|LDL")
print(f"total samples {tot}, warp-instructions executed {texec}")
for b in range(0, len(data), blk):
    seg = data[b:b + blk]
    s = sum(d[0] for d in seg)
    e = sum(d[1] for d in seg)
    marks = {}
    for d in seg:
        for m in LAND.findall(d[2]):
            marks[m] = marks.get(m, 0) + 1
    ms = " ".join(f"{k}x{v}" for k, v in sorted(marks.items(), key=lambda kv: -kv[1])[:6])
    print(f"  [{b:5d}..{b+len(seg):5d})  samples {100*s/tot:5.1f}%  exec {100*e/texec:5.1f}%   {ms}")
