"""Top SASS lines of one kernel of an .ncu-rep by a given stall column (e.g. stall_long_sb).
python scripts/ncu_stall.py report.ncu-rep stall_long_sb [kernel substring] [n]"""
import csv, io, subprocess, sys
rep, colname = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else None
n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
pick = starts[-1]
if want:
    pick = next((i for i in starts if want in rows[i][1]), pick)
end = next((i for i in starts if i > pick), len(rows))
rows = rows[pick:end]
hi = next(i for i, r in enumerate(rows) if "Source" in r)
h = rows[hi]
c = h.index(colname); cs = h.index("Source"); ca = h.index("Address"); ce = h.index("Instructions Executed"); ct = h.index("# Samples")
data = []
for idx, r in enumerate(rows[hi + 1:]):
    try: data.append((int(r[c] or 0), idx, r[ca][-5:], int(r[ce] or 0), int(r[ct] or 0), r[cs]))
    except Exception: pass
tot = sum(d[0] for d in data) or 1
alls = sum(d[4] for d in data) or 1
print(f"{colname}: {tot} samples of {alls} total")
for d in sorted(data, reverse=True)[:n]:
    print(f"  {100*d[0]/tot:5.1f}%  line {d[1]:5d}  {d[2]}  exec {d[3]:8d}  {d[5][:90]}")
