"""Top warp-stall sites of one kernel of an ncu report (source page, SASS level, needs -lineinfo + --import-source on).
    python scripts/ncu_source.py <file.ncu-rep> <kernel regex> <launch skip> [top N]"""
import csv
import subprocess
import sys

rep, rx, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:120])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or not r[ix["# Samples"]].isdigit():
        if r and r[0] in ("Kernel Name", "Address", "#"):
            if data:
                break  # a second view (source level) follows the SASS view
        continue
    data.append(r)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
print("samples %d, warp instructions %d, SASS lines %d" % (tot, inst, len(data)))
agg = {h: sum(int(r[ix[h]]) for r in data if r[ix[h]].isdigit()) for h in stall_cols}
print("stall reasons:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
    reasons = sorted(((int(r[ix[h]]), h[6:]) for h in stall_cols if r[ix[h]].isdigit() and int(r[ix[h]]) > 0), reverse=True)
    print(r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9), r[ix["Source"]].strip()[:64].ljust(64), reasons[:3])
