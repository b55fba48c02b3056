"""Top stalled SASS instructions of one kernel from `ncu -i rep --page source --csv` output (file argument)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
S = idx["# Samples"]
tot = sum(int(r[S] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for r in sorted(data, key=lambda r: -int(r[S] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    st = {k[6:]: r[idx[k]] for k in keys if r[idx[k]] not in ("0", "")}
    print(r[S].rjust(6), r[idx["Instructions Executed"]].rjust(9), r[idx["Source"]][:100].ljust(100), st)
