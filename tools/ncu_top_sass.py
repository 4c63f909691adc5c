"""Top stall-sample SASS lines per kernel section of an `ncu --page source --csv --print-source sass` dump. Dev tool."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
which = int(sys.argv[3]) if len(sys.argv) > 3 else -1
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; secs.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
for k, s in enumerate(secs):
    if which >= 0 and k != which:
        continue
    h = s["hdr"]; iS = h.index("# Samples"); iE = h.index("Instructions Executed")
    data = [(int(r[iS] or 0), int(r[iE] or 0), i, r[1].strip()) for i, r in enumerate(s["rows"]) if len(r) > iE]
    tot = sum(d[0] for d in data)
    print(f"== section {k}: {s['name']} samples={tot} instrs={len(data)}")
    for d in sorted(data, reverse=True)[:top]:
        print(f"  {d[0]:6d} {100*d[0]/max(tot,1):5.1f}%  exec={d[1]:8d} line={d[2]:5d}  {d[3]}")
