"""Per-kernel launch counts / time shares from an `ncu --metrics gpu__time_duration.sum --csv --log-file X` launch list.
usage: python tools/launch_shares.py launches.csv"""
import collections, csv, re, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    us = float(r[iV].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iU], 1e-3)
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "")
    a = acc.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(v[1] for v in acc.values())
print(f"{'kernel':60s} {'launches':>8s} {'total us':>10s} {'share':>7s}")
for k, (n, us) in sorted(acc.items(), key=lambda t: -t[1][1]):
    print(f"{k:60s} {n:8d} {us:10.1f} {100*us/tot:6.1f}%")
conv = sum(v[1] for k, v in acc.items() if k.startswith("conv_tc_kernel"))
print(f"\nconv_tc_kernel share of the captured GPU time: {100*conv/tot:.1f} %")
