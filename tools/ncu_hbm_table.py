"""Table of the HBM-bound launches in an ncu report: duration, DRAM bytes, achieved GB/s against the measured copy peak.
usage: python tools/ncu_hbm_table.py report.ncu-rep [last_n]   (needs ncu on PATH; reads `ncu -i ... --page raw --csv`)"""
import csv, io, json, os, subprocess, sys
rep = sys.argv[1]
last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
def find(d, pat):
    for k, v in d.items():
        if isinstance(v, dict):
            r = find(v, pat)
            if r: return r
        elif pat in k.lower() and isinstance(v, (int, float)): return v
    return None
peak = find(peaks, "hbm") or find(peaks, "copy") or 6548.0
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
def val(r, name):
    i = col[name]
    v = float(r[i].replace(",", "")) if r[i] else 0.0
    u = units[i]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}.get(u, 1.0)
    return v * scale
if last: data = data[-last:]
print(f"(DRAM GB/s = (read + write) / duration; HBM copy peak {peak:.0f} GB/s)")
for r in data:
    name = r[col["Kernel Name"]].split("(")[0]
    us = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    gbs = (rd + wr) / us / 1e3
    wa = val(r, "sm__warps_active.avg.pct_of_peak_sustained_active")
    print(f"{name:42s} {us:7.1f} us  dram r {rd/1e6:8.1f} MB  w {wr/1e6:8.1f} MB  {gbs:7.0f} GB/s  ({100*gbs/peak:4.0f}% of copy peak)  warps active {wa:4.0f}%")
