"""SASS opcode census of the shipped library (no GPU needed): python tools/sass_census.py > profiles/rNN_sass_census.txt
Counts, per kernel, the instructions that prove which hardware path it uses (B200_PROFILING.md mnemonics)."""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "se_unet_airseg_b200", "libseunet_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEY = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCATOMSWS", "UTMACCTL", "LDTM", "STTM", "UTCCP", "HMMA", "LDSM", "LDGSTS",
       "RED", "ATOM", "ATOMG", "DADD", "DFMA", "MUFU", "SHFL")
funcs, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        funcs[cur]["instrs"] += 1
        op = m.group(1)
        if op in KEY: funcs[cur][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
print("SASS opcode census of se_unet_airseg_b200/libseunet_b200.so (cuobjdump -sass, sm_100a; tools/sass_census.py).")
print("UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, UBLKCP = bulk copy (weights), UTCBAR = tcgen05.commit, LDTM = tcgen05.ld,")
print("HMMA/LDSM = warp-level mma.sync/ldmatrix (only the fused apply + 1x1x1 CAT pass), LDGSTS = cp.async, RED/ATOM = global atomics.\n")
tot = collections.Counter()
for (mangled, c), name in sorted(zip(funcs.items(), names), key=lambda t: -t[0][1]["instrs"]):
    tot.update(c)
    tags = " ".join(f"{k}={c[k]}" for k in KEY if c[k])
    print(f"{name[:110]:110s} instrs={c['instrs']:6d} {tags}")
print("\nwhole library: " + " ".join(f"{k}={tot[k]}" for k in ("instrs",) + KEY if tot[k]))
