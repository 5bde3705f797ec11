"""Per-kernel counts of the SASS mnemonics that show a Blackwell-native path (cuobjdump -sass of libsoapb200.so):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor copies, UBLKCP = 1-D bulk copies,
DMMA = fp64 mma.sync, SYNCS = mbarrier.   python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cpctools_b200", "libsoapb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "DMMA", "HMMA", "IMMA", "SYNCS", "LDGSTS", "DFMA", "LDS", "ATOMS", "MATCH"]
per, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("soapb::", "")
        per[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        for k in keys:
            if op == k or (k == "UTCIMMA" and op.startswith("UTCIMMA")):
                per[name][k] += 1
print(f"# {os.path.basename(lib)}: {len(per)} kernels; SASS mnemonic counts per kernel (only kernels with at least one of them)")
print("# " + " ".join(f"{k:>8s}" for k in keys) + "  kernel")
tot = collections.Counter()
for n, c in per.items():
    tot.update(c)
    if any(c[k] for k in keys[:10]):
        print("  " + " ".join(f"{c[k]:8d}" for k in keys) + "  " + n[:110])
print("# totals")
print("  " + " ".join(f"{tot[k]:8d}" for k in keys))
