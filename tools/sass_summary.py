"""Per-kernel counts of the Blackwell-native SASS mnemonics in the shipped library (cuobjdump -sass):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP/UTMAPF = TMA, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, LDGMC = multimem.ld_reduce (NVLink SHARP), HMMA = legacy mma.sync (must be absent).   python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = Path(__file__).resolve().parents[1] / "sinddm_b200" / "lib" / "libsinddm_b200.so"
out = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTMACCTL",
        "SYNCS", "LDGMC", "HMMA", "HGMMA", "FFMA", "LDG", "STG", "LDS", "STS", "MUFU", "SHFL", "REDUX", "LD.E", "ST.E"]
kern, counts, arch = None, collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\((?:bool|int)\)", "", kern)            # template arguments print as (bool)0, (int)1120
        kern = re.sub(r"\(.*", "", kern).replace("sinddm::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        counts[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    if kern:
        for p in pats:
            if re.search(r"\b" + re.escape(p) + r"\b|\b" + re.escape(p) + r"\.", line):
                counts[kern][p] += 1
print(f"# cuobjdump -sass {lib.name}: arch {arch}; instruction counts per kernel (static SASS)")
hdr = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "LDGMC", "HMMA", "FFMA", "MUFU"]
print(f"{'kernel':58s}" + "".join(f"{h:>9s}" for h in hdr))
for k, c in counts.items():
    print(f"{k[:58]:58s}" + "".join(f"{c.get(h, 0):9d}" for h in hdr))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print(f"{'TOTAL':58s}" + "".join(f"{tot.get(h, 0):9d}" for h in hdr))
assert tot.get("HMMA", 0) == 0 and tot.get("HGMMA", 0) == 0, "legacy tensor-core instructions found"
