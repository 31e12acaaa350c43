"""Counts of the Blackwell-native SASS mnemonics per kernel of the built library (profiles/r2_sass_summary.txt):
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor loads / stores, UBLKCP = bulk copies,
VHMNMX = packed 3-input FP16 minimum, FMNMX3 = 3-input FP32 minimum, SYNCS = mbarrier operations.

  python scripts/sass_summary.py [reductive_b200/lib/libreductive_b200.so]
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "reductive_b200/lib/libreductive_b200.so"
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "VHMNMX", "HMNMX2", "FMNMX3", "HSET2", "FFMA2",
        "SYNCS", "HMMA"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts, total = collections.OrderedDict(), collections.Counter()
fn = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::", "", fn).split("(")[0]
        counts.setdefault(fn, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["_all"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[fn][k] += 1
                total[k] += 1
print(f"# {lib}: SASS mnemonic counts per kernel (cuobjdump -sass), sm_100a")
print("kernel".ljust(64) + "instr".rjust(7) + "".join(k.rjust(9) for k in KEYS))
for fn, c in counts.items():
    if any(c[k] for k in KEYS if k not in ("SYNCS", "FFMA2")):
        print(fn[:63].ljust(64) + str(c["_all"]).rjust(7) + "".join(str(c[k] or "").rjust(9) for k in KEYS))
print("total".ljust(71) + "".join(str(total[k]).rjust(9) for k in KEYS))
