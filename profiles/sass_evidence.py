"""Counts the Blackwell-specific SASS mnemonics per kernel of libss2.so (cuobjdump, no GPU needed):
UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor loads, FFMA2/FMUL2/FADD2 = packed fp32."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "stabstitch2_b200/libss2.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
pat = re.compile(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)")
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    m = pat.match(line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        if op.startswith("UTC") and op.endswith("MMA"):
            counts[cur]["UTC*MMA"] += 1
        elif op in ("LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "FFMA2", "FMUL2", "FADD2", "HMMA", "SYNCS", "ELECT"):
            counts[cur][op] += 1
keys = ["total", "UTC*MMA", "LDTM", "UTMALDG", "SYNCS", "ELECT", "FFMA2", "FMUL2", "FADD2", "HMMA"]
print("%-72s %s" % ("kernel", " ".join("%8s" % k for k in keys)))
for k, c in counts.items():
    if any(c[x] for x in keys[1:]):
        print("%-72s %s" % (k[:72], " ".join("%8d" % c[x] for x in keys)))
