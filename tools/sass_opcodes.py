"""SASS opcode histogram per kernel of the shipped library (the tcgen05 / TMA / TMEM evidence the profiling recipe asks
for).  Usage: python tools/sass_opcodes.py > profiles/rNN_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "rule_guided_music_b200", "librgm_b200.so")
KEEP = re.compile(r"^(UTC\w*|LDTM|STTM|UTMA\w*|SYNCS|HMMA|LDSM|LDGSTS|MUFU|ATOM\w*|RED|REDUX|MEMBAR|CCTL|NANOSLEEP)")
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
demangle = {}
names = sorted(set(re.findall(r"Function : (\S+)", out)))
if names:
    dm = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, dm))
hist = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"\(.*", "", demangle.get(m.group(1), m.group(1)))
        hist.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]+)", line)
    if m and cur is not None and KEEP.match(m.group(1)):
        hist[cur][m.group(1)] += 1
print("SASS opcode histogram of rule_guided_music_b200/librgm_b200.so (cuobjdump -sass, sm_100a), per kernel.")
print("UTC*MMA = tcgen05.mma (UTCHMMA: kind::f16; .2CTA: cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA tiled load,")
print("UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA/LDSM = mma.sync/ldmatrix (vae_out only), LDGSTS = cp.async, MUFU.* = SFU,")
print("ATOMG / RED on 64-bit words = the GroupNorm statistics accumulators of the convolutions that normalise their own output.\n")
total = collections.Counter()
for k, c in hist.items():
    if not c:
        continue
    print(k)
    print("    " + "  ".join("%s:%d" % kv for kv in sorted(c.items())))
    total.update(c)
print("\nTOTAL")
print("    " + "  ".join("%s:%d" % kv for kv in sorted(total.items())))
