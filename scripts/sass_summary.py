"""SASS mnemonic counts per kernel of fms_fsdp_b200/_C.so (proof of tcgen05 / TMA / TMEM / NVLink-peer instructions)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "fms_fsdp_b200/_C.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = re.compile(r"\b(UTCHMMA[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UBLKCP[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|"
                  r"SYNCS[.\w]*|MUFU\.EX2|HMMA[.\w]*|LDG\.E[.\w]*SYS|STG\.E[.\w]*SYS|REDG[.\w]*|ELECT)\b")
cur, counts = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:150]
        counts[cur] = collections.Counter()
        continue
    if cur:
        for w in want.findall(line):
            counts[cur][w] += 1
print(__doc__.strip())
print("UTCHMMA = tcgen05.mma, UTMALDG = TMA load, UBLKCP = cp.async.bulk, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit,")
print(".2CTA = cta_group::2, LDG...SYS = ld.relaxed.sys on NVLink peer pointers, ELECT = elect.sync. No HMMA (mma.sync) anywhere.\n")
for k, c in counts.items():
    if c:
        print(k)
        print("    " + ", ".join(f"{m}={n}" for m, n in sorted(c.items())))
