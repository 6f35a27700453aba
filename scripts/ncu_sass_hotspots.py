"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump: samples by opcode / stall reason / hottest SASS."""
import collections
import csv
import sys


def num(s):
    try:
        return int(s.replace(",", ""))
    except Exception:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
hdr = rows[hi[0]]
src = rows[hi[0] + 1:(hi[1] if len(hi) > 1 else len(rows))]
ci = hdr.index("# Samples")
ie = hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
sass = [r for r in src if len(r) > ci and r[2].strip()]
tot = sum(num(r[ci]) for r in sass)
print("kernel:", rows[1][1][:120])
print("sass rows", len(sass), "samples", tot, "warp-insts", sum(num(r[ie]) for r in sass))
byop, byst, byopi = collections.Counter(), collections.Counter(), collections.Counter()
for r in sass:
    n = num(r[ci])
    t = r[3].strip().split()
    op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
    byop[op] += n
    byopi[op] += num(r[ie])
    for i, h in stall_cols:
        byst[h[6:]] += num(r[i])
print("samples by opcode :", [(k, f"{100 * v / tot:.1f}%") for k, v in byop.most_common(18)])
print("insts by opcode   :", [(k, v) for k, v in byopi.most_common(18)])
print("samples by stall  :", [(k, f"{100 * v / tot:.1f}%") for k, v in byst.most_common(12)])
for r in sorted(sass, key=lambda r: -num(r[ci]))[:topn]:
    st = sorted([(num(r[i]), h[6:]) for i, h in stall_cols], reverse=True)[:2]
    print(f"{100 * num(r[ci]) / tot:5.1f}% L{r[0]:>5s} {r[3][:84]:84s} {st}")
