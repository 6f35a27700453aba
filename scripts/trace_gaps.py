"""Where does the compute stream idle?  Reads a chrome trace exported by `bench.py --profile P --trace T`, takes the
stream that carries the GEMMs, and lists the largest gaps between consecutive kernels with their neighbours."""
import collections
import gzip
import json
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
op = gzip.open if path.endswith(".gz") else open
ev = json.load(op(path, "rt"))["traceEvents"]
kern = [e for e in ev if e.get("cat") == "kernel"]
by_stream = collections.defaultdict(list)
for e in kern:
    by_stream[e["args"].get("stream")].append(e)
main = max(by_stream, key=lambda s: sum(e["dur"] for e in by_stream[s] if "gemm" in e["name"]))
ks = sorted(by_stream[main], key=lambda e: e["ts"])
span = ks[-1]["ts"] + ks[-1]["dur"] - ks[0]["ts"]
busy = sum(e["dur"] for e in ks)
print(f"streams: { {s: round(sum(e['dur'] for e in v) / 1e3, 1) for s, v in by_stream.items()} } (ms busy)")
print(f"compute stream {main}: span {span / 1e3:.1f} ms, busy {busy / 1e3:.1f} ms, idle {(span - busy) / 1e3:.1f} ms over {len(ks)} kernels")
gaps = []
for a, b in zip(ks, ks[1:]):
    g = b["ts"] - (a["ts"] + a["dur"])
    if g > 0:
        gaps.append((g, a, b))
hist = collections.Counter()
for g, a, b in gaps:
    hist["<5us" if g < 5 else "<20us" if g < 20 else "<100us" if g < 100 else "<1ms" if g < 1000 else ">=1ms"] += g
print("idle time by gap size (ms):", {k: round(v / 1e3, 2) for k, v in hist.items()})
agg = collections.Counter()
for g, a, b in gaps:
    agg[(a["name"][:50], b["name"][:50])] += g
print("\nidle time by (previous kernel -> next kernel), ms:")
for (x, y), g in agg.most_common(topn):
    print(f"  {g / 1e3:8.2f}  {x}  ->  {y}")
print("\nlargest single gaps:")
t0 = ks[0]["ts"]
for g, a, b in sorted(gaps, key=lambda t: -t[0])[:topn]:
    print(f"  {g / 1e3:8.3f} ms at t={(a['ts'] - t0) / 1e3:8.1f} ms  {a['name'][:48]}  ->  {b['name'][:48]}")
