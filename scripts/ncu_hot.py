"""Top SASS lines of an ncu source-page export (scripts/r2_ncu.sh: <tag>_source.csv.gz) by stall samples, with the
dominant stall reason and execution counts; plus an instruction-mix tally weighted by executions."""
import collections, csv, gzip, io, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.TextIOWrapper(gzip.open(path))))
hdr = rows[1]; ix = {k: i for i, k in enumerate(hdr)}
stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    samples = int(r[ix["# Samples"]] or 0); ex = int(r[ix["Instructions Executed"]] or 0)
    st = {k: int(r[ix[k]] or 0) for k in stall_cols}
    data.append((samples, ex, r[ix["Source"]].strip(), st, rows.index(r) if False else 0))
tot = sum(d[0] for d in data); totex = sum(d[1] for d in data)
print(f"total samples {tot}, warp instructions executed {totex}, SASS lines {len(data)}")
agg = collections.Counter()
for d in data:
    for k, v in d[3].items(): agg[k] += v
print("stall totals:", ", ".join(f"{k[6:]}={v} ({100*v/max(tot,1):.1f}%)" for k, v in agg.most_common(8)))
print(f"{'line':>5} {'samples':>8} {'pct':>6} {'executed':>10}  top-stall        SASS")
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
for i in sorted(order):
    s, ex, src, st, _ = data[i]
    k = max(st, key=st.get) if st else ""
    print(f"{i:5d} {s:8d} {100*s/max(tot,1):6.2f} {ex:10d}  {k[6:]:<14}  {src[:90]}")
mix = collections.Counter()
for s, ex, src, st, _ in data:
    op = src.split()[0] if src else "?"
    if op.startswith("@"): op = src.split()[1]
    mix[op.split(".")[0]] += ex
print("instruction mix (executed):", ", ".join(f"{k}={100*v/max(totex,1):.1f}%" for k, v in mix.most_common(16)))
