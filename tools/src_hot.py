"""Hottest SASS instructions of an exported `ncu --page source --csv` file (.csv or .csv.gz)."""
import csv, gzip, io, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
i_src, i_s, i_n = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_")]
data = []
for k, r in enumerate(rows[hi + 1:]):
    if len(r) > i_n and r[i_s].isdigit():
        data.append((int(r[i_s]), r[i_src].strip(), int(r[i_n] or 0), k, r))
half = len(data) // 2
if len(data) % 2 == 0 and [d[:3] for d in data[:half]] == [d[:3] for d in data[half:]]:
    data = data[:half]                      # some exports list the kernel's instructions twice
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for s, src, n, k, r in data:
    for i, h in stall_cols:
        if r[i].isdigit(): agg[h] = agg.get(h, 0) + int(r[i])
print("stall reasons:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(1, sum(agg.values()))) for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for s, src, n, k, r in sorted(data, key=lambda d: -d[0])[:top]:
    why = sorted(((int(r[i]), h[6:]) for i, h in stall_cols if r[i].isdigit() and int(r[i])), reverse=True)[:2]
    print("%6d %5.1f%%  n=%9d  #%-5d %-70s %s" % (s, 100.0 * s / max(tot, 1), n, k, src[:70], why))
