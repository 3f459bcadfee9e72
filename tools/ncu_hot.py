"""Hottest SASS instructions (warp stall samples) of one launch in an .ncu-rep."""
import csv, io, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "0"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
i_src, i_s, i_n = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[hi + 1:]):
    if len(r) > i_n and r[i_s].isdigit():
        data.append((int(r[i_s]), r[i_src].strip(), int(r[i_n] or 0), k))
tot = sum(d[0] for d in data)
print(rows[0][:2], "total samples", tot, "instructions", len(data))
for s, src, n, k in sorted(data, reverse=True)[:top]:
    print("%6d %5.1f%%  n=%9d  #%-5d %s" % (s, 100.0 * s / max(tot, 1), n, k, src[:100]))
