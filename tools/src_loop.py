"""Stall-reason and opcode breakdown of the instructions of an exported ncu source page whose execution count lies in
[lo, hi] (isolates one loop body, e.g. the producers' per-K-block code)."""
import csv, gzip, io, re, sys, collections
path, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
rows = list(csv.reader(io.StringIO(raw)))
hi_ = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi_]; body = rows[hi_ + 1:]
i_src, i_s, i_n, i_a = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed"), hdr.index("Address")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
seen = set(); ops = collections.Counter(); opsamp = collections.Counter(); agg = collections.Counter(); tot = 0; ninstr = 0; nexec = 0
for r in body:
    if len(r) <= i_n or r[i_a] in seen or not r[i_n].isdigit(): continue
    seen.add(r[i_a])
    n = int(r[i_n])
    if not (lo <= n <= hi): continue
    op = re.sub(r'^@!?U?P\d+\s+', '', r[i_src].strip()).split()[0].rstrip(';')
    s = int(r[i_s]) if r[i_s].isdigit() else 0
    ops[op] += 1; opsamp[op] += s; tot += s; ninstr += 1; nexec += n
    for i, h in stall_cols:
        if r[i].isdigit(): agg[h[6:]] += int(r[i])
print("static instr %d, executed %d, samples %d" % (ninstr, nexec, tot))
print("stalls:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, sum(agg.values()))) for k, v in agg.most_common(10)))
for op, c in ops.most_common(40):
    print("%-28s static %4d  samples %6d (%.1f%%)" % (op, c, opsamp[op], 100.0 * opsamp[op] / max(1, tot)))
