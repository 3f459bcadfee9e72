"""Stall samples of an exported ncu source page per block of SASS instructions, with the marker opcodes in each block."""
import csv, gzip, io, re, collections, sys
path = sys.argv[1]; BIN = int(sys.argv[2]) if len(sys.argv) > 2 else 400; MIN = int(sys.argv[3]) if len(sys.argv) > 3 else 300
rows = list(csv.reader(io.StringIO(gzip.open(path, 'rt').read() if path.endswith('.gz') else open(path).read())))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
body = body[:len(body) // 2] if len(body) > 2 and body[0][0] == body[len(body) // 2][0] else body
i_src, i_s, i_n = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[i_s]) for r in body if r[i_s].isdigit())
print("total samples", tot, "instructions", len(body))
for b0 in range(0, len(body), BIN):
    seg = body[b0:b0 + BIN]
    s = sum(int(r[i_s]) for r in seg if r[i_s].isdigit())
    n = sum(int(r[i_n] or 0) for r in seg if r[i_n].isdigit())
    ops = collections.Counter()
    for r in seg:
        m = re.sub(r'^@!?U?P\d+\s+', '', r[i_src].strip()).split()
        if m and re.match(r'LDTM|STG|LDG|UTC|UBLKCP|SYNCS|BAR|STS|LDS|F2FP|LDL|STL|MUFU|SHFL|USETMAXREG', m[0]): ops[m[0].split('.')[0]] += 1
    why = collections.Counter()
    for r in seg:
        for i, h in stall:
            if r[i].isdigit(): why[h[6:]] += int(r[i])
    if s >= MIN:
        print("%5d-%5d %6d %4.1f%% n=%9d %s | %s" % (b0, b0 + BIN, s, 100 * s / tot, n, dict(ops), ", ".join("%s %d" % (k, v) for k, v in why.most_common(3))))
