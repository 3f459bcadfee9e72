"""Stall samples of an exported source page grouped into regions delimited by marker instructions."""
import csv, gzip, io, re, sys
path = sys.argv[1]
raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]
i_src, i_s, i_n = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
body = rows[hi + 1:]
body = body[:len(body) // 2] if len(body) > 2 and body[0][0] == body[len(body) // 2][0] else body
MARK = re.compile(r"LDTM|STG|LDG|UTC\w*MMA|UBLKCP|SYNCS|BAR\.|UTCBAR|STS|LDS|EXIT|ELECT|MUFU")
acc = 0; accn = 0; start = 0
tot = sum(int(r[i_s]) for r in body if len(r) > i_n and r[i_s].isdigit())
print("total samples", tot)
for k, r in enumerate(body):
    if len(r) <= i_n or not r[i_s].isdigit(): continue
    s, n = int(r[i_s]), int(r[i_n] or 0)
    src = r[i_src].strip()
    if MARK.search(src):
        if acc: print("        ... #%d-%d: %6d samples %5.1f%% (%d instr exec)" % (start, k - 1, acc, 100.0 * acc / tot, accn))
        print("#%-5d %6d %5.1f%% n=%10d %s" % (k, s, 100.0 * s / tot, n, src[:90]))
        acc = 0; accn = 0; start = k + 1
    else:
        acc += s; accn += n
if acc: print("        ... #%d-end: %6d samples %5.1f%%" % (start, acc, 100.0 * acc / tot))
