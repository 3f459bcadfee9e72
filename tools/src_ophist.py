"""Dynamic opcode histogram (warp instructions executed) of an exported ncu source page."""
import csv, gzip, io, re, collections, sys
path = sys.argv[1]
raw = gzip.open(path, 'rt').read() if path.endswith('.gz') else open(path).read()
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
hdr = rows[hi]; body = rows[hi + 1:]
seen = set(); hist = collections.Counter(); tot = 0
i_src, i_n, i_a = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Address")
for r in body:
    if len(r) <= i_n or r[i_a] in seen or not r[i_n].isdigit(): continue
    seen.add(r[i_a])
    n = int(r[i_n])
    op = re.sub(r'^@!?U?P\d+\s+', '', r[i_src].strip()).split()[0].rstrip(';')
    hist[op] += n; tot += n
print("total warp instr", tot)
for op, n in hist.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 30): print("%-28s %12d %5.1f%%" % (op, n, 100 * n / tot))
