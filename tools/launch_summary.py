"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, re, sys
path = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 150.0
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(float); cnt = collections.Counter(); seq = []
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
    short = re.sub(r"\(.*", "", row["Kernel Name"])[:72]
    tot[short] += v; cnt[short] += 1; seq.append((short, v, row.get("Grid Size"), row.get("Block Size")))
T = sum(tot.values())
print("total %.1f us over %d launches" % (T, len(seq)))
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:22]:
    print("%9.1f us %5.1f%% n=%3d  %s" % (v, 100 * v / T, cnt[k], k))
print()
for i, s in enumerate(seq):
    if s[1] > thr:
        print(i, "%8.1f" % s[1], s[0][-40:], s[2])
