"""DRAM traffic of the DCN launches of ONE step from an exported `ncu --set full --page raw --csv` capture
(tools/ncu_capture.sh <tag> "dcn_tile_kernel|conv_gather_kernel<0" 16 ...): sums dram__bytes_read.sum +
dram__bytes_write.sum over the captured launches and records the figure in profiles/dcn_traffic.json, which bench.py
reads at run time for `roofline.traffic`.
usage: python tools/ncu_traffic.py <raw.csv> <mode> <batch> <source label>"""
import csv, json, os, sys
path, mode, batch, label = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
C = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = {"dram__bytes_read.sum": 0.0, "dram__bytes_write.sum": 0.0}
n = 0
for r in rows[2:]:
    name = r[C["Kernel Name"]]
    if "dcn_tile_kernel" not in name and "conv_gather_kernel<0" not in name and "conv_gather_kernel<(sgta::PROD)0" not in name:
        continue
    n += 1
    for k in tot:
        tot[k] += float(r[C[k]].replace(",", "")) * scale[units[C[k]]]
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "dcn_traffic.json")
d = json.load(open(out)) if os.path.exists(out) else {}
d["%s_b%d" % (mode, batch)] = {"bytes_per_step": tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"],
                               "read": tot["dram__bytes_read.sum"], "write": tot["dram__bytes_write.sum"], "launches": n,
                               "source": label}
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d["%s_b%d" % (mode, batch)]))
