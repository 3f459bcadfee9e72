#!/bin/bash
# Runs ON THE GPU BOX (under gpurun).  One `ncu --set full` capture of the launches matching a
# kernel regex inside ONE un-graphed bench step, exported as CSV because the .ncu-rep (2 MB per
# launch) does not fit gpurun_out's 64 MiB:
#   gpurun_out/<tag>_raw.csv            all metrics of every captured launch (--page raw)
#   gpurun_out/<tag>_src_<i>.csv.gz     per-instruction page of the TOP launches by duration
# usage: tools/ncu_capture.sh <tag> <kernel regex> <count> [n_top_source] [extra bench args...]
tag=$1; rx=$2; n=$3; top=${4:-4}; shift 4
rep=/tmp/$tag
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$rx" --launch-skip "${SKIP:-0}" -c "$n" \
    -f -o $rep python bench.py --profile-pass --no-cpu-baseline "$@" > gpurun_out/${tag}_ncu.log 2>&1
ncu -i $rep.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python - "$tag" "$top" <<'EOF' > /tmp/${tag}_top.txt
import csv, sys
tag, top = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open("gpurun_out/%s_raw.csv" % tag)))
hdr = rows[0]
name, dur = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
best = {}
for i, r in enumerate(rows[2:]):
    key = (r[name], r[hdr.index("Grid Size")])
    v = float(r[dur].replace(",", ""))
    if key not in best or v > best[key][0]:
        best[key] = (v, i)
for v, i in sorted(best.values(), reverse=True)[:top]:
    print(i)
EOF
for i in $(cat /tmp/${tag}_top.txt); do
  ncu -i $rep.ncu-rep --page source --csv --launch-skip $i --launch-count 1 2>/dev/null | gzip -9 > gpurun_out/${tag}_src_$i.csv.gz
done
ls -la gpurun_out | head -40
