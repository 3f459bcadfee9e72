"""Per-launch table from an exported `ncu --page raw --csv` file (tools/ncu_capture.sh)."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
C = {h: i for i, h in enumerate(hdr)}
cols = [("us", "gpu__time_duration.sum"), ("dramR_MB", "dram__bytes_read.sum"), ("dramW_MB", "dram__bytes_write.sum"),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
        ("tensor_rt%", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
        ("dram%", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1%", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l2%", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("l1hit%", "l1tex__t_sector_hit_rate.pct"), ("l2hit%", "lts__t_sector_hit_rate.pct"),
        ("regs", "launch__registers_per_thread"), ("smem_KB", "launch__shared_mem_per_block_dynamic")]
def val(r, name):
    if name not in C: return float("nan")
    s = r[C[name]].replace(",", "")
    try: v = float(s)
    except ValueError: return float("nan")
    u = units[C[name]]
    if u == "ns": v /= 1e3
    elif u == "ms": v *= 1e3
    elif u == "second" or u == "s": v *= 1e6
    elif u == "byte": v /= 1e6
    elif u == "Kbyte": v /= 1e3
    elif u == "Gbyte": v *= 1e3
    if name.startswith("launch__shared"): v = v * 1e6 / 1024 if u == "byte" else v * 1e3 / 1.024 if u == "Kbyte" else v
    return v
print("%3s %-34s %-9s " % ("#", "kernel", "grid") + " ".join("%9s" % c[0] for c in cols))
for i, r in enumerate(rows[2:]):
    name = re.sub(r"^void |sgta::|\(.*", "", r[C["Kernel Name"]])[:34]
    print("%3d %-34s %-9s " % (i, name, r[C["Grid Size"]].replace(" ", "")[:9]) + " ".join("%9.1f" % val(r, c[1]) for c in cols))
