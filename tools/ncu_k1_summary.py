"""Summarise an `ncu --set full` raw page (CSV) of ONE tick's K1 kernels: per kernel time, DRAM bytes, registers,
occupancy; and the tick's total DRAM traffic, which bench.py reports as roofline.traffic.
  python tools/ncu_k1_summary.py gpurun_out/r02_sort_k1_full_raw.csv profiles/r02_k1_traffic.json [profiles/r02_k1_summary.md]"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(name):
    return hdr.index(name)


def val(r, name, scale_to=None):
    i = col(name)
    v = float(r[i].replace(",", "")) if r[i] else 0.0
    u = units[i]
    if scale_to == "byte":
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    if scale_to == "us":
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}[u]
    return v


kern = []
for r in data:
    name = r[col("Kernel Name")]
    short = name.replace("void gcra::", "").split("(")[0]
    kern.append({"kernel": short, "us": val(r, "gpu__time_duration.sum", "us"),
                 "dram_read": val(r, "dram__bytes_read.sum", "byte"), "dram_write": val(r, "dram__bytes_write.sum", "byte"),
                 "regs": val(r, "launch__registers_per_thread"),
                 "occupancy_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                 "ipc": val(r, "sm__inst_executed.avg.per_cycle_elapsed"),
                 "l2_hit_pct": val(r, "lts__t_sector_hit_rate.pct")})
total = sum(k["dram_read"] + k["dram_write"] for k in kern)
out = {"source": sys.argv[1], "kernels_of_one_tick": kern, "dram_bytes_per_tick": total,
       "dram_read_per_tick": sum(k["dram_read"] for k in kern), "dram_write_per_tick": sum(k["dram_write"] for k in kern),
       "sum_of_kernel_times_us": sum(k["us"] for k in kern)}
json.dump(out, open(sys.argv[2], "w"), indent=1)
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        f.write("| kernel | us | DRAM read MB | DRAM write MB | regs | achieved occupancy % | IPC / SM | L2 hit % |\n|---|---|---|---|---|---|---|---|\n")
        for k in kern:
            f.write("| `%s` | %.1f | %.1f | %.1f | %d | %.0f | %.2f | %.0f |\n" % (k["kernel"], k["us"], k["dram_read"] / 1e6, k["dram_write"] / 1e6,
                                                                          k["regs"], k["occupancy_pct"], k["ipc"], k["l2_hit_pct"]))
        f.write("\nDRAM traffic of the tick: %.1f MB (read %.1f + write %.1f); kernel times sum to %.1f us (cold-cache, serialised by ncu).\n"
                % (total / 1e6, out["dram_read_per_tick"] / 1e6, out["dram_write_per_tick"] / 1e6, out["sum_of_kernel_times_us"]))
print("dram bytes per tick: %.1f MB over %d kernels" % (total / 1e6, len(kern)))
