"""profiles/r02_top_kernel_traffic.json from an `ncu --set full` raw CSV of tools/nb_profile_target.py (4 launches:
W K/4, W K, Zc K/4, Zc K iterations): DRAM bytes per BiCGStab iteration and per launch set-up, per system.
usage: python tools/ncu_traffic.py raw.csv K out.json [n]"""
import csv
import json
import sys

raw, K, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
n_mesh = int(sys.argv[4]) if len(sys.argv) > 4 else 192
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    v = float(r[ix[name]].replace(",", ""))
    u = units[ix[name]].lower()
    return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "msecond": 1e-3,
                "usecond": 1e-6, "second": 1.0}.get(u, 1.0)


sel = [r for r in rows[2:] if "k_bicgstab_nb" in r[ix["Kernel Name"]]][-4:]  # the capture may include the time step's solves
assert len(sel) == 4 and "<2" in sel[0][ix["Kernel Name"]] and "<3" in sel[3][ix["Kernel Name"]], [r[ix["Kernel Name"]] for r in sel]
res = {"kernel": "k_bicgstab_nb", "n": n_mesh, "source": raw, "K": K, "dram_bytes_per_iteration": {}, "dram_bytes_setup": {},
       "launches": []}
for sp, (a, b) in (("W", sel[0:2]), ("Zc", sel[2:4])):
    by = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in (a, b)]
    per = (by[1] - by[0]) / (K - K // 4)
    res["dram_bytes_per_iteration"][sp] = per
    res["dram_bytes_setup"][sp] = by[0] - per * (K // 4)
    for r, k in ((a, K // 4), (b, K)):
        res["launches"].append({"space": sp, "iterations": k, "kernel": r[ix["Kernel Name"]],
                                "duration_s_under_ncu": val(r, "gpu__time_duration.sum"),
                                "dram_read": val(r, "dram__bytes_read.sum"), "dram_write": val(r, "dram__bytes_write.sum"),
                                "dram_throughput_pct": float(r[ix["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
                                "warps_active_pct": float(r[ix["sm__warps_active.avg.pct_of_peak_sustained_active"]]),
                                "registers": int(float(r[ix["launch__registers_per_thread"]]))})
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res["dram_bytes_per_iteration"]), json.dumps(res["dram_bytes_setup"]))
