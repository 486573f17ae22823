"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep, read here without a GPU) into the small text files committed
under profiles/ and into profiles/ncu_traffic.json (bench.py's roofline.traffic: DRAM read + write bytes of ONE launch
of the dominant kernel).
    python tools/ncu_summary.py <report.ncu-rep> <out.txt> [--traffic workload n_gpus]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v.replace(",", "")) * m.get(unit, 1)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none capture: {os.path.basename(rep)} (per-launch, cold-cache, serialised: never a bench value)"]
    traffic = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"kernel: {d.get('Kernel Name', '?')}")
        for k in KEYS:
            if k in d:
                lines.append(f"  {k:70s} {d[k]} {u.get(k, '')}")
        rd = to_bytes(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"])
        wr = to_bytes(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"])
        lines.append(f"  {'dram read + write bytes of this launch':70s} {rd + wr:.0f}")
        if traffic is None:
            traffic = rd + wr
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if "--traffic" in sys.argv:
        i = sys.argv.index("--traffic")
        wl, n = sys.argv[i + 1], sys.argv[i + 2]
        p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        t = json.load(open(p)) if os.path.exists(p) else {}
        t.setdefault(wl, {})[n] = traffic
        t.setdefault("_source", {})[f"{wl}/{n}"] = os.path.basename(out)
        json.dump(t, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
