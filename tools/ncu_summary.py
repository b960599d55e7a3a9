"""Summarise an `ncu --page raw --csv` export (one row per profiled launch) into one line per launch and one block
per kernel: duration, DRAM bytes read / written, DRAM throughput %, tensor-pipe active %, issue-slot %, occupancy,
registers.  Usage: python tools/ncu_summary.py gpurun_out/prof_users_raw.csv [--per-launch]"""
import csv
import re
import sys
from collections import OrderedDict

COLS = {
    "dur_ms": r"^gpu__time_duration\.sum$",
    "rd": r"^dram__bytes_read\.sum$",
    "wr": r"^dram__bytes_write\.sum$",
    "dram_pct": r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    "tensor_pct": r"sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed$",
    "issue_pct": r"^sm__issue_active\.avg\.pct_of_peak_sustained_elapsed$",
    "warps_pct": r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    "regs": r"^launch__registers_per_thread$",
    "l2_hit": r"^lts__t_sector_hit_rate\.pct$",
    "sm_clk": r"^sm__cycles_elapsed\.avg\.per_second$",
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6,
              "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    path = sys.argv[1]
    per_launch = "--per-launch" in sys.argv
    rd = csv.reader(open(path))
    hdr = next(rd)
    units = next(rd)
    idx = {}
    for key, pat in COLS.items():
        for i, h in enumerate(hdr):
            if re.search(pat, h):
                idx[key] = i
                break
    ki, gi, bi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size")
    groups = OrderedDict()
    for row in rd:
        if len(row) < len(hdr):
            continue
        v = {}
        for key, i in idx.items():
            x = num(row[i])
            v[key] = x * UNIT_SCALE.get(units[i], 1.0)
        name = re.sub(r"\(.*", "", row[ki]).replace("unirec::", "")
        name = re.sub(r"^void ", "", name)
        tag = f"{name} grid={row[gi]} block={row[bi]}"
        if per_launch:
            print(f"{tag:90s} {v.get('dur_ms', 0):8.3f} ms  rd {v.get('rd', 0) / 1e6:9.1f} MB  wr {v.get('wr', 0) / 1e6:9.1f} MB  "
                  f"dram {v.get('dram_pct', 0):5.1f}%  tensor {v.get('tensor_pct', 0):5.1f}%  issue {v.get('issue_pct', 0):5.1f}%")
        groups.setdefault(tag, []).append(v)
    print(f"{'kernel':90s} {'n':>4s} {'avg ms':>8s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>7s} {'dram%':>6s} {'tens%':>6s} "
          f"{'issue%':>6s} {'warps%':>6s} {'regs':>4s} {'L2hit%':>6s}")
    tot = sum(sum(x.get("dur_ms", 0) for x in g) for g in groups.values())
    for tag, g in sorted(groups.items(), key=lambda kv: -sum(x.get("dur_ms", 0) for x in kv[1])):
        n = len(g)
        avg = lambda k: sum(x.get(k, 0) for x in g) / n
        gbs = (avg("rd") + avg("wr")) / (avg("dur_ms") * 1e-3) / 1e9 if avg("dur_ms") > 0 else 0
        print(f"{tag:90s} {n:4d} {avg('dur_ms'):8.3f} {avg('rd') / 1e6:9.1f} {avg('wr') / 1e6:9.1f} {gbs:7.0f} {avg('dram_pct'):6.1f} "
              f"{avg('tensor_pct'):6.1f} {avg('issue_pct'):6.1f} {avg('warps_pct'):6.1f} {avg('regs'):4.0f} {avg('l2_hit'):6.1f}"
              f"   share {100 * sum(x.get('dur_ms', 0) for x in g) / tot:5.1f}%")
    print(f"total profiled time {tot:.3f} ms over {sum(len(g) for g in groups.values())} launches")


if __name__ == "__main__":
    main()
