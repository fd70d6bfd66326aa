"""Summarise an `ncu --set full` capture for profiles/ (not a test).

python tests/ncu_summary.py <raw.csv> <out_prefix> [source note]
  <raw.csv>     : `ncu -i capture.ncu-rep --page raw --csv > raw.csv`
  writes <out_prefix>_kernels.csv  one row per captured launch: kernel, grid, block, registers, duration, DRAM bytes read /
                                   written, DRAM and tensor-pipe utilisation, L2 bytes
         profiles/ncu_traffic.json {"bytes_per_launch": {kernel function: mean dram read + write bytes per launch}, "source": ...}
                                   -- what bench.py reports as roofline.traffic (read at run time, never a constant there)
"""
import csv
import json
import sys
from collections import defaultdict
from pathlib import Path

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3,
        "second": 1e6, "%": 1.0, "": 1.0}


def main():
    raw, prefix = Path(sys.argv[1]), sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else raw.name
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        i = col.get(name)
        if i is None or r[i] == "":
            return float("nan")
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    out = []
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0]
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        us = val(r, "gpu__time_duration.sum")
        out.append({"kernel": name, "grid": r[col["Grid Size"]], "block": r[col["Block Size"]],
                    "regs": r[col["launch__registers_per_thread"]], "us": round(us, 3), "dram_read_B": int(rd), "dram_write_B": int(wr),
                    "dram_pct": val(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                    "tensor_pipe_pct_active": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    "tensor_pipe_pct_elapsed": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                    "sm_pct": val(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                    "l2_bytes": val(r, "lts__t_bytes.sum")})
        a = agg[name]
        a[0] += 1; a[1] += rd + wr; a[2] += us
    with open(prefix + "_kernels.csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=list(out[0].keys()))
        w.writeheader()
        w.writerows(out)
    traffic = {k: v[1] / v[0] for k, v in agg.items()}
    p = Path(__file__).resolve().parents[1] / "profiles" / "ncu_traffic.json"
    p.write_text(json.dumps({"bytes_per_launch": traffic, "launches": {k: v[0] for k, v in agg.items()},
                             "us_per_launch_cold": {k: v[2] / v[0] for k, v in agg.items()}, "source": note}, indent=1) + "\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        print(f"{k:34s} launches {v[0]:3d}  {v[2] / v[0]:8.2f} us/launch  {v[1] / v[0] / 1e6:8.2f} MB DRAM/launch")


if __name__ == "__main__":
    main()
