"""DRAM traffic per launch of the hot-path kernels, from `ncu --set full` reports of the CURRENT build -> profiles/r2_ncu_traffic.json
(what bench.py reports as `roofline*.traffic`, with the report it came from).  Run where ncu is installed; no GPU needed.

    python tools/ncu_traffic.py name=report.ncu-rep [name=report.ncu-rep ...]
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")


def read(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]

    def val(name, row):
        i = hdr.index(name)
        v, u = float(row[i].replace(",", "")), units[i]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6,
                 "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
        return v * scale

    res = []
    for row in data:
        res.append({"kernel": row[hdr.index("Kernel Name")], "dram_read_bytes": val("dram__bytes_read.sum", row),
                    "dram_write_bytes": val("dram__bytes_write.sum", row), "duration_s_under_ncu": val("gpu__time_duration.sum", row),
                    "tensor_pipe_active_pct": float(row[hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")])
                    if "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active" in hdr else None})
    return res


def main():
    table = json.load(open(OUT)) if os.path.isfile(OUT) else {}
    for arg in sys.argv[1:]:
        name, path = arg.split("=", 1)
        launches = read(path)
        k = launches[-1]
        table[name] = dict(k, traffic_bytes=k["dram_read_bytes"] + k["dram_write_bytes"], report=os.path.basename(path),
                           launches_in_report=len(launches))
        print(name, table[name])
    with open(OUT, "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
