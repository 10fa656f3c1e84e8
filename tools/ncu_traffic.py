"""DRAM traffic per launch of the hot kernels from `ncu --set full` captures -> profiles/r2_ncu_traffic.json, the file
bench.py reads for `roofline.traffic` (so that the figure comes from a committed capture through a tool, not from a
literal in bench.py).

    python tools/ncu_traffic.py gpurun_out/r2_update.ncu-rep [more.ncu-rep ...]

Every kernel found in the reports gets {launches, duration_us (mean), dram_read_bytes / dram_write_bytes /
dram_bytes_per_launch (mean), tensor_pipe_pct_elapsed, source report}."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0,
        "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6, "%": 1.0}


def read(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in body:
        name = r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "").replace("void ", "")

        def val(metric):
            if metric not in col or not r[col[metric]]:
                return None
            return float(r[col[metric]].replace(",", "")) * UNIT.get(units[col[metric]], 1.0)
        d = out.setdefault(name, {"launches": 0, "duration_us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0,
                                  "tensor_pipe_pct_elapsed": 0.0, "source": os.path.basename(rep)})
        d["launches"] += 1
        d["duration_us"] += val("gpu__time_duration.sum") or 0.0
        d["dram_read_bytes"] += val("dram__bytes_read.sum") or 0.0
        d["dram_write_bytes"] += val("dram__bytes_write.sum") or 0.0
        d["tensor_pipe_pct_elapsed"] += val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed") or 0.0
    for d in out.values():
        n = d["launches"]
        for k in ("duration_us", "dram_read_bytes", "dram_write_bytes", "tensor_pipe_pct_elapsed"):
            d[k] /= n
        d["dram_bytes_per_launch"] = d["dram_read_bytes"] + d["dram_write_bytes"]
    return out


if __name__ == "__main__":
    merged = {}
    for rep in sys.argv[1:]:
        merged.update(read(rep))
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    with open(path, "w") as f:
        json.dump(merged, f, indent=1, sort_keys=True)
    for k, d in sorted(merged.items()):
        print(f"{k:36s} x{d['launches']}  {d['duration_us']:9.1f} us  dram {d['dram_bytes_per_launch'] / 1e6:9.2f} MB/launch  "
              f"tensor pipe {d['tensor_pipe_pct_elapsed']:5.1f} %")
    print("wrote", path)
