"""Warp-stall samples of an .ncu-rep aggregated by CUDA source line (needs -lineinfo and --import-source on).
usage: python tools/ncu_lines.py report.ncu-rep [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname = ""
out = []
tot = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 7 and r[0] not in ("", "Line No") and r[2] == "-":
        try:
            n = int(r[6])
        except ValueError:
            continue
        tot += n
        out.append((n, fname, r[0], r[1].strip(), int(r[7] or 0), r[17], r[18]))
print("total samples", tot)
for n, f, l, s, ex, conf, exc in sorted(out, reverse=True)[:top]:
    print(f"{n:6d} {100.0 * n / max(tot, 1):5.1f}%  inst {ex:9d}  smem-excess-wavefronts {exc:>8s}  {f}:{l}  {s[:110]}")
