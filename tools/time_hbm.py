"""HBM-bound helpers (K2/K3/K6/K7/K8 + gate) at the sizes of a chi = 1024 bulk bond: achieved GB/s against the
measured copy bandwidth of MEASURED_PEAKS.json.  Algorithmic bytes = bytes read + bytes written once."""
import json
import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    peak = 6538.3
rng = np.random.default_rng(0)
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
def crand(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
rows = []
def run(name, fn, nbytes, reps=7):
    fn(); ctx.synchronize()
    best = 1e9
    for _ in range(reps):
        ctx.timer_begin(); fn(); best = min(best, ctx.timer_end())
    gbs = nbytes / best / 1e6
    rows.append({"kernel": name, "ms": best, "bytes": nbytes, "GB/s": gbs, "frac_of_copy_peak": gbs / peak})
    print(f"{name:44s} {best:8.4f} ms {gbs:8.1f} GB/s  {gbs / peak:6.1%}", flush=True)
for dt, eb in ((np.complex128, 16), (np.complex64, 8)):
    tag = "c128" if eb == 16 else "c64"
    site = ctx.array(crand(chi, 2, chi).astype(dt))            # (l, o, r) site tensor: 32 MiB in c128
    lam = ctx.array(rng.random(chi) + 0.5)
    n = site.size
    run(f"{tag} scale_mode (mode 0, fastest)", lambda: qb.scale_mode(site, 0, lam, inplace=True), 2 * n * eb)
    run(f"{tag} scale_mode (mode 2, slowest)", lambda: qb.scale_mode(site, 2, lam, inplace=True), 2 * n * eb)
    run(f"{tag} scale_mode pinv (mode 2)", lambda: qb.scale_mode(site, 2, lam, inverse=True, atol=1e-32, inplace=True), 2 * n * eb)
    run(f"{tag} conj", lambda: qb.conj(site), 2 * n * eb)
    run(f"{tag} permute (o,l,r) -> (l,o,r)", lambda: qb.permute(site, (1, 0, 2)), 2 * n * eb)
    run(f"{tag} permute reverse (r,o,l)", lambda: qb.permute(site, (2, 1, 0)), 2 * n * eb)
    run(f"{tag} slice bond {chi} -> {chi // 2} (last mode)", lambda: qb.slice_mode(site, 2, chi // 2), n * eb)
    run(f"{tag} slice bond {chi} -> {chi // 2} (first mode)", lambda: qb.slice_mode(site, 0, chi // 2), n * eb)
    run(f"{tag} select (view ind => 1) mode 1", lambda: qb.select_mode(site, 1, 1), n * eb)
    run(f"{tag} norm2", lambda: qb.norm2(site), n * eb)
    run(f"{tag} copy", lambda: site.copy(), 2 * n * eb)
json.dump({"hbm_copy_peak_gbs": peak, "chi": chi, "rows": rows}, open("gpurun_out/hbm_kernels.json", "w"), indent=1)
