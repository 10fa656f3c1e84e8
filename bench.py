#!/usr/bin/env python
"""bench.py -- TEBD sweeps/s (n=64, chi=1024, ComplexF64) on B200, BASELINE.json's headline metric.

One "step" = one TEBD sweep: 32 odd-bond + 31 even-bond `evolve!(psi, G; maxdim=chi, iscanonical=true,
renormalize=true)` calls on a Vidal-form MPS (BASELINE.json configs[3]).  The sweep is sequential along
the chain, so at N > 1 GPUs the TEBD line is N independent replicas ("replicas only"); the two paths
that genuinely shard -- the sliced circuit-TN contraction with one NCCL sum, and batched independent expectation values
on an MPS replicated by ncclBroadcast -- are reported in the same JSON line under "sliced_contraction" (40 qubits, depth
6: the round-1 workload), "sliced_contraction_depth7" (one layer deeper: 1024 slices, the workload whose 1 -> 8 GPU curve
is meaningful) and "expect_batch", each with its own in-run check.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--sites 64] [--bond-dim 1024]

`--impl reference` times the CPU oracle (the restated reference path, NumPy/SciPy -> OpenBLAS zgesdd/zgemm; the Julia
reference itself cannot run in this image) on the box's host cores.  Each of its steps is ONE `evolve!` on a bulk bond
(theta 2048 x 2048) of a canonized `rand` MPS -- a bounded sample of the sweep, scaled to sweeps/s by the algorithmic
flop share of that bond; with `--steps 1 --warmup 0` it runs one real full 63-bond sweep instead.
That arm imports nothing of the product package.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "TEBD sweeps/s (n=64, chi=1024, ComplexF64)"
SIG_TOL = 1e-12
POOL_DMMA_PEAK_TFLOPS = 37.06   # best FP64 DMMA issue rate measured on this pool's B200s (profiles/r1b_peaks.txt)


# ------------------------------------------------------------------------------------------------------
# workload definition shared by both arms (no product import: the reference arm must not map the product .so)
def bond_dims(n, chi, p=2):
    """Bond dimensions of `rand(Chain, Open, State)` (Chain.jl:230-236): bond b (1-based) has min(chi, p^b, p^(n-b))."""
    return [min(chi, p ** b, p ** (n - b)) for b in range(1, n)]


def sweep_bonds(n):
    """odd bonds then even bonds, 1-based left site of each bond (a user-level loop)."""
    return list(range(1, n, 2)) + list(range(2, n, 2))


def haar_matrix(layer, bond):
    """Haar-random 4 x 4 unitary of (layer, bond): QR of complex Ginibre from default_rng(2000 + 64 layer + bond), phases
    fixed.  Reshaped column-major to (o1, o2, i1, i2) it is the reference's gate array."""
    rng = np.random.default_rng(2000 + layer * 64 + bond)
    z = (rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))) / np.sqrt(2)
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def gate_for(layer, bond):
    return np.reshape(haar_matrix(layer, bond), (2, 2, 2, 2), order="F")


def bond_flops(cl, cb, cr):
    """Algorithmic flops of one evolve!: theta GEMM 8MNK, gate, thin SVD 4(14 m n^2 + 8 n^3)."""
    m_, n_ = sorted((2 * cl, 2 * cr), reverse=True)
    return 8.0 * (2 * cl) * (2 * cr) * cb + 8.0 * 16 * cl * cr + 4.0 * (14.0 * m_ * n_ * n_ + 8.0 * n_ ** 3)


def svd_flops(cl, cr):
    m_, n_ = sorted((2 * cl, 2 * cr), reverse=True)
    return 4.0 * (14.0 * m_ * n_ * n_ + 8.0 * n_ ** 3)


def sweep_flops(n, chi):
    d = [1] + bond_dims(n, chi) + [1]
    return sum(bond_flops(d[b - 1], d[b], d[b + 1]) for b in range(1, n))


def sweep_svd_flops(n, chi):
    d = [1] + bond_dims(n, chi) + [1]
    return sum(svd_flops(d[b - 1], d[b + 1]) for b in range(1, n))


def config_dict(n, chi):
    """The `config` object: identical in both arms (the driver compares them)."""
    return {"workload": f"TEBD sweep n={n} chi={chi} ComplexF64: {n - 1} evolve! calls (odd then even bonds), "
                        f"maxdim={chi}, iscanonical, renormalize, Vidal-form rand MPS (BASELINE configs[3])",
            "sites": n, "bond_dim": chi, "eltype": "ComplexF64",
            "gates": "Haar-random two-site unitaries, default_rng(2000 + 64 layer + bond)",
            "algorithmic_tflop_per_step": sweep_flops(n, chi) / 1e12,
            "l2": "inputs larger than L2: the MPS is 1451 MiB resident in HBM at chi=1024 and every bulk bond touches "
                  ">= 192 MiB (theta, X, B0), against 126 MB of L2"}


def set_blas_threads():
    try:  # torchrun exports OMP_NUM_THREADS=1: give OpenBLAS every host core back (BLAS threads = core count)
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []
        # sampling period; QB200_BENCH_CLOCK_MS overrides (0: no sampling), see the note where the sampler is started
        self.period_ms = int(os.environ.get("QB200_BENCH_CLOCK_MS", "500"))

    def start(self):
        if self.period_ms <= 0:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# CPU legs (the oracle: only place bench.py touches oracle/)
def oracle_bulk_chain(sites_lor, lams):
    """4-site Vidal chain around one bond for the oracle: `sites_lor` = the two site tensors (l, o, r) of the bond,
    `lams` = (Λ_left, Λ_bond, Λ_right).  The two outer sites are fillers (evolve! on sites (2, 3) never reads them)."""
    from oracle import chain as oc

    gl, gr = sites_lor
    cl, cr = gl.shape[0], gr.shape[2]
    edge_l = np.ones((1, 2, cl)) / np.sqrt(2.0 * cl)
    edge_r = np.ones((cr, 2, 1)) / np.sqrt(2.0 * cr)
    return oc.chain_from_vidal([edge_l.astype(complex), gl, gr, edge_r.astype(complex)], list(lams))


def cpu_bulk_bond_on_device_state(psi, probe, n, chi, layer, bonds):
    """cpu_baseline + parity_check in one: for each bond b in `bonds` (bulk bonds of the benchmark state `psi`, HBM
    resident), the oracle runs evolve!(...; maxdim, iscanonical, renormalize) on the CPU from the device's own
    Γ_b, Γ_{b+1}, Λ_{b-1}, Λ_b, Λ_{b+1} (timed: the CPU baseline), the device runs the same call on `probe` (a copy),
    and the new Schmidt vectors are compared."""
    from oracle import chain as oc

    set_blas_threads()
    times, worst, kept_eq = [], 0.0, True
    for b in bonds:
        lams = psi.lambdas()
        q = oracle_bulk_chain((psi.site(b - 1), psi.site(b)), (lams[b - 2], lams[b - 1], lams[b]))
        u = haar_matrix(layer, b)
        t0 = time.perf_counter()
        q.evolve(oc.gate(u, [2, 3]), iscanonical=True, maxdim=chi, renormalize=True)
        times.append(time.perf_counter() - t0)
        want = q.lambdas()[1]
        kept, _ = probe.evolve(gate_for(layer, b), [b, b + 1], maxdim=chi, iscanonical=True, renormalize=True)
        got = probe.lambdas()[b - 1]
        kept_eq = kept_eq and (kept == len(want) == len(got))
        k = min(len(want), len(got))
        worst = max(worst, float(np.abs(got[:k] - want[:k]).max() / want[0]))
    return times, {"what": "after the timed region: evolve! on bulk bonds of the benchmark state, device vs CPU oracle "
                           "(LAPACK zgesdd) on the same Γ/Λ and gate",
                   "bonds": list(bonds), "kept_equal": bool(kept_eq), "max_dsigma_over_sigma1": worst,
                   "tol": SIG_TOL, "ok": bool(kept_eq and worst <= SIG_TOL)}


def reference_state(n_small, chi, seed):
    """The reference arm's state: `rand` MPS + canonize! on the CPU oracle, on the shortest chain whose middle bonds
    are bulk bonds of the benchmark (dims chi, chi, chi): same distribution, same bulk shape as the n=64 state."""
    from oracle import chain as oc

    o = oc.Chain(oc.rand_mps_arrays(np.random.default_rng(seed), n_small, chi, fast=True))
    return o.canonize()


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path on the host cores (imports nothing of the product)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import chain as oc

    set_blas_threads()
    n, chi, cores = args.n, args.chi, os.cpu_count()
    full = (args.steps == 1 and args.warmup == 0) or args.full_sweep
    d = [1] + bond_dims(n, chi) + [1]
    t_setup = time.perf_counter()
    if full:
        o = reference_state(n, chi, 1000 + 4)
        bulk = []
        equiv = 1.0
        sample = (f"one REAL full sweep per step: {n - 1} oracle evolve! calls (odd then even bonds) on the canonized "
                  f"n={n} chi={chi} rand MPS")
    else:
        # shortest chain with three bulk bonds (all three bond dims = chi)
        lg = int(np.ceil(np.log2(chi)))
        ns = 2 * lg + 4
        o = reference_state(ns, chi, 1000 + 4)
        ds = [1] + bond_dims(ns, chi) + [1]
        bulk = [b for b in range(1, ns) if ds[b - 1] == ds[b] == ds[b + 1] == chi]
        equiv = sweep_flops(n, chi) / bond_flops(chi, chi, chi)
        sample = (f"each step = ONE oracle evolve! (maxdim={chi}, iscanonical, renormalize) on a bulk bond (theta "
                  f"{2 * chi} x {2 * chi}) of a canonized n={ns} chi={chi} rand MPS, cycling over bonds {bulk}; a sweep "
                  f"is {equiv:.2f} such bonds by algorithmic flops (SURVEY §8d), value = 1 / (step time x {equiv:.2f})")
    setup_s = time.perf_counter() - t_setup
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if full:
            for b in sweep_bonds(n):
                o.evolve(oc.gate(haar_matrix(it, b), [b, b + 1]), iscanonical=True, maxdim=chi, renormalize=True)
        else:
            b = bulk[it % len(bulk)]
            o.evolve(oc.gate(haar_matrix(it, b), [b, b + 1]), iscanonical=True, maxdim=chi, renormalize=True)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    sec = float(np.mean(times))
    val = 1.0 / (sec * equiv)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "sweeps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64 arithmetic)", "data": "synthetic",
            "config": config_dict(n, chi),
            "detail": {"note": "Julia/Tenet cannot run in this image: CPU oracle (NumPy/SciPy -> OpenBLAS zgesdd, zgemm), "
                               "the restated reference path",
                       "steps_per_sweep_equivalent": equiv, "setup_s": setup_s,
                       "step_s": [round(t, 3) for t in times]},
            "cpu_baseline": {"value": val, "unit": "sweeps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
def ncu_traffic():
    """DRAM bytes per launch of the Jacobi kernels from the committed `ncu --set full` capture, as written by
    tools/ncu_traffic.py (profiles/r2_ncu_traffic.json).  None when no capture has been committed."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def median_peak(ctx, reps=5):
    vals = [ctx.dmma_peak_tflops() for _ in range(reps)]
    return float(np.median(vals)), [round(v, 2) for v in vals]


def run_b200(args):
    import torch
    import torch.distributed as dist
    import qrochet_b200 as qb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = qb.Context(local)
    n, chi = args.n, args.chi

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---- untimed set-up: rand MPS (Chain.jl:223-256 restated) -> device -> canonize! (Vidal form) ----
    t0 = time.perf_counter()
    arrays = qb.rand_mps_arrays(np.random.default_rng(1000 + 4), n, chi)
    psi = qb.B200MPS(ctx, arrays)
    del arrays
    psi.canonize()
    ctx.synchronize()
    setup_s = time.perf_counter() - t0
    bonds = sweep_bonds(n)
    mps_bytes = sum(int(np.prod(psi.site_dims(s))) * 16 for s in range(n))

    def sweep(state, layer):
        """One TEBD sweep = the 63 evolve! calls in program order (odd bonds, then even bonds), issued as ONE gate
        list: an update starts when the earlier updates on its two sites are done (same results as the loop)."""
        kept, dw = state.evolve_circuit([gate_for(layer, b) for b in bonds], bonds, maxdim=chi, iscanonical=True,
                                        renormalize=True)
        return sum(kept), sum(dw)

    layer = 0
    for _ in range(args.warmup):
        sweep(psi, layer)
        layer += 1
    peak_before, peak_before_runs = median_peak(ctx)

    # ---- timed region: K sweeps, state resident in HBM, NO profiler ----
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = ctx.launches
    svd0 = ctx.svd_totals()
    step_wall_ms = []
    ctx.timer_begin()
    for _ in range(args.steps):
        t_s = time.perf_counter()
        sweep(psi, layer)  # returns the kept counts: the call ends with every worker stream drained
        step_wall_ms.append((time.perf_counter() - t_s) * 1e3)
        layer += 1
    ms = ctx.timer_end()
    launches = ctx.launches - l0
    svd1 = ctx.svd_totals()
    barrier()
    clocks = sampler.stop()
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = world * 1e3 / ms_per_step
    peak_after, peak_after_runs = median_peak(ctx)
    # The micro-benchmark under-reads on some boxes of the pool (29.7 against 37.0 TFLOP/s at identical clocks, while the
    # sweep itself runs at the same speed: BENCH_r01 vs SCALE_r01, gpurun_out/r2multi): the denominator is never taken
    # below the pool's best measured value, so that a slow reading cannot inflate the fraction.
    peak_tf = max(peak_before, peak_after, POOL_DMMA_PEAK_TFLOPS)
    peak_stable = abs(peak_before - peak_after) <= 0.05 * max(peak_before, peak_after)
    if not peak_stable:
        print(f"[bench] WARNING: DMMA peak micro-benchmark unstable: {peak_before:.2f} before vs {peak_after:.2f} "
              f"TFLOP/s after the timed region", file=sys.stderr, flush=True)
    norm_after = psi.norm()

    # ---- e2e: the same sweep through the public API with HOST (pinned) buffers: upload, sweep, download ----
    host_sites = []
    for s in range(n):
        d = psi.site_dims(s)
        tbuf = torch.empty(int(np.prod(d)), dtype=torch.complex128).pin_memory()
        a = tbuf.numpy().reshape(d, order="F")
        psi.site_into(s, a)
        host_sites.append((tbuf, a))
    host_lams = psi.lambdas()
    e2e_steps = max(3, min(args.steps, 5))

    def e2e_step(layer):
        st = qb.B200MPS.from_sites(ctx, [a for _, a in host_sites], host_lams, form=1)  # H2D from pinned memory
        sweep(st, layer)
        for s in range(n):  # D2H of the result into pinned memory (bond dims are stationary at maxdim)
            d = st.site_dims(s)
            if tuple(d) == host_sites[s][1].shape:
                st.site_into(s, host_sites[s][1])
            else:
                st.site(s)
        lams = st.lambdas()
        del st
        return lams

    host_lams = e2e_step(layer)  # untimed warm-up of the host path (memory pool, pinned staging)
    layer += 1
    barrier()
    ctx.timer_begin()
    for _ in range(e2e_steps):
        host_lams = e2e_step(layer)
        layer += 1
    e2e_ms = max_over_ranks(ctx.timer_end())
    e2e_value = world * 1e3 / (e2e_ms / e2e_steps)
    lam_bytes = sum(0 if l is None else l.size * 8 for l in host_lams)
    gate_bytes = len(bonds) * 256
    del host_sites

    # ---- after the timed region: one profiled sweep (all worker streams) and one profiled bulk bond (single stream)
    #      give the per-kernel table; the headline roofline is §8(d)'s algorithmic sweep count over the timed region ----
    ctx.profile(True)
    ctx.profile_read()
    sweep(psi, layer)
    layer += 1
    prof_sweep = ctx.profile_read()
    d = [1] + psi.bond_dims() + [1]
    bulk = [b for b in range(1, n) if d[b - 1] == d[b] == d[b + 1] == chi]
    prof_bond = None
    if bulk:
        probe = psi.copy()
        ctx.profile_read()
        probe.evolve(gate_for(layer, bulk[len(bulk) // 2]), [bulk[len(bulk) // 2], bulk[len(bulk) // 2] + 1],
                     maxdim=chi, iscanonical=True, renormalize=True)
        prof_bond = ctx.profile_read()
        del probe
    ctx.profile(False)

    def table(prof, scale=1.0):
        out = {}
        for k, (cnt, pms, work) in prof.items():
            if cnt:
                out[k] = {"launches_or_calls": cnt, "ms": round(pms * scale, 3),
                          "algorithmic_tflops": round(work / (pms * 1e-3) / 1e12, 2) if pms > 0 and work > 0 else None}
        return out

    alg_tf = sweep_flops(n, chi) / 1e12
    ach = alg_tf / (ms_per_step * 1e-3)
    traffic = ncu_traffic()
    roof = {"bound": "tensor", "unit": "TFLOP/s",
            "kernel": "Jacobi SVD chain of evolve! (jacobi_update / jacobi_gram / jacobi_evd + QR preconditioner): "
                      f"{sweep_svd_flops(n, chi) / sweep_flops(n, chi):.3f} of a sweep's algorithmic flops are its 63 SVDs",
            "achieved": ach, "peak": peak_tf, "frac": ach / peak_tf,
            "achieved_note": "ALGORITHMIC flops of one sweep (SURVEY §8d: theta GEMM 8MNK + gate + thin SVD 4(14mn^2+8n^3) "
                             "per bond on the true bond profile = %.2f TFLOP) / measured ms_per_step of the timed region; "
                             "independent of the Jacobi sweep count actually executed" % alg_tf,
            "peak_source": "FP64 DMMA m8n8k4 issue-bound micro-benchmark of libqrochet_b200_diag.so, median of 5 runs "
                           "before and after the timed region; the larger of the two and of the pool's best measured value "
                           "(37.06 TFLOP/s, profiles/r1b_peaks.txt) so that a slow reading cannot inflate the fraction "
                           "(MEASURED_PEAKS.json has no FP64 figure)",
            "peak_before_after": [peak_before, peak_after], "peak_runs": [peak_before_runs, peak_after_runs],
            "peak_stable": bool(peak_stable),
            "traffic": next((v.get("dram_bytes_per_launch") for k, v in (traffic or {}).items()
                             if k.startswith("jacobi_update_kernel")), None),
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of jacobi_update_kernel from the "
                            "committed ncu --set full capture (profiles/r2_ncu_traffic.json, written by "
                            "tools/ncu_traffic.py); algorithmic: 2 x 32 MiB of X + 2 MiB of W per launch",
            "jacobi_sweeps_per_svd": (svd1[1] - svd0[1]) / max(1, svd1[0] - svd0[0]),
            "kernels_one_sweep_all_streams": table(prof_sweep),
            "kernels_one_bulk_bond_single_stream": table(prof_bond) if prof_bond else None}
    if prof_bond and prof_bond["svd"][0]:
        _, svd_ms, svd_work = prof_bond["svd"]
        roof["svd_single_stream"] = {"flops": svd_work, "ms": svd_ms,
                                     "achieved_tflops": svd_work / (svd_ms * 1e-3) / 1e12,
                                     "frac_of_peak": svd_work / (svd_ms * 1e-3) / 1e12 / peak_tf}

    line = {"metric": METRIC, "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128 (f64 arithmetic)", "data": "synthetic",
            "config": config_dict(n, chi),
            "detail": {"parallelism": "replicas only (sequential sweep)" if world > 1 else "single GPU",
                       "mps_mib": mps_bytes / 2 ** 20, "setup_s": setup_s,
                       "step_wall_ms": [round(x, 1) for x in step_wall_ms],
                       "norm_after": norm_after,
                       "norm_note": "renormalize=true normalises each new Schmidt vector (Chain.jl:653-654), not the "
                                    "state: truncating a bond leaves its neighbours non-canonical, so |psi| drifts "
                                    "from 1 exactly as in the reference (tests/test_gpu_configs.py compares it with "
                                    "the oracle's)"},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "sweeps/s", "h2d_bytes_per_step": mps_bytes + lam_bytes + gate_bytes,
                    "d2h_bytes_per_step": mps_bytes + lam_bytes, "steps": e2e_steps},
            "roofline": roof}

    # ---- parity check + CPU baseline on the benchmark state itself (rank 0, N = 1) ----
    if rank == 0 and world == 1 and bulk and not args.no_cpu_baseline:
        probe = psi.copy()
        pick = [bulk[len(bulk) // 4], bulk[len(bulk) // 2], bulk[3 * len(bulk) // 4]]
        pick = sorted(set(pick))
        times, parity = cpu_bulk_bond_on_device_state(psi, probe, n, chi, layer, pick)
        del probe
        equiv = sweep_flops(n, chi) / bond_flops(chi, chi, chi)
        sec = float(np.mean(times))
        line["parity_check"] = parity
        line["cpu_baseline"] = {"value": 1.0 / (sec * equiv), "unit": "sweeps/s", "cores": os.cpu_count(),
                                "kind": "port",
                                "sample": f"oracle evolve! (NumPy/SciPy -> OpenBLAS zgesdd) on {len(pick)} bulk bonds "
                                          f"{pick} of the benchmark state downloaded from HBM, {sec:.2f} s each; a sweep "
                                          f"is {equiv:.2f} bulk bonds by algorithmic flops"}
        if not parity["ok"]:
            print(f"[bench] PARITY CHECK FAILED: {parity}", file=sys.stderr, flush=True)
    del psi

    if not args.no_sliced:
        try:
            # reference value: the same amplitude with the on-GPU slice target (2^28 elements = 4 GiB per intermediate
            # instead of the 2^24 of examples/distributed.jl:46): the round-2 tree needs no cut at all there
            big, amp_ref = run_sliced(ctx, qb, 0, world=1, peak_tf=peak_tf, barrier=lambda: ctx.synchronize(),
                                      max_over_ranks=lambda x: x, target=2 ** 28, reps=2)
            line["sliced_contraction"], _ = run_sliced(ctx, qb, rank, world, peak_tf, barrier, max_over_ranks,
                                                       check_against=amp_ref)
            if world == 1:
                line["sliced_contraction_large_target"] = big
                line["sliced_contraction_round1_planner"], _ = run_sliced(ctx, qb, rank, world, peak_tf, barrier,
                                                                          max_over_ranks, reps=1, optimizer=0,
                                                                          check_against=amp_ref)
                line["sliced_contraction_time_model_planner"], _ = run_sliced(ctx, qb, rank, world, peak_tf, barrier,
                                                                              max_over_ranks, optimizer=2,
                                                                              check_against=amp_ref)
        except Exception as e:  # never lose the headline line to the extra measurement
            line["sliced_contraction"] = {"error": str(e)[:300]}
        try:
            # one layer deeper: 57x the work of the depth-6 network with the round-2 planner (1024 slices at the 2^24
            # target), i.e. enough slices per GPU for the 1 -> 8 GPU curve to mean something; checked against the same
            # amplitude from the 64-slice plan of the 2^28 target
            big7, amp7 = run_sliced(ctx, qb, 0, world=1, peak_tf=peak_tf, barrier=lambda: ctx.synchronize(),
                                    max_over_ranks=lambda x: x, depth=7, target=2 ** 28, reps=1)
            line["sliced_contraction_depth7"], _ = run_sliced(ctx, qb, rank, world, peak_tf, barrier, max_over_ranks,
                                                              depth=7, reps=2, check_against=amp7)
            line["sliced_contraction_depth7"]["amplitude_check"]["against"] = (
                "the same amplitude from the 64-slice plan of the 2^28 slice target, contracted on one GPU")
            if world == 1:
                line["sliced_contraction_depth7_large_target"] = big7
        except Exception as e:
            line["sliced_contraction_depth7"] = {"error": str(e)[:300]}
    if not args.no_expect:
        try:
            line["expect_batch"] = run_expect_batch(ctx, qb, rank, world, n, chi, barrier, max_over_ranks)
        except Exception as e:
            line["expect_batch"] = {"error": str(e)[:300]}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def ensure_comm(ctx, qb, rank, world):
    """libqrochet_b200's own NCCL communicator (one per process), bootstrapped through torch.distributed."""
    import torch.distributed as dist

    if world > 1 and not getattr(ctx, "_comm_ready", False):
        uid = [qb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        qb.comm_init(ctx, world, rank, uid[0])
        ctx._comm_ready = True


def random_product_bra(n, seed=4000):
    """n normalised random local vectors (the bra of the benchmark amplitude)."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    return list(v / np.linalg.norm(v, axis=1, keepdims=True))


def run_sliced(ctx, qb, rank, world, peak_tf, barrier, max_over_ranks, qubits=40, depth=6, target=2 ** 24, reps=3,
               optimizer=1, check_against=None):
    """Second half of BASELINE.json's metric: sliced contraction of the <b|U|0..0> network of a 40-qubit, depth-6
    random FSim circuit (examples/distributed.jl:11-53 pattern; slice target 2^24 elements, :46) with a seeded random
    product bra b (FSim gates leave |0..0> alone, so <0|U|0> = 1 would exercise no phases).  Slices are dealt s mod W
    to the ranks (no data-path communication), each rank accumulates on its device, ONE NCCL sum (:101).
    optimizer: 1 = the round-2 planner, 0 = the round-1 single greedy tree (the "before" column)."""
    gates = qb.random_fsim_circuit(qubits, depth)
    arrays, modes = qb.amplitude_network(qubits, gates, bra=random_product_bra(qubits))
    t0 = time.perf_counter()
    sc = qb.SlicedContraction(ctx, arrays, modes, target, optimizer=optimizer)
    plan_s = time.perf_counter() - t0
    ensure_comm(ctx, qb, rank, world)
    reducer = (lambda v: qb.comm_allreduce_sum(ctx, v)) if world > 1 else (lambda v: v)
    # warm-up: one slice per rank (builds the offset tables, the arena and the slice graph)
    sc.contract(first_slice=rank % sc.nslices, stride=sc.nslices)
    reducer(0j)
    best, amp = None, None
    for _ in range(reps):
        barrier()
        ctx.timer_begin()
        amp = qb.contract_sliced_distributed(sc, rank, world, reducer)
        ms = max_over_ranks(ctx.timer_end())
        best = ms if best is None else min(best, ms)
    flops = sc.nslices * sc.flops_per_slice + sc.flops_invariant * min(world, sc.nslices)
    tf = flops / (best * 1e-3) / 1e12
    out = {"metric": "sliced TN contraction TFLOP/s", "value": tf, "unit": "TFLOP/s", "n_gpus": world,
           "ms": best, "time_to_amplitude_ms": best, "scaling": "strong",
           "config": {"workload": f"{qubits}-qubit depth-{depth} random FSim circuit amplitude <b|U|0..0>, random product "
                                  f"bra (seed 4000), slice target 2^{int(np.log2(target))} elements (BASELINE configs[4])",
                      "planner": {0: "round 1: one greedy tree",
                                  1: "round 2: ContractSimplification + multi-start greedy + sub-tree reconfiguration "
                                     "(objective: flops)",
                                  2: "round 2, objective = time model max(macs, 10 x elements moved) per node"}[optimizer],
                      "plan_s": plan_s, "nslices": sc.nslices, "cut_indices": len(sc.sliced_modes),
                      "flops_per_slice": sc.flops_per_slice, "flops_slice_invariant": sc.flops_invariant,
                      "total_flops": sc.nslices * sc.flops_per_slice + sc.flops_invariant,
                      "max_intermediate_elements": sc.max_intermediate,
                      "flop_count": "8 x complex MACs (EinExprs flops x 8): per-slice nodes x slices + slice-invariant "
                                    "nodes once per rank that holds a slice",
                      "executor": "one CUDA graph launch per slice (leaf gather + tree replay + cursor advance)"},
           "amplitude": [amp.real, amp.imag], "frac_of_dmma_peak": tf / (peak_tf * world),
           "collective": "one ncclAllReduce(sum) of 2 doubles" if world > 1 else "none"}
    if check_against is not None:
        err = abs(amp - check_against) / abs(check_against)
        out["amplitude_check"] = {"against": "the same amplitude contracted UNSLICED on one GPU (slice target 2^28)",
                                  "rel_err": err, "tol": 1e-10, "ok": bool(err <= 1e-10)}
    return out, amp


def run_expect_batch(ctx, qb, rank, world, n, chi, barrier, max_over_ranks, nobs=None):
    """Batched independent expectation values (north star's second sharding path; reference semantics Chain.jl:724-735):
    rank 0 holds the chi-bond MPS, ONE ncclBroadcast replicates it (timed separately), observable i goes to rank
    i mod W, every rank sweeps its own environments once and evaluates its share, ONE allreduce gathers the values."""
    nobs = nobs or 2 * n
    ensure_comm(ctx, qb, rank, world)
    psi = None
    if rank == 0:
        psi = qb.B200MPS(ctx, qb.rand_mps_arrays(np.random.default_rng(1000 + 5), n, chi)).canonize()
    mps_bytes = 0
    bcast_ms = 0.0
    if world > 1:
        barrier()
        ctx.timer_begin()
        psi = qb.broadcast_mps(ctx, psi, 0)
        bcast_ms = max_over_ranks(ctx.timer_end())
    mps_bytes = sum(int(np.prod(psi.site_dims(s))) * 16 for s in range(n))
    rng = np.random.default_rng(77)
    paulis = [np.array([[0, 1], [1, 0]], complex), np.array([[0, -1j], [1j, 0]]), np.diag([1.0, -1.0]).astype(complex)]
    ops = [paulis[int(rng.integers(3))] for _ in range(nobs)]
    sites = [1 + (i * 7) % n for i in range(nobs)]
    reducer = (lambda v: qb.comm_allreduce_sum_vec(ctx, v)) if world > 1 else (lambda v: v)
    qb.expect_batch_distributed(psi, ops, sites, rank, world, reducer)  # warm-up
    best, vals = None, None
    for _ in range(2):
        barrier()
        ctx.timer_begin()
        vals = qb.expect_batch_distributed(psi, ops, sites, rank, world, reducer)
        ms = max_over_ranks(ctx.timer_end())
        best = ms if best is None else min(best, ms)
    return {"metric": "batched single-site expectation values / s", "value": nobs / (best * 1e-3), "unit": "observables/s",
            "n_gpus": world, "ms": best, "scaling": "strong", "observables": nobs,
            "config": {"workload": f"{nobs} independent expect(psi, [O_s]) on one canonized n={n} chi={chi} MPS "
                                   f"(Chain.jl:724-735), Pauli observables, sites 1 + 7i mod n"},
            "mps_broadcast": {"bytes": mps_bytes, "ms": bcast_ms,
                              "gb_per_s": (mps_bytes / 1e9) / (bcast_ms * 1e-3) if bcast_ms > 0 else None,
                              "how": "qb200_mps_broadcast: ncclBroadcast of every site tensor over NVLink" if world > 1
                              else "single GPU: nothing to replicate"},
            "checksum": [float(np.sum(vals).real), float(np.sum(vals).imag)],
            "collective": f"one ncclAllReduce(sum) of {2 * nobs} doubles" if world > 1 else "none"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sites", dest="n", type=int, default=64)
    ap.add_argument("--bond-dim", dest="chi", type=int, default=1024)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sliced", action="store_true", help="skip the sliced circuit-TN contraction part")
    ap.add_argument("--no-expect", action="store_true", help="skip the batched expectation-value part")
    ap.add_argument("--full-sweep", action="store_true", help="reference arm: real full sweeps instead of the sample")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
