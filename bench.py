#!/usr/bin/env python
"""bench.py -- TEBD sweeps/s (n=64, chi=1024, ComplexF64) on B200, BASELINE.json's headline metric.

One "step" = one TEBD sweep: 32 odd-bond + 31 even-bond `evolve!(psi, G; maxdim=chi, iscanonical=true,
renormalize=true)` calls on a Vidal-form MPS (BASELINE.json configs[3]).  The sweep is
sequential along the chain, so at N > 1 GPUs the TEBD line is N independent replicas ("replicas only",
DESIGN.md §multi-GPU); the path that genuinely shards -- the sliced circuit-TN contraction with one NCCL sum --
is reported in the same JSON line under "sliced_contraction".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--sites 64] [--bond-dim 1024]

`--impl reference` times the CPU oracle (the restated reference path, NumPy/SciPy -> OpenBLAS zgesdd/zgemm;
the Julia reference itself cannot run in this image) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from collections import Counter

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "TEBD sweeps/s (n=64, chi=1024, ComplexF64)"


# ------------------------------------------------------------------------------------------------------
def sweep_bonds(n):
    """odd bonds then even bonds, 1-based left site of each bond (a user-level loop)."""
    return list(range(1, n, 2)) + list(range(2, n, 2))


def gate_for(layer, bond):
    import qrochet_b200 as qb
    return qb.haar_gate(np.random.default_rng(2000 + layer * 64 + bond))


def workload_name(n, chi):
    return (f"TEBD sweep n={n} chi={chi} ComplexF64: {n - 1} evolve! calls (odd then even bonds), maxdim={chi}, "
            f"renormalize, Vidal form (BASELINE configs[3])")


def sweep_flops(n, chi):
    """Algorithmic flops of one sweep on the true bond profile: theta GEMM 8MNK, gate, thin SVD
    4(14 m n^2 + 8 n^3)."""
    import qrochet_b200 as qb
    d = [1] + qb.bond_dims(n, chi) + [1]
    total = 0.0
    for b in range(1, n):
        cl, cb, cr = d[b - 1], d[b], d[b + 1]
        m_, n_ = sorted((2 * cl, 2 * cr), reverse=True)
        total += 8.0 * (2 * cl) * (2 * cr) * cb + 8.0 * 16 * cl * cr + 4.0 * (14.0 * m_ * n_ * n_ + 8.0 * n_ ** 3)
    return total


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "500"],
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
def cpu_tebd_sample(n, chi, bulk_reps):
    """CPU oracle (restated reference path) on a bounded sample of the workload: one `evolve!` per distinct
    bond shape (chi_l, chi_b, chi_r) of the sweep -- the bulk shape `bulk_reps` times -- extrapolated to the 63
    bonds with the true bond profile.  Returns (sweep seconds, description)."""
    from oracle import chain as oc
    from oracle.tenet import Tensor
    import qrochet_b200 as qb

    try:  # torchrun exports OMP_NUM_THREADS=1: give OpenBLAS every host core back (BLAS threads = core count)
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    d = [1] + qb.bond_dims(n, chi) + [1]
    classes = Counter((d[b - 1], d[b], d[b + 1]) for b in range(1, n))
    rng = np.random.default_rng(4242)
    total, detail = 0.0, []
    for (cl, cb, cr), count in sorted(classes.items()):
        def rnd(*s):
            return (rng.standard_normal(s) + 1j * rng.standard_normal(s)) / np.sqrt(s[-1])
        arrays = [rnd(2, cl), rnd(2, cl, cb), rnd(2, cb, cr), rnd(2, cr)]
        q = oc.Chain(arrays)
        for k, dim in zip((1, 2, 3), (cl, cb, cr)):
            lam = np.sort(rng.random(dim))[::-1] + 0.1
            q.tn.push(Tensor(lam / np.linalg.norm(lam), [q.bond_ind(oc.site(k), oc.site(k + 1))]))
        reps = bulk_reps if (cl, cb, cr) == (chi, chi, chi) else 1
        best = None
        for r in range(reps):
            qq = q.copy()
            g = oc.gate(oc.haar_unitary(rng), [2, 3])
            t0 = time.perf_counter()
            qq.evolve(g, iscanonical=True, maxdim=chi, renormalize=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        total += best * count
        detail.append((cl, cb, cr, count, best))
    bulk = [x for x in detail if x[:3] == (chi, chi, chi)]
    desc = (f"oracle evolve! timed once per distinct bond shape ({len(classes)} shapes, bulk "
            f"{chi}^3 x{bulk_reps} best-of, {bulk[0][4]:.2f} s each) and summed over the {n - 1} bonds of the true "
            f"bond profile" if bulk else f"oracle evolve! once per distinct bond shape ({len(classes)} shapes)")
    return total, desc


def run_reference(args):
    """--impl reference: the CPU restatement of the reference path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    times = []
    desc = ""
    for it in range(args.warmup + args.steps):
        t, desc = cpu_tebd_sample(args.n, args.chi, 1)
        if it >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    val = 1.0 / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "sweeps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "c128 (f64 arithmetic)", "data": "synthetic",
            "config": {"workload": workload_name(args.n, args.chi),
                       "note": "Julia/Tenet cannot run in this image: CPU oracle (NumPy/SciPy -> OpenBLAS zgesdd, "
                               "zgemm), each step a bounded sample extrapolated to the full sweep"},
            "cpu_baseline": {"value": val, "unit": "sweeps/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
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

    odd, even = list(range(1, n, 2)), list(range(2, n, 2))

    def sweep(state, layer):
        """One TEBD sweep = the 63 evolve! calls in program order (odd bonds, then even bonds), issued as ONE gate
        list: an update starts when the earlier updates on its two sites are done (same results as the loop)."""
        order = odd + even
        kept, dw = state.evolve_circuit([gate_for(layer, b) for b in order], order, maxdim=chi, renormalize=True)
        return sum(kept), sum(dw)

    layer = 0
    for _ in range(args.warmup):
        sweep(psi, layer)
        layer += 1
    peak_tf = ctx.dmma_peak_tflops()

    # ---- timed region: K sweeps, state resident in HBM ----
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ctx.profile(True)
    l0 = ctx.launches
    svd0 = ctx.svd_totals()
    step_wall_ms = []
    ctx.timer_begin()
    for _ in range(args.steps):
        t_s = time.perf_counter()
        sweep(psi, layer)  # returns the kept counts: the call ends with the stream drained
        step_wall_ms.append((time.perf_counter() - t_s) * 1e3)
        layer += 1
    ms = ctx.timer_end()
    launches = ctx.launches - l0
    svd1 = ctx.svd_totals()
    prof = ctx.profile_read()
    ctx.profile(False)
    barrier()
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * 1e3 / ms_per_step
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
    e2e_steps = 1

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
    e2e_ms = ctx.timer_end()
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = world * 1e3 / (e2e_ms / e2e_steps)
    lam_bytes = sum(0 if l is None else l.size * 8 for l in host_lams)
    gate_bytes = len(bonds) * 256

    # ---- roofline of the dominant kernel (Jacobi update: X_p <- X_p W_p, complex GEMM 8MNK per pair).  Inside the
    # timed region up to 8 bond updates run concurrently, so event pairs on one stream also see the other streams'
    # kernels; the per-launch duration is therefore measured live right after the timed region on ONE bulk bond
    # update issued on a single stream (same kernels, same shapes, CUDA events on the launching stream). ----
    roof = None
    bulk = [b for b in bonds if psi.site_dims(b - 1)[0] == chi and psi.site_dims(b)[2] == chi]
    if bulk:
        probe = psi.copy()
        ctx.profile(True)
        ctx.profile_read()
        probe.evolve(gate_for(layer, bulk[len(bulk) // 2]), [bulk[len(bulk) // 2]] * 1 + [bulk[len(bulk) // 2] + 1],
                     maxdim=chi, renormalize=True)
        pp = ctx.profile_read()
        ctx.profile(False)
        del probe
        cnt, pms, work = pp["jacobi_update"]
        tot_ms = sum(v[1] for k, v in pp.items() if k != "svd")
        if cnt:
            ach = work / (pms * 1e-3) / 1e12
            svd_cnt, svd_ms, svd_work = pp["svd"]
            roof = {"bound": "tensor", "kernel": "jacobi_update_kernel (FP64 DMMA)", "achieved": ach, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": ach / peak_tf,
                    "traffic": 78.5e6, "traffic_note": "dram read (69.26 MB) + write (9.28 MB) bytes per launch from the "
                                                      "ncu --set full capture profiles/r1b_ncu_jacobi_summary.txt "
                                                      "(algorithmic: 2 x 32 MiB of X + 2 MiB of W per launch; most "
                                                      "of the written X stays in the 126 MB L2)",
                    "flop_note": "achieved = ALGORITHMIC flops (8 per complex multiply-add) / time.  The kernel uses the "
                                 "3M complex product: it executes 6 DMMA flops + 3/32 FP64 adds per complex "
                                 "multiply-add, so the executed-DMMA fraction of the pipe is 0.75 x frac (ncu: tensor "
                                 "pipe 65.8 % of elapsed) and frac can exceed 1 only above 4/3",
                    "peak_source": "measured here: DMMA m8n8k4 issue-bound micro-benchmark (qb200_bench_dmma_peak); "
                                   "MEASURED_PEAKS.json has no FP64 figure",
                    "launches": cnt, "avg_launch_ms": pms / cnt, "share_of_step": pms / tot_ms,
                    "share_note": "share of the kernel time of one bulk bond update run alone on one stream",
                    "phases_ms_one_bulk_bond": {k: v[1] for k, v in pp.items() if v[0]},
                    "phases_ms_per_step_all_streams": {k: v[1] / args.steps for k, v in prof.items() if v[0]},
                    "svd_algorithmic": {"flops": svd_work, "ms": svd_ms,
                                        "achieved_tflops": svd_work / (svd_ms * 1e-3) / 1e12 if svd_ms else None,
                                        "frac_of_peak": (svd_work / (svd_ms * 1e-3) / 1e12) / peak_tf if svd_ms else None}}

    line = {"metric": METRIC, "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "c128 (f64 arithmetic)", "data": "synthetic",
            "config": {"workload": workload_name(n, chi),
                       "parallelism": "replicas only (sequential sweep)" if world > 1 else "single GPU",
                       "l2": f"inputs larger than L2: MPS {mps_bytes / 2**20:.0f} MiB resident in HBM, "
                             f"each bulk bond touches >= 192 MiB",
                       "algorithmic_tflop_per_step": sweep_flops(n, chi) / 1e12, "setup_s": setup_s,
                       "step_wall_ms": [round(x, 1) for x in step_wall_ms],
                       "jacobi_sweeps_per_svd": (svd1[1] - svd0[1]) / max(1, svd1[0] - svd0[0]),
                       "norm_after": norm_after},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "sweeps/s", "h2d_bytes_per_step": mps_bytes + lam_bytes + gate_bytes,
                    "d2h_bytes_per_step": mps_bytes + lam_bytes, "steps": e2e_steps},
            "roofline": roof}

    if not args.no_sliced:
        line["sliced_contraction"] = run_sliced(ctx, qb, rank, world, peak_tf, barrier)
        if world == 1:
            # the same network with the slice target an on-GPU budget allows (2^28 elements = 4 GiB per intermediate
            # instead of the 2^24 of examples/distributed.jl:46): fewer cuts, larger GEMMs.  Reported separately.
            try:
                line["sliced_contraction_large_target"] = run_sliced(ctx, qb, rank, world, peak_tf, barrier,
                                                                     target=2 ** 28, reps=2)
            except Exception as e:  # never lose the headline line to the extra measurement
                line["sliced_contraction_large_target"] = {"error": str(e)[:200]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sec, desc = cpu_tebd_sample(n, chi, 3 if chi >= 512 else 1)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "sweeps/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": desc}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_sliced(ctx, qb, rank, world, peak_tf, barrier, qubits=40, depth=6, target=2 ** 24, reps=3):
    """Second half of BASELINE.json's metric: sliced contraction of the <0..0|U|0..0> network of a 40-qubit, depth-6
    random FSim circuit (examples/distributed.jl:11-53 pattern; slice target 2^24 elements, :46).  Slices are dealt
    s mod W to the ranks (no data-path communication), each rank accumulates on its device, ONE NCCL sum (:101)."""
    import torch.distributed as dist

    gates = qb.random_fsim_circuit(qubits, depth)
    arrays, modes = qb.amplitude_network(qubits, gates)
    sc = qb.SlicedContraction(ctx, arrays, modes, target)
    if world > 1:
        uid = [qb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        qb.comm_init(ctx, world, rank, uid[0])
        reducer = lambda v: qb.comm_allreduce_sum(ctx, v)  # noqa: E731
    else:
        reducer = lambda v: v  # noqa: E731
    # warm-up: one slice per rank (builds the offset tables, contracts the slice-invariant sub-trees once)
    sc.contract(first_slice=rank % sc.nslices, stride=sc.nslices)
    reducer(0j)
    best, amp = None, None
    for _ in range(reps):
        barrier()
        ctx.timer_begin()
        amp = qb.contract_sliced_distributed(sc, rank, world, reducer)
        ms = ctx.timer_end()
        if world > 1:
            import torch
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        best = ms if best is None else min(best, ms)
    flops = sc.nslices * sc.flops_per_slice
    tf = flops / (best * 1e-3) / 1e12
    return {"metric": "sliced TN contraction TFLOP/s", "value": tf, "unit": "TFLOP/s", "n_gpus": world,
            "ms": best, "scaling": "strong",
            "config": {"workload": f"{qubits}-qubit depth-{depth} random FSim circuit amplitude <0|U|0>, greedy path, "
                                   f"slice target 2^{int(np.log2(target))} elements (BASELINE configs[4])",
                       "nslices": sc.nslices, "cut_indices": len(sc.sliced_modes),
                       "flops_per_slice": sc.flops_per_slice, "max_intermediate_elements": sc.max_intermediate,
                       "flop_count": "8 x complex MACs over all tree nodes x slices (EinExprs flops x 8)"},
            "amplitude": [amp.real, amp.imag], "frac_of_dmma_peak": tf / (peak_tf * world),
            "collective": "one ncclAllReduce(sum) of 2 doubles" if world > 1 else "none"}


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
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
