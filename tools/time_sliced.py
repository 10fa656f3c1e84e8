"""Time-to-amplitude of the sliced circuit-TN contraction as bench.py runs it (slice graph, no profiler):
python tools/time_sliced.py <qubits> <depth> <log2 slice target> [max slices]"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
n, depth, target = int(sys.argv[1]), int(sys.argv[2]), 2 ** int(sys.argv[3])
gates = qb.random_fsim_circuit(n, depth)
arrays, modes = qb.amplitude_network(n, gates)
t0 = time.perf_counter()
sc = qb.SlicedContraction(ctx, arrays, modes, target)
plan_s = time.perf_counter() - t0
nsl = sc.nslices if len(sys.argv) < 5 else min(sc.nslices, int(sys.argv[4]))
stride = sc.nslices // nsl
print("plan %.2f s, nslices %d (timing %d), flops/slice %.3e, invariant %.3e, max intermediate 2^%d" %
      (plan_s, sc.nslices, nsl, sc.flops_per_slice, sc.flops_invariant, np.log2(sc.max_intermediate)))
sc.contract(0, sc.nslices)
for rep in range(3):
    ctx.synchronize(); ctx.timer_begin()
    v = sc.contract(0, stride)
    ms = ctx.timer_end()
    print("rep %d: %.2f ms for %d slices = %.3f ms per slice, %.2f TFLOP/s, amplitude %s" %
          (rep, ms, nsl, ms / nsl, nsl * sc.flops_per_slice / ms / 1e9, v))
