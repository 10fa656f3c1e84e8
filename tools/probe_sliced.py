import sys, time
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
n, depth, target = int(sys.argv[1]), int(sys.argv[2]), 2 ** int(sys.argv[3])
gates = qb.random_fsim_circuit(n, depth)
arrays, modes = qb.amplitude_network(n, gates)
sc = qb.SlicedContraction(ctx, arrays, modes, target)
print("nslices", sc.nslices, "flops/slice %.3e" % sc.flops_per_slice, "max inter 2^%d" % np.log2(sc.max_intermediate))
ctx.profile(True)
t0 = time.time(); ctx.timer_begin()
v = sc.contract(0, sc.nslices)  # first slice only
ms = ctx.timer_end(); print("first slice", ms, "ms", v, "TF/s", sc.flops_per_slice / ms / 1e9)
ctx.timer_begin()
v = sc.contract(1, max(sc.nslices // 4, 1)) if sc.nslices > 1 else 0
ms = ctx.timer_end(); print("4 slices", ms, "ms per slice", ms / 4, "TF/s", 4 * sc.flops_per_slice / ms / 1e9)
