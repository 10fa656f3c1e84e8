"""Per-node timing of ONE slice of the sliced circuit-TN contraction (QB200_DEBUG_TN=1 prints the lines): run as
QB200_DEBUG_TN=1 python tools/probe_sliced_nodes.py 40 6 24"""
import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
n, depth, target = int(sys.argv[1]), int(sys.argv[2]), 2 ** int(sys.argv[3])
gates = qb.random_fsim_circuit(n, depth)
arrays, modes = qb.amplitude_network(n, gates)
sc = qb.SlicedContraction(ctx, arrays, modes, target)
print("nslices", sc.nslices, "flops/slice %.3e" % sc.flops_per_slice, "max inter 2^%d" % np.log2(sc.max_intermediate))
for rep in range(2):
    print("== rep", rep, file=sys.stderr, flush=True)
    ctx.timer_begin()
    v = sc.contract(0, sc.nslices)  # first slice only
    print("slice", ctx.timer_end(), "ms", v)
