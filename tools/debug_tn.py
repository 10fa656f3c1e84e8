import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
gates = qb.random_fsim_circuit(40, 6)
arrays, modes = qb.amplitude_network(40, gates)
sc = qb.SlicedContraction(ctx, arrays, modes, 2 ** 24)
sc.contract(0, sc.nslices)      # warm: tables, invariants
sc.contract(1, sc.nslices)
