"""qrochet_b200: B200-native backend for Qrochet.jl's tensor-network hot path (host-side Python mirror of the
Julia package extension; all numerics are hand-written sm_100a CUDA behind libqrochet_b200.so)."""
from . import _capi
from ._capi import MissingSchmidtCoefficientsException, QB200Error
from .device import (Context, DeviceArray, conj, contract, norm2, permute, qr, scale, scale_mode, select_mode,
                     slice_mode, svd)
from . import chain
from . import gates
from .gates import Gate
from .mps import B200MPS
from .product import Product
from .tn import SlicedContraction, amplitude_network, circuit_network, fsim, random_fsim_circuit
from .parallel import (broadcast_mps, comm_allreduce_sum, comm_allreduce_sum_vec, comm_init, comm_unique_id,
                       contract_sliced_distributed, expect_batch_distributed, my_slices, torch_allreduce_sum,
                       torch_allreduce_sum_vec)
from .rand import bond_dims, haar_gate, heisenberg_mpo_arrays, rand_mpo_arrays, rand_mps_arrays

__all__ = ["Context", "DeviceArray", "B200MPS", "contract", "scale_mode", "slice_mode", "select_mode", "conj",
           "permute", "norm2", "scale", "qr", "svd", "QB200Error", "MissingSchmidtCoefficientsException"]
