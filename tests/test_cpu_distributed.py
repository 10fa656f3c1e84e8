"""world_size-2 gloo test (CPU) of the N > 1 host logic of the sliced contraction: slices are dealt s mod W, every
slice is contracted exactly once, and one all-reduce of the complex partial sums gives the amplitude.  The per-rank
partials come from the oracle (this test has no GPU); the dealing and the reduction are the product's code."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import circuit as ocirc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, depth, maxel, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import qrochet_b200 as qb
    gates = qb.random_fsim_circuit(n, depth)
    arrays, modes = qb.amplitude_network(n, gates, ocirc.random_product_state(n, 1), ocirc.random_product_state(n, 2))
    plan = qb.SlicedContraction(None, arrays, modes, maxel)          # planner runs without a GPU
    mine = list(qb.my_slices(plan.nslices, rank, world))
    oplan = ocirc.plan(modes, {x: 2 for m in modes for x in m}, maxel)
    assert plan.sliced_modes == oplan["sliced"]

    class OracleSlices:                                              # stands in for the GPU executor
        nslices = plan.nslices

        def contract(self, first_slice, stride):
            return ocirc.contract_sliced(arrays, modes, oplan, first_slice, stride)[0]

    total = qb.contract_sliced_distributed(OracleSlices(), rank, world, qb.torch_allreduce_sum)
    out[rank] = (total, mine)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slices_dealt_round_robin_and_summed(world):
    n, depth, maxel = 10, 4, 2 ** 5
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, depth, maxel, out), nprocs=world, join=True)
    exact = ocirc.statevector_amplitude(n, ocirc.random_fsim_circuit(n, depth), ocirc.random_product_state(n, 1),
                                        ocirc.random_product_state(n, 2))
    seen = []
    for r in range(world):
        total, mine = out[r]
        assert abs(total - exact) < 1e-12           # every rank holds the full amplitude after the reduce
        seen += mine
    assert sorted(seen) == list(range(len(seen))) and len(seen) > world   # every slice exactly once


def _worker_expect(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import qrochet_b200 as qb
    from oracle import chain as oc
    n = 8
    arrays = oc.rand_mps_arrays(np.random.default_rng(3), n, 8)
    psi = oc.Chain(arrays)
    Z = np.diag([1.0, -1.0]).astype(complex)
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    ops = [Z, X, Z, X, Z, X, Z]
    sites = [1, 2, 3, 4, 5, 6, 8]

    class OracleMPS:                                   # stands in for the device MPS on a box without GPU
        def expect(self, ops_, sites_):
            return np.array([psi.expect([oc.gate(o, [s])]) for o, s in zip(ops_, sites_)])

    vals = qb.expect_batch_distributed(OracleMPS(), ops, sites, rank, world, qb.torch_allreduce_sum_vec)
    want = np.array([psi.expect([oc.gate(o, [s])]) for o, s in zip(ops, sites)])
    out[rank] = float(np.abs(vals - want).max())
    dist.destroy_process_group()


def test_batched_expectation_values_dealt_and_gathered():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_expect, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] < 1e-12 and out[1] < 1e-12
