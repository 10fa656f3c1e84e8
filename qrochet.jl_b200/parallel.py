"""Multi-GPU plumbing for the one path that shards: the sliced contraction (examples/distributed.jl:55-101).

The reference spawns one Distributed.jl worker per CPU, deals slices round-robin (`takenth(drop(product(cuttings...),
i-1), nworkers())`, :69 -- intended semantics: slice s on worker s mod W, every slice exactly once) and finishes with
`sum(fetch.(partial_results))` (:101).  Here: one process per GPU, no data-path communication, and ONE sum of a
complex scalar at the end -- through libqrochet_b200's NCCL communicator on GPUs, or through the caller's
torch.distributed group (gloo on CPU: used by the tests for the host-side logic)."""
from __future__ import annotations

import ctypes as C

from . import _capi as capi
from ._capi import check, lib


def my_slices(nslices: int, rank: int, world: int):
    """Slices owned by `rank`: s = rank, rank + world, ... (first cut index fastest inside the slice number)."""
    return range(rank, nslices, world)


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    code = lib.qb200_comm_unique_id(buf)
    if code != 0:
        raise capi.QB200Error(code, "ncclGetUniqueId failed (libnccl not loadable?)")
    return buf.raw


def comm_init(ctx, world: int, rank: int, uid: bytes):
    check(ctx.h, lib.qb200_comm_init(ctx.h, world, rank, C.create_string_buffer(uid, 128)))


def comm_allreduce_sum(ctx, value: complex) -> complex:
    """`sum(fetch.(partial_results))`: NCCL all-reduce of (re, im)."""
    v = (C.c_double * 2)(value.real, value.imag)
    check(ctx.h, lib.qb200_comm_allreduce_sum(ctx.h, v, 2))
    return complex(v[0], v[1])


def torch_allreduce_sum(value: complex, group=None) -> complex:
    """Same reduction through an already initialised torch.distributed process group (any backend)."""
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([value.real, value.imag], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return complex(float(t[0]), float(t[1]))


def contract_sliced_distributed(sc, rank: int, world: int, reducer) -> complex:
    """Each rank contracts its share of the slices of `sc` (a SlicedContraction); `reducer(partial)` sums the
    per-rank partial amplitudes (comm_allreduce_sum / torch_allreduce_sum)."""
    partial = sc.contract(first_slice=rank, stride=world) if rank < sc.nslices else 0j
    return reducer(partial)


def broadcast_mps(ctx, mps, root: int = 0):
    """Replicate a device-resident MPS on every rank with ncclBroadcast over NVLink (`qb200_mps_broadcast`): `mps` is
    the chain on `root` and None elsewhere; every rank gets a B200MPS back (the root its own)."""
    from .mps import B200MPS

    h = C.c_void_p(mps.h.value if mps is not None else None)
    check(ctx.h, lib.qb200_mps_broadcast(ctx.h, C.byref(h), root))
    return mps if mps is not None else B200MPS(ctx, _handle=h)


def expect_batch_distributed(mps, ops, sites, rank: int, world: int, reducer_vec):
    """Batched independent expectation values, the second path of the north star that shards: the MPS is replicated
    on every rank (`broadcast_mps`: one ncclBroadcast of the state), observable i goes to rank i mod W, every rank reuses its own
    left/right environments for its share (`qb200_mps_expect1_batch`), and ONE sum-reduction of a zero-initialised
    vector (2 doubles per observable) gathers the results -- `reducer_vec(list_of_floats) -> list_of_floats`
    (comm_allreduce_sum_vec on GPUs, torch_allreduce_sum_vec for any torch.distributed group)."""
    import numpy as np

    n = len(ops)
    mine = list(range(rank, n, world))
    buf = [0.0] * (2 * n)
    if mine:
        vals = mps.expect([ops[i] for i in mine], [sites[i] for i in mine])
        for i, v in zip(mine, vals):
            buf[2 * i], buf[2 * i + 1] = float(np.real(v)), float(np.imag(v))
    out = reducer_vec(buf)
    return np.array(out[0::2]) + 1j * np.array(out[1::2])


def comm_allreduce_sum_vec(ctx, values):
    """NCCL sum of a short vector of doubles (<= 4096) through libqrochet_b200's communicator."""
    n = len(values)
    v = (C.c_double * n)(*values)
    check(ctx.h, lib.qb200_comm_allreduce_sum(ctx.h, v, n))
    return [float(v[i]) for i in range(n)]


def torch_allreduce_sum_vec(values, group=None):
    import torch
    import torch.distributed as dist

    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor(values, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(x) for x in t.cpu()]
