"""`Product` ansatz (/root/reference/src/Ansatz/Product.jl): one small tensor per lane and no bonds.  Everything here is
O(n p) host arithmetic -- in the Julia package it stays in Julia too -- and the way onto the device is
`to_chain(ctx)` = `convert(Chain, ::Product)` (Chain.jl:174-183): a bond-dimension-1 B200MPS whose `overlap` with a
chain is the reference's `overlap(::Product, ::Chain)` (Chain.jl:751-752)."""
from __future__ import annotations

import numpy as np


class Product:
    """`Product(arrays)`: vectors -> State (Product.jl:24-33), matrices with array dims (output, input) -> Operator
    (:35-46)."""

    def __init__(self, arrays):
        self.arrays = [np.array(a) for a in arrays]
        nd = {a.ndim for a in self.arrays}
        if nd == {1}:
            self.socket = "state"
        elif nd == {2}:
            self.socket = "operator"
        else:
            raise TypeError("Product takes a list of vectors (State) or of matrices (Operator)")

    @classmethod
    def zeros(cls, n: int, p: int = 2, dtype=bool) -> "Product":
        """`zeros(Product, n; p, eltype)` (Product.jl:48-50): |0...0>."""
        v = np.zeros(p, dtype=dtype)
        v[0] = 1
        return cls([v.copy() for _ in range(n)])

    @classmethod
    def ones(cls, n: int, p: int = 2, dtype=bool) -> "Product":
        """`ones(Product, n; p, eltype)` (Product.jl:52-58): |1...1>."""
        v = np.zeros(p, dtype=dtype)
        v[1] = 1
        return cls([v.copy() for _ in range(n)])

    def nlanes(self) -> int:
        return len(self.arrays)

    def copy(self) -> "Product":
        return Product([a.copy() for a in self.arrays])

    def norm(self, p: float = 2):
        """Product.jl:60-65 exactly as written: (prod_i ||t_i||_p)^(1/p) (for p = 2 the square root of the product of
        the site norms: a quirk of the reference, exact for normalised sites -- which is what `normalize!` produces
        and the reference tests)."""
        prod = 1.0
        for a in self.arrays:
            prod *= np.linalg.norm(np.ravel(a), p)
        return prod ** (1.0 / p)

    def opnorm(self, p: float = 2):
        """Product.jl:67-72 (operators)."""
        if self.socket != "operator":
            raise TypeError("opnorm needs an Operator")
        prod = 1.0
        for a in self.arrays:
            prod *= np.linalg.norm(a, p)
        return prod ** (1.0 / p)

    def normalize_(self, p: float = 2) -> "Product":
        """`normalize!` (Product.jl:74-80)."""
        self.arrays = [a / np.linalg.norm(np.ravel(a), p) for a in self.arrays]
        return self

    def overlap(self, other: "Product"):
        """`overlap(a::Product, b::Product)` (Product.jl:82-90): prod_i dot(a_i, conj(b_i)), Julia's `dot` conjugating
        its first argument."""
        if self.socket != "state" or other.socket != "state":
            raise TypeError("overlap needs two States")
        assert self.nlanes() == other.nlanes(), "Ansatzes must have the same sites"
        out = 1.0 + 0.0j
        for a, b in zip(self.arrays, other.arrays):
            out *= np.vdot(a, np.conj(b))
        return out

    def to_chain(self, ctx):
        """`convert(Chain, ::Product)` (Chain.jl:174-183) onto the device: a bond-dimension-1 chain."""
        from .mps import B200MPS
        if self.socket != "state":
            raise TypeError("only a State converts to an MPS")
        return B200MPS.from_product(ctx, self.arrays)
