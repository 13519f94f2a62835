"""Packed Pauli-term arrays and their GPU handle.

A packed term is (x, z, c): the operator ``c * i^{popc(x&z)} * X^x Z^z`` with qubit ``q`` of an
``n``-qubit register at mask bit ``n-1-q`` (qubit 0 = most significant index bit, as in
/root/reference/qat/fermion/hamiltonians.py:231-235); ``c`` equals ``Term.coeff`` of the reference.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_ptr, check

_LETTER = {(1, 0): "X", (0, 1): "Z", (1, 1): "Y"}


def pack_terms(nqbits: int, terms):
    """[(coeff, op, qbits)] / Term objects -> (x, z, c) arrays.  Letters multiply left to right, so
    repeated qubits behave like the reference's operator chain (hamiltonians.py:198-204)."""
    T = len(terms)
    xs = np.zeros(T, dtype=np.uint64)
    zs = np.zeros(T, dtype=np.uint64)
    cs = np.zeros(T, dtype=np.complex128)
    for i, t in enumerate(terms):
        coeff, op, qbits = (t.coeff, t.op, t.qbits) if hasattr(t, "coeff") else t
        x = z = 0
        k = 0  # total power of i of the running XZ-form product
        ny = 0
        for letter, q in zip(op, qbits):
            if letter == "I":
                continue
            if letter not in "XYZ":
                raise TypeError("SpinHamiltonian does not support fermionic operators. Please use FermionHamiltonian instead.")
            if not 0 <= q < nqbits:
                raise ValueError(f"qubit index {q} outside 0..{nqbits - 1}")
            b = 1 << (nqbits - 1 - q)
            lx, lz = (b if letter in "XY" else 0), (b if letter in "ZY" else 0)
            k += (1 if letter == "Y" else 0) + 2 * bin(z & lx).count("1")
            x ^= lx
            z ^= lz
        # operator = coeff * i^k X^x Z^z ; stored c satisfies c * i^{popc(x&z)} = coeff * i^k
        ny = bin(x & z).count("1")
        xs[i], zs[i] = x, z
        cs[i] = complex(coeff) * (1j) ** ((k - ny) % 4)
    return xs, zs, cs


def unpack_term(nqbits: int, x: int, z: int):
    """(x, z) -> (op letters on ascending qubits, qubit list)."""
    op, qb = [], []
    for q in range(nqbits):
        b = 1 << (nqbits - 1 - q)
        key = (1 if x & b else 0, 1 if z & b else 0)
        if key != (0, 0):
            op.append(_LETTER[key])
            qb.append(q)
    return "".join(op), qb


class PackedTerms:
    """Owner of a ``qfe_terms`` handle: a canonical term set living in libqfe."""

    def __init__(self, handle: C.c_void_p):
        self._h = handle
        n = C.c_int()
        T = C.c_int64()
        nx = C.c_int64()
        ns = C.c_int64()
        check(_lib.load().qfe_terms_info(self._h, C.byref(n), C.byref(T), C.byref(nx), C.byref(ns)))
        self.nqbits, self.n_terms, self.n_xmasks, self.n_slots = n.value, T.value, nx.value, ns.value
        self._arrays = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().qfe_terms_destroy(h)
            except Exception:
                pass

    @property
    def handle(self):
        return self._h

    def arrays(self):
        """(x, z, c, owner, constant) numpy arrays, canonical order (slot, x, z)."""
        if self._arrays is None:
            T, S = self.n_terms, self.n_slots
            x = np.zeros(max(T, 1), dtype=np.uint64)
            z = np.zeros(max(T, 1), dtype=np.uint64)
            c = np.zeros(max(T, 1), dtype=np.complex128)
            o = np.zeros(max(T, 1), dtype=np.int32)
            k = np.zeros(S, dtype=np.complex128)
            check(
                _lib.load().qfe_terms_export(
                    self._h, as_ptr(x, C.c_uint64), as_ptr(z, C.c_uint64), as_ptr(c.view(np.float64), C.c_double), as_ptr(o, C.c_int32),
                    as_ptr(k.view(np.float64), C.c_double),
                )
            )
            self._arrays = (x[:T], z[:T], c[:T], o[:T], k)
        return self._arrays

    @property
    def constant(self) -> complex:
        return complex(self.arrays()[4][0])
