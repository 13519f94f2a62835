"""
TEST INFRASTRUCTURE (see oracle/__init__.py).  Fock-space side of the oracle.

Restates /root/reference/qat/fermion/util.py:238-281 (``init_creation_ops``, ``dag``) and
/root/reference/qat/fermion/hamiltonians.py:389-425 (``FermionHamiltonian.get_matrix``): the
operator is built directly from creation / annihilation matrices, with no spin transform
involved, which makes it the independent check of the three fermion->spin encodings.

Conventions (util.py:251,267-269): orbital i <-> index bit ``2**(n-1-i)``; |1> = occupied;
c_i^dagger |j> = (-1)^{number of occupied orbitals among 0..i-1} |j + 2^{n-1-i}> when i is empty.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def creation_operators(n: int, sparse: bool = False) -> dict:
    """{i: c_i^dagger} as (2^n, 2^n) matrices."""
    dim = 2**n
    states = np.arange(dim, dtype=np.int64)
    ops = {}
    for i in range(n):
        occ_bit = 1 << (n - 1 - i)
        empty = states[(states & occ_bit) == 0]
        # occupation of orbitals 0..i-1 = the i most significant bits of the n-bit index
        ahead = empty >> (n - i)
        sign = 1 - 2 * (np.bitwise_count(ahead.astype(np.uint64)).astype(np.int64) & 1)
        m = sp.coo_matrix((sign.astype(float), (empty + occ_bit, empty)), shape=(dim, dim))
        ops[i] = m.tocsr().astype(complex) if sparse else m.toarray().astype(complex)
    return ops


def fermion_get_matrix(n: int, terms, constant=0.0, sparse: bool = False):
    """Matrix of sum_t coeff_t * prod(ladder operators) + constant * I.

    ``terms``: iterable of (coeff, op, qbits) with op letters in {"C", "c"} (C = creation)."""
    dim = 2**n
    cre = creation_operators(n, sparse=sparse)
    ann = {i: (m.conj().T.tocsr() if sparse else m.conj().T) for i, m in cre.items()}
    table = {"C": cre, "c": ann}
    eye = (lambda: sp.identity(dim, dtype=complex, format="csr")) if sparse else (lambda: np.identity(dim, dtype=complex))
    total = None
    for coeff, op, qbits in terms:
        m = eye()
        for letter, q in zip(op, qbits):
            if letter not in table:
                raise TypeError("FermionicTerm only accepts fermionic operators C and c.")
            m = m.dot(table[letter][q])
        total = coeff * m if total is None else total + coeff * m
    ident = eye()
    return constant * ident if total is None else total + constant * ident
