"""
TEST INFRASTRUCTURE (see oracle/__init__.py).  Spin-operator side of the oracle.

Two things live here:

1. ``PauliSum`` -- dictionary algebra over packed Pauli strings, restating the published
   behaviour of the closed-source ``qat.core.Observable`` arithmetic the reference leans on
   (call sites: /root/reference/qat/fermion/transforms.py:128,130,181,183,253,255 and
   qat/fermion/hamiltonians.py:110-135).  A string is a key ``(x, z)`` of two Python ints and
   stands for ``X^x Z^z`` (Z applied first); qubit ``q`` of an ``n``-qubit register is mask bit
   ``n-1-q`` because the reference makes qubit 0 the most significant index bit
   (qat/fermion/hamiltonians.py:231-235).  ``Y = i X Z``.

2. ``spin_get_matrix`` -- the reference's matrix builder, ``SpinHamiltonian.get_matrix``
   (qat/fermion/hamiltonians.py:163-250): every term is a chain of products of
   ``kron(I, sigma, I)`` operators, summed with its coefficient, plus ``constant * I``.
"""

from __future__ import annotations

import numpy as np
import scipy.sparse as sp

# Pauli matrices with the reference's sign convention for Y (qat/fermion/hamiltonians.py:20-25)
_SIGMA = {
    "I": np.array([[1, 0], [0, 1]], dtype=complex),
    "X": np.array([[0, 1], [1, 0]], dtype=complex),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=complex),
    "Z": np.array([[1, 0], [0, -1]], dtype=complex),
}
_I_POW = (1, 1j, -1, -1j)


def bit(n: int, q: int) -> int:
    """Mask bit of qubit ``q`` in an ``n``-qubit register (qubit 0 = MSB)."""
    return 1 << (n - 1 - q)


def mask(n: int, qubits) -> int:
    m = 0
    for q in qubits:
        m |= bit(n, q)
    return m


def popcount(v: int) -> int:
    return bin(v).count("1")


def string_to_packed(n: int, op: str, qbits) -> tuple[int, int, complex]:
    """``Term(1, op, qbits)`` -> (x, z, phase) with the operator equal to phase * X^x Z^z.

    Letters are multiplied left to right, so a repeated qubit is handled like the reference's
    matrix chain does (qat/fermion/hamiltonians.py:201-204)."""
    x = z = 0
    phase = 0  # power of i
    for letter, q in zip(op, qbits):
        b = bit(n, q)
        if letter == "I":
            continue
        if letter == "X":
            lx, lz, lp = b, 0, 0
        elif letter == "Z":
            lx, lz, lp = 0, b, 0
        elif letter == "Y":
            lx, lz, lp = b, b, 1
        else:
            raise TypeError(f"'{letter}' is not a Pauli letter")
        # (X^x Z^z)(X^lx Z^lz) = (-1)^{|z & lx|} X^{x^lx} Z^{z^lz}
        phase = (phase + lp + 2 * popcount(z & lx)) % 4
        x ^= lx
        z ^= lz
    return x, z, _I_POW[phase]


def packed_to_string(n: int, x: int, z: int) -> tuple[str, list[int], complex]:
    """(x, z) -> (op letters on ascending qubits, qubits, factor) with X^x Z^z = factor * letters.

    X Z = -i Y, hence factor = (-i)^{#Y}."""
    op = []
    qb = []
    ny = 0
    for q in range(n):
        b = bit(n, q)
        hx, hz = bool(x & b), bool(z & b)
        if hx and hz:
            op.append("Y")
            ny += 1
        elif hx:
            op.append("X")
        elif hz:
            op.append("Z")
        else:
            continue
        qb.append(q)
    return "".join(op), qb, (-1j) ** ny


class PauliSum:
    """sum_k c_k X^{x_k} Z^{z_k} on n qubits; the identity key (0, 0) carries the constant."""

    __slots__ = ("n", "d")

    def __init__(self, n: int, d: dict | None = None):
        self.n = n
        self.d = {} if d is None else d

    # -- construction -------------------------------------------------------------------------
    @classmethod
    def from_terms(cls, n: int, terms, constant=0.0) -> "PauliSum":
        """``terms`` = iterable of (coeff, op, qbits), the reference's Term triple."""
        out = cls(n)
        for coeff, op, qbits in terms:
            x, z, ph = string_to_packed(n, op, qbits)
            out.d[(x, z)] = out.d.get((x, z), 0.0) + coeff * ph
        if constant != 0:
            out.d[(0, 0)] = out.d.get((0, 0), 0.0) + constant
        return out

    def copy(self) -> "PauliSum":
        return PauliSum(self.n, dict(self.d))

    # -- algebra -------------------------------------------------------------------------------
    def __add__(self, other: "PauliSum") -> "PauliSum":
        out = self.copy()
        out += other
        return out

    def __iadd__(self, other: "PauliSum") -> "PauliSum":
        for k, c in other.d.items():
            self.d[k] = self.d.get(k, 0.0) + c
        return self

    def __sub__(self, other: "PauliSum") -> "PauliSum":
        return self + other.scale(-1.0)

    def scale(self, s) -> "PauliSum":
        return PauliSum(self.n, {k: s * c for k, c in self.d.items()})

    def __mul__(self, other):
        if not isinstance(other, PauliSum):
            return self.scale(other)
        out: dict = {}
        for (x1, z1), c1 in self.d.items():
            for (x2, z2), c2 in other.d.items():
                sgn = -1.0 if popcount(z1 & x2) & 1 else 1.0
                k = (x1 ^ x2, z1 ^ z2)
                out[k] = out.get(k, 0.0) + sgn * c1 * c2
        return PauliSum(self.n, out)

    __rmul__ = __mul__

    def commutator(self, other: "PauliSum") -> "PauliSum":
        """``A | B`` of the reference (qat/fermion/hamiltonians.py:134-135)."""
        return self * other - other * self

    def dag(self) -> "PauliSum":
        # (c X^x Z^z)^dagger = conj(c) Z^z X^x = conj(c) (-1)^{|x&z|} X^x Z^z
        return PauliSum(
            self.n, {(x, z): np.conj(c) * (-1.0 if popcount(x & z) & 1 else 1.0) for (x, z), c in self.d.items()}
        )

    # -- canonical views -----------------------------------------------------------------------
    def pruned(self, eps: float = 0.0) -> "PauliSum":
        return PauliSum(self.n, {k: c for k, c in self.d.items() if abs(c) > eps})

    @property
    def constant(self):
        return self.d.get((0, 0), 0.0)

    def canonical(self, eps: float = 1e-12):
        """Ascending (x, z) list of non-identity strings with |c| > eps, Y-form coefficients
        (what ``Term.coeff`` would hold), plus the constant.  This is the comparison format."""
        keys = sorted(k for k, c in self.d.items() if k != (0, 0) and abs(c) > eps)
        xs = np.array([k[0] for k in keys], dtype=np.uint64)
        zs = np.array([k[1] for k in keys], dtype=np.uint64)
        cs = np.array(
            [self.d[k] * (-1j) ** (popcount(k[0] & k[1]) % 4) for k in keys], dtype=np.complex128
        )
        return xs, zs, cs, complex(self.constant)

    def term_triples(self, eps: float = 1e-12):
        """[(coeff, op, qbits)] in the Term convention (Y-form coefficient), ascending (x, z)."""
        out = []
        for (x, z) in sorted(self.d):
            c = self.d[(x, z)]
            if (x, z) == (0, 0) or abs(c) <= eps:
                continue
            op, qb, fac = packed_to_string(self.n, x, z)
            out.append((c * fac, op, qb))
        return out


# ---------------------------------------------------------------------------------------------
# matrix builder -- restates SpinHamiltonian.get_matrix / _make_spin_op
# ---------------------------------------------------------------------------------------------
def _one_qubit_operator(letter: str, q: int, n: int, sparse: bool):
    """kron(I_{2^q}, sigma, I_{2^{n-q-1}})  (qat/fermion/hamiltonians.py:213-250)."""
    if sparse:
        return sp.kron(sp.identity(2**q, dtype=complex), sp.kron(sp.csr_matrix(_SIGMA[letter]), sp.identity(2 ** (n - q - 1), dtype=complex)), format="csr")
    return np.kron(np.identity(2**q, dtype=complex), np.kron(_SIGMA[letter], np.identity(2 ** (n - q - 1), dtype=complex)))


def spin_get_matrix(n: int, terms, constant=0.0, sparse: bool = False):
    """Matrix of sum_t coeff_t * prod(letters) + constant * I, built the way the reference
    builds it (qat/fermion/hamiltonians.py:163-211): a chain of full-size operator products
    per term.  ``terms`` is an iterable of (coeff, op, qbits)."""
    dim = 2**n
    eye = (lambda: sp.identity(dim, dtype=complex, format="csr")) if sparse else (lambda: np.identity(dim, dtype=complex))
    cache: dict = {}
    total = None
    for coeff, op, qbits in terms:
        m = eye()
        for letter, q in zip(op, qbits):
            if letter == "I":
                continue
            if letter not in _SIGMA:
                raise TypeError("SpinHamiltonian does not support fermionic operators.")
            key = (letter, q)
            if key not in cache:
                cache[key] = _one_qubit_operator(letter, q, n, sparse)
            m = m.dot(cache[key])
        total = coeff * m if total is None else total + coeff * m
    ident = eye()
    total = constant * ident if total is None else total + constant * ident
    return total


# ---------------------------------------------------------------------------------------------
# vectorised matrix-free action of packed terms (validated against spin_get_matrix in the tests);
# used as the checker at sizes where the kron chain is too slow
# ---------------------------------------------------------------------------------------------
def _parity(v: np.ndarray) -> np.ndarray:
    return (np.bitwise_count(v) & 1).astype(np.int8)


def apply_packed(n: int, xs, zs, cs_yform, constant, psi: np.ndarray) -> np.ndarray:
    """H|psi> for H = constant + sum_t c_t * i^{|x&z|} X^x Z^z  (c_t = Y-form / Term coefficient).

    (H psi)[b ^ x] += c_XZ * (-1)^{|z & b|} * psi[b]"""
    dim = 2**n
    idx = np.arange(dim, dtype=np.uint64)
    out = constant * psi.astype(np.complex128)
    psi = psi.astype(np.complex128)
    for x, z, c in zip(xs, zs, cs_yform):
        x = int(x)
        z = int(z)
        cxz = complex(c) * (1j) ** (popcount(x & z) % 4)
        sgn = 1.0 - 2.0 * _parity(idx & np.uint64(z))
        # out[b ^ x] += cxz * sgn[b] * psi[b]   <=>   out[b'] += cxz * sgn[b'^x] * psi[b'^x]
        src = idx ^ np.uint64(x)
        out += cxz * (sgn[src] * psi[src])
    return out


def expectation_packed(n: int, xs, zs, cs_yform, constant, psi: np.ndarray, bra: np.ndarray | None = None) -> complex:
    """<bra|H|psi> (bra = psi when omitted)."""
    phi = apply_packed(n, xs, zs, cs_yform, constant, psi)
    return complex(np.vdot(psi if bra is None else bra.astype(np.complex128), phi))


def csr_from_packed(n: int, xs, zs, cs_yform, constant) -> sp.csr_matrix:
    """Sparse matrix of a packed term list, assembled directly (one COO block per term).
    Equal to ``spin_get_matrix(..., sparse=True)`` (checked in the tests) but O(T 2^n) memory-light."""
    dim = 2**n
    idx = np.arange(dim, dtype=np.int64)
    acc = sp.identity(dim, dtype=complex, format="csr") * constant
    rows, cols, vals = [], [], []
    pending = 0
    for x, z, c in zip(xs, zs, cs_yform):
        x = int(x)
        z = int(z)
        cxz = complex(c) * (1j) ** (popcount(x & z) % 4)
        sgn = 1.0 - 2.0 * _parity(idx.astype(np.uint64) & np.uint64(z))
        rows.append(idx ^ x)
        cols.append(idx)
        vals.append(cxz * sgn)
        pending += dim
        if pending >= (1 << 24):
            acc = acc + sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(dim, dim)).tocsr()
            rows, cols, vals, pending = [], [], [], 0
    if rows:
        acc = acc + sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(dim, dim)).tocsr()
    return acc
