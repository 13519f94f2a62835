"""
TEST INFRASTRUCTURE (see oracle/__init__.py).  Fermion -> spin encodings of the oracle.

Restates /root/reference/qat/fermion/transforms.py:
  * :19-79   Fenwick tree and the P / C / U qubit sets of the Bravyi-Kitaev encoding
             (the reference uses ``anytree`` nodes; here the tree is a parent array),
  * :82-257  ``transform_to_{jw,parity,bk}_basis`` -- every ladder operator becomes a two-string
             spin operator ``1/2 A_j -+ i/2 B_j`` and a fermionic term becomes the ordered product
             of those, summed over terms (the products/sums are ``qat.core`` algebra in the
             reference; here they are oracle.spin.PauliSum dictionary algebra),
  * :260-407 code matrices, ``recode_integer``, ``change_encoding`` (the reference's own test
             criterion, tests/unitary/test_unitary_transforms.py:44-52).
"""

from __future__ import annotations

import numpy as np

from .spin import PauliSum, bit, mask

ENCODINGS = ("jordan-wigner", "bravyi-kitaev", "parity")


# ---------------------------------------------------------------------------------------------
# Fenwick tree (transforms.py:19-32): fenwick(l, r): m = floor((l+r)/2); parent[m] = r; recurse
# on (l, m) and (m+1, r); root is n-1.
# ---------------------------------------------------------------------------------------------
def fenwick_parents(n: int) -> list[int]:
    parent = [-1] * n

    def split(lo: int, hi: int) -> None:
        if lo == hi:
            return
        mid = (lo + hi) // 2
        parent[mid] = hi
        split(lo, mid)
        split(mid + 1, hi)

    if n > 0:
        split(0, n - 1)
    return parent


def pcu_sets(n: int) -> list[tuple[set, set, set]]:
    """[(P_j, C_j, U_j)] (transforms.py:35-79).
    U = ancestors, F = children, C = children of ancestors with index < j, P = C | F."""
    parent = fenwick_parents(n)
    children: list[set] = [set() for _ in range(n)]
    for k, p in enumerate(parent):
        if p >= 0:
            children[p].add(k)
    out = []
    for j in range(n):
        anc = set()
        a = parent[j]
        while a >= 0:
            anc.add(a)
            a = parent[a]
        c_set = {k for a in anc for k in children[a] if k < j}
        out.append((c_set | children[j], c_set, anc))
    return out


# ---------------------------------------------------------------------------------------------
# ladder operators as packed strings
# ---------------------------------------------------------------------------------------------
def ladder_strings(n: int, encoding: str) -> list[tuple[tuple[int, int, complex], tuple[int, int, complex]]]:
    """For every mode j: (A_j, B_j) as (x, z, factor) with a_j = 1/2 A_j + i/2 B_j and
    a_j^dagger = 1/2 A_j - i/2 B_j.  ``factor`` makes (x, z) the XZ-form of the Pauli string
    (a Y contributes i)."""
    out = []
    if encoding == "jordan-wigner":  # transforms.py:116-126
        for j in range(n):
            zs = mask(n, range(j))
            out.append(((bit(n, j), zs, 1.0), (bit(n, j), zs | bit(n, j), 1j)))
    elif encoding == "parity":  # transforms.py:169-179
        for j in range(n):
            xs = mask(n, range(j, n))
            za = bit(n, j - 1) if j > 0 else 0
            out.append(((xs, za, 1.0), (xs, bit(n, j), 1j)))
    elif encoding == "bravyi-kitaev":  # transforms.py:227-251
        sets = pcu_sets(n)
        for j in range(n):
            p_set, c_set, u_set = sets[j]
            xs = bit(n, j) | mask(n, u_set)
            out.append(((xs, mask(n, p_set), 1.0), (xs, mask(n, c_set) | bit(n, j), 1j)))
    else:
        raise KeyError(encoding)
    return out


def transform(n: int, fermion_terms, constant, encoding: str) -> PauliSum:
    """Spin image of sum_t coeff_t prod(ladder ops) + constant (transforms.py:82-257).

    ``fermion_terms``: iterable of (coeff, op, qbits), letters "C"/"c"."""
    lad = ladder_strings(n, encoding)
    total = PauliSum(n)
    if constant != 0:
        total.d[(0, 0)] = constant
    for coeff, op, qbits in fermion_terms:
        cur = PauliSum(n, {(0, 0): coeff})
        for letter, q in zip(op, qbits):
            (ax, az, af), (bx, bz, bf) = lad[q]
            sign = -1.0 if letter == "C" else 1.0
            mini = PauliSum(n)
            mini.d[(ax, az)] = 0.5 * af
            mini.d[(bx, bz)] = mini.d.get((bx, bz), 0.0) + 0.5j * sign * bf
            cur = cur * mini
        total += cur
    return total


# ---------------------------------------------------------------------------------------------
# code matrices / basis relabelling (transforms.py:260-407)
# ---------------------------------------------------------------------------------------------
def jw_code(n: int) -> np.ndarray:
    return np.identity(n, dtype=int)


def parity_code(n: int) -> np.ndarray:
    return np.triu(np.ones((n, n), dtype=int))


def bk_code(n: int) -> np.ndarray:
    """Column i has ones at rows {i} and all Fenwick descendants of i (transforms.py:324-346)."""
    parent = fenwick_parents(n)
    mat = np.zeros((n, n), dtype=int)
    for j in range(n):
        a = j
        while a >= 0:
            mat[j, a] = 1
            a = parent[a]
    return mat


def code_for(n: int, encoding: str) -> np.ndarray:
    return {"jordan-wigner": jw_code, "parity": parity_code, "bravyi-kitaev": bk_code}[encoding](n)


def recode_integer(integer: int, code: np.ndarray) -> int:
    """p_i = sum_j C_ji f_j mod 2 on MSB-first bit strings (transforms.py:349-379)."""
    n = code.shape[0]
    f = np.array([(integer >> (n - 1 - k)) & 1 for k in range(n)], dtype=int)
    p = code.T.dot(f) % 2
    out = 0
    for k in range(n):
        out = (out << 1) | int(p[k])
    return out


def change_encoding(mat: np.ndarray, code: np.ndarray) -> np.ndarray:
    """B[C(i), C(j)] = A[i, j] (transforms.py:382-407)."""
    n = code.shape[0]
    table = np.array([recode_integer(i, code) for i in range(2**n)], dtype=np.int64)
    out = np.zeros(mat.shape, dtype=mat.dtype)
    out[np.ix_(table, table)] = mat
    return out
