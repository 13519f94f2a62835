"""Fermion -> spin transforms with the reference's names
(/root/reference/qat/fermion/transforms.py:82-257), executed by kernel (1) of libqfe.

``transform_to_jw_basis / transform_to_bk_basis / transform_to_parity_basis`` accept a
``FermionHamiltonian`` or an ``ElectronicStructureHamiltonian`` and return a ``SpinHamiltonian``
whose terms are the canonical packed set produced on the GPU (merged, identity folded into
``constant_coeff``, ascending (x, z); the reference's first-appearance order is not part of the
contract).  An ``ElectronicStructureHamiltonian`` is fed as its integral
arrays, so the n^2 + n^4 Python term objects of the reference (hamiltonians.py:625-648) are never built.

The code-matrix helpers (transforms.py:260-379) are integer bookkeeping on n x n matrices and
stay on the host.
"""

from __future__ import annotations

import numpy as np

from .hamiltonians import TOL, ElectronicStructureHamiltonian, FermionHamiltonian, SpinHamiltonian

DROP_EPS = 1e-12


def _engine():
    from .engine import get_engine

    return get_engine()


def _transform_integrals(h: ElectronicStructureHamiltonian, method: str) -> SpinHamiltonian:
    from ._lib import ENC

    if method not in ENC:
        raise KeyError(method)  # transform_map[method] of the reference (hamiltonians.py:454)
    packed = _engine().terms_from_integrals(h.hpq, h.hpqrs, constant=h.constant_coeff, encoding=method, tol=TOL, drop_eps=DROP_EPS)
    return SpinHamiltonian(h.nbqbits, _packed=packed)


def _transform(h, method: str) -> SpinHamiltonian:
    if isinstance(h, ElectronicStructureHamiltonian):
        return _transform_integrals(h, method)
    if not isinstance(h, FermionHamiltonian):
        raise TypeError("the transforms accept FermionHamiltonian / ElectronicStructureHamiltonian objects")
    packed = _engine().terms_from_fermion(h.nbqbits, h.terms, constant=h.constant_coeff, encoding=method, drop_eps=DROP_EPS)
    return SpinHamiltonian(h.nbqbits, _packed=packed)


def transform_to_jw_basis(fermion_hamiltonian) -> SpinHamiltonian:
    """Jordan-Wigner image (transforms.py:82-132): a_j -> 1/2 Z_0..Z_{j-1} (X_j + i Y_j)."""
    return _transform(fermion_hamiltonian, "jordan-wigner")


def transform_to_parity_basis(fermion_hamiltonian) -> SpinHamiltonian:
    """Parity-basis image (transforms.py:135-185)."""
    return _transform(fermion_hamiltonian, "parity")


def transform_to_bk_basis(fermion_hamiltonian) -> SpinHamiltonian:
    """Bravyi-Kitaev image on the reference's recursive-midpoint Fenwick tree (transforms.py:19-79, 188-257)."""
    return _transform(fermion_hamiltonian, "bravyi-kitaev")


# name -> transform, as TransformObservable.transforms (/root/reference/qat/plugins/transform_observable.py:29-33)
TRANSFORMS = {
    "jordan-wigner": transform_to_jw_basis,
    "bravyi-kitaev": transform_to_bk_basis,
    "parity-basis": transform_to_parity_basis,
}


def get_transform(name: str):
    """ValueError for an unknown name, like TransformObservable.__init__ (transform_observable.py:36-40)."""
    try:
        return TRANSFORMS[name]
    except KeyError as excpt:
        raise ValueError(f"Unknown transformation {name!r}") from excpt


# ---------------------------------------------------------------------------------------------
# code matrices (transforms.py:19-32, 260-379)
# ---------------------------------------------------------------------------------------------
def _fenwick_parents(n: int):
    parent = [-1] * n
    stack = [(0, n - 1)] if n > 0 else []
    while stack:
        lo, hi = stack.pop()
        if lo == hi:
            continue
        mid = (lo + hi) // 2
        parent[mid] = hi
        stack.append((lo, mid))
        stack.append((mid + 1, hi))
    return parent


def _pcu_of_mode(j: int, parent) -> tuple:
    anc = []
    a = parent[j]
    while a >= 0:
        anc.append(a)
        a = parent[a]
    children = {k for k in range(len(parent)) if parent[k] == j}
    c_set = {k for k in range(j) if parent[k] in anc}
    return c_set | children, c_set, set(anc)


def make_PCU_sets(nqbits: int) -> list:
    """[(P, C, U)] qubit sets of every mode 0..nqbits-1, the reference's signature and return value (transforms.py:65-79); the
    Fenwick tree is built once."""
    parent = _fenwick_parents(nqbits)
    return [_pcu_of_mode(j, parent) for j in range(nqbits)]


def get_jw_code(nbits: int) -> np.ndarray:
    return np.identity(nbits, dtype=int)


def get_parity_code(nbits: int) -> np.ndarray:
    return np.triu(np.ones((nbits, nbits), dtype=int))


def get_bk_code(nbits: int) -> np.ndarray:
    parent = _fenwick_parents(nbits)
    mat = np.zeros((nbits, nbits), dtype=int)
    for j in range(nbits):
        a = j
        while a >= 0:
            mat[j, a] = 1
            a = parent[a]
    return mat


def recode_integer(integer: int, code: np.ndarray) -> int:
    """p = C^T f mod 2 on MSB-first bit strings (transforms.py:349-379)."""
    n = code.shape[0]
    f = np.array([(integer >> (n - 1 - k)) & 1 for k in range(n)], dtype=int)
    p = code.T.dot(f) % 2
    out = 0
    for k in range(n):
        out = (out << 1) | int(p[k])
    return out


def change_encoding(mat: np.ndarray, code: np.ndarray) -> np.ndarray:
    """B[C(i), C(j)] = A[i, j] with C = recode_integer(., code): the basis relabelling the reference's own tests use to compare an
    encoded spin Hamiltonian with the Fock-space matrix (/root/reference/qat/fermion/transforms.py:380-407, a 4^n Python loop
    there).  Host-side index permutation of a caller-supplied matrix, not part of the device path."""
    n = code.shape[0]
    table = np.array([recode_integer(i, code) for i in range(1 << n)], dtype=np.int64)
    out = np.zeros(mat.shape, mat.dtype)
    out[np.ix_(table, table)] = mat
    return out
