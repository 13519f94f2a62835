"""Hamiltonian objects with the reference's names and signatures
(/root/reference/qat/fermion/hamiltonians.py), backed by packed terms and the GPU engine.

What is kept: ``Term(coeff, op, qbits)``, ``SpinHamiltonian(nqbits, terms, constant_coeff)``,
``FermionHamiltonian(nqbits, terms, constant_coeff, normal_order)``,
``ElectronicStructureHamiltonian(hpq, hpqrs, constant_coeff)`` with ``.hpq/.hpqrs`` setters,
``to_spin(method)``, ``to_electronic()``, ``to_fermion()``, ``dag()``, ``copy()``, ``get_matrix(sparse)``,
``+ - * |``, ``make_hubbard_model``, ``make_anderson_model``, ``make_tot_density_op``; the same
exceptions (TypeError for a fermionic letter in a spin term and vice versa, KeyError for an unknown
``to_spin`` method).  What changes: no 2^n x 2^n matrix is needed to measure or apply a Hamiltonian
(see engine.py); ``get_matrix`` itself is one GPU apply on vec(identity).
"""

from __future__ import annotations

import itertools
from copy import deepcopy
from numbers import Number
from typing import List, Optional

import numpy as np

from .packed import PackedTerms, pack_terms, unpack_term

TOL = 1e-12  # hamiltonians.py:541


class Term:
    """``qat.core.Term`` look-alike: coefficient, operator letters, qubits."""

    __slots__ = ("coeff", "op", "qbits")

    def __init__(self, coefficient, pauli_op: str, list_qbits, do_validity_check: bool = True):
        if do_validity_check and len(pauli_op) != len(list_qbits):
            raise ValueError("operator string and qubit list differ in length")
        self.coeff = coefficient
        self.op = pauli_op
        self.qbits = list(list_qbits)

    def copy(self):
        return deepcopy(self)

    def __repr__(self):
        return f"{self.coeff} * ({self.op}|{self.qbits})"

    def __eq__(self, other):
        return isinstance(other, Term) and self.coeff == other.coeff and self.op == other.op and self.qbits == other.qbits


class FermionicTerm(Term):
    """Term restricted to ladder letters C (creation) / c (annihilation)
    (/root/reference/qat/fermion/fermion_algebra.py:22-68)."""

    __slots__ = ()

    def __init__(self, coefficient, op: str, qbits, do_validity_check: bool = True):
        if any(v in "IXYZ" for v in op):
            raise TypeError("FermionicTerm only accepts fermionic operators C and c.")
        super().__init__(coefficient, op, qbits, do_validity_check)

    def __mul__(self, other):
        if isinstance(other, Number):
            return FermionicTerm(self.coeff * other, self.op, self.qbits)
        return FermionicTerm(self.coeff * other.coeff, self.op + other.op, self.qbits + other.qbits)

    @staticmethod
    def from_term(term: Term) -> "FermionicTerm":
        return FermionicTerm(term.coeff, term.op, term.qbits)


# ---------------------------------------------------------------------------------------------
# Wick ordering (fermion_algebra.py:71-251): creations left, annihilations right, indices ascending
# ---------------------------------------------------------------------------------------------
def normal_order_fermionic_term(term: Term) -> List[FermionicTerm]:
    """All creation operators to the left of all annihilation operators, each block sorted by
    ascending index, using {a_p, a+_q} = delta_pq; terms with a repeated index in a block vanish."""
    out: dict = {}
    stack = [(term.coeff, list(zip(term.op, term.qbits)))]
    while stack:
        coeff, ops = stack.pop()
        pos = next((i for i in range(len(ops) - 1) if ops[i][0] == "c" and ops[i + 1][0] == "C"), None)
        if pos is not None:
            swapped = ops[:pos] + [ops[pos + 1], ops[pos]] + ops[pos + 2 :]
            stack.append((-coeff, swapped))
            if ops[pos][1] == ops[pos + 1][1]:
                stack.append((coeff, ops[:pos] + ops[pos + 2 :]))
            continue
        n_c = sum(1 for o in ops if o[0] == "C")
        sign = 1
        blocks = []
        vanish = False
        for block in (ops[:n_c], ops[n_c:]):
            idx = [q for _, q in block]
            if len(set(idx)) != len(idx):
                vanish = True
                break
            inv = sum(1 for a in range(len(idx)) for b in range(a + 1, len(idx)) if idx[a] > idx[b])
            sign *= -1 if inv & 1 else 1
            blocks.append(sorted(idx))
        if vanish:
            continue
        key = ("C" * n_c + "c" * (len(ops) - n_c), tuple(blocks[0] + blocks[1]))
        out[key] = out.get(key, 0) + sign * coeff
    return [FermionicTerm(c, op, list(qb)) for (op, qb), c in out.items()]


def _merge_terms(terms):
    """Like-term merge in first-appearance order (what qat.core's ``+=`` does at hamiltonians.py:43)."""
    acc: dict = {}
    for t in terms:
        key = (t.op, tuple(t.qbits))
        acc[key] = acc.get(key, 0) + t.coeff
    return acc


class _Base:
    nbqbits: int
    constant_coeff: complex

    def copy(self):
        return deepcopy(self)

    def __neg__(self):
        return self * -1


# ---------------------------------------------------------------------------------------------
class SpinHamiltonian(_Base):
    """Sum of Pauli strings (/root/reference/qat/fermion/hamiltonians.py:63-258)."""

    def __init__(self, nqbits: int, terms: Optional[List[Term]] = None, constant_coeff=0.0, *, _packed: PackedTerms | None = None):
        self.nbqbits = nqbits
        self.matrix = None
        self._terms = list(terms) if terms is not None else []
        self.constant_coeff = constant_coeff
        self._packed = _packed
        if _packed is not None:
            self._terms = None
            self.constant_coeff = _packed.constant
        elif self._terms:
            for letter in ("C", "c"):  # hamiltonians.py:137-147: only the first term is checked
                if letter in self._terms[0].op:
                    raise TypeError("SpinHamiltonian does not support fermionic operators. Please use FermionHamiltonian instead.")

    # -- term views ----------------------------------------------------------------------------
    @property
    def terms(self) -> List[Term]:
        if self._terms is None:
            x, z, c, _, _ = self._packed.arrays()
            self._terms = []
            for xi, zi, ci in zip(x.tolist(), z.tolist(), c.tolist()):
                op, qb = unpack_term(self.nbqbits, xi, zi)
                self._terms.append(Term(ci, op, qb))
        return self._terms

    def add_term(self, term: Term) -> None:
        self.terms.append(term)
        self._packed = None
        self.matrix = None

    def set_terms(self, terms) -> None:
        self._terms = list(terms)
        self._packed = None
        self.matrix = None

    def packed_arrays(self):
        """(x, z, c, constant) -- canonical when the object came from a transform, raw otherwise."""
        if self._packed is not None:
            x, z, c, _, k = self._packed.arrays()
            return x, z, c, complex(k[0])
        x, z, c = pack_terms(self.nbqbits, self._terms)
        return x, z, c, complex(self.constant_coeff)

    def packed(self, engine=None) -> PackedTerms:
        """Canonical term set in libqfe (merged, sorted, identity folded into the constant)."""
        if self._packed is None or complex(self.constant_coeff) != self._packed.constant:
            from .engine import get_engine

            eng = engine or get_engine()
            x, z, c, _k = self.packed_arrays()
            self._packed = eng.terms_from_pauli(self.nbqbits, x, z, c, constant=self.constant_coeff, drop_eps=0.0)
        return self._packed

    # -- algebra (reference: arithmetic of qat.core.Observable re-wrapped, hamiltonians.py:110-135) ----
    def _combine(self, other, sign):
        if isinstance(other, Number):
            out = SpinHamiltonian(self.nbqbits, [t.copy() for t in self.terms], self.constant_coeff + sign * other)
            return out
        if other.nbqbits != self.nbqbits:
            raise ValueError("operands act on registers of different size")
        x1, z1, c1, k1 = self.packed_arrays()
        x2, z2, c2, k2 = other.packed_arrays()
        return _spin_from_arrays(self.nbqbits, np.concatenate([x1, x2]), np.concatenate([z1, z2]), np.concatenate([c1, sign * c2]), k1 + sign * k2)

    def __add__(self, other):
        return self._combine(other, 1)

    __radd__ = __add__

    def __sub__(self, other):
        return self._combine(other, -1)

    def __rsub__(self, other):
        return (self * -1) + other

    def __mul__(self, other):
        if isinstance(other, Number):
            x, z, c, k = self.packed_arrays()
            return _spin_from_arrays(self.nbqbits, x, z, c * other, k * other, merge=False)
        if other.nbqbits != self.nbqbits:
            raise ValueError("operands act on registers of different size")
        # include the constants as identity strings, multiply all pairs
        x1, z1, c1, k1 = self.packed_arrays()
        x2, z2, c2, k2 = other.packed_arrays()
        x1, z1, c1 = np.append(x1, np.uint64(0)), np.append(z1, np.uint64(0)), np.append(c1, k1)
        x2, z2, c2 = np.append(x2, np.uint64(0)), np.append(z2, np.uint64(0)), np.append(c2, k2)
        pc = lambda a: np.bitwise_count(a).astype(np.int64)
        # XZ-form: (i^{a} X^x1 Z^z1)(i^{b} X^x2 Z^z2) = i^{a+b} (-1)^{|z1&x2|} X^{x1^x2} Z^{z1^z2}
        X = x1[:, None] ^ x2[None, :]
        Z = z1[:, None] ^ z2[None, :]
        k = pc(x1 & z1)[:, None] + pc(x2 & z2)[None, :] + 2 * pc(z1[:, None] & x2[None, :]) - pc(X & Z)
        coeff = c1[:, None] * c2[None, :] * np.array([1, 1j, -1, -1j])[k % 4]
        return _spin_from_arrays(self.nbqbits, X.ravel(), Z.ravel(), coeff.ravel(), 0.0)

    def __rmul__(self, other):
        if isinstance(other, Number):
            return self * other
        return other * self

    def __or__(self, other):
        return self * other - other * self

    def dag(self) -> "SpinHamiltonian":
        return SpinHamiltonian(self.nbqbits, [Term(np.conj(t.coeff), t.op, t.qbits) for t in self.terms], np.conj(self.constant_coeff))

    # -- matrix (hamiltonians.py:163-211), computed on the GPU ---------------------------------
    def get_matrix(self, sparse: bool = False):
        self.matrix = _matrix_on_gpu(self, sparse)
        return self.matrix

    def __repr__(self):
        head = f"{self.constant_coeff} * I^{self.nbqbits}"
        return " +\n".join([head] + [repr(t) for t in self.terms])


def _spin_from_arrays(n, x, z, c, constant, merge=True) -> SpinHamiltonian:
    from .engine import get_engine

    if merge:
        packed = get_engine().terms_from_pauli(n, x, z, c, constant=constant, drop_eps=0.0)
        return SpinHamiltonian(n, _packed=packed)
    terms = []
    for xi, zi, ci in zip(np.asarray(x).tolist(), np.asarray(z).tolist(), np.asarray(c).tolist()):
        op, qb = unpack_term(n, xi, zi)
        terms.append(Term(ci, op, qb))
    return SpinHamiltonian(n, terms, constant)


def _matrix_on_gpu(h: SpinHamiltonian, sparse: bool):
    """H as a dense matrix = (H (x) 1) applied to vec(identity): ONE matrix-free apply on a
    2n-qubit register whose high n qubits carry H."""
    import torch

    from .engine import get_engine

    n = h.nbqbits
    if n > 14:
        raise ValueError("get_matrix should not be used if the Hamiltonian is too large (use the matrix-free engine)")
    eng = get_engine()
    x, z, c, k = h.packed_arrays()
    big = eng.terms_from_pauli(2 * n, x << np.uint64(n), z << np.uint64(n), c, constant=k, drop_eps=0.0)
    dim = 1 << n
    eye = torch.eye(dim, dtype=torch.complex128, device=f"cuda:{eng.device}").reshape(-1).contiguous()
    mat = eng.apply(big, eye).reshape(dim, dim).cpu().numpy()
    if sparse:
        import scipy.sparse as sp

        return sp.csr_matrix(mat)
    return mat


# ---------------------------------------------------------------------------------------------
class FermionHamiltonian(_Base):
    """Sum of products of ladder operators (/root/reference/qat/fermion/hamiltonians.py:261-499)."""

    def __init__(self, nqbits: int, terms: List[Term], constant_coeff=0.0, normal_order: bool = True):
        self.nbqbits = nqbits
        self.matrix = None
        self.normal_order = normal_order
        self.constant_coeff = constant_coeff
        terms = list(terms or [])
        if terms:
            terms = [t if isinstance(t, FermionicTerm) else FermionicTerm.from_term(t) for t in terms]
            if normal_order:
                ordered = []
                for t in terms:
                    ordered.extend(normal_order_fermionic_term(t))
                # a scalar left over by contractions (empty operator string) belongs to the constant
                merged = _merge_terms(ordered)
                terms = []
                for (op, qb), coeff in merged.items():
                    if op == "":
                        self.constant_coeff = self.constant_coeff + coeff
                    else:
                        terms.append(FermionicTerm(coeff, op, list(qb)))
        self.terms = terms

    def set_terms(self, terms) -> None:
        self.terms = list(terms)
        self.matrix = None

    def __add__(self, other):
        if isinstance(other, Number):
            return FermionHamiltonian(self.nbqbits, self.terms, self.constant_coeff + other, normal_order=False)
        merged = _merge_terms(list(self.terms) + list(other.terms))
        return FermionHamiltonian(self.nbqbits, [FermionicTerm(c, op, list(qb)) for (op, qb), c in merged.items()], self.constant_coeff + other.constant_coeff)

    __radd__ = __add__

    def __sub__(self, other):
        return self + (other * -1)

    def __rsub__(self, other):
        return (self * -1) + other

    def __mul__(self, other):
        if isinstance(other, Number):
            return FermionHamiltonian(self.nbqbits, [FermionicTerm(t.coeff * other, t.op, t.qbits) for t in self.terms], self.constant_coeff * other, normal_order=False)
        out = [FermionicTerm(t.coeff * other.constant_coeff, t.op, t.qbits) for t in self.terms]
        out += [FermionicTerm(self.constant_coeff * t.coeff, t.op, t.qbits) for t in other.terms]
        out += [FermionicTerm.from_term(a) * FermionicTerm.from_term(b) for a, b in itertools.product(self.terms, other.terms)]
        return FermionHamiltonian(self.nbqbits, out, self.constant_coeff * other.constant_coeff, normal_order=True)

    def __rmul__(self, other):
        if isinstance(other, Number):
            return self * other
        return other * self

    def __or__(self, other):
        return self * other - other * self

    def dag(self) -> "FermionHamiltonian":
        # (a+_p ... a_s)^+ reverses the operator order and swaps C <-> c
        swap = {"C": "c", "c": "C"}
        terms = [FermionicTerm(np.conj(t.coeff), "".join(swap[o] for o in reversed(t.op)), list(reversed(t.qbits))) for t in self.terms]
        return FermionHamiltonian(self.nbqbits, terms, np.conj(self.constant_coeff))

    def to_spin(self, method: Optional[str] = "jordan-wigner") -> SpinHamiltonian:
        from .transforms import transform_to_bk_basis, transform_to_jw_basis, transform_to_parity_basis

        transform_map = {"jordan-wigner": transform_to_jw_basis, "bravyi-kitaev": transform_to_bk_basis, "parity": transform_to_parity_basis}
        return transform_map[method](self)

    def to_electronic(self) -> "ElectronicStructureHamiltonian":
        n = self.nbqbits
        hpq = np.zeros((n,) * 2, dtype=complex)
        hpqrs = np.zeros((n,) * 4, dtype=complex)
        for term in self.terms:
            nc, na = term.op.count("C"), term.op.count("c")
            if not ((nc == 1 and na == 1) or (nc == 2 and na == 2)):
                raise TypeError(
                    "The Hamiltonian contains fermionic operators incompatible with a transformation to a electronic-structure Hamiltonian."
                    "The terms need to have either one creation and one annihilation operator, or two creation and two annihilation operators"
                )
            for t in normal_order_fermionic_term(term):
                if t.op == "Cc":
                    hpq[tuple(t.qbits)] += t.coeff
                elif t.op == "CCcc":
                    hpqrs[tuple(t.qbits)] += 2 * t.coeff
        return ElectronicStructureHamiltonian(hpq, hpqrs, self.constant_coeff)

    def get_matrix(self, sparse: bool = False):
        """Fock-space matrix (hamiltonians.py:389-425) = the Jordan-Wigner spin matrix (JW keeps the
        occupation-number basis, transforms.py:260-278), built on the GPU."""
        self.matrix = self.to_spin("jordan-wigner").get_matrix(sparse)
        return self.matrix

    def __repr__(self):
        return " +\n".join(([f"{self.constant_coeff} * I^{self.nbqbits}"] if self.constant_coeff else []) + [repr(t) for t in self.terms])


class ElectronicStructureHamiltonian(FermionHamiltonian):
    """H = sum h_pq a+_p a_q + 1/2 sum h_pqrs a+_p a+_q a_r a_s + c (hamiltonians.py:502-660).

    Unlike the reference, the n^2 + n^4 fermionic Term objects are only materialised when
    ``.terms`` is read; ``to_spin`` feeds the integral arrays straight to the GPU kernel."""

    TOL = TOL

    def __init__(self, hpq: np.ndarray, hpqrs: np.ndarray = None, constant_coeff=0.0):
        hpq = np.asarray(hpq)
        if hpqrs is None:
            hpqrs = np.zeros((hpq.shape[0],) * 4)
        self._hpq = hpq
        self._hpqrs = np.asarray(hpqrs)
        self.nbqbits = hpq.shape[0]
        self.constant_coeff = constant_coeff
        self.matrix = None
        self.normal_order = True
        self._terms_cache = None

    # integrals with invalidation (hamiltonians.py:560-596)
    @property
    def hpq(self):
        return self._hpq

    @hpq.setter
    def hpq(self, value):
        self._hpq = np.asarray(value)
        self.nbqbits = self._hpq.shape[0]
        self._terms_cache = None
        self.matrix = None

    @property
    def hpqrs(self):
        return self._hpqrs

    @hpqrs.setter
    def hpqrs(self, value):
        self._hpqrs = np.asarray(value)
        self._terms_cache = None
        self.matrix = None

    def _get_fermionic_terms(self) -> List[Term]:
        terms = []
        for p, q in np.argwhere(np.abs(self._hpq) > TOL):
            terms.append(FermionicTerm(self._hpq[p, q], "Cc", [int(p), int(q)]))
        for p, q, r, s in np.argwhere(np.abs(self._hpqrs) > TOL):
            terms.append(FermionicTerm(0.5 * self._hpqrs[p, q, r, s], "CCcc", [int(p), int(q), int(r), int(s)]))
        return terms

    @property
    def terms(self):
        if self._terms_cache is None:
            self._terms_cache = FermionHamiltonian(self.nbqbits, self._get_fermionic_terms(), 0.0).terms
        return self._terms_cache

    @terms.setter
    def terms(self, value):
        self._terms_cache = list(value)

    def to_fermion(self) -> FermionHamiltonian:
        return FermionHamiltonian(self.nbqbits, self._get_fermionic_terms(), self.constant_coeff)

    def to_electronic(self):
        return self.copy()

    def dag(self) -> "ElectronicStructureHamiltonian":
        # as the reference (hamiltonians.py:598-606): element-wise conjugation of the integral arrays
        return ElectronicStructureHamiltonian(np.conj(self.hpq), np.conj(self.hpqrs), np.conj(self.constant_coeff))

    def __add__(self, other):
        if isinstance(other, ElectronicStructureHamiltonian):
            return ElectronicStructureHamiltonian(self.hpq + other.hpq, self.hpqrs + other.hpqrs, self.constant_coeff + other.constant_coeff)
        return self.to_fermion() + other

    def __mul__(self, other):
        if isinstance(other, Number):
            return ElectronicStructureHamiltonian(self.hpq * other, self.hpqrs * other, self.constant_coeff * other)
        return self.to_fermion() * other

    def to_spin(self, method: Optional[str] = "jordan-wigner") -> SpinHamiltonian:
        from .transforms import _transform_integrals

        return _transform_integrals(self, method)


# ---------------------------------------------------------------------------------------------
# model builders (hamiltonians.py:663-941)
# ---------------------------------------------------------------------------------------------
def make_hubbard_model(t_mat: np.ndarray, U: float, mu: float) -> ElectronicStructureHamiltonian:
    """H = sum (t_ij - mu delta_ij) c+_{i s} c_{j s} + U sum n_{i up} n_{i dn}; spin-orbital 2*site + sigma."""
    t_mat = np.asarray(t_mat)
    ns = t_mat.shape[0]
    n = 2 * ns
    hpq = np.zeros((n, n), dtype=t_mat.dtype if np.iscomplexobj(t_mat) else float)
    for sig in (0, 1):
        hpq[sig::2, sig::2] = t_mat
    hpq[np.arange(n), np.arange(n)] += -mu
    hpqrs = np.zeros((n, n, n, n))
    for i in range(ns):
        for sig in (0, 1):
            a, b = 2 * i + sig, 2 * i + 1 - sig
            hpqrs[a, b, a, b] = -U
    return ElectronicStructureHamiltonian(hpq=hpq, hpqrs=hpqrs)


def make_anderson_model(u: float, mu: float, v: np.ndarray, epsilon: np.ndarray) -> ElectronicStructureHamiltonian:
    """Single impurity (modes 0 = up, 1 = down) coupled to n_b bath levels (modes 2k+2+sigma)."""
    nb = len(v)
    if len(epsilon) != nb:
        raise Exception("Error: The bath modes energies vector must be the same size as the tunneling energies vector.")
    n = 2 * nb + 2
    hpq = np.zeros((n, n))
    hpqrs = np.zeros((n, n, n, n))
    hpq[0, 0] = hpq[1, 1] = -mu
    for i in range(nb):
        for sig in (0, 1):
            b = 2 * (i + 1) + sig
            hpq[b, b] += epsilon[i]
            hpq[sig, b] += v[i]
            hpq[b, sig] += v[i]
    hpqrs[0, 1, 0, 1] = hpqrs[1, 0, 1, 0] = -u
    return ElectronicStructureHamiltonian(hpq, hpqrs)


def ind_clusters_ord(i: int, n_orb: int) -> int:
    """Cluster-ordered index (up, dn, ..., up, dn) of the spin-orbital with spin-ordered index i (hamiltonians.py:833-855)."""
    if i < n_orb:
        return 2 * i
    if i >= 2 * n_orb:
        print("index must be lesser than 2*n_orb")
        return None
    return 2 * i - (2 * n_orb - 1)


def ind_spins_ord(i: int, n_orb: int) -> int:
    """Spin-ordered index (all up, then all down) of the spin-orbital with cluster-ordered index i (hamiltonians.py:858-871)."""
    return i // 2 if i % 2 == 0 else (i - 1) // 2 + n_orb


def make_embedded_model(u: float, mu: float, D: np.ndarray, lambda_c: np.ndarray, t_loc: Optional[np.ndarray] = None,
                        int_kernel: Optional[np.ndarray] = None, grouping: Optional[str] = "spins") -> ElectronicStructureHamiltonian:
    """Impurity cluster (M spin-orbitals) hybridised with an uncorrelated bath cluster (M spin-orbitals), as
    /root/reference/qat/fermion/hamiltonians.py:732-830 fills it: h_pq = -mu on the impurity levels (+ t_loc), D and D^dagger between
    the clusters, -lambda_c on the bath (lambda_c f f^dagger = tr(lambda_c) - lambda_c f^dagger f: the trace goes to constant_coeff);
    h_pqrs = -u * int_kernel, or -u on the (up, down) pairs of the impurity orbitals.  ``grouping="spins"`` reorders h_pq from the
    cluster order (up, dn, up, dn, ...) to (all up, all down), exactly as the reference does (h_pqrs entries are then placed with
    ind_clusters_ord)."""
    M = np.shape(lambda_c)[0]
    D = np.asarray(D)
    lambda_c = np.asarray(lambda_c)
    h_pq = np.zeros((2 * M, 2 * M), dtype=np.complex128)
    h_pq[:M, :M] = -mu * np.eye(M)
    if t_loc is not None:
        h_pq[:M, :M] += np.asarray(t_loc)
    h_pq[M:, M:] = -lambda_c
    h_pq[:M, M:] = D
    h_pq[M:, :M] = np.conj(D).T
    const_coeff = np.trace(lambda_c)
    h_pqrs = np.zeros((2 * M, 2 * M, 2 * M, 2 * M)) if int_kernel is None else -u * np.asarray(int_kernel)
    if grouping == "spins":
        perm = np.zeros((2 * M, 2 * M))  # from spin order to cluster order
        perm[2 * np.arange(M), np.arange(M)] = 1
        perm[2 * np.arange(M) + 1, np.arange(M) + M] = 1
        if int_kernel is None and u != 0:
            for i in range(M // 2):
                a, b = ind_clusters_ord(2 * i, M), ind_clusters_ord(2 * i + 1, M)
                h_pqrs[a, b, a, b] = h_pqrs[b, a, b, a] = -u
        h_pq = perm @ h_pq @ perm.T
    elif grouping == "clusters":
        if int_kernel is None and u != 0:
            for i in range(M // 2):
                h_pqrs[2 * i, 2 * i + 1, 2 * i, 2 * i + 1] = h_pqrs[2 * i + 1, 2 * i, 2 * i + 1, 2 * i] = -u  # sign of the c+ c+ c c form
    else:
        print("Grouping must be either clusters or spins.")
    return ElectronicStructureHamiltonian(h_pq, h_pqrs, const_coeff)


def make_tot_density_op(n_sites: int) -> ElectronicStructureHamiltonian:
    return ElectronicStructureHamiltonian(hpq=np.identity(2 * n_sites))
