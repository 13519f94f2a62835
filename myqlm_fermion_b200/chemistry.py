"""Input generators of the UCC / ADAPT workflows, restated from /root/reference/qat/fermion/chemistry/ucc.py so that a reference
script finds the same names: ``convert_to_h_integrals`` (:128-229), ``get_hf_ket`` (:779-798), ``get_cluster_ops`` (:801-852).  They
produce small host objects (index tuples, FermionHamiltonians of two terms each); the arithmetic on them happens in libqfe."""

from __future__ import annotations

from typing import List, Optional

from .hamiltonians import FermionHamiltonian, Term
from .workloads import convert_to_h_integrals, uccsd_pool_fermion_terms  # noqa: F401  (convert_to_h_integrals is re-exported)


def get_hf_ket(n_electrons: int, nqbits: int) -> int:
    """Hartree-Fock determinant as an integer: the first n_electrons qubits (most significant bits) set."""
    return sum(1 << (nqbits - 1 - i) for i in range(n_electrons))


def get_cluster_ops(n_electrons: int, nqbits: Optional[int] = None, noons: Optional[List[float]] = None) -> List[FermionHamiltonian]:
    """The UCCSD pool {i(a+_a a_i - h.c.)} U {i(a+_b a+_a a_j a_i - h.c.)} over occupied i, j and unoccupied a, b spin-orbitals, spin
    preserving, in the reference's order (singles, then doubles, each sorted descending).  ``noons`` only fixes the register size
    here (2 * len(noons)), as in the reference when no operator limit is given."""
    if noons is None and nqbits is None:
        raise TypeError("Missing input nqbits/noons. One is needed to compute the cluster operators.")
    if noons is not None:
        nqbits = 2 * len(noons)
    elif nqbits % 2 != 0:
        raise ValueError("Only even values of nqbits are allowed.")
    if n_electrons % 2 != 0:
        raise ValueError("Only even values of n_electrons are allowed.")
    return [FermionHamiltonian(nqbits, [Term(c, op, idx) for c, op, idx in terms]) for terms in uccsd_pool_fermion_terms(n_electrons, nqbits)]


def construct_ucc_ansatz(cluster_ops, ket_hf: int, n_steps: int = 1, engine=None):
    """State preparation of the UCC ansatz as ``construct_ucc_ansatz`` builds it (ucc.py:291-343): the Hartree-Fock determinant, then per
    Trotter step the first-order slice of H_theta = sum_k theta_k (iT_k) with coeff = 1 / n_steps, i.e. prod_t exp(-i c_t P_t / n_steps)
    over the merged Pauli terms of H_theta.  Returns ``theta -> device state`` (len(theta) = len(cluster_ops) * n_steps); ``cluster_ops``
    are the spin-transformed cluster operators."""
    from .engine import get_engine

    eng = engine or get_engine()
    nqbits = cluster_ops[0].nbqbits

    def state(theta):
        if len(theta) != len(cluster_ops) * n_steps:
            raise ValueError(f"expected {len(cluster_ops) * n_steps} angles, got {len(theta)}")
        psi = eng.basis_state(nqbits, ket_hf)
        idx = 0
        for _ in range(n_steps):
            h_theta = None
            for th, op in zip(theta[idx : idx + len(cluster_ops)], cluster_ops):
                h_theta = float(th) * op if h_theta is None else h_theta + float(th) * op
            eng.trotter_slice(h_theta, psi, coeff=1.0 / n_steps)
            idx += len(cluster_ops)
        return psi

    return state
