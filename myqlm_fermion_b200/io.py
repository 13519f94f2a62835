"""On-disk formats.

* Molecule files: ``.npz`` with the keys of the reference's fixtures ``tests/resources/{h2,lih}_data.npz`` (written by
  /root/reference/qat/fermion/chemistry/pyscf_tools.py:36-139): ``one_body_integrals`` (m, m), ``two_body_integrals`` (m, m, m, m),
  ``nuclear_repulsion``, ``n_electrons``, optionally ``rdm1``, ``orbital_energies``, ``info``.  ``molecule_hamiltonian`` turns one
  into the spin-orbital ``ElectronicStructureHamiltonian`` the reference builds with ``convert_to_h_integrals``
  (chemistry/ucc.py:128-229), nuclear repulsion as ``constant_coeff``.
* Packed term sets: ``.npz`` with ``nqbits``, ``x``, ``z`` (uint64 masks, bit n-1-q <-> qubit q), ``c`` (complex128, ``Term.coeff``),
  ``owner`` (operator slot of each term), ``constant`` (one per slot): the arrays ``qfe_terms_export`` returns, so a 10^5-term
  Hamiltonian is stored and reloaded without rebuilding Python ``Term`` objects.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

from .hamiltonians import ElectronicStructureHamiltonian
from .workloads import convert_to_h_integrals

MOLECULE_KEYS = ("one_body_integrals", "two_body_integrals", "nuclear_repulsion", "n_electrons")


def load_molecule_npz(path) -> dict:
    """The arrays of a molecule file as a dict; raises KeyError when a mandatory key is missing."""
    with np.load(Path(path), allow_pickle=True) as d:
        out = {k: d[k] for k in d.files}
    for k in MOLECULE_KEYS:
        if k not in out:
            raise KeyError(f"{path}: missing key '{k}' (expected the layout of the reference's tests/resources/*_data.npz)")
    return out


def save_molecule_npz(path, one_body_integrals, two_body_integrals, nuclear_repulsion, n_electrons, **extra) -> None:
    np.savez(Path(path), one_body_integrals=np.asarray(one_body_integrals, dtype=np.float64), two_body_integrals=np.asarray(two_body_integrals, dtype=np.float64),
             nuclear_repulsion=np.float64(nuclear_repulsion), n_electrons=np.int64(n_electrons), **extra)


def molecule_hamiltonian(molecule) -> tuple[ElectronicStructureHamiltonian, int]:
    """(H on 2m spin-orbitals, n_electrons) from a molecule file or the dict ``load_molecule_npz`` returns."""
    d = molecule if isinstance(molecule, dict) else load_molecule_npz(molecule)
    hpq, hpqrs = convert_to_h_integrals(np.asarray(d["one_body_integrals"], dtype=np.float64), np.asarray(d["two_body_integrals"], dtype=np.float64))
    return ElectronicStructureHamiltonian(hpq, hpqrs, constant_coeff=float(d["nuclear_repulsion"])), int(d["n_electrons"])


def save_terms_npz(path, terms) -> None:
    """``terms``: a PackedTerms (device-side canonical set, possibly a batch of operators) or a SpinHamiltonian."""
    if hasattr(terms, "arrays"):
        x, z, c, owner, constant = terms.arrays()
        nqbits = terms.nqbits
    else:
        x, z, c, k = terms.packed_arrays()
        owner, constant, nqbits = np.zeros(len(x), dtype=np.int32), np.array([k], dtype=np.complex128), terms.nbqbits
    np.savez(Path(path), nqbits=np.int64(nqbits), x=np.asarray(x, dtype=np.uint64), z=np.asarray(z, dtype=np.uint64), c=np.asarray(c, dtype=np.complex128),
             owner=np.asarray(owner, dtype=np.int32), constant=np.asarray(constant, dtype=np.complex128))


def load_terms_npz(path, engine=None):
    """PackedTerms on ``engine`` (default: the process-wide engine of the current device); batches come back as batches."""
    from .engine import get_engine

    eng = engine or get_engine()
    with np.load(Path(path)) as d:
        n, x, z, c, owner, constant = int(d["nqbits"]), d["x"], d["z"], d["c"], d["owner"], d["constant"]
    if len(x) != len(z) or len(x) != len(c) or len(x) != len(owner):
        raise ValueError(f"{path}: x, z, c and owner must have the same length")
    ops = [eng.terms_from_pauli(n, x[owner == s], z[owner == s], c[owner == s], constant=complex(constant[s])) for s in range(len(constant))]
    return ops[0] if len(ops) == 1 else eng.concat(ops)
