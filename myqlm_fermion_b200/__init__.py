"""myqlm_fermion_b200 -- B200-native (sm_100a) Pauli-sum engine behind myqlm-fermion's Hamiltonian hot path.

Host side in Python with the reference's object names; every number is computed by libqfe.so
(hand-written CUDA behind the C ABI of include/qfe.h).  There is no CPU fallback: a missing
library raises ImportError at first use, a missing GPU raises RuntimeError.
"""

from ._lib import QfeError, load as load_library
from .engine import Engine, Plan, get_engine
from .hamiltonians import (
    ElectronicStructureHamiltonian,
    FermionHamiltonian,
    FermionicTerm,
    SpinHamiltonian,
    Term,
    ind_clusters_ord,
    ind_spins_ord,
    make_anderson_model,
    make_embedded_model,
    make_hubbard_model,
    make_tot_density_op,
    normal_order_fermionic_term,
)
from .packed import PackedTerms, pack_terms, unpack_term
from . import plugins, workloads
from .chemistry import construct_ucc_ansatz, convert_to_h_integrals, get_cluster_ops, get_hf_ket
from .io import load_molecule_npz, load_terms_npz, molecule_hamiltonian, save_molecule_npz, save_terms_npz
from .transforms import (
    change_encoding,
    get_bk_code,
    get_jw_code,
    get_parity_code,
    make_PCU_sets,
    recode_integer,
    transform_to_bk_basis,
    transform_to_jw_basis,
    transform_to_parity_basis,
)

__all__ = [
    "QfeError", "load_library", "Engine", "Plan", "get_engine",
    "ElectronicStructureHamiltonian", "FermionHamiltonian", "FermionicTerm", "SpinHamiltonian", "Term",
    "ind_clusters_ord", "ind_spins_ord", "make_anderson_model", "make_embedded_model", "make_hubbard_model", "make_tot_density_op", "normal_order_fermionic_term",
    "PackedTerms", "pack_terms", "unpack_term", "plugins", "construct_ucc_ansatz", "convert_to_h_integrals", "get_cluster_ops", "get_hf_ket",
    "load_molecule_npz", "load_terms_npz", "molecule_hamiltonian", "save_molecule_npz", "save_terms_npz",
    "change_encoding", "get_bk_code", "get_jw_code", "get_parity_code", "make_PCU_sets", "recode_integer",
    "transform_to_bk_basis", "transform_to_jw_basis", "transform_to_parity_basis",
]
