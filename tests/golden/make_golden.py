"""
Generates the committed fixtures in tests/golden/ from the reference checkout.

Run ONCE in the build container (``python tests/golden/make_golden.py``); it reads
``/root/reference`` (data files and stored notebook outputs only -- the reference package itself
cannot be imported there, ``qat`` is missing) and nothing at test time reads the reference again.

Outputs
  reference_goldens.json  numbers and term lists the reference's own tests / executed notebooks hold
                          for the Hamiltonian hot path, each with its file:line
  h2_integrals.npz        one/two-body integrals + scalars of tests/resources/h2_data.npz
  lih_integrals.npz       same for tests/resources/lih_data.npz
"""

from __future__ import annotations

import json
import re
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent


def notebook_output(nb_path: Path, cell: int) -> str:
    nb = json.loads(nb_path.read_text())
    texts = []
    for o in nb["cells"][cell].get("outputs", []):
        if "text" in o:
            texts.append("".join(o["text"]))
    return "".join(texts)


_TERM_RE = re.compile(r"^\(?([^()]*?j?)\)? \* (?:I\^(\d+)|\((\w+)\|\[([\d, ]*)\]\))")


def parse_observable_print(text: str):
    """'H_spin = (c) * I^n +\n c * (OPS|[q, ..]) + ...' -> (constant, [(coeff, op, qbits)])."""
    body = text.split("=", 1)[1]
    const = 0.0
    terms = []
    for line in body.split("+\n"):
        line = line.strip()
        if not line:
            continue
        m = _TERM_RE.match(line)
        if not m:
            raise ValueError(f"cannot parse {line!r}")
        coeff = complex(m.group(1).replace(" ", ""))
        if m.group(2) is not None:
            const = coeff
        else:
            qb = [int(v) for v in m.group(4).split(",") if v.strip()]
            terms.append(([coeff.real, coeff.imag], m.group(3), qb))
    return [const.real, const.imag], terms


def main() -> None:
    g = {}
    nb = REF / "doc/notebooks/qat_fermion_spin_fermion_transforms.ipynb"
    c3, t3 = parse_observable_print(notebook_output(nb, 5))
    c4, t4 = parse_observable_print(notebook_output(nb, 12))
    g["parity_n3"] = {
        "source": "doc/notebooks/qat_fermion_spin_fermion_transforms.ipynb cells 3,5,7",
        "nqbits": 3,
        "fermion_terms": [[1.0, "Cc", [0, 1]], [0.5, "CCcc", [0, 1, 0, 1]]],
        "encoding": "parity",
        "constant": c3,
        "terms": t3,
    }
    g["parity_n4_electronic"] = {
        "source": "doc/notebooks/qat_fermion_spin_fermion_transforms.ipynb cells 10,12,14",
        "nqbits": 4,
        "hpq": [[0.0, 1.0, 0.0, 0.0], [1.0, 0.0, 1.0, 0.0], [0.0, 1.0, 0.0, 1.0], [0.0, 0.0, 1.0, 0.0]],
        "hpqrs_nonzero": [[[0, 1, 1, 0], 0.6], [[1, 0, 0, 1], 0.6], [[2, 0, 0, 2], 0.6]],
        "encoding": "parity",
        "constant": c4,
        "terms": t4,
    }
    c1, t1 = parse_observable_print(notebook_output(REF / "doc/notebooks/qat_fermion_vqe_hubbard.ipynb", 4))
    g["hubbard_one_site_jw"] = {
        "source": "doc/notebooks/qat_fermion_vqe_hubbard.ipynb cells 2,4",
        "U": 2.0,
        "mu": 1.0,
        "constant": c1,
        "terms": t1,
    }
    g["scalars"] = {
        "zne_energy_shallow_seed0": {
            "value": 1.482703437040279,
            "decimals": 12,
            "source": "tests/integration/test_integration_zne_plugin.py:16,23-37,43,49",
        },
        "seqoptim_energy_seed1": {
            "value": -0.28053014718543934,
            "decimals": 12,
            "source": "tests/integration/test_integration_seqoptim.py:16-33",
        },
        "h2_ucc_vqe_energy": {
            "value": -1.1372701746609022,
            "decimals": 8,
            "source": "tests/integration/test_integration_ucc.py:112,161",
        },
        "spin_min_eigenvalue": {
            "value": -1.25,
            "decimals": 13,
            "source": "tests/unitary/test_unitary_hamiltonians.py:14-34",
        },
        "hubbard_dimer_ground": {
            "value": -2.1180339887498945,
            "source": "doc/notebooks/qat_fermion_gaussian_state_preparation.ipynb cell 4 (t=0.25, U=2, mu=1)",
        },
        "hubbard_dimer_ground_no_interaction": {
            "value": -4.0,
            "source": "doc/notebooks/qat_fermion_gaussian_state_preparation.ipynb cell 6",
        },
    }
    g["spin_matrix_4x4"] = {
        "source": "tests/unitary/test_unitary_hamiltonians.py:252-276",
        "terms": [[0.5, "X", [0]], [0.25, "Y", [1]], [1.0, "ZZ", [0, 1]]],
        "matrix_re": [[1.0, 0.0, 0.5, 0.0], [0.0, -1.0, 0.0, 0.5], [0.5, 0.0, -1.0, 0.0], [0.0, 0.5, 0.0, 1.0]],
        "matrix_im": [[0.0, -0.25, 0.0, 0.0], [0.25, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, -0.25], [0.0, 0.0, 0.25, 0.0]],
    }
    g["fermion_matrix_4x4"] = {
        "source": "tests/unitary/test_unitary_hamiltonians.py:279-293",
        "terms": [[1.0, "Cc", [0, 1]], [0.5, "CCcc", [0, 1, 0, 1]]],
        "matrix_re": [[0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, -0.5]],
    }
    for name in ("h2", "lih"):
        d = np.load(REF / f"tests/resources/{name}_data.npz", allow_pickle=True)
        info = d["info"].tolist()
        np.savez_compressed(
            OUT / f"{name}_integrals.npz",
            one_body_integrals=d["one_body_integrals"],
            two_body_integrals=d["two_body_integrals"],
            nuclear_repulsion=float(d["nuclear_repulsion"]),
            n_electrons=int(d["n_electrons"]),
            rdm1=d["rdm1"],
            orbital_energies=d["orbital_energies"],
            fci=float(info["FCI"]),
            hf=float(info["HF"]),
        )
        g["scalars"][f"{name}_fci"] = {"value": float(info["FCI"]), "source": f"tests/resources/{name}_data.npz field info['FCI']"}
    (OUT / "reference_goldens.json").write_text(json.dumps(g, indent=1))
    print("wrote", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    main()
