"""Target program for the ncu captures of the kernels next to the headline one (28 qubits):
  python tools/ncu_rows.py apply   -> only APPLY sweeps of the molecular Hamiltonian are launched (phi = H|psi>)
  python tools/ncu_rows.py rotate  -> one Trotter slice of the 14-site Hubbard chain (fused Pauli-rotation passes)
e.g. ncu --set full --clock-control none --import-source on -k regex:tile_sweep --launch-skip 100 -c 1 -o gpurun_out/apply python tools/ncu_rows.py apply"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import myqlm_fermion_b200 as qfe
from myqlm_fermion_b200 import make_hubbard_model, workloads

mode = sys.argv[1] if len(sys.argv) > 1 else "apply"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
eng = qfe.get_engine(0)
psi = eng.random_state(n, seed=0)
if mode == "apply":
    hpq, hpqrs = workloads.synthetic_h_integrals_n(n, seed=1234)
    plan = eng.plan(eng.terms_from_integrals(hpq, hpqrs))
    phi = eng.apply(plan, psi)
    print("apply done", plan.n_sweeps, "sweeps", eng.dot(psi, phi))
else:
    hub = make_hubbard_model(workloads.hubbard_chain_t(n // 2), 4.0, 2.0).to_spin("jordan-wigner").packed(eng)
    l0 = eng.launch_count()
    eng.trotter_slice(hub, psi, coeff=0.05)
    eng.sync()
    print("slice done", eng.launch_count() - l0, "passes", eng.dot(psi, psi))
