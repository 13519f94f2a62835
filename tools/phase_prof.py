"""Development aid: time <psi|H|psi> of the synthetic molecular JW Hamiltonian at n = 2M qubits and, with QFE_PROF=1 in the
environment, print where the warps of the sweep kernel spend their cycles (phases of the tile loop, summed over CTAs and sweeps).
usage: [QFE_EXPECT2=0|1] [QFE_PIPELINED=0|1] [QFE_PROF=1] python tools/phase_prof.py M [reps] [transition]"""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import myqlm_fermion_b200 as qfe
from myqlm_fermion_b200.workloads import synthetic_h_integrals

m = int(sys.argv[1]) if len(sys.argv) > 1 else 14
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
transition = "transition" in sys.argv[3:]  # <phi|H|psi> with phi != psi: the path of the non-local classes (bra tile != ket tile, full visits)
n = 2 * m
eng = qfe.get_engine(0)
eng.use_torch_stream()
hpq, hpqrs = synthetic_h_integrals(m, seed=1234)
terms = eng.terms_from_integrals(hpq, hpqrs)
plan = eng.plan(terms)
psi = eng.random_state(n, dtype=torch.complex128, seed=0)
bra = eng.random_state(n, dtype=torch.complex128, seed=1) if transition else None
e = eng.expectation(plan, psi, bra=bra)  # warm-up
cyc = np.zeros(16 * 8, dtype=np.uint64)
eng.lib.qfe_dev_phase_cycles(eng.ctx, cyc.ctypes.data_as(C.POINTER(C.c_uint64)), 1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
ev0.record()
for _ in range(reps):
    e = eng.expectation(plan, psi, bra=bra)
ev1.record()
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / reps
eng.lib.qfe_dev_phase_cycles(eng.ctx, cyc.ctypes.data_as(C.POINTER(C.c_uint64)), 1)
out = {"transition": transition, "n": n, "terms": terms.n_terms, "sweeps": plan.n_sweeps, "env": {k: v for k, v in os.environ.items() if k.startswith("QFE_")},
       "ms": ms, "terms_amps_per_s": terms.n_terms * float(1 << n) / (ms / 1e3), "energy": e.real}
if cyc.any():
    c = cyc.reshape(16, 8).astype(np.float64)
    names = ["wait", "issue+own", "groups", "skew", "load", "stage"]
    tot = c[:, :6].sum()
    out["phase_share"] = {k: float(c[:, i].sum() / tot) for i, k in enumerate(names)}
    out["groups_cycles_by_warp_rel"] = [round(float(v), 4) for v in c[:, 2] / c[:, 2].mean()]
    out["skew_cycles_by_warp_rel"] = [round(float(v), 3) for v in c[:, 3] / max(c[:, 3].mean(), 1.0)]
    tiles = max(c[:, 6].mean(), 1.0)
    out["cycles_per_tile"] = {k: float(c[:, i].mean() / tiles) for i, k in enumerate(names)}
print(json.dumps(out), flush=True)
