"""Multi-GPU parity check (run under torchrun, one rank per GPU): the sharded engine over NCCL must reproduce the
single-GPU results of the same library to 1e-12, which in turn are checked against the oracle by tests/.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 tools/check_sharded.py [n_qubits]"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import myqlm_fermion_b200 as qfe
from myqlm_fermion_b200 import workloads
from myqlm_fermion_b200.sharded import ShardedEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = qfe.get_engine(local)
se = ShardedEngine(eng, n)

for enc in ("jordan-wigner", "bravyi-kitaev"):
    hpq, hpqrs = workloads.synthetic_h_integrals_n(n, seed=1234)
    terms = eng.terms_from_integrals(hpq, hpqrs, constant=0.37, encoding=enc)
    full = eng.random_state(n, seed=3)  # every rank builds the same full vector (counter-based generator)
    lo, hi = rank << se.n_local, (rank + 1) << se.n_local
    shard = full[lo:hi].clone()
    plan1 = eng.plan(terms)
    plans = se.plan(terms)
    e1 = eng.expectation(plan1, full)
    es = se.expectation(plans, shard)
    assert abs(es - e1) <= 1e-12 * abs(e1), (enc, es, e1)
    landing = torch.zeros_like(shard)  # end-to-end form: shard from host memory, uploaded behind the sweeps of the local class
    eh = se.expectation_host(plans, shard.cpu().numpy(), landing)
    assert abs(eh - e1) <= 1e-12 * abs(e1) and torch.equal(landing, shard), (enc, eh, e1)
    phi1 = eng.apply(plan1, full)
    phis = se.apply(plans, shard)
    err = (phis - phi1[lo:hi]).abs().max().item()
    assert err <= 1e-12 * phi1.abs().max().item(), (enc, err)
    if rank == 0:
        print(f"{enc}: n={n} world={world} classes={plans.classes} E={es.real:.12f} |dE|={abs(es - e1):.2e} max|dphi|={err:.2e} exchanges={se.exchanges}", flush=True)

# ADAPT gradient sweep sharded vs single GPU (JW, UCCSD pool)
ne = 4
pool_f = workloads.uccsd_pool_fermion_terms(ne, n - (n % 2))[::9]
if n % 2 == 0:
    hpq, hpqrs = workloads.synthetic_h_integrals_n(n, seed=1234)
    h = eng.terms_from_integrals(hpq, hpqrs)
    pool = eng.concat([eng.terms_from_fermion(n, op) for op in pool_f])
    full = eng.random_state(n, seed=5)
    lo, hi = rank << se.n_local, (rank + 1) << se.n_local
    shard = full[lo:hi].clone()
    g1 = eng.commutator_gradients(h, pool, full)
    gs = se.commutator_gradients(se.plan(h), se.plan(pool), shard)
    assert np.max(np.abs(gs - g1)) <= 1e-11 * np.max(np.abs(g1)), np.max(np.abs(gs - g1))
    if rank == 0:
        print(f"commutator gradients: K={len(pool_f)} max|dg|={np.max(np.abs(gs - g1)):.2e} argmax={int(np.argmax(np.abs(gs)))}", flush=True)

# sharded Lanczos vs single-GPU Lanczos on a Hubbard chain
ns = n // 2
t_mat = workloads.hubbard_chain_t(ns)
hub = qfe.make_hubbard_model(t_mat, 4.0, 2.0)
hterms = eng.terms_from_integrals(hub.hpq, hub.hpqrs)
se2 = ShardedEngine(eng, 2 * ns)
hplan = se2.plan(hterms)
e_s, v_s, info_s = se2.ground_state(hplan, max_iter=300, tol=1e-10, want_vector=True)
e_1, _, info_1 = eng.ground_state(hterms, max_iter=300, tol=1e-10, want_vector=False)
assert abs(e_s - e_1) <= 1e-9 * abs(e_1), (e_s, e_1)
# the sharded Ritz vector: unit norm, <v|H|v> = E0, small eigen-residual
w_s = se2.apply(hplan, v_s)
eng.axpy(-e_s, v_s, w_s)
resid = abs(se2.dot(w_s, w_s)) ** 0.5
assert abs(se2.dot(v_s, v_s) - 1.0) < 1e-10 and abs(se2.expectation(hplan, v_s) - e_s) <= 1e-9 * abs(e_s) and resid <= 1e-4, resid
if rank == 0:
    print(f"Lanczos {2 * ns}q Hubbard chain: sharded E0={e_s:.10f} ({info_s['iterations']} it, |Hv-Ev|={resid:.1e}) single E0={e_1:.10f} ({info_1['iterations']} it)", flush=True)
if rank == 0:
    print("comm:", se.comm_info(), flush=True)
dist.barrier()
if rank == 0:
    print("sharded parity ok", flush=True)
dist.destroy_process_group()
