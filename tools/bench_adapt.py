"""Config 5 of BASELINE.json as a measurement: one ADAPT-VQE gradient sweep g_k = -i<psi|[H, tau_k]|psi> over the whole UCCSD pool, on a
statevector sharded over the ranks of torchrun (or one GPU).  Prints one JSON line on rank 0.
With "check" the sweep is verified on the spot: g_k of a few operators against single-operator plans (another kernel mode, their own exchanges),
and the reference's selection numbers (norm, argmax |g|: qat/plugins/adapt_vqe.py:186-195) are reported.
usage: [python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29556] tools/bench_adapt.py N_QUBITS [steps] [c128|c64] [check]"""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import myqlm_fermion_b200 as qfe
from myqlm_fermion_b200 import workloads
from myqlm_fermion_b200.sharded import ShardedEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tdtype = torch.complex64 if "c64" in sys.argv[3:] else torch.complex128
check = "check" in sys.argv[3:]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = qfe.get_engine(local)
n_e = (n // 2) & ~1
hpq, hpqrs = workloads.synthetic_h_integrals_n(n, seed=1234)
h = eng.terms_from_integrals(hpq, hpqrs)
del hpqrs
pool_f = workloads.uccsd_pool_fermion_terms(n_e, n)
t0 = time.perf_counter()
ops = [eng.terms_from_fermion(n, op) for op in pool_f]
pool = eng.concat(ops)
t_pool = time.perf_counter() - t0
if world > 1:
    se = ShardedEngine(eng, n)
    hp, pp = se.plan(h), se.plan(pool)
    psi = se.random_state(dtype=tdtype, seed=1)
    phi = torch.empty_like(psi)
    step = lambda: se.commutator_gradients(hp, pp, psi, scratch=phi)
else:
    hp, pp = eng.plan(h), eng.plan(pool)
    psi = eng.random_state(n, dtype=tdtype, seed=1)
    phi = torch.empty_like(psi)
    step = lambda: eng.commutator_gradients(hp, pp, psi, scratch=phi)
if "nowarm" not in sys.argv[3:]:
    g = step()  # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(steps):
    g = step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / steps
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t[0])
checked = None
if check:
    # phi holds H|psi> of the last sweep: g_k = 2 Im <phi|tau_k|psi> from single-operator plans
    scale = float(np.max(np.abs(g)))
    worst = 0.0
    ks = sorted({0, 1, len(ops) // 7, len(ops) // 3, len(ops) // 2, len(ops) - 2, len(ops) - 1, int(np.argmax(np.abs(g)))})
    for kk in ks:
        single = se.expectation(se.plan(ops[kk]), psi, bra_shard=phi) if world > 1 else eng.expectation(eng.plan(ops[kk]), psi, bra=phi)
        worst = max(worst, abs(g[kk] - 2.0 * single.imag) / scale)
    tol = 1e-10 if tdtype == torch.complex128 else 1e-4
    assert worst <= tol, worst
    checked = {"operators": ks, "max_rel_dev_vs_single_operator_plans": worst, "tolerance": tol}
if rank == 0:
    print(json.dumps({
        "workload": "ADAPT-VQE gradient sweep: H|psi> then <H psi|tau_k|psi> for the whole UCCSD pool", "n_qubits": n, "n_gpus": world, "n_electrons": n_e,
        "pauli_terms_H": h.n_terms, "pool_operators": len(pool_f), "pool_pauli_terms": pool.n_terms, "seconds_per_sweep": dt,
        "pool_operators_per_s": len(pool_f) / dt, "grad_norm": float(np.linalg.norm(g)), "argmax": int(np.argmax(np.abs(g))), "pool_build_s": t_pool,
        "dtype": "c64" if tdtype == torch.complex64 else "c128", "checked": checked,
        "comm": se.comm_info() if world > 1 else None, "exchanges_total": se.exchanges if world > 1 else 0,
    }), flush=True)
if world > 1:
    dist.destroy_process_group()
