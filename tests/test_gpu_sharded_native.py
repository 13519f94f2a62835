"""The multi-GPU path behind the C ABI (csrc/sharded.cu: qfe_comm_*, qfe_*_sharded) on the B200 box.

* one rank: a single-rank NCCL communicator in this process -- the sharded entry points must give the numbers of the unsharded ones and of
  the oracle (no exchange happens, but the communicator, the class walk, the all-reduce and the sharded Lanczos all run);
* two or more GPUs visible: tools/check_sharded.py under torchrun (N-GPU result == 1-GPU result to 1e-12 for <H>, H|psi>, the ADAPT
  sweep and Lanczos, over real ncclSend/ncclRecv exchanges).  Skipped on a one-GPU box; the logs of the 2- and 8-GPU runs of the round
  are under profiles/."""

import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import models, spin

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_single_rank_communicator_matches_the_unsharded_calls():
    import torch
    import torch.distributed as dist

    import myqlm_fermion_b200 as qfe
    from myqlm_fermion_b200.sharded import ShardedEngine

    own_group = not dist.is_initialized()
    if own_group:
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29611", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        eng = qfe.get_engine(0)
        n = 12
        se = ShardedEngine(eng, n)
        assert se.native and se.comm_info()["nranks"] == 1 and se.comm_info()["nccl_version"] > 20000
        one, two = models.synthetic_integrals(n // 2, seed=1234)
        hpq, hpqrs = models.convert_to_h_integrals(one, two)
        terms = eng.terms_from_integrals(hpq, hpqrs, constant=0.3)
        x, z, c, _, k = terms.arrays()
        rng = np.random.default_rng(4)
        psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
        psi /= np.linalg.norm(psi)
        want_phi = spin.apply_packed(n, x, z, c, complex(k[0]), psi)
        d = torch.from_numpy(psi).to("cuda:0")
        plan = se.plan(terms)
        e = se.expectation(plan, d)
        assert abs(e - np.vdot(psi, want_phi)) <= 1e-10 * abs(e)
        assert abs(se.expectation_host(plan, psi, torch.empty_like(d)) - e) <= 1e-12 * abs(e)
        phi = se.apply(plan, d).cpu().numpy()
        assert np.max(np.abs(phi - want_phi)) <= 1e-10 * np.max(np.abs(want_phi))
        assert abs(se.dot(d, d) - 1.0) < 1e-12
        pool = eng.concat([eng.terms_from_fermion(n, op) for op in models.cluster_ops(4, n)[::6]])
        g = se.commutator_gradients(plan, se.plan(pool), d)
        assert np.max(np.abs(g - eng.commutator_gradients(terms, pool, d))) <= 1e-12 * np.max(np.abs(g))
        e0, v0, info = se.ground_state(plan, max_iter=200, tol=1e-11)
        e1, _, _ = eng.ground_state(terms, max_iter=200, tol=1e-11)
        assert abs(e0 - e1) <= 1e-10 * abs(e1) and abs(se.dot(v0, v0) - 1.0) < 1e-10
        assert se.exchanges == 0
    finally:
        if own_group:
            dist.destroy_process_group()


def test_world_size_n_against_one_gpu_under_torchrun():
    import torch

    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("one GPU visible: the N-rank check needs at least two (logs of the round's 2- and 8-GPU runs: profiles/r02_sharded_*)")
    world = 1 << (min(n_dev, 8).bit_length() - 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29612",
           str(ROOT / "tools" / "check_sharded.py"), "20"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ))
    assert r.returncode == 0 and "sharded parity ok" in r.stdout, (r.stdout[-2000:], r.stderr[-2000:])
