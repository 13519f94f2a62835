#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: <psi|H|psi> of a synthetic molecular
Jordan-Wigner Hamiltonian (BASELINE.json config 4 family) in terms*amps/s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU algorithm

One "step" = one evaluation of <psi|H|psi> over the whole (sharded) statevector.
N=1: 32 qubits (random 8-fold-symmetric integrals on 32 spin-orbitals, 94 848 Pauli terms, complex128,
68.7 GB state).  N=2/4/8: 33/34/34 qubits, state sharded by its high qubits, one process per GPU
(torchrun), NCCL send/recv per x_hi class.  Prints ONE JSON line on rank 0.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "<psi|H|psi> terms*amps/s"
UNIT = "terms*amps/s"
BYTES_PER_SWEEP_AMP = {"c128": 16, "c64": 8}  # every amplitude is read once per sweep


def qubits_for(n_gpus: int) -> int:
    # BASELINE.json: 32 qubits on 1 B200 ... 34 qubits on 8 B200
    return {1: 32, 2: 33, 4: 34, 8: 34}.get(n_gpus, 32)


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL,
            )
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in Path(self.path).read_text().splitlines():
            f = [v.strip() for v in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(pw)), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port), bounded sample, all host cores
# ------------------------------------------------------------------------------------------------
def cpu_sample(n: int, workers: int, steps: int, warmup: int):
    from concurrent.futures import ProcessPoolExecutor

    from oracle import baseline

    triples, const = baseline.sample_terms(n)
    rng = np.random.default_rng(0)
    psi = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    psi /= np.linalg.norm(psi)
    pool = ProcessPoolExecutor(workers) if workers > 1 else None
    try:
        for _ in range(warmup):
            baseline.reference_energy(n, triples, const, psi, workers, pool)
        t0 = time.perf_counter()
        for _ in range(steps):
            e, _, _ = baseline.reference_energy(n, triples, const, psi, workers, pool)
        dt = time.perf_counter() - t0
    finally:
        if pool is not None:
            pool.shutdown()
    T = len(triples)
    return {"T": T, "n": n, "seconds_per_step": dt / steps, "value": T * float(1 << n) * steps / dt, "energy": e.real}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import baseline

    cores = baseline.usable_cores()
    n = 14 if cores >= 8 else 12
    r = cpu_sample(n, cores, max(1, args.steps), max(0, min(args.warmup, 1)))
    sample = f"{n}-qubit member of the same generator ({r['T']} Pauli terms): get_matrix(sparse=True) kron chain + vdot per step, term list split over {cores} processes"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * r["seconds_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
        "data": "synthetic", "config": bench_config(args.gpus, qubits_for(args.gpus), None, None),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bench_config(n_gpus, n, T, plan):
    cfg = {
        "workload": f"config4 family: random 8-fold-symmetric integrals on {n} spin-orbitals (seed 1234), Jordan-Wigner, <psi|H|psi> on a random complex state",
        "n_qubits": n, "n_gpus": n_gpus, "sharding": f"statevector split by its {max(n_gpus.bit_length() - 1, 0)} high qubits, one process per GPU",
        "l2": "inputs larger than L2 (each sweep streams the whole shard, >= 34 GB)",
    }
    if T is not None:
        cfg["pauli_terms"] = T
    if plan is not None:
        cfg.update(x_masks=plan.n_groups, sweeps_per_step=plan.n_sweeps, tile_bits=plan.tile_bits, x_hi_classes=plan.n_classes)
    return cfg


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="qfe", choices=["qfe", "reference"])
    ap.add_argument("--qubits", type=int, default=0, help="override the register size (development only)")
    ap.add_argument("--dtype", default="c128", choices=["c128", "c64"])
    ap.add_argument("--tile-bits", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (development only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (development only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    import myqlm_fermion_b200 as qfe
    from myqlm_fermion_b200 import workloads
    from myqlm_fermion_b200.sharded import ShardedEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # rank 0 prints exactly one JSON line: keep NCCL's version banner (written to stdout when the communicator is created) off it
        sys.stdout.flush()
        saved_fd, null_fd = os.dup(1), os.open(os.devnull, os.O_WRONLY)
        os.dup2(null_fd, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
            os.close(null_fd)

    n = args.qubits or qubits_for(world)
    tdtype = torch.complex128 if args.dtype == "c128" else torch.complex64
    elt = BYTES_PER_SWEEP_AMP[args.dtype]

    eng = qfe.get_engine(local_rank)
    eng.use_torch_stream()  # kernels on torch's current stream, so torch.cuda.Event brackets them

    # ---- kernel 1: integrals -> packed Pauli terms (on every rank; term tables are replicated) ----
    hpq, hpqrs = workloads.synthetic_h_integrals_n(n, seed=1234)
    t0 = time.perf_counter()
    terms = eng.terms_from_integrals(hpq, hpqrs, encoding="jordan-wigner")
    eng.sync()
    t_terms = time.perf_counter() - t0
    del hpqrs
    T = terms.n_terms

    if world > 1:
        se = ShardedEngine(eng, n)
        plan = se.plan(terms)
        psi = se.random_state(dtype=tdtype, seed=0)
        step = lambda: se.expectation(plan, psi)
    else:
        se = None
        plan = eng.plan(terms, tile_bits=args.tile_bits)
        psi = eng.random_state(n, dtype=tdtype, seed=0)
        step = lambda: eng.expectation(plan, psi)
    n_local = plan.n_local
    shard_bytes = (1 << n_local) * elt

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        ev0.record()
        val = None
        for _ in range(steps):
            val = fn()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        launches = eng.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms, float(launches)], dtype=torch.float64, device=dev)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, launches = float(tmax[0]), int(t[1])
        return ms, launches, val

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches, energy = timed(step, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    sec = ms / 1e3
    amps = float(1 << n)
    value = T * amps * args.steps / sec

    # ---- roofline of the dominant kernel (tile_sweep_kernel): HBM, 16 B (c128) per amplitude per sweep ----
    peak, peak_src = measured_peak()
    sweeps_per_step = plan.n_sweeps  # per rank; every rank launches the same number of sweeps over its own shard
    sweep_launches = sweeps_per_step * args.steps
    avg_launch_s = sec / max(sweep_launches, 1)  # upper bound: includes the (tiny) reduce kernels and, for N>1, the exchanges
    achieved = shard_bytes / avg_launch_s / 1e9
    # DRAM bytes of one sweep launch from the ncu --set full capture under profiles/ (dram__bytes_read.sum + dram__bytes_write.sum;
    # r01_sweep_v6_32q_ncu.txt: 68.742878 GB + 8.92 MB for a 2^32-amplitude complex128 shard); None where no capture exists
    ncu_traffic = {(32, "c128"): 68.742878e9 + 8.920320e6}.get((n_local, args.dtype))
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic,
        "kernel": "tile_sweep_kernel<EXPECT>", "peak_source": peak_src, "bytes_per_launch": shard_bytes, "avg_launch_ms": 1e3 * avg_launch_s,
        "terms_per_sweep": T / max(sweeps_per_step, 1),
        "note": "each sweep evaluates terms_per_sweep Pauli terms per amplitude from one HBM pass, so the kernel is issue/shared-memory bound by design; "
        "equiv_gbs_one_sweep_per_xmask is the HBM rate a one-pass-per-X-mask schedule would need for the same throughput",
        "equiv_gbs_one_sweep_per_xmask": plan.n_groups * shard_bytes * args.steps / sec / 1e9,
    }
    # the resource that actually binds the sweep kernel: shared-memory bandwidth.  Algorithmic shared-memory bytes per step =
    # one 16 B partner read per visited (X mask, amplitude) -- Hermitian X masks visit half the amplitudes -- plus one write and
    # one read of every staged amplitude per sweep.  Peak = 128 B/clk/SM (measured: tools/ubench/smem_paths.cu) x SMs x SM clock.
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    smem_bytes = shard_bytes * (0.5 * plan.half_groups + plan.full_groups + 2.0 * sweeps_per_step)
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    smem_peak = 128.0 * sm_count * sm_mhz * 1e6 / 1e9
    smem_achieved = smem_bytes * args.steps / sec / 1e9
    roofline["smem"] = {
        "achieved": smem_achieved, "peak": smem_peak, "unit": "GB/s", "frac": smem_achieved / smem_peak,
        "half_visit_x_masks": plan.half_groups, "full_visit_x_masks": plan.full_groups,
        "note": "algorithmic shared-memory bytes (partner rows + tile staging) / time; peak = 128 B/clk/SM x SMs x sampled SM clock",
    }

    # ---- e2e: the same step through the C-ABI call with HOST buffers (pinned), H2D inside the timed region ----
    e2e = None
    e2e_skip = None
    if not args.no_e2e:
        # every rank stages its shard in host memory: make sure the box has room for all of them before pinning
        try:
            import psutil

            avail = psutil.virtual_memory().available
        except Exception:
            avail = None
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        ok = not (avail is not None and avail < 1.15 * shard_bytes * local_world)
        if world > 1:  # one decision for all ranks: the end-to-end leg contains collectives
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok = bool(flag.item())
        if not ok:
            avail = avail or 0
            e2e_skip = f"host memory: {avail / 2**30:.0f} GiB available, {shard_bytes * local_world / 2**30:.0f} GiB needed for the shards of {local_world} ranks"
    if not args.no_e2e and e2e_skip is None:
        pinned = True
        try:
            host = torch.empty(1 << n_local, dtype=tdtype, pin_memory=True)
        except RuntimeError:
            host = torch.empty(1 << n_local, dtype=tdtype)
            pinned = False
        host.copy_(psi)
        torch.cuda.synchronize()
        host_np = host.numpy()
        if world == 1:
            del psi
            torch.cuda.empty_cache()
            e2e_step = lambda: eng.expectation_host(plan, host_np)
        else:
            def e2e_step():  # the shard is uploaded into psi inside the library call, behind the sweeps of the local class
                return se.expectation_host(plan, host_np, psi)
        e2e_step()  # warm-up (allocates the landing buffer)
        e2e_steps = max(1, min(args.steps, 2))
        ms2, _, energy2 = timed(e2e_step, e2e_steps)
        e2e = {
            "value": T * amps * e2e_steps / (ms2 / 1e3), "unit": UNIT, "h2d_bytes_per_step": shard_bytes * world, "d2h_bytes_per_step": 16 * world,
            "steps": e2e_steps, "ms_per_step": ms2 / e2e_steps, "energy_matches_device": bool(abs(energy2 - energy) <= 1e-10 * abs(energy)),
            "host_memory": "pinned" if pinned else "pageable",
        }

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu and world == 1:
        r = cpu_sample(13, 1, 1, 0)
        cpu = {
            "value": r["value"], "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"13-qubit member of the same generator ({r['T']} Pauli terms): oracle port of get_matrix(sparse=True) + vdot, one evaluation, {r['seconds_per_step']:.1f} s",
        }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": bench_config(world, n, T, plan), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "energy": energy.real, "energy_imag": energy.imag, "term_expansion_s": t_terms,
        "hbm_gbs_algorithmic": roofline["achieved"] * world,
    }
    if e2e_skip is not None:
        line["e2e_skipped"] = e2e_skip
    if se is not None:
        line["exchanges_per_step"] = se.exchanges // max(args.warmup + args.steps + (0 if e2e is None else 1 + e2e["steps"]), 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
