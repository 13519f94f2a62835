"""Development aid (CPU only): dump the sweep plan of a synthetic molecular JW Hamiltonian and print the work the tile
kernel will do per category (kind, half-visit or not), as a model of shared-memory wavefronts and issue slots.
usage: python tools/plan_stats.py M [tile_bits] [n_gpus]   (n = 2M qubits; n_gpus > 1: the plan of rank 0 of a register sharded by its high qubits,
with the sweeps, X masks and terms of every x_hi class)"""
import ctypes as C
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np
import pandas as pd

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import models, vectorised  # noqa: E402  (development tool: the oracle supplies term sets without a GPU)

m = int(sys.argv[1]) if len(sys.argv) > 1 else 13
tile_bits = int(sys.argv[2]) if len(sys.argv) > 2 else 13
n_gpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1
n = 2 * m
n_local = n - (n_gpus.bit_length() - 1)
one, two = models.synthetic_integrals(m, seed=1234)
hpq, hpqrs = models.convert_to_h_integrals(one, two)
x, z, c, k = vectorised.expand_electronic(hpq, hpqrs, 0.0, "jordan-wigner")
print(f"n={n} T={len(x)} xmasks={len(np.unique(x))}")
subprocess.run(["make", "-C", str(ROOT / "tests" / "csrc")], check=True, capture_output=True)
lib = C.CDLL(str(ROOT / "tests" / "csrc" / "libplan_emul.so"))
P = C.POINTER
owner = np.zeros(len(x), np.int32)
const = np.array([k], np.complex128)
c = np.ascontiguousarray(c, np.complex128)
with tempfile.NamedTemporaryFile(suffix=".csv") as f:
    rc = lib.plan_emul_dump(C.c_int(n), C.c_int64(len(x)), x.ctypes.data_as(P(C.c_uint64)), z.ctypes.data_as(P(C.c_uint64)),
                            c.view(np.float64).ctypes.data_as(P(C.c_double)), owner.ctypes.data_as(P(C.c_int32)), C.c_int64(1),
                            const.view(np.float64).ctypes.data_as(P(C.c_double)), C.c_int(n_local), C.c_int(0), C.c_int(tile_bits), C.c_int(1920), C.c_int(512),
                            f.name.encode())
    assert rc == 0, rc
    df = pd.read_csv(f.name)
df["half"] = df.skipbit > 0
df["visits"] = np.where(df.half, 0.5, 1.0)
if n_gpus > 1:
    per_class = df.groupby("x_hi").agg(sweeps=("sweep", "nunique"), x_masks=("nt", "size"), terms=("nt", "sum"), half_visit_masks=("half", "sum"))
    per_class["masks_per_sweep"] = per_class.x_masks / per_class.sweeps
    print(per_class)
print(f"sweeps={df.sweep.nunique()} groups={len(df)} terms(padded)={df.nt.sum()}  ranges/sweep={df.groupby('sweep').range.nunique().mean():.1f}")
g = df.groupby(["kind", "half"]).agg(groups=("nt", "size"), terms=("nt", "sum"))
g["group_visits"] = g.groups * np.where(g.index.get_level_values("half"), 0.5, 1.0)
g["term_visits"] = g.terms * np.where(g.index.get_level_values("half"), 0.5, 1.0)
print(g)
print("total group visits (per amplitude):", g.group_visits.sum(), " term visits:", g.term_visits.sum())
# X bits in the register rows / lanes / row blocks
xw = df.xword.values
print("groups with no X bit on a row-block position (no half-visit possible):", int(((xw >> 9) == 0).sum()), " with X bits on register rows:", int(((xw & 15) != 0).sum()))
per = df.groupby("sweep").agg(groups=("nt", "size"), terms=("nt", "sum"))
print(per.describe().T)
df["xweight"] = [bin(int(v)).count("1") for v in xw]
print(df.groupby(["xweight", "kind"]).agg(groups=("nt", "size"), terms=("nt", "sum"), nt_mean=("nt", "mean"), nt_max=("nt", "max")))
