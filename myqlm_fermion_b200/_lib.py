"""ctypes binding of libqfe.so (the C ABI in include/qfe.h).

The CUDA library is the only compute path of this package: if it is missing or a call fails the
error is raised, never papered over with a CPU fallback.
"""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libqfe.so"

QFE_OK = 0
QFE_ERR_ARG, QFE_ERR_CUDA, QFE_ERR_OOM, QFE_ERR_UNSUPPORTED, QFE_ERR_STATE, QFE_ERR_COMM = -1, -2, -3, -4, -5, -6
ENC = {"jordan-wigner": 0, "bravyi-kitaev": 1, "parity": 2}
C128, C64 = 0, 1

c_void_pp = C.POINTER(C.c_void_p)
c_double_p = C.POINTER(C.c_double)
c_i64_p = C.POINTER(C.c_int64)
c_i32_p = C.POINTER(C.c_int32)
c_u64_p = C.POINTER(C.c_uint64)
c_u8_p = C.POINTER(C.c_uint8)
c_int_p = C.POINTER(C.c_int)

# name -> (restype, argtypes); mirrors include/qfe.h one to one
SIGNATURES = {
    "qfe_last_error": (C.c_char_p, []),
    "qfe_version": (C.c_int, []),
    "qfe_ctx_create": (C.c_int, [c_void_pp, C.c_int]),
    "qfe_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "qfe_ctx_sync": (C.c_int, [C.c_void_p]),
    "qfe_ctx_destroy": (C.c_int, [C.c_void_p]),
    "qfe_ctx_launch_count": (C.c_int, [C.c_void_p, c_i64_p]),
    "qfe_terms_from_integrals": (C.c_int, [C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double, C.c_int, c_void_pp]),
    "qfe_terms_from_fermion": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_u8_p, c_i32_p, c_i64_p, c_double_p, c_double_p, C.c_double, C.c_int, c_void_pp]),
    "qfe_terms_from_pauli": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_u64_p, c_u64_p, c_double_p, c_double_p, C.c_double, c_void_pp]),
    "qfe_terms_concat": (C.c_int, [C.c_void_p, c_void_pp, C.c_int64, c_void_pp]),
    "qfe_terms_info": (C.c_int, [C.c_void_p, c_int_p, c_i64_p, c_i64_p, c_i64_p]),
    "qfe_terms_export": (C.c_int, [C.c_void_p, c_u64_p, c_u64_p, c_double_p, c_i32_p, c_double_p]),
    "qfe_terms_destroy": (C.c_int, [C.c_void_p]),
    "qfe_plan_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_void_pp]),
    "qfe_plan_info": (C.c_int, [C.c_void_p, c_i64_p, c_i64_p, c_i64_p, c_i32_p, c_i32_p]),
    "qfe_plan_hermitian": (C.c_int, [C.c_void_p, c_int_p]),
    "qfe_plan_work": (C.c_int, [C.c_void_p, c_i64_p, c_i64_p, c_i64_p, c_i64_p]),
    "qfe_plan_geometry": (C.c_int, [C.c_void_p, c_int_p, c_int_p]),
    "qfe_plan_slots": (C.c_int, [C.c_void_p, c_i64_p]),
    "qfe_plan_classes": (C.c_int, [C.c_void_p, c_i32_p, C.c_int32]),
    "qfe_plan_destroy": (C.c_int, [C.c_void_p]),
    "qfe_expectation_class": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, c_double_p]),
    "qfe_expectation_class_part": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p]),
    "qfe_expectation": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_double_p]),
    "qfe_expectation_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_double_p]),
    "qfe_expectation_class0_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_double_p]),
    "qfe_apply_class": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "qfe_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "qfe_commutator_gradients": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_double_p]),
    "qfe_dev_phase_cycles": (C.c_int, [C.c_void_p, c_u64_p, C.c_int]),
    "qfe_lanczos": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, c_double_p, c_int_p, c_double_p]),
    "qfe_vec_dot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, c_double_p]),
    "qfe_vec_axpy": (C.c_int, [C.c_void_p, c_double_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "qfe_vec_scale": (C.c_int, [C.c_void_p, c_double_p, C.c_void_p, C.c_int64, C.c_int]),
    "qfe_vec_gram": (C.c_int, [C.c_void_p, c_void_pp, C.c_int, c_void_pp, C.c_int, C.c_int64, C.c_int, c_double_p]),
    "qfe_vec_fill_random": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_uint64]),
    "qfe_vec_product_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, c_double_p, C.c_int]),
    "qfe_vec_basis_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64]),
    "qfe_comm_unique_id": (C.c_int, [c_u8_p]),
    "qfe_comm_init": (C.c_int, [C.c_void_p, c_u8_p, C.c_int, C.c_int]),
    "qfe_comm_info": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int_p, c_i64_p, c_i64_p]),
    "qfe_comm_class_times": (C.c_int, [C.c_void_p, c_i32_p, c_double_p, C.c_int32, c_i32_p]),
    "qfe_comm_destroy": (C.c_int, [C.c_void_p]),
    "qfe_allreduce_sum": (C.c_int, [C.c_void_p, c_double_p, C.c_int]),
    "qfe_exchange_shard": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int]),
    "qfe_expectation_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, c_double_p]),
    "qfe_expectation_sharded_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, c_double_p]),
    "qfe_apply_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "qfe_commutator_gradients_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, c_double_p]),
    "qfe_vec_dot_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, c_double_p]),
    "qfe_lanczos_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int, c_double_p,
                                      c_int_p, c_double_p]),
    "qfe_pauli_rotations": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, c_u64_p, c_u64_p, c_double_p, C.c_void_p, C.c_int]),
}


class QfeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libqfe error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def load() -> C.CDLL:
    """Load libqfe.so (built in-tree by ``__graft_entry__.build()`` / ``make -C myqlm_fermion_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("QFE_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"{path} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()'). "
            "There is no CPU fallback."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != QFE_OK:
        msg = load().qfe_last_error()
        raise QfeError(rc, msg.decode() if msg else "")


def as_ptr(arr, ctype):
    """Pointer to a C-contiguous numpy array's data."""
    return arr.ctypes.data_as(C.POINTER(ctype))
