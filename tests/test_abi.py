"""CPU-side checks of the drop-in boundary: libqfe.so loads, exports exactly the entry points
include/qfe.h declares, the ctypes table mirrors the header, and argument errors come back as
status codes + messages (no compute call is made: there is no GPU here)."""

import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "qfe.h"


def _declared():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qfe_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from myqlm_fermion_b200 import _lib

    if not _lib.LIB_PATH.exists():
        import __graft_entry__ as g

        g.build()
    return _lib.load()


def test_header_and_ctypes_table_agree():
    from myqlm_fermion_b200 import _lib

    assert _declared() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    from myqlm_fermion_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [name for name in _declared() if name not in exported]
    assert not missing, missing
    # nothing but the C ABI is exported with C linkage under the qfe_ prefix
    extra = [s for s in exported if s.startswith("qfe_") and s not in _declared()]
    assert not extra, extra
    assert lib.qfe_version() >= 100


def test_header_compiles_as_c():
    """The boundary is plain C: the header must compile with a C compiler."""
    src = '#include "qfe.h"\nint main(void) { return QFE_OK + QFE_ENC_PARITY - 2 + QFE_C64 - 1; }\n'
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", str(HEADER.parent), "-x", "c", "-"], input=src, text=True, check=True)


def test_argument_errors_are_status_codes(lib):
    from myqlm_fermion_b200 import _lib

    # null handles never reach the device
    assert lib.qfe_terms_info(None, None, None, None, None) == _lib.QFE_ERR_ARG
    assert b"null" in lib.qfe_last_error()
    assert lib.qfe_plan_info(None, None, None, None, None, None) == _lib.QFE_ERR_ARG
    assert lib.qfe_terms_destroy(None) == 0
    assert lib.qfe_plan_destroy(None) == 0
    assert lib.qfe_ctx_destroy(None) == 0
    out = C.c_void_p()
    assert lib.qfe_ctx_create(None, 0) == _lib.QFE_ERR_ARG
    # without a GPU the context cannot be created: an error code, not a crash, and no fallback
    rc = lib.qfe_ctx_create(C.byref(out), 0)
    import torch

    if not torch.cuda.is_available():
        assert rc in (_lib.QFE_ERR_CUDA, _lib.QFE_ERR_ARG) and not out.value


def test_engine_refuses_to_run_without_gpu():
    import torch

    import myqlm_fermion_b200 as qfe

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        qfe.Engine(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing in the package may import it."""
    for path in (ROOT / "myqlm_fermion_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
