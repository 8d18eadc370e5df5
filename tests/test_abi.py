"""The C-ABI library loads on a CPU-only box and exports every symbol include/turdb_cuda.h declares."""
import ctypes
import os
import re

from turdb_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "turdb_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(turdb_cuda_\w+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    L = _lib.load()
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(L, n), f"{n} declared in turdb_cuda.h but not exported"
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    assert L.turdb_cuda_abi_version() == 2


def test_errors_without_device_or_arguments():
    L = _lib.load()
    assert L.turdb_cuda_index_create(None, 0, None) == _lib.ERR_INVALID_ARGUMENT
    assert b"null" in L.turdb_cuda_last_error()
    assert L.turdb_cuda_index_destroy(None) == _lib.OK
    c = ctypes.c_int32(-1)
    rc = L.turdb_cuda_device_count(ctypes.byref(c))
    assert rc in (_lib.OK, _lib.ERR_NO_DEVICE) and c.value >= 0


def test_no_cpu_fallback_in_product_package():
    """The product package must not import the oracle (tier rule: oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "turdb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "libturdb_oracle" not in text and "hnsw_oracle" not in text, f"{f} links the oracle"


def test_sass_shows_the_blackwell_native_paths():
    """SASS evidence (B200_PROFILING.md): tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA tensor loads -> UTMALDG, bulk
    copies / prefetch of the traversal -> UBLKCP / UBLKPF; the library is built for sm_100a only."""
    import shutil
    import subprocess
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        import pytest
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in out
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UBLKPF", "FFMA2"):
        assert mnemonic in out, f"{mnemonic} missing from the SASS of libturdb_cuda.so"
    import re
    # no Hopper wgmma, no legacy mma.sync tensor path (UTCHMMA[.2CTA] is the tcgen05 instruction, not HMMA)
    assert "HGMMA" not in out and not re.search(r"(?<![A-Z])HMMA\.", out)
    assert "UTCHMMA.2CTA" in out and "UTMALDG.2D.2CTA" in out  # the exact path's two-CTA (cta_group::2) form
