"""The C++17 host mirror (include/turdb_cuda.hpp) — the native-language host side over the C ABI (the reference is
compiled code; its Rust toolchain is absent here).  Compiled with g++ against the in-tree library."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "turdb_b200")
BIN = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.bin")


def build_binary():
    from turdb_b200 import build as tb
    tb.build_library()
    src = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
    hdrs = [os.path.join(ROOT, "include", h) for h in ("turdb_cuda.hpp", "turdb_cuda.h")]
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < max(map(os.path.getmtime, [src, *hdrs])):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", BIN,
                               "-L", LIBDIR, "-lturdb_cuda", f"-Wl,-rpath,{LIBDIR}"])
    return BIN


def test_cpp_mirror_compiles_and_fails_loudly_without_a_device():
    exe = build_binary()
    import ctypes
    from turdb_b200 import _lib
    c = ctypes.c_int32(0)
    if _lib.load().turdb_cuda_device_count(ctypes.byref(c)) == 0 and c.value > 0:
        pytest.skip("a device is present: the full run below covers it")
    out = subprocess.run([exe, "--no-device"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "no-device checks ok" in out.stdout


@pytest.mark.gpu
def test_cpp_mirror_end_to_end(gpu_required):
    exe = build_binary()
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr + out.stdout
    assert "all checks ok" in out.stdout
