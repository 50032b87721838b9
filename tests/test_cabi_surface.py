"""CPU checks of the C-ABI boundary: the library builds for sm_100a, loads, and exports every symbol
include/tatt_b200.h declares; argument validation reports through tatt_last_error() (no compute)."""
import ctypes
import os
import subprocess

import pytest


@pytest.fixture(scope="module")
def libpath():
    from tatt_b200 import build
    return build.build()


def test_every_declared_symbol_is_exported(libpath):
    from tatt_b200 import _cabi
    protos = _cabi.parse_header()
    assert len(protos) >= 40
    L = ctypes.CDLL(libpath)
    for name in protos:
        assert hasattr(L, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("tatt_")}
    assert set(protos) <= exported
    assert exported <= set(protos), "exported but undeclared: %s" % (exported - set(protos))


def test_version_arch_and_error_channel(libpath):
    from tatt_b200 import _cabi
    L = _cabi.lib()
    assert L.tatt_version() >= 100
    assert L.tatt_arch() == 1                       # compiled with -gencode arch=compute_100a,code=sm_100a
    rc = L.tatt_gemm(7, 0, None, 1, None, 1, None, 1, None, 4, 4, 4, 1, 0, 0, 0, 0, 0, 0, 0, None, 0, None)
    assert rc != 0 and "amode" in _cabi.last_error()
    with pytest.raises(RuntimeError, match="Lk"):
        _cabi.call("tatt_mha64_fwd", None, None, None, None, None, 1, 8, 40, 0.0, None, 0, None)
    with pytest.raises(RuntimeError, match="multiple of 4"):
        _cabi.call("tatt_conv2d_igemm", None, None, None, None, 1, 4, 4, 3, 8, 3, 3, 1, 1, 0, None, 0, None)


def test_sass_is_sm100a_only(libpath):
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "--list-elf", libpath], capture_output=True, text=True).stdout
    archs = {tok for l in out.splitlines() for tok in l.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs
