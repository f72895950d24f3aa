"""The C-ABI library loads and exports every symbol include/zigp.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "zigp.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(zigp_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported_and_bound():
    import __graft_entry__ as ge
    ge.build()
    from onoffgpf import _cabi
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "libzigp.so does not export %s" % s
        assert s in _cabi.SYMBOLS, "ctypes binding missing for %s" % s
    assert sorted(_cabi.SYMBOLS) == syms


def test_no_cpu_fallback():
    import torch
    from onoffgpf import _cabi, synth
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    X, Y = synth.make_data(32, 1, 0)
    with pytest.raises(_cabi.ZigpError):
        synth.build_model(X, Y, synth.make_params(4, 4, 1, 0))
    # the C ABI itself refuses as well
    h = ctypes.c_void_p()
    rc = _cabi.lib().zigp_create(ctypes.byref(h), 0, 4, 4, 1, 1e-6, 0)
    assert rc == -2 and b"no CPU fallback" in _cabi.lib().zigp_last_error(None)


def test_create_rejects_bad_arguments():
    from onoffgpf import _cabi
    h = ctypes.c_void_p()
    for args in [(0, 0, 4, 1, 1e-6, 0), (0, 4, 4, 0, 1e-6, 0), (0, 4, 4, 9, 1e-6, 0), (0, 4, 4, 1, -1.0, 0)]:
        assert _cabi.lib().zigp_create(ctypes.byref(h), *args) == -1
    assert _cabi.lib().zigp_theta_len(None) == 0 and _cabi.lib().zigp_destroy(None) == 0


def test_product_does_not_import_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/"""
    pkg = os.path.join(ROOT, "zero-inflated-gp_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dp, f)).read(), f
