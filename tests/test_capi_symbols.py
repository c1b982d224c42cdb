"""The C-ABI library builds for sm_100a, loads and exports every symbol include/b200_solid.h declares
(no compute without a GPU)."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
b200 = importlib.import_module("4c_b200")


def _declared():
    src = open(os.path.join(ROOT, "include", "b200_solid.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200s_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build_library()
    L = ctypes.CDLL(b200.capi.LIB_PATH)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    assert sorted(b200.capi.SYMBOLS) == names


def test_version_and_no_device_behaviour():
    L = b200.capi.load()
    assert b"sm_100a" in L.b200s_version()
    import torch
    if not torch.cuda.is_available():
        # the product has no CPU fallback: creating a context without a GPU fails loudly
        with pytest.raises(b200.capi.B200SolidError):
            b200.capi.Context(0)


def test_library_contains_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", b200.capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
