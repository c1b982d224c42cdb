"""The C++ host adapter above the C ABI (4c_b200/host): builds, links against the in-tree library and
(a) without a GPU fails loudly instead of falling back, (b) on a GPU evaluates a one-element hex8
block through Action::struct_calc_nlnstiff and rethrows det J <= 0 as an exception like the
reference."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "4c_b200", "host")


def _build():
    import __graft_entry__ as g
    g.build_library()
    subprocess.check_call(["make", "-C", HOST], stdout=subprocess.DEVNULL)


def test_host_adapter_builds_and_has_no_cpu_fallback():
    import torch
    _build()
    r = subprocess.run([os.path.join(HOST, "host_check")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert "HOST_ADAPTER_NO_GPU" in r.stdout


@pytest.mark.gpu
def test_host_adapter_one_element_on_gpu():
    _build()
    r = subprocess.run([os.path.join(HOST, "host_check")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "HOST_ADAPTER_OK" in r.stdout and "expected throw" in r.stdout
