"""GPU: the typed C++17 host façade (include/tagg.hpp) reproduces the reference's unit tests through the C ABI."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_facade_reference_tests():
    exe = os.path.join(ROOT, "tests", "cpp", "test_reference.bin")
    assert os.path.exists(exe), "tests/cpp/test_reference.bin not built (build.sh)"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference tests passed" in r.stdout
