"""The C++ host classes (include/percept_b200/ReferenceMeshSmootherConjugateGradient.hpp) keep the reference's
smoother class API; tests/cpp/test_host_api.cpp reads like the reference's own gtest cases."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "test_host_api.bin")


def _compile():
    lib = os.path.join(ROOT, "percept_b200", "lib")
    cmd = ["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_host_api.cpp"), "-L" + lib, "-lpms", "-Wl,-rpath," + lib, "-o", EXE]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return EXE


def test_host_api_compiles_against_the_c_abi():
    _compile()


@pytest.mark.gpu
def test_host_api_reference_style_tests_pass_on_gpu(tmp_path):
    exe = _compile()
    p = subprocess.run([exe], capture_output=True, text=True, cwd=tmp_path, timeout=900)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "ALL PASSED" in p.stdout
