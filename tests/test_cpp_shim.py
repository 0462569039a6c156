"""The header-only C++ shim (include/cloudini_b200/cloudini.hpp) mirrors the reference's Cloudini:: API on the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "_shim_roundtrip")


def _compile(lib_built):
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "shim_roundtrip.cpp"),
           "-L" + os.path.dirname(lib_built), "-lcloudini_b200", "-Wl,-rpath," + os.path.dirname(lib_built), "-o", EXE]
    subprocess.check_call(cmd)


def test_shim_compiles_and_links(lib_built):
    _compile(lib_built)
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_shim_roundtrip_on_gpu(lib_built):
    _compile(lib_built)
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "shim_roundtrip: ok" in out.stdout
