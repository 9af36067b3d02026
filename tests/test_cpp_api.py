"""The reference's Go tests restated in C++ against include/gkrb200.hpp (the C++ mirror of the Go API over the C ABI):
tests/cpp/test_api.cpp.  The CPU suite compiles it and runs the host-only part; the GPU suite runs everything."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_api.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(OUT_DIR, "test_api")
LIB_DIR = os.path.join(ROOT, "gkr-mimc_b200")


@pytest.fixture(scope="module")
def test_api_exe():
    import gkrb200
    gkrb200.lib()  # builds libgkrb200.so if it is missing
    deps = [SRC, os.path.join(ROOT, "include", "gkrb200.hpp"), os.path.join(ROOT, "include", "gkrb200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        os.makedirs(OUT_DIR, exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-o", EXE,
                               "-L" + LIB_DIR, "-lgkrb200", "-Wl,-rpath," + LIB_DIR])
    return EXE


def _run(exe, *args):
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = LIB_DIR + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    return subprocess.run([exe, *args], capture_output=True, text=True, timeout=600, env=env)


def test_cpp_mirror_compiles_and_host_tests_pass(test_api_exe):
    r = _run(test_api_exe, "--host-only")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host-only: 0 failure(s)" in r.stdout


@pytest.mark.gpu
def test_cpp_mirror_reference_tests_on_device(test_api_exe):
    """TestFold, TestFolding, TestWithCipherGate, TestWithMultiIdentity, TestGKR of the reference through the C++ API"""
    r = _run(test_api_exe)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all: 0 failure(s)" in r.stdout
