"""Builds and runs tests/cpp/test_api.cpp: the reference's own test pattern (construct . init . write . run . read)
against the C++ drop-in classes of include/ICP/algorithms.hpp, checked with the CPU oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_cpp_test():
    exe = os.path.join(ROOT, "tests", "cpp", "test_api")
    src = os.path.join(ROOT, "tests", "cpp", "test_api.cpp")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
           "-L", os.path.join(ROOT, "icp_b200"), "-licp_b200", "-L", os.path.join(ROOT, "oracle"), "-licp_oracle",
           "-Wl,-rpath," + os.path.join(ROOT, "icp_b200"), "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-ldl", "-lpthread", "-lrt"]
    subprocess.check_call(cmd)
    return exe


def test_cpp_header_compiles_on_cpu(po):
    """not-gpu: the drop-in header and its test program compile and link against the C ABI."""
    assert os.path.exists(build_cpp_test())


@pytest.mark.gpu
def test_cpp_api_matches_oracle(po):
    exe = build_cpp_test()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "ALL C++ API CHECKS PASSED" in r.stdout
