"""examples/registration.cpp (headless counterpart of the reference's icp_registration app):
pc8d .bin I/O -> ICPLMs x2 -> ICP<RC,WC> -> full-cloud ICPTransform, checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "registration")


def build_example():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples"), "CXX=g++"], stdout=subprocess.DEVNULL)
    return EXE


def test_example_compiles_on_cpu():
    assert os.path.exists(build_example())


def test_example_usage_and_io_errors(tmp_path):
    exe = build_example()
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0 and "usage" in r.stderr
    r = subprocess.run([exe, str(tmp_path / "nope1.bin"), str(tmp_path / "nope2.bin")], capture_output=True, text=True)
    assert r.returncode != 0 and "cannot open" in r.stderr
    short = tmp_path / "short.bin"
    short.write_bytes(b"\0" * 1024)
    r = subprocess.run([exe, str(short), str(short)], capture_output=True, text=True)
    assert r.returncode != 0 and "not a 640x480 pc8d frame" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("flags,rot,weighted", [((), "power", True), (("--svd",), "svd", True), (("--regular",), "power", False)])
def test_example_registers_a_room_pair_like_the_oracle(po, tmp_path, flags, rot, weighted):
    from icp_b200 import synth
    exe = build_example()
    c1, c2, _, _ = synth.room_pair(seed=1001)
    f1, f2 = tmp_path / "pc8d_1.bin", tmp_path / "pc8d_2.bin"
    synth.save_pc8d(str(f1), c1)
    synth.save_pc8d(str(f2), c2)
    pose, out = tmp_path / "pose.txt", tmp_path / "registered.bin"
    r = subprocess.run([exe, str(f1), str(f2), "--pose", str(pose), "--out", str(out), *flags], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Iterations" in r.stdout and "Rotation angle" in r.stdout and "Translation vector" in r.stdout
    lines = pose.read_text().split("\n")
    k = int(lines[0])
    T8 = np.array(lines[1].split(), np.float32)
    T16 = np.array([l.split() for l in lines[2:6]], np.float32)
    # oracle: same pipeline on the CPU
    F, M = po.get_lms(c1), po.get_lms(c2)
    ref = po.icp_register(F, M, 128, 128, 256, a=2e2, c=1e-6, rot=rot, weighted=weighted)
    assert k == ref["k"]
    assert np.array_equal(T8.view(np.uint32), ref["T"].view(np.uint32)), (T8, ref["T"])
    assert np.abs(T16 - ref["T16"]).max() <= 1e-5
    reg = np.fromfile(str(out), np.float32).reshape(-1, 8)
    want = po.transform_q(c2.reshape(-1, 8), ref["T"])
    assert np.array_equal(reg.view(np.uint32), want.view(np.uint32))
