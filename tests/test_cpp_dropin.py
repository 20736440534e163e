"""The C++ drop-in class CCC::DdpCentroidal (centroidalcontrolcollection_b200/include/CCC/) driven by the
reference's own closed-loop test, restated in tests/cpp/TestDdpCentroidal.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "centroidalcontrolcollection_b200")
EXE = os.path.join(ROOT, "tests", "cpp", "test_ddp_centroidal.bin")


def _build():
    from centroidalcontrolcollection_b200 import build

    build.build()
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", EXE, os.path.join(ROOT, "tests", "cpp", "TestDdpCentroidal.cpp"),
           "-L" + PKG, "-lccc_b200", "-Wl,-rpath," + PKG, "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.check_call(cmd)


def test_cpp_dropin_compiles_and_refuses_without_gpu():
    """CPU tier: the header-only host classes compile against the C-ABI; without a device the
    program must fail loudly (no CPU fallback)."""
    _build()
    from centroidalcontrolcollection_b200 import engine

    if engine.lib().ccc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


@pytest.mark.gpu
def test_cpp_plan_once_closed_loop():
    """reference tests/src/TestDdpCentroidal.cpp:15-163 through CCC::DdpCentroidal::planOnce, plus
    planBatch == repeated planOnce, on the GPU."""
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    print(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ALL CHECKS PASSED" in r.stdout
