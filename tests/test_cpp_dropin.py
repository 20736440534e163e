"""The C++ drop-in host classes (centroidalcontrolcollection_b200/include/CCC/*.h) driven by the reference's own
test scenarios, restated in tests/cpp/*.cpp (GoogleTest and Eigen are absent: plain checks, exit code).

Two tiers, the same test programs in both:
  * CPU tier — TestHostModels needs no engine at all (discretisation, condensing, preview gains and the whole
    PreviewControlZmp closed loop, BASELINE config 1).  The other programs are linked against
    tests/cpp/oracle_engine_shim.cpp, which implements the C-ABI on the CPU oracle, so that the host-side logic
    (callback sampling, QP coefficient assembly, batching, post-processing) is checked without a GPU; and linked
    against the real libccc_b200.so they must refuse to run without a device (no CPU fallback in the product).
  * GPU tier — linked against libccc_b200.so and run on the device.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "centroidalcontrolcollection_b200")
CPP = os.path.join(ROOT, "tests", "cpp")
ORACLE = os.path.join(ROOT, "oracle")

ENGINE_TESTS = ["TestDdpCentroidal", "TestDdpSingleRigidBody", "TestDdpZmp", "TestZmpMpc", "TestLinearMpcXY", "TestLinearMpcZ", "TestClosedForm", "TestPreviewControlCentroidal", "TestStepMpc"]


def _compile(name, engine):
    """engine: None (host only), "gpu" (libccc_b200.so) or "oracle" (CPU shim)."""
    exe = os.path.join(CPP, f"{name}_{engine or 'host'}.bin")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, os.path.join(CPP, name + ".cpp")]
    if engine == "gpu":
        from centroidalcontrolcollection_b200 import build

        build.build()
        cmd += ["-L" + PKG, "-lccc_b200", "-Wl,-rpath," + PKG, "-L/usr/local/cuda/lib64", "-lcudart"]
    elif engine == "oracle":
        from oracle import binding

        binding.build()
        cmd += [os.path.join(CPP, "oracle_engine_shim.cpp"), "-L" + ORACLE, "-lccc_oracle", "-Wl,-rpath," + ORACLE, "-pthread"]
    subprocess.check_call(cmd)
    return exe


def _run(exe, timeout=900):
    r = subprocess.run([exe], capture_output=True, text=True, timeout=timeout)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "ALL CHECKS PASSED" in r.stdout


def test_host_models_and_preview_control_closed_loop():
    """StateSpaceModel / sequential extensions (reference TestStateSpaceModel, TestInvariantSequentialExtension,
    TestVariantSequentialExtension) and BASELINE config 1: TestPreviewControlZmp, single problem on the CPU."""
    _run(_compile("TestHostModels", None))


@pytest.mark.parametrize("mode", ["standin", "absent", "disabled"])
def test_eigen_interop_header(mode):
    """CCC/EigenInterop.h (conversions between the drop-in classes' std::array / std::vector and Eigen's vector types) in its
    three states: <Eigen/Core> found (a minimal stand-in, tests/cpp/eigen_standin: Eigen is not in this image), not found,
    and found but switched off with CCC_B200_NO_EIGEN."""
    exe = os.path.join(CPP, f"TestEigenInterop_{mode}.bin")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-o", exe, os.path.join(CPP, "TestEigenInterop.cpp")]
    if mode != "absent":
        cmd += ["-I" + os.path.join(CPP, "eigen_standin")]
    if mode == "disabled":
        cmd += ["-DCCC_B200_NO_EIGEN"]
    subprocess.check_call(cmd)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "ALL CHECKS PASSED" in r.stdout, r.stdout + r.stderr
    assert ("EigenInterop active" in r.stdout) == (mode == "standin")


@pytest.mark.parametrize("name", ENGINE_TESTS)
def test_cpp_host_logic_with_oracle_engine(name):
    """The reference scenarios through the C++ classes with the CPU oracle standing in for the engine."""
    _run(_compile(name, "oracle"))


def test_cpp_dropin_refuses_without_gpu():
    """Linked against the product library, a drop-in program must fail loudly without a device."""
    from centroidalcontrolcollection_b200 import engine

    exe = _compile("TestDdpCentroidal", "gpu")
    if engine.lib().ccc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ENGINE_TESTS)
def test_cpp_dropin_on_gpu(name):
    """reference tests/src/Test{DdpCentroidal,DdpSingleRigidBody,DdpZmp,LinearMpcZmp,IntrinsicallyStableMpc,
    LinearMpcXY,LinearMpcZ}.cpp closed loops through CCC::<Method>::planOnce on the GPU, plus planBatch == repeated planOnce."""
    _run(_compile(name, "gpu"))
