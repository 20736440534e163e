"""CPU tier: the C-ABI library loads and exports every symbol include/ccc_b200.h declares."""
import ctypes as C
import os
import re

import pytest

from centroidalcontrolcollection_b200 import _abi, engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "ccc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ccc_[a-z0-9_]+)\s*\(", src)))


def test_header_functions_are_exported():
    from centroidalcontrolcollection_b200 import build

    build.build()
    L = C.CDLL(engine.LIB_PATH)
    names = _declared_functions()
    assert "ccc_ddp_centroidal_solve" in names
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ccc_b200.h but not exported"


def test_struct_sizes_match_header():
    # sizes computed by hand from the header (LP64): keeps the ctypes mirror honest
    assert C.sizeof(_abi.DdpConfig) == 4 * 4 + 9 * 8 + 16 * 8 + 2 * 4 + 5 * 8
    assert C.sizeof(_abi.DdpResult) == 5 * 8 + 2 * 4 + 3 * 8
    assert C.sizeof(_abi.DdpCentroidalBatch) == 4 * 4 + 2 * 8 + 5 * 8 + 19 * 8 + 2 * 8 + 2 * 8
    # ccc_ddp_centroidal_loop_t: 4 int32, 3 double, 4 int32, 5 pointers, 10 + 9 + 2 double, pointer, 2 int32, 3 double
    assert C.sizeof(_abi.DdpCentroidalLoop) == 4 * 4 + 3 * 8 + 4 * 4 + 5 * 8 + 21 * 8 + 8 + 2 * 4 + 3 * 8
    assert C.sizeof(_abi.DdpCentroidalLoopResult) == 3 * 8
    # ccc_linear_mpc_xy_batch_t: 4 int32, 2 double, 7 pointers, 6 + 1 + 2 double, pointer; result: 9 pointers
    assert C.sizeof(_abi.LinearMpcXyBatch) == 4 * 4 + 2 * 8 + 7 * 8 + 9 * 8 + 8
    assert C.sizeof(_abi.LinearMpcXyResult) == 9 * 8


def test_default_config_matches_host_mirror():
    from centroidalcontrolcollection_b200 import build, problem

    build.build()
    c = _abi.DdpConfig()
    engine.lib().ccc_ddp_config_default(C.addressof(c))
    d = problem.ddp_config()
    for name, _ in _abi.DdpConfig._fields_:
        a, b = getattr(c, name), getattr(d, name)
        if name == "alpha":
            assert list(a) == list(b)
        else:
            assert a == b, name


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the engine must refuse loudly instead of computing on the CPU."""
    from centroidalcontrolcollection_b200 import build

    build.build()
    if engine.lib().ccc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(engine.EngineError):
        engine.DdpCentroidalEngine(10, 4)
