"""ctypes binding of the CPU oracle (oracle/libccc_oracle.so).

TEST INFRASTRUCTURE ONLY: import this from tests/, bench.py's cpu_baseline / --impl reference
legs and __graft_entry__.smoke() — never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libccc_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.ccc_oracle_ddp_centroidal_solve.restype = C.c_int32
        _LIB.ccc_oracle_ddp_centroidal_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        _LIB.ccc_oracle_ddp_config_default.argtypes = [C.c_void_p]
        _LIB.ccc_oracle_ddp_config_default.restype = None
        _LIB.ccc_oracle_centroidal_eval.restype = C.c_int32
        _LIB.ccc_oracle_centroidal_eval.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 10
        _LIB.ccc_oracle_hardware_threads.restype = C.c_int32
        _LIB.ccc_oracle_ddp_srb_solve.restype = C.c_int32
        _LIB.ccc_oracle_ddp_srb_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        _LIB.ccc_oracle_srb_eval.restype = C.c_int32
        _LIB.ccc_oracle_srb_eval.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 10
    return _LIB


def hardware_threads():
    return int(lib().ccc_oracle_hardware_threads())


def ddp_centroidal_solve(problem_set, cfg, trace_len=0, n_threads=1):
    """Run the oracle on a centroidalcontrolcollection_b200.problem.DdpCentroidalProblemSet."""
    res = problem_set.new_result(trace_len)
    bs, rs = problem_set.as_struct(), res.as_struct()
    rc = lib().ccc_oracle_ddp_centroidal_solve(C.addressof(bs), C.addressof(cfg), C.addressof(rs), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return res


def ddp_srb_solve(problem_set, cfg, trace_len=0, n_threads=1):
    """Run the oracle on a centroidalcontrolcollection_b200.problem.DdpSrbProblemSet."""
    res = problem_set.new_result(trace_len)
    bs, rs = problem_set.as_struct(), res.as_struct()
    rc = lib().ccc_oracle_ddp_srb_solve(C.addressof(bs), C.addressof(cfg), C.addressof(rs), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return res


def qp_solve(problem_set, n_threads=1):
    """Run the oracle on a centroidalcontrolcollection_b200.qp.QpProblemSet."""
    L = lib()
    L.ccc_oracle_qp_solve.restype = C.c_int32
    L.ccc_oracle_qp_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    res = problem_set.new_result()
    bs, rs = problem_set.as_struct(), res.as_struct()
    rc = L.ccc_oracle_qp_solve(C.addressof(bs), C.addressof(rs), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return res


def ddp_zmp_solve(problem_set, cfg, trace_len=0, n_threads=1):
    """Run the oracle on a centroidalcontrolcollection_b200.problem.DdpZmpProblemSet."""
    L = lib()
    L.ccc_oracle_ddp_zmp_solve.restype = C.c_int32
    L.ccc_oracle_ddp_zmp_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    res = problem_set.new_result(trace_len)
    bs, rs = problem_set.as_struct(), res.as_struct()
    rc = L.ccc_oracle_ddp_zmp_solve(C.addressof(bs), C.addressof(cfg), C.addressof(rs), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return res


def ddp_centroidal_closed_loop(loop, cfg, n_threads=1):
    """Run the oracle's closed loop on a centroidalcontrolcollection_b200.closed_loop.CentroidalLoop."""
    L = lib()
    L.ccc_oracle_ddp_centroidal_closed_loop.restype = C.c_int32
    L.ccc_oracle_ddp_centroidal_closed_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
    res = loop.new_result()
    ls, rs = loop.as_struct(), res.as_struct()
    rc = L.ccc_oracle_ddp_centroidal_closed_loop(C.addressof(ls), C.addressof(cfg), C.addressof(rs), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return res
