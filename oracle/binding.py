"""ctypes binding of the CPU oracle (oracle/libccc_oracle.so).

TEST INFRASTRUCTURE ONLY: import this from tests/, bench.py's cpu_baseline / --impl reference
legs and __graft_entry__.smoke() — never from the product package.

Two builds of the same sources exist (oracle/Makefile): the canonical arithmetic the engine
reproduces bit for bit (module-level functions below) and the textbook arithmetic
(`textbook()` returns the same functions bound to libccc_oracle_textbook.so).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# num.hpp `Choice` bits
CHOICE_K_REL_ELEMENTWISE = 1
CHOICE_BOXQP_COLD_START = 2
CHOICE_VXX_REGULARISED = 4
CHOICE_DESCENT_TOL = 8
CHOICE_BOXQP_WARM_SAME_STAGE = 16


def build():
    import fcntl

    with open(os.path.join(_HERE, ".build.lock"), "w") as lock:  # concurrent ranks / test workers build once
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            subprocess.check_call(["make", "-s", "-C", _HERE])
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


class Oracle:
    """One loaded oracle library."""

    def __init__(self, name):
        self.path = os.path.join(_HERE, name)
        self._lib = None

    def lib(self):
        if self._lib is None:
            if not os.path.exists(self.path):
                build()
            L = C.CDLL(self.path)
            for fn in ("ccc_oracle_ddp_centroidal_solve", "ccc_oracle_ddp_srb_solve", "ccc_oracle_ddp_zmp_solve",
                       "ccc_oracle_ddp_centroidal_closed_loop"):
                getattr(L, fn).restype = C.c_int32
                getattr(L, fn).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
            L.ccc_oracle_ddp_config_default.argtypes = [C.c_void_p]
            L.ccc_oracle_ddp_config_default.restype = None
            for fn in ("ccc_oracle_centroidal_eval", "ccc_oracle_srb_eval"):
                getattr(L, fn).restype = C.c_int32
                getattr(L, fn).argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 10
            L.ccc_oracle_hardware_threads.restype = C.c_int32
            L.ccc_oracle_qp_solve.restype = C.c_int32
            L.ccc_oracle_qp_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
            L.ccc_oracle_linear_mpc_xy_solve.restype = C.c_int32
            L.ccc_oracle_linear_mpc_xy_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
            L.ccc_oracle_set_choices.argtypes = [C.c_uint32]
            L.ccc_oracle_set_choices.restype = None
            L.ccc_oracle_is_textbook.restype = C.c_int32
            self._lib = L
        return self._lib

    def hardware_threads(self):
        return int(self.lib().ccc_oracle_hardware_threads())

    def set_choices(self, bits):
        """Select unpinned algorithmic alternatives (CHOICE_* bits); 0 restores the oracle's definition."""
        self.lib().ccc_oracle_set_choices(int(bits))

    def _ddp(self, fn, problem_set, cfg, trace_len, n_threads):
        res = problem_set.new_result(trace_len)
        bs, rs = problem_set.as_struct(), res.as_struct()
        rc = getattr(self.lib(), fn)(C.addressof(bs), C.addressof(cfg), C.addressof(rs), int(n_threads))
        if rc != 0:
            raise RuntimeError(f"oracle returned {rc}")
        return res

    def ddp_centroidal_solve(self, problem_set, cfg, trace_len=0, n_threads=1):
        """Run the oracle on a centroidalcontrolcollection_b200.problem.DdpCentroidalProblemSet."""
        return self._ddp("ccc_oracle_ddp_centroidal_solve", problem_set, cfg, trace_len, n_threads)

    def ddp_srb_solve(self, problem_set, cfg, trace_len=0, n_threads=1):
        """Run the oracle on a centroidalcontrolcollection_b200.problem.DdpSrbProblemSet."""
        return self._ddp("ccc_oracle_ddp_srb_solve", problem_set, cfg, trace_len, n_threads)

    def ddp_zmp_solve(self, problem_set, cfg, trace_len=0, n_threads=1):
        """Run the oracle on a centroidalcontrolcollection_b200.problem.DdpZmpProblemSet."""
        return self._ddp("ccc_oracle_ddp_zmp_solve", problem_set, cfg, trace_len, n_threads)

    def qp_solve(self, problem_set, n_threads=1):
        """Run the oracle on a centroidalcontrolcollection_b200.qp.QpProblemSet."""
        res = problem_set.new_result()
        bs, rs = problem_set.as_struct(), res.as_struct()
        rc = self.lib().ccc_oracle_qp_solve(C.addressof(bs), C.addressof(rs), int(n_threads))
        if rc != 0:
            raise RuntimeError(f"oracle returned {rc}")
        return res

    def linear_mpc_xy_solve(self, sweep, n_threads=1, intermediates=True):
        """Run the oracle on a centroidalcontrolcollection_b200.linear_mpc_xy.XySweepProblemSet."""
        res = sweep.new_result(intermediates)
        bs, rs = sweep.as_struct(), res.as_struct()
        rc = self.lib().ccc_oracle_linear_mpc_xy_solve(C.addressof(bs), C.addressof(rs), int(n_threads))
        if rc != 0:
            raise RuntimeError(f"oracle returned {rc}")
        return res

    def ddp_centroidal_closed_loop(self, loop, cfg, n_threads=1):
        """Run the oracle's closed loop on a centroidalcontrolcollection_b200.closed_loop.CentroidalLoop."""
        res = loop.new_result()
        ls, rs = loop.as_struct(), res.as_struct()
        rc = self.lib().ccc_oracle_ddp_centroidal_closed_loop(C.addressof(ls), C.addressof(cfg), C.addressof(rs),
                                                              int(n_threads))
        if rc != 0:
            raise RuntimeError(f"oracle returned {rc}")
        return res


_CANON = Oracle("libccc_oracle.so")
_TEXTBOOK = Oracle("libccc_oracle_textbook.so")


def textbook():
    """The textbook-arithmetic build (no fma, plain sums, LLT with sqrt, divided Armijo ...; num.hpp)."""
    assert _TEXTBOOK.lib().ccc_oracle_is_textbook() == 1
    return _TEXTBOOK


lib = _CANON.lib
hardware_threads = _CANON.hardware_threads
set_choices = _CANON.set_choices
ddp_centroidal_solve = _CANON.ddp_centroidal_solve
ddp_srb_solve = _CANON.ddp_srb_solve
ddp_zmp_solve = _CANON.ddp_zmp_solve
qp_solve = _CANON.qp_solve
linear_mpc_xy_solve = _CANON.linear_mpc_xy_solve
ddp_centroidal_closed_loop = _CANON.ddp_centroidal_closed_loop
