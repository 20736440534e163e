"""Loader of the warp emulator (tests/emu): runs the CUDA solver-core *source* on the CPU."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "emu")])
        _LIB = C.CDLL(os.path.join(_HERE, "emu", "libccc_emu.so"))
        _LIB.ccc_emu_ddp_centroidal_solve.restype = C.c_int32
        _LIB.ccc_emu_ddp_centroidal_solve.argtypes = [C.c_void_p] * 3
        _LIB.ccc_emu_ddp_srb_solve.restype = C.c_int32
        _LIB.ccc_emu_ddp_srb_solve.argtypes = [C.c_void_p] * 3
    return _LIB


def ddp_centroidal_solve(problem_set, cfg, trace_len=0, chunk=0, feat=1, team=0):
    """feat: 0 = no feature bits, 1 = product default, 2 = every feature of the solver core (ddp_warp_core.cuh kFeat*).
    team = 1: the small-batch kernel (ddp_team.cuh, a CTA of 8 warps per problem) instead of a warp per problem."""
    lib().ccc_emu_set_chunk(int(chunk))
    lib().ccc_emu_set_feat(int(feat))
    lib().ccc_emu_set_team(int(team))
    res = problem_set.new_result(trace_len)
    bs, rs = problem_set.as_struct(), res.as_struct()
    rc = lib().ccc_emu_ddp_centroidal_solve(C.addressof(bs), C.addressof(cfg), C.addressof(rs))
    assert rc == 0
    return res


def ddp_srb_solve(problem_set, cfg, trace_len=0, chunk=0, feat=1, team=0):
    lib().ccc_emu_set_chunk(int(chunk))
    lib().ccc_emu_set_feat(int(feat))
    lib().ccc_emu_set_team(int(team))
    res = problem_set.new_result(trace_len)
    bs, rs = problem_set.as_struct(), res.as_struct()
    rc = lib().ccc_emu_ddp_srb_solve(C.addressof(bs), C.addressof(cfg), C.addressof(rs))
    assert rc == 0
    return res


def qp_solve(problem_set, rcap=0):
    """rcap > 0: as the engine, a first pass with a packed R of `rcap` columns and the full-R pass for the problems
    that outgrow it (last_qp_overflows() tells how many did)."""
    L = lib()
    L.ccc_emu_qp_set_rcap(int(rcap))
    L.ccc_emu_qp_solve.restype = C.c_int32
    L.ccc_emu_qp_solve.argtypes = [C.c_void_p] * 2
    res = problem_set.new_result()
    bs, rs = problem_set.as_struct(), res.as_struct()
    assert L.ccc_emu_qp_solve(C.addressof(bs), C.addressof(rs)) == 0
    return res


def ddp_zmp_solve(problem_set, cfg, trace_len=0, chunk=0, feat=1):
    L = lib()
    L.ccc_emu_set_feat(int(feat))
    L.ccc_emu_set_team(0)
    L.ccc_emu_ddp_zmp_solve.restype = C.c_int32
    L.ccc_emu_ddp_zmp_solve.argtypes = [C.c_void_p] * 3
    L.ccc_emu_set_chunk(int(chunk))
    res = problem_set.new_result(trace_len)
    bs, rs = problem_set.as_struct(), res.as_struct()
    assert L.ccc_emu_ddp_zmp_solve(C.addressof(bs), C.addressof(cfg), C.addressof(rs)) == 0
    return res


def ddp_zmp_thread_solve(problem_set, cfg, trace_len=0):
    """csrc/ddp_thread_zmp.cuh (one thread per problem) compiled for the host."""
    L = lib()
    L.ccc_emu_ddp_zmp_thread_solve.restype = C.c_int32
    L.ccc_emu_ddp_zmp_thread_solve.argtypes = [C.c_void_p] * 3
    res = problem_set.new_result(trace_len)
    bs, rs = problem_set.as_struct(), res.as_struct()
    assert L.ccc_emu_ddp_zmp_thread_solve(C.addressof(bs), C.addressof(cfg), C.addressof(rs)) == 0
    return res


def last_qp_overflows():
    return int(lib().ccc_emu_qp_last_overflows())
