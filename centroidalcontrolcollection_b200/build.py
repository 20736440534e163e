"""Build libccc_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libccc_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # only explicit fma() may fuse: canonical arithmetic, DESIGN.md §4
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(HERE, "csrc", "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(HERE, "csrc", "*.cuh")) + [os.path.join(HERE, "..", "include", "ccc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, ab_variants=None):
    """ab_variants (or CCC_AB_VARIANTS=1 in the environment): also compile the A/B builds of the DDP solver core's
    feature bits (ddp_host.cuh Variants), selectable with ccc_ddp_centroidal_set_variant — measurement builds only."""
    if ab_variants is None:
        ab_variants = os.environ.get("CCC_AB_VARIANTS", "0") not in ("", "0")
    if not force and not needs_build():
        return LIB
    # Several processes may get here at once (one rank per GPU under torchrun): one builds, the others wait on the
    # lock and find the library up to date; the library appears by an atomic rename, never half written.
    import fcntl

    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            cmd = [nvcc] + NVCC_FLAGS + (["-DCCC_AB_VARIANTS"] if ab_variants else []) + (["-Xptxas", "-v"] if verbose else [])
            cmd += os.environ.get("CCC_EXTRA_NVCC_FLAGS", "").split()  # A/B measurement builds (e.g. -DCCC_NO_GAIN_PREFETCH)
            tmp = f"{LIB}.tmp.{os.getpid()}"
            cmd += ["-t", "0", "-o", tmp] + sources()
            try:
                subprocess.check_call(cmd)
                os.replace(tmp, LIB)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
