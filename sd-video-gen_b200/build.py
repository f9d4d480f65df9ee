"""Build libsdvg.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsdvg.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + [os.path.join(HERE, "..", "include", "sdvg.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    """Compile into a temporary file and rename it over libsdvg.so, under a file lock: in a torchrun launch every rank
    calls this, and nobody may dlopen a half-written library."""
    import fcntl
    if not force and not stale():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale():      # another process built it while we waited
                return LIB
            tmp = f"{LIB}.tmp.{os.getpid()}"
            cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                   "-shared", "-Xcompiler", "-fPIC,-O2", "-Xptxas", "-v" if verbose else "-O3",
                   "-o", tmp, os.path.join(CSRC, "sdvg_api.cu")]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed building libsdvg.so")
            os.replace(tmp, LIB)
            if verbose:
                sys.stderr.write(r.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
