"""Builds asvd4llm_b200/csrc/libasvd_b200.so (sm_100a only) with nvcc.  In-tree so the .so travels with the
repository snapshot to the GPU box; rebuilt automatically when a source is newer than the library."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libasvd_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v", "-lcuda"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = (sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))
            + glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    # several ranks may find the library stale at once (torchrun): one builds, the others wait on the lock and then
    # see a fresh library; the result is moved into place atomically
    import fcntl
    with open(os.path.join(CSRC, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not stale():
                return LIB
            tmp = LIB + f".tmp{os.getpid()}"
            cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp] + sources()
            proc = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or proc.returncode != 0:
                sys.stderr.write(proc.stdout + proc.stderr)
            if proc.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed building libasvd_b200.so:\n" + proc.stderr[-4000:])
            os.replace(tmp, LIB)
            with open(os.path.join(CSRC, "build.log"), "w") as f:
                f.write(" ".join(cmd).replace(tmp, LIB) + "\n" + proc.stdout + proc.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
