"""Build recipe for liboclcg.so (nvcc, sm_100a only) -- the CMakeLists.txt:15-19 of the reference.

    python conjugate-gradient-pyopencl_b200/build.py [--force]

Outputs (all git-ignored, all travel to the GPU box with the snapshot):
    conjugate-gradient-pyopencl_b200/liboclcg.so   the library the package loads
    build/liboclcg.so                              where p_h-PY_C-CL.py:38 looks for it
    build/oclcgex                                  the example executable (main.c), from csrc/oclcgex.c
    liboclcg.so                                    where p_helmholtz.py:29 looks for it
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "liboclcg.so")
SOURCES = ["cgb200.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared",
    "--cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: liboclcg.so cannot be built (there is no CPU fallback)")


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include"))]
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def _replace_copy(src, dst):
    """Copy through a temporary name + os.replace: a reader never sees a half-written file."""
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    tmp = f"{dst}.tmp.{os.getpid()}"
    shutil.copy2(src, tmp)
    os.replace(tmp, dst)


def build(force=False, verbose=False):
    """Compile (if stale) and place the copies the reference drivers expect.  Returns the path.

    Safe under torchrun: the ranks of one box serialise on a file lock, the first one in builds into a
    temporary file and renames it into place, the others find a fresh library."""
    import fcntl
    import hashlib
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    with open(os.path.join(ROOT, "build", ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if force or _stale():
                tmp = f"{LIB}.tmp.{os.getpid()}"
                cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                      ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
                # the image exports CC/CXX pointing at a wrapper without OpenMP specs; nvcc wants plain g++
                env = dict(os.environ)
                env.pop("CC", None)
                env.pop("CXX", None)
                subprocess.check_call(cmd, cwd=CSRC, env=env)
                os.replace(tmp, LIB)
                # so that a build log proves which binary a run used
                with open(LIB, "rb") as f:
                    digest = hashlib.sha256(f.read()).hexdigest()[:16]
                print("cgb200 build:", " ".join(cmd[:-3] + ["-o", LIB] + cmd[-1:]), f"-> sha256 {digest}", file=sys.stderr)
            for dst in (os.path.join(ROOT, "build", "liboclcg.so"), os.path.join(ROOT, "liboclcg.so")):
                if not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(LIB):
                    _replace_copy(LIB, dst)
            build_example()
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


EXE = os.path.join(ROOT, "build", "oclcgex")


def build_example(force=False):
    """build/oclcgex -- the reference's example executable (main.c; CMakeLists.txt:16), linked against
    build/liboclcg.so (found at run time through $ORIGIN)."""
    src = os.path.join(CSRC, "oclcgex.c")
    if not force and os.path.exists(EXE) and os.path.getmtime(EXE) >= max(os.path.getmtime(src), os.path.getmtime(LIB)):
        return EXE
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    tmp = f"{EXE}.tmp.{os.getpid()}"
    subprocess.check_call([gcc, "-O2", "-std=c11", "-Wall", "-o", tmp, src, "-L" + os.path.join(ROOT, "build"),
                           "-loclcg", "-lm", "-Wl,-rpath,$ORIGIN"])
    os.replace(tmp, EXE)
    return EXE


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
