"""Builds libilqr_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libilqr_b200.so")
SOURCES = ["capi.cu", "kernels_lpt.cu", "kernels_chain.cu", "kernels_chain_fl.cu", "layout.cu", "pool.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libilqr_b200.so cannot be built (there is no CPU fallback)")
    return exe


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ilqr_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source into ilqr.jl_b200/libilqr_b200.so.  Returns the path."""
    if not force and not _stale():
        return LIB
    extra = os.environ.get("ILQR_NVCC_EXTRA", "").split()   # experiments: -DILQR_FWD1_STAGES=2 …
    # one nvcc per translation unit, in parallel, then one link step
    objdir = os.path.join(HERE, "..", "build")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas", "-v"] if verbose else [])

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        res = subprocess.run([_nvcc()] + compile_flags + ["-c", "-o", obj, os.path.join(CSRC, src)], capture_output=True, text=True)
        return src, obj, res

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(compile_one, SOURCES))
    for src, obj, res in results:
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, res.stdout, res.stderr))
        if verbose:
            print(res.stderr)
    res = subprocess.run([_nvcc(), "-shared", "-o", LIB] + [obj for _, obj, _ in results], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
