"""Builds libia3b200.so (sm_100a) in-tree with nvcc.  `python -m imageanalysis3_b200.build`.

The library is the product: there is no CPU fallback, the Python layer raises if it is missing.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libia3b200.so")
SOURCES = ["capi.cu", "seed_kernels.cu", "fit_kernels.cu", "aux_kernels.cu", "corr_kernels.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant / defines: a developer build with extra -D flags into libia3b200_<variant>.so (select it with
    the IA3_LIB environment variable); the product library is the plain build."""
    nvcc = _nvcc()
    bdir = os.path.join(HERE, "_build" if not variant else os.path.join("_build", variant))
    lib = LIB if not variant else os.path.join(HERE, f"libia3b200_{variant}.so")
    os.makedirs(bdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")]
    headers.append(os.path.join(HERE, "..", "include", "ia3b200.h"))
    objs, jobs = [], []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(bdir, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or not _newer(obj, [sp] + headers):
            cmd = [nvcc, "-c", sp, "-o", obj, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                   "-Xptxas", "-v" if verbose else "-O3"] + ARCH + [f"-D{d}" for d in defines]
            jobs.append(cmd)
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if verbose or res.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr + "\n")
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed for " + cmd[2])
    if jobs or force or not os.path.exists(lib):
        cmd = [nvcc, "-shared", "-o", lib] + objs + ARCH + ["-Xcompiler", "-fPIC", "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return lib


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var[0] if var else None, defines=defs))
