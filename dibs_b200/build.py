"""Build recipe for the native library (nvcc, sm_100a only, in-tree output).

    python -m dibs_b200.build [--force] [-v]      # -> dibs_b200/libdibs_b200.so

The register-tiled Monte-Carlo kernels are instantiated per (family, DMAX) in separate objects
(`mc_*_inst.cu` compiled with -DDIBS_DMAX=n) so the build parallelises over the host cores.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
OUT = os.path.join(HERE, "libdibs_b200.so")
HEADER = os.path.join(HERE, "..", "include", "dibs_b200.h")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]

DMAX = {"lingauss": [8, 16, 20, 32, 64, 128], "nn": [8, 16, 20, 32, 64, 128], "bge": [8, 16, 20, 32, 64]}


def units():
    u = [("dibs_abi.o", "dibs_abi.cu", [])]
    for fam, dms in DMAX.items():
        for dm in dms:
            u.append((f"mc_{fam}_{dm}.o", f"mc_{fam}_inst.cu", [f"-DDIBS_DMAX={dm}"]))
    return u


def _newest_dep():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER, os.path.abspath(__file__)]
    return max(os.path.getmtime(d) for d in deps)


def build(force=False, verbose=False, jobs=None):
    os.makedirs(OBJ, exist_ok=True)
    newest = _newest_dep()
    todo = []
    for obj, src, defs in units():
        o = os.path.join(OBJ, obj)
        if force or not os.path.exists(o) or os.path.getmtime(o) < newest:
            todo.append((o, os.path.join(CSRC, src), defs))

    def compile_one(job):
        o, src, defs = job
        cmd = [NVCC] + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + defs + ["-c", src, "-o", o]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return job, res

    if todo:
        with ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            for job, res in ex.map(compile_one, todo):
                if res.returncode != 0:
                    sys.stderr.write(res.stdout + res.stderr)
                    raise RuntimeError(f"nvcc failed on {job[1]} {job[2]}")
                if verbose:
                    sys.stderr.write(f"== {os.path.basename(job[0])}\n{res.stderr}")
    objs = [os.path.join(OBJ, u[0]) for u in units()]
    if todo or not os.path.exists(OUT):
        cmd = [NVCC] + ARCH + ["-shared", "-cudart", "static"] + objs + ["-o", OUT, "-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc link failed for libdibs_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
