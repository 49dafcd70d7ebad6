"""Build libklara_b200.so (sm_100a) in-tree with nvcc.  No GPU is needed to build.

    python klara.jl_b200/build.py [--force] [-j N]

Objects go to klara.jl_b200/_build/, the shared library to klara.jl_b200/lib/libklara_b200.so
(git-ignored, but shipped to the GPU box by gpurun).
"""
import argparse
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIBDIR = os.path.join(HERE, "lib")
# experiments: KLB_VARIANT=name KLB_EXTRA_FLAGS="-DKLB_RANDN_UNROLL=4" builds lib/libklara_b200_name.so
VARIANT = os.environ.get("KLB_VARIANT", "")
if VARIANT:
    OBJ = os.path.join(HERE, "_build_" + VARIANT)
LIB = os.path.join(LIBDIR, "libklara_b200%s.so" % ("_" + VARIANT if VARIANT else ""))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false",                      # never contract a*b+c behind our back: fma is always explicit
         "-Xcompiler", "-fPIC,-ffp-contract=off,-O2"] + os.environ.get("KLB_EXTRA_FLAGS", "").split()

HEADERS = ["klb_kernels.cuh", "klb_dense.cuh", "klb_dense_mma.cuh", "klb_hmc_ws.cuh", "klb_glm.cuh", "klb_math.h", "klb_tables.h", "../../include/klara_b200.h"]


def units():
    u = [("klb_api", "klb_api.cu", []), ("klb_aux", "klb_aux.cu", []),
         ("klb_init", "klb_kernels_inst.cu", ["-DKLB_INST_INIT"]), ("klb_dense", "klb_dense_inst.cu", []),
         ("klb_dense_mma", "klb_dense_mma_inst.cu", []), ("klb_glm", "klb_glm_inst.cu", [])]
    for smp in (0, 1, 2):
        for fma in (0, 1):
            u.append(("klb_chain_%d_%d" % (smp, fma), "klb_kernels_inst.cu",
                      ["-DKLB_INST_SAMPLER=%d" % smp, "-DKLB_INST_FMA=%d" % fma]))
    for fma in (0, 1):
        u.append(("klb_hmc_ws_%d" % fma, "klb_hmc_ws_inst.cu", ["-DKLB_INST_FMA=%d" % fma]))
    return u


def stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def compile_one(name, src, defs, force):
    obj = os.path.join(OBJ, name + ".o")
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and not stale(obj, deps):
        return obj, ""
    cmd = [NVCC] + FLAGS + defs + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (name, " ".join(cmd), r.stderr))
    return obj, r.stderr


def build(force=False, jobs=None, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    jobs = jobs or min(8, os.cpu_count() or 1)
    objs = []
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        futs = [ex.submit(compile_one, n, s, d, force) for n, s, d in units()]
        for f in futs:
            obj, err = f.result()
            objs.append(obj)
            if verbose and err:
                sys.stderr.write(err)
    if force or stale(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s" % r.stderr)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("-j", type=int, default=None)
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.j, a.v))
