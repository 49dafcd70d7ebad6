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
# objects live under gpurun_out/ (scratch: never pushed to the GPU box, never committed)
OBJROOT = os.path.join(os.path.dirname(HERE), "gpurun_out", "klb_build")
OBJ = os.path.join(OBJROOT, "default")
LIBDIR = os.path.join(HERE, "lib")
# experiments: KLB_VARIANT=name KLB_EXTRA_FLAGS="-DKLB_RANDN_UNROLL=4" builds lib/libklara_b200_name.so
VARIANT = os.environ.get("KLB_VARIANT", "")
if VARIANT:
    OBJ = os.path.join(OBJROOT, VARIANT)
LIB = os.path.join(LIBDIR, "libklara_b200%s.so" % ("_" + VARIANT if VARIANT else ""))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-fmad=false",                      # never contract a*b+c behind our back: fma is always explicit
         "-Xcompiler", "-fPIC,-ffp-contract=off,-O2"] + os.environ.get("KLB_EXTRA_FLAGS", "").split()

HEADERS = ["klb_kernels.cuh", "klb_dense.cuh", "klb_dense_mma.cuh", "klb_hmc_ws.cuh", "klb_nuts.cuh", "klb_glm.cuh", "klb_math.h", "klb_tables.h", "../../include/klara_b200.h"]


def units():
    u = [("klb_api", "klb_api.cu", []), ("klb_aux", "klb_aux.cu", []), ("klb_multi", "klb_multi.cu", []),
         ("klb_init", "klb_kernels_inst.cu", ["-DKLB_INST_INIT"]), ("klb_dense", "klb_dense_inst.cu", []),
         ("klb_dense_mma", "klb_dense_mma_inst.cu", []), ("klb_glm", "klb_glm_inst.cu", [])]
    for smp in (0, 1, 2):
        for fma in (0, 1):
            u.append(("klb_chain_%d_%d" % (smp, fma), "klb_kernels_inst.cu",
                      ["-DKLB_INST_SAMPLER=%d" % smp, "-DKLB_INST_FMA=%d" % fma]))
    for fma in (0, 1):
        u.append(("klb_hmc_ws_%d" % fma, "klb_hmc_ws_inst.cu", ["-DKLB_INST_FMA=%d" % fma]))
        u.append(("klb_nuts_%d" % fma, "klb_nuts_inst.cu", ["-DKLB_INST_FMA=%d" % fma]))
    return u


def stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


# Units whose cubin gets a scheduling-control post-pass between ptxas and fatbinary (tools/sass_patch.py):
# {unit-name prefix: sass_patch.py arguments}.  KLB_SASS_PATCH="" disables it, KLB_SASS_PATCH="--yield --stall 2" overrides.
# Default for the warp-specialised HMC kernels: every integer-pipe instruction (the producers' Philox / ziggurat work)
# gets a stall count of at least 2, so that a producer warp is not eligible in back-to-back cycles and the consumer
# warps of the same scheduler keep their every-other-cycle fp64 issue: +6 % leapfrog-steps/s on C3, bit-identical
# results (a larger stall count only delays an instruction).  Measured alternatives: profiles/r2_summary.md.
SASS_PATCH = {"klb_hmc_ws_": os.environ.get("KLB_SASS_PATCH", "--select intalu --stall 2").split()}
PATCHER = os.path.join(os.path.dirname(HERE), "tools", "sass_patch.py")


def compile_patched(name, cmd, patch_args):
    """nvcc -dryrun lists the sub-commands (cicc, ptxas, fatbinary, host gcc); they are replayed one by one and the
    cubin ptxas wrote is rewritten in place before fatbinary embeds it."""
    import re
    import shlex
    import tempfile
    tmp = tempfile.mkdtemp(prefix="klb_" + name + "_")
    r = subprocess.run(cmd + ["-dryrun", "--keep", "--keep-dir", tmp], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc -dryrun failed for %s:\n%s" % (name, r.stderr))
    env = dict(os.environ)
    log = ""
    for line in r.stderr.splitlines():
        if not line.startswith("#$ "):
            continue
        line = line[3:]
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)=(.*)$", line)
        if m and not line.startswith(("gcc", "cicc", "ptxas", "fatbinary", "cudafe++", "rm", '"')):
            val = m.group(2).strip()
            for k, v in env.items():
                val = val.replace("$" + k, v)
            env[m.group(1)] = val.strip('"') if m.group(1) in ("PATH", "LD_LIBRARY_PATH", "CICC_PATH", "TOP", "NVVMIR_LIBRARY_DIR") else val
            continue
        if line.startswith("rm "):
            continue
        rr = subprocess.run(line, shell=True, env=env, capture_output=True, text=True, cwd=CSRC)
        log += rr.stderr
        if rr.returncode != 0:
            raise RuntimeError("sub-command failed for %s:\n%s\n%s" % (name, line, rr.stderr))
        if line.startswith("ptxas "):
            cubin = shlex.split(line)[shlex.split(line).index("-o") + 1]
            pr = subprocess.run([sys.executable, PATCHER, cubin, cubin] + patch_args, capture_output=True, text=True)
            log += pr.stderr
            if pr.returncode != 0:
                raise RuntimeError("sass_patch failed for %s:\n%s" % (name, pr.stderr))
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)
    return log


def compile_one(name, src, defs, force):
    obj = os.path.join(OBJ, name + ".o")
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__), PATCHER]
    if not force and not stale(obj, deps):
        return obj, ""
    cmd = [NVCC] + FLAGS + defs + ["-c", os.path.join(CSRC, src), "-o", obj]
    patch = next((a for pre, a in SASS_PATCH.items() if name.startswith(pre) and a), None)
    if patch:
        try:
            return obj, compile_patched(name, cmd, patch)
        except Exception as e:           # the post-pass is an optimisation: never let it break the build
            sys.stderr.write("klb build: cubin post-pass failed for %s (%s); compiling without it\n" % (name, e))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (name, " ".join(cmd), r.stderr))
    return obj, r.stderr


def build(force=False, jobs=None, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    jobs = jobs or min(8, os.cpu_count() or 1)
    objs = []
    # experiments: KLB_VARIANT_UNITS="klb_chain_0_0,klb_chain_1_0" recompiles only those units for the variant and
    # links the default build's objects for the rest
    only = [u for u in os.environ.get("KLB_VARIANT_UNITS", "").split(",") if u] if VARIANT else []
    with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
        futs = []
        for n, s, d in units():
            if only and n not in only:
                objs.append(os.path.join(OBJROOT, "default", n + ".o"))
                continue
            futs.append(ex.submit(compile_one, n, s, d, force))
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
