"""Build mcxcl_b200/libmcxb200.so with nvcc for sm_100a (no GPU needed: nvcc cross-compiles).

    python -m mcxcl_b200.build [--force] [--jobs N]

The kernel instantiation groups (csrc/kernels_inst.cu, -DMCXB_INST_GROUP=k) are compiled in parallel.
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmcxb200.so")
OBJDIR = os.path.join(HERE, "build")
NGROUPS = 12
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I" + CSRC]


def _sources():
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    files.append(os.path.join(os.path.dirname(HERE), "include", "mcxb200.h"))
    files.append(os.path.abspath(__file__))
    return files


def _digest():
    h = hashlib.sha256()
    for f in _sources():
        h.update(open(f, "rb").read())
    return h.hexdigest()


def _run(job):
    cmd, log = job
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("nvcc failed: see " + log)
    return log


def build(force=False, jobs=None, verbose=True, defines=(), variant=None):
    """Build the engine.  `variant` + `defines` build an experimental copy (tuning knobs such as
    -DMCXB_BLOCK=128 -DMCXB_MINBLOCKS=8) into build/variants/<variant>/libmcxb200.so; load it by setting
    MCXB200_LIB (mcxcl_b200.abi).  The shipped library is always the default build."""
    objdir = OBJDIR if variant is None else os.path.join(OBJDIR, "variants", variant)
    lib = LIB if variant is None else os.path.join(objdir, "libmcxb200.so")
    stamp = os.path.join(objdir, "build.stamp")
    digest = _digest() + " " + " ".join(defines)
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == digest:
        if verbose:
            print("mcxcl_b200.build: up to date")
        return lib
    os.makedirs(objdir, exist_ok=True)
    flags = FLAGS + ["-D" + d for d in defines]
    work, objs = [], []
    for g in range(NGROUPS):
        obj = os.path.join(objdir, "kernels_g%d.o" % g)
        work.append(([NVCC] + flags + ["-DMCXB_INST_GROUP=%d" % g, "-c", os.path.join(CSRC, "kernels_inst.cu"), "-o", obj],
                     os.path.join(objdir, "kernels_g%d.log" % g)))
        objs.append(obj)
    for name in ("engine", "engine_multi", "testhooks"):
        obj = os.path.join(objdir, name + ".o")
        work.append(([NVCC] + flags + ["-c", os.path.join(CSRC, name + ".cu"), "-o", obj], os.path.join(objdir, name + ".log")))
        objs.append(obj)
    with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
        list(ex.map(_run, work))
    _run(([NVCC, "-shared", "-o", lib] + objs + ["-lcudart", "-ldl"], os.path.join(objdir, "link.log")))
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("mcxcl_b200.build: built", lib)
    return lib


def ptxas_summary():
    """registers / spills per kernel from the last build's logs (profiles/ keeps a copy)."""
    out = []
    for g in range(NGROUPS):
        log = os.path.join(OBJDIR, "kernels_g%d.log" % g)
        if not os.path.exists(log):
            continue
        name = None
        for line in open(log):
            if "Compiling entry function" in line:
                name = line.split("'")[1]
            elif "Used" in line and name:
                out.append("%s: %s" % (name, line.strip().replace("ptxas info    : ", "")))
                name = None
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("--variant", default=None)
    ap.add_argument("-D", dest="defines", action="append", default=[])
    a = ap.parse_args()
    build(force=a.force, jobs=a.jobs, defines=tuple(a.defines), variant=a.variant)
