#!/usr/bin/env python3
"""Build oracle/_ref/libmcxref.so: the REFERENCE kernel source compiled for the host CPU.

TEST INFRASTRUCTURE ONLY.  The reference sources are read where they lie under /root/reference
(never copied into the repository); the only derived text, a copy of mcx_core.cl with three
constructor macros rewritten for C++, is generated into oracle/_ref/, which is
git-ignored.  The resulting .so travels to the GPU box with the snapshot; /root/reference does not.

Usage: python oracle/build_ref.py [--force] [--jobs N]
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_KERNEL = os.environ.get("MCX_REFERENCE_KERNEL", "/root/reference/src/mcx_core.cl")

# same order as MCX_SRC_* (reference src/mcx_const.h:75-92) and sourceflag[] (src/mcx_host.cpp:55-60)
SOURCES = [
    ("pencil", "MCX_SRC_PENCIL"), ("isotropic", "MCX_SRC_ISOTROPIC"), ("cone", "MCX_SRC_CONE"),
    ("gaussian", "MCX_SRC_GAUSSIAN"), ("planar", "MCX_SRC_PLANAR"), ("pattern", "MCX_SRC_PATTERN"),
    ("fourier", "MCX_SRC_FOURIER"), ("arcsine", "MCX_SRC_ARCSINE"), ("disk", "MCX_SRC_DISK"),
    ("fourierx", "MCX_SRC_FOURIERX"), ("fourierx2d", "MCX_SRC_FOURIERX2D"), ("zgaussian", "MCX_SRC_ZGAUSSIAN"),
    ("line", "MCX_SRC_LINE"), ("slit", "MCX_SRC_SLIT"), ("pencilarray", "MCX_SRC_PENCILARRAY"),
    ("pattern3d", "MCX_SRC_PATTERN3D"), ("hyperboloid", "MCX_SRC_HYPERBOLOID_GAUSSIAN"), ("ring", "MCX_SRC_RING"),
]

# what the reference passes at its default optlevel for label media with atomics on
# (src/mcx_host.cpp:857-891), minus USE_MACRO_CONST so gcfg-> fields are read from the struct
BASEFLAGS = ["-DMED_TYPE=1", "-DUSE_ATOMIC", "-DMCX_USE_NATIVE"]
MEDIA_FORMATS = [97, 99, 100, 101, 102, 103, 104]    # MEDIA_2LABEL_SPLIT (SVMC), MEDIA_LABEL_HALF .. MEDIA_AS_SHORT (src/mcx_const.h:57-64)
CXX = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fpermissive", "-w", "-fPIC", "-fopenmp", "-x", "c++"]

PATCHES = [
    ("#define FLOAT4(a,b,c,d) ((float4)((a),(b),(c),(d)))", "#define FLOAT4(a,b,c,d) float4((a),(b),(c),(d))"),
    ("#define FLOAT3(a,b,c)   ((float3)((a),(b),(c)))", "#define FLOAT3(a,b,c)   float3((a),(b),(c))"),
    ("#define SHORT4(a,b,c,d) ((short4)((a),(b),(c),(d)))", "#define SHORT4(a,b,c,d) short4((a),(b),(c),(d))"),
]


def patched_kernel_text():
    text = open(REF_KERNEL).read()
    for old, new in PATCHES:
        if text.count(old) != 1:
            raise SystemExit("build_ref: expected exactly one occurrence of %r in %s" % (old, REF_KERNEL))
        text = text.replace(old, new)
    return text


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit("build_ref: command failed")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 4)
    args = ap.parse_args()
    lib = os.path.join(OUT, "libmcxref.so")
    if not os.path.exists(REF_KERNEL):
        if os.path.exists(lib):
            print("build_ref: reference tree absent, keeping prebuilt", lib)
            return 0
        print("build_ref: reference tree absent and no prebuilt library; nothing to do")
        return 0
    os.makedirs(OUT, exist_ok=True)
    text = patched_kernel_text()
    h = hashlib.sha256()
    h.update(text.encode())
    for f in ("clshim.h", "ref_variant.cpp", "ref_driver.cpp", "oracle_api.h", "build_ref.py", "../include/mcxb200.h"):
        h.update(open(os.path.join(HERE, f), "rb").read())
    stamp = os.path.join(OUT, "build.stamp")
    if not args.force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        print("build_ref: up to date")
        return 0
    with open(os.path.join(OUT, "mcx_core_patched.cl"), "w") as f:
        f.write(text)
    jobs, objs = [], []
    inc = ["-I" + HERE, "-I" + OUT]
    for name, macro in SOURCES:
        for refl in (0, 1):
            for det in (0, 1):
                suffix = "%s_r%d_d%d" % (name, refl, det)
                obj = os.path.join(OUT, "k_%s.o" % suffix)
                flags = BASEFLAGS + ["-D" + macro, "-DREF_SUFFIX=" + suffix]
                if refl:
                    flags.append("-DMCX_DO_REFLECTION")
                if det:
                    flags.append("-DMCX_SAVE_DETECTORS")
                jobs.append(CXX + flags + inc + ["-c", os.path.join(HERE, "ref_variant.cpp"), "-o", obj])
                objs.append(obj)
    # continuous media: the reference JIT passes -DMED_TYPE=<cfg->mediabyte> (src/mcx_host.cpp:882); pencil source only
    for med in MEDIA_FORMATS:
        for refl in (0, 1):
            for det in (0, 1):
                suffix = "pencil_m%d_r%d_d%d" % (med, refl, det)
                obj = os.path.join(OUT, "k_%s.o" % suffix)
                flags = [f for f in BASEFLAGS if not f.startswith("-DMED_TYPE")] + ["-DMED_TYPE=%d" % med, "-DMCX_SRC_PENCIL", "-DREF_SUFFIX=" + suffix]
                if refl:
                    flags.append("-DMCX_DO_REFLECTION")
                if det:
                    flags.append("-DMCX_SAVE_DETECTORS")
                jobs.append(CXX + flags + inc + ["-c", os.path.join(HERE, "ref_variant.cpp"), "-o", obj])
                objs.append(obj)
    drv = os.path.join(OUT, "ref_driver.o")
    jobs.append(CXX + BASEFLAGS + ["-DMCX_SRC_PENCIL", "-DMCX_DO_REFLECTION", "-DMCX_SAVE_DETECTORS"] + inc +
                ["-c", os.path.join(HERE, "ref_driver.cpp"), "-o", drv])
    objs.append(drv)
    with cf.ThreadPoolExecutor(max_workers=args.jobs) as ex:
        list(ex.map(run, jobs))
    run(["g++", "-shared", "-fopenmp", "-o", lib] + objs)
    for o in objs:
        os.remove(o)
    with open(stamp, "w") as f:
        f.write(h.hexdigest())
    print("build_ref: built", lib)
    return 0


if __name__ == "__main__":
    sys.exit(main())
