#!/usr/bin/env python3
"""Relink the reference's UNCHANGED command-line front-end against the B200 engine.

    python integration/build_cli.py [--force]     ->  integration/_build/mcxcl

What is compiled, and from where:
  * from /root/reference/src, where the files lie (nothing is copied into this repository):
      mcxcl.c (main), mcx_utils.c, mcx_shapes.c, mcx_lang.c, mcx_tictoc.c, mcx_neurojson.cpp, mcx_mie.cpp,
      cjson/cJSON.c, ubj/ubjw.c and the zmat codecs -- i.e. FILES_COMMON of src/Makefile:41 plus libzmat,
      with the flags of src/Makefile:22-25 minus -DMCX_EMBED_CL (no OpenCL kernel text is embedded);
  * from this repository: integration/mcx_cuda_host.cpp, which takes the place of src/mcx_host.cpp, and
    integration/clstub/CL/cl.h, which takes the place of the OpenCL SDK header;
  * linked against mcxcl_b200/libmcxb200.so (rpath $ORIGIN/../../mcxcl_b200) instead of libOpenCL.

The reference's own build system is not used.  The output directory is git-ignored; the binary travels to
the GPU box with the snapshot (it has no /root/reference).
"""
import argparse
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("MCX_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "src")
OUT = os.path.join(HERE, "_build")

C_FILES = ["mcxcl.c", "mcx_utils.c", "mcx_shapes.c", "mcx_lang.c", "mcx_tictoc.c", "cjson/cJSON.c", "ubj/ubjw.c"]
CXX_FILES = ["mcx_neurojson.cpp", "mcx_mie.cpp"]
LZMA = ["LzmaEnc", "LzmaDec", "LzmaLib", "LzFind", "Bra", "BraIA64", "Alloc", "7zCrc", "7zCrcOpt", "CpuArch", "Lzma2Dec",
        "Lzma2DecMt", "Lzma2Enc", "MtCoder", "MtDec", "Xz", "XzCrc64", "XzCrc64Opt", "XzEnc", "XzDec", "Sha256", "Sha256Opt",
        "Delta", "Bra86", "7zStream", "LzFindMt", "LzFindOpt", "Threads"]
ZMAT_FILES = (["zmat/zmatlib.c", "zmat/miniz/miniz.c", "zmat/lz4/lz4.c", "zmat/lz4/lz4hc.c"] +
              ["zmat/easylzma/%s.c" % f for f in ("compress", "decompress", "lzma_header", "lzip_header", "common_internal")] +
              ["zmat/easylzma/lzma/%s.c" % f for f in LZMA])
# src/Makefile:21-25 (+ what `make all` adds at :47) without MCX_EMBED_CL
DEFS = ["-DUSE_OS_TIMER", "-DCL_SILENCE_DEPRECATION", "-DUSE_OPENCL", "-DMCX_OPENCL"]
INC = ["-I" + os.path.join(HERE, "clstub"), "-I" + SRC, "-I" + os.path.join(SRC, "zmat"), "-I" + os.path.join(SRC, "zmat", "easylzma"),
       "-I" + os.path.join(SRC, "zmat", "easylzma", "lzma"), "-I" + os.path.join(SRC, "ubj"), "-I" + os.path.join(ROOT, "include")]
# src/zmat/Makefile defaults: HAVE_ZLIB=no HAVE_LZMA=yes HAVE_LZ4=yes HAVE_ZSTD=no HAVE_BLOSC2=no
ZMAT_DEFS = ["-DNO_ZLIB", "-D_LARGEFILE64_SOURCE=1", "-DZMAT_USE_LZMA_SDK", "-DCOMPRESS_MF_MT", "-DNO_BLOSC2", "-DNO_ZSTD"]
ZMAT_INC = ["-I" + os.path.join(SRC, "zmat", d) for d in ("", "miniz", "easylzma", "easylzma/lzma", "lz4")]


def run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit("build_cli: command failed")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()
    exe = os.path.join(OUT, "mcxcl")
    if not os.path.isdir(SRC):
        print("build_cli: reference tree absent;", "keeping prebuilt " + exe if os.path.exists(exe) else "nothing to do")
        return 0
    os.makedirs(OUT, exist_ok=True)
    h = hashlib.sha256()
    for f in [os.path.join(HERE, "mcx_cuda_host.cpp"), os.path.join(HERE, "clstub", "CL", "cl.h"), os.path.abspath(__file__),
              os.path.join(ROOT, "include", "mcxb200.h"), os.path.join(HERE, "mexstub", "mex.h")] + \
            [os.path.join(SRC, f) for f in C_FILES + CXX_FILES + ["pmcxcl.cpp", "mcxlabcl.cpp"]]:
        h.update(open(f, "rb").read())
    stamp = os.path.join(OUT, "build.stamp")
    if not args.force and os.path.exists(exe) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        print("build_cli: up to date")
        return 0
    jobs, objs = [], []

    def add(cmd, src, tag=""):
        obj = os.path.join(OUT, tag + src.replace("/", "_").rsplit(".", 1)[0] + ".o")
        jobs.append(cmd + ["-c", os.path.join(SRC, src) if not os.path.isabs(src) else src, "-o", obj])
        objs.append(obj)

    for f in C_FILES:
        add(["gcc", "-std=c99", "-O2", "-w", "-m64"] + DEFS + INC, f)
    for f in CXX_FILES:
        add(["g++", "-O2", "-w", "-m64"] + DEFS + INC, f)
    for f in ZMAT_FILES:
        add(["gcc", "-O2", "-w", "-fPIC"] + ZMAT_DEFS + ZMAT_INC, f, tag="z_")
    obj = os.path.join(OUT, "mcx_cuda_host.o")
    jobs.append(["g++", "-std=c++17", "-O2", "-Wall", "-m64"] + DEFS + INC + ["-c", os.path.join(HERE, "mcx_cuda_host.cpp"), "-o", obj])
    objs.append(obj)
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(run, jobs))
    run(["g++", "-o", exe] + objs + ["-L" + os.path.join(ROOT, "mcxcl_b200"), "-lmcxb200", "-Wl,-rpath,$ORIGIN/../../mcxcl_b200",
                                    "-lm", "-pthread"])
    for o in objs:
        os.remove(o)
    print("build_cli: built", exe)
    mod = build_pmcxcl()
    print("build_cli: built", mod)
    print("build_cli: built", build_mexcheck())
    with open(stamp, "w") as f:
        f.write(h.hexdigest())
    return 0


def build_pmcxcl():
    """The reference's UNCHANGED Python front-end (src/pmcxcl.cpp, a pybind11 module) against the B200 engine:
    the `_pmcxcl` target of src/CMakeLists.txt:116-144 with mcx_host.cpp -> integration/mcx_cuda_host.cpp and
    OpenCL::OpenCL -> libmcxb200.so.  -DMCX_CONTAINER makes mcx_error raise instead of exit (src/mcx_utils.c:1298-1305)."""
    import sysconfig

    import pybind11
    defs = DEFS + ["-DMCX_CONTAINER", "-DPYBIND11_VERSION_MAJOR"]
    inc = INC + ["-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"]]
    mod = os.path.join(OUT, "_pmcxcl" + sysconfig.get_config_var("EXT_SUFFIX"))
    jobs, objs = [], []

    def add(cmd, src, tag="py_"):
        obj = os.path.join(OUT, tag + os.path.basename(src).rsplit(".", 1)[0] + ".o")
        jobs.append(cmd + ["-c", os.path.join(SRC, src) if not os.path.isabs(src) else src, "-o", obj])
        objs.append(obj)

    for f in ["mcx_utils.c", "mcx_shapes.c", "mcx_lang.c", "mcx_tictoc.c", "cjson/cJSON.c", "ubj/ubjw.c"]:
        add(["gcc", "-std=c99", "-O2", "-w", "-m64", "-fPIC"] + defs + inc, f)
    for f in ["mcx_mie.cpp", "mcx_neurojson.cpp", "pmcxcl.cpp"]:
        add(["g++", "-O2", "-w", "-m64", "-fPIC", "-fvisibility=hidden"] + defs + inc, f)
    for f in ZMAT_FILES:
        add(["gcc", "-O2", "-w", "-fPIC"] + ZMAT_DEFS + ZMAT_INC, f, tag="pyz_")
    add(["g++", "-std=c++17", "-O2", "-Wall", "-m64", "-fPIC"] + defs + inc, os.path.join(HERE, "mcx_cuda_host.cpp"))
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(run, jobs))
    run(["g++", "-shared", "-o", mod] + objs + ["-L" + os.path.join(ROOT, "mcxcl_b200"), "-lmcxb200", "-Wl,-rpath,$ORIGIN/../../mcxcl_b200",
                                               "-lm", "-pthread"])
    for o in objs:
        os.remove(o)
    return mod


def build_mexcheck():
    """The reference's UNCHANGED MATLAB / Octave front-end (src/mcxlabcl.cpp) compiled and linked against the B200 binding:
    the `mcxlabcl` target of src/CMakeLists.txt:146-188 (MCX_CONTAINER MATLAB_MEX_FILE) with mcx_host.cpp ->
    integration/mcx_cuda_host.cpp.  This image has neither MATLAB nor Octave, so <mex.h> is integration/mexstub/mex.h
    (declarations of the public MEX API only) and the result, integration/_build/mcxlabcl_check.so, keeps its mx* / mex*
    symbols undefined -- what a MEX file looks like before MATLAB loads it.  It cannot be run here; it shows that the file
    compiles and that the only symbols it needs from the host layer are the three the binding exports."""
    defs = DEFS + ["-DMCX_CONTAINER", "-DMATLAB_MEX_FILE"]
    inc = ["-I" + os.path.join(HERE, "mexstub")] + INC
    out = os.path.join(OUT, "mcxlabcl_check.so")
    jobs, objs = [], []

    def add(cmd, src, tag="mex_"):
        obj = os.path.join(OUT, tag + os.path.basename(src).rsplit(".", 1)[0] + ".o")
        jobs.append(cmd + ["-c", os.path.join(SRC, src) if not os.path.isabs(src) else src, "-o", obj])
        objs.append(obj)

    for f in ["mcx_utils.c", "mcx_shapes.c", "mcx_lang.c", "mcx_tictoc.c", "cjson/cJSON.c", "ubj/ubjw.c"]:
        add(["gcc", "-std=c99", "-O2", "-w", "-m64", "-fPIC"] + defs + inc, f)
    for f in ["mcx_mie.cpp", "mcx_neurojson.cpp", "mcxlabcl.cpp"]:
        add(["g++", "-O2", "-w", "-m64", "-fPIC"] + defs + inc, f)
    for f in ZMAT_FILES:
        add(["gcc", "-O2", "-w", "-fPIC"] + ZMAT_DEFS + ZMAT_INC, f, tag="mexz_")
    add(["g++", "-std=c++17", "-O2", "-Wall", "-m64", "-fPIC"] + defs + inc, os.path.join(HERE, "mcx_cuda_host.cpp"))
    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        list(ex.map(run, jobs))
    run(["g++", "-shared", "-o", out] + objs + ["-L" + os.path.join(ROOT, "mcxcl_b200"), "-lmcxb200", "-Wl,-rpath,$ORIGIN/../../mcxcl_b200",
                                               "-lm", "-pthread"])
    for o in objs:
        os.remove(o)
    return out


if __name__ == "__main__":
    sys.exit(main())
