"""The reference's UNCHANGED MATLAB / Octave front-end (src/mcxlabcl.cpp) compiled and linked against the B200 binding
(integration/build_cli.py::build_mexcheck).  Neither MATLAB nor Octave exists in this image, so this is a link-level check:
the object exports mexFunction, takes mcx_run_simulation / mcx_list_gpu / ocl_assess from the binding (the three symbols
src/mcxlabcl.cpp:139, 264, 345 call), every engine symbol it needs is exported by libmcxb200.so, and everything else it
leaves open belongs to the MEX API that MATLAB provides at load time (declared in integration/mexstub/mex.h)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MEX = os.path.join(ROOT, "integration", "_build", "mcxlabcl_check.so")


def symbols(path):
    out = subprocess.run(["nm", "-D", path], capture_output=True, text=True, check=True).stdout
    defined, undefined = set(), set()
    for line in out.splitlines():
        parts = line.split()
        if len(parts) == 2 and parts[0] in ("U", "w"):
            undefined.add(parts[1].split("@")[0])
        elif len(parts) == 3 and parts[1] in "TtWwBbDd":
            defined.add(parts[2].split("@")[0])
    return defined, undefined


def test_mcxlabcl_links_against_the_binding():
    if not os.path.exists(MEX):
        pytest.skip("integration/_build/mcxlabcl_check.so not built (python integration/build_cli.py needs /root/reference)")
    defined, undefined = symbols(MEX)
    assert {"mexFunction", "mcx_run_simulation", "mcx_list_gpu", "ocl_assess"} <= defined
    engine_defined, _ = symbols(os.path.join(ROOT, "mcxcl_b200", "libmcxb200.so"))
    need = {s for s in undefined if s.startswith("mcxb_")}
    assert need and need <= engine_defined, need - engine_defined
    declared = set(re.findall(r"\b(mx[A-Z]\w+|mex[A-Z]\w+)\s*\(", open(os.path.join(ROOT, "integration", "mexstub", "mex.h")).read()))
    mexapi = {s for s in undefined if re.match(r"mx[A-Z]|mex[A-Z]", s)}
    assert len(mexapi) > 25 and mexapi <= declared, mexapi - declared
    # nothing of OpenCL is left in the front-end once the binding replaces mcx_host.cpp
    assert not {s for s in undefined if s.startswith("cl") and s[2:3].isupper()}
