#!/usr/bin/env python3
"""Extract the two atlas volumes the BASELINE.json configs run on into tests/golden/volumes/*.npz.

Runs in the build container only (needs /root/reference); the GPU box uses the committed .npz files.

  colin27    the uint8 181x217x181 atlas embedded in the reference's benchmark table as an LZMA-alone /
             base64 JData array (src/mcx_bench.h:572-626, "Shapes" of the colin27 deck)
  digimouse  the uint8 190x496x104 atlas of example/digimouse/digimouse.json (zlib / base64 JData array),
             with its media table and LengthUnit

Both are stored as numpy arrays of shape (Nx, Ny, Nz) (x fastest in the reference's file order == Fortran order).
"""
import base64
import json
import lzma
import os
import re
import sys
import zlib

import numpy as np

REF = os.environ.get("MCX_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "volumes")


def jdata_array(node):
    size = [int(x) for x in node["_ArraySize_"]]
    raw = base64.b64decode(node["_ArrayZipData_"])
    kind = node["_ArrayZipType_"]
    if kind == "lzma":
        raw = lzma.decompress(raw, format=lzma.FORMAT_ALONE)
    elif kind == "zlib":
        raw = zlib.decompress(raw)
    else:
        raise SystemExit("unsupported _ArrayZipType_ " + kind)
    a = np.frombuffer(raw, dtype=np.dtype(node["_ArrayType_"]))
    assert a.size == int(np.prod(size)), (a.size, size)
    # JData stores the array row-major over _ArraySize_; the reference reads it as vol[x][y][z] and
    # transposes to x-fastest (mcx_loadjson, src/mcx_utils.c: "Shapes" with _ArraySize_)
    return a.reshape(size)


def colin27():
    text = open(os.path.join(REF, "src", "mcx_bench.h")).read()
    at = text.index('"ID":       "colin27"')
    blob_at = text.index('"_ArrayZipData_"', at)
    m = re.compile(r'"_ArrayZipData_":\s*"([^"]*)"').match(text, blob_at)
    b64 = m.group(1).replace("\\\n", "").replace("\n", "")
    node = {"_ArrayType_": "uint8", "_ArraySize_": [181, 217, 181], "_ArrayZipType_": "lzma", "_ArrayZipData_": b64}
    return jdata_array(node)


def digimouse():
    d = json.loads(open(os.path.join(REF, "example", "digimouse", "digimouse.json")).read(), strict=False)
    vol = jdata_array(d["Shapes"])
    prop = np.array([[m["mua"], m["mus"], m["g"], m["n"]] for m in d["Domain"]["Media"]], dtype=np.float32)
    return vol, prop, float(d["Domain"]["LengthUnit"])


def main():
    os.makedirs(OUT, exist_ok=True)
    v = colin27()
    hist = np.bincount(v.ravel(), minlength=7)
    # label histogram probed from the reference blob
    assert hist.tolist() == [3068647, 1457090, 498234, 308195, 990129, 660550, 126292], hist
    np.savez_compressed(os.path.join(OUT, "colin27.npz"), vol=v)
    vol, prop, unit = digimouse()
    assert vol.shape == (190, 496, 104) and prop.shape[0] == 22
    np.savez_compressed(os.path.join(OUT, "digimouse.npz"), vol=vol, prop=prop, unitinmm=unit)
    for n in ("colin27", "digimouse"):
        print(n, os.path.getsize(os.path.join(OUT, n + ".npz")), "bytes")


if __name__ == "__main__":
    sys.exit(main())
