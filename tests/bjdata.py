"""Minimal reader of binary JData / UBJSON -- enough to read back the .bnii volumes the reference's mcx_savebnii writes
(src/mcx_utils.c:598-735) in the output-format tests.  The file announces "SerialFormat: bjdata/draft2" (little-endian)
but the reference's vendored ubj writer emits every multi-byte number BIG-endian, the UBJSON convention; `load` takes
the byte order as an argument and defaults to what the reference actually writes."""
import struct

import numpy as np

_FIXED = {b"i": ("b", 1), b"U": ("B", 1), b"I": ("h", 2), b"u": ("H", 2), b"l": ("i", 4), b"m": ("I", 4),
          b"L": ("q", 8), b"M": ("Q", 8), b"h": ("e", 2), b"d": ("f", 4), b"D": ("d", 8)}
_NP = {b"i": "i1", b"U": "u1", b"I": "i2", b"u": "u2", b"l": "i4", b"m": "u4", b"L": "i8", b"M": "u8", b"h": "f2", b"d": "f4", b"D": "f8"}


class _Reader:
    def __init__(self, buf, order=">"):
        self.b, self.i, self.order = buf, 0, order

    def take(self, n):
        out = self.b[self.i:self.i + n]
        if len(out) != n:
            raise ValueError("truncated BJData stream")
        self.i += n
        return out

    def peek(self):
        return self.b[self.i:self.i + 1]

    def number(self, marker):
        fmt, n = _FIXED[marker]
        return struct.unpack(self.order + fmt, self.take(n))[0]

    def length(self):
        m = self.take(1)
        if m not in _FIXED:
            raise ValueError("bad length marker %r" % m)
        return int(self.number(m))

    def string(self):
        return self.take(self.length()).decode("utf-8")

    def value(self, marker=None):
        m = marker or self.take(1)
        if m == b"Z":
            return None
        if m == b"T":
            return True
        if m == b"F":
            return False
        if m == b"N":
            return self.value()
        if m in _FIXED:
            return self.number(m)
        if m == b"C":
            return self.take(1).decode("latin1")
        if m == b"S":
            return self.string()
        if m == b"[":
            return self.array()
        if m == b"{":
            return self.object()
        raise ValueError("unknown BJData marker %r at %d" % (m, self.i))

    def header(self):
        typ = cnt = None
        if self.peek() == b"$":
            self.take(1)
            typ = self.take(1)
        if self.peek() == b"#":
            self.take(1)
            cnt = self.length()
        return typ, cnt

    def array(self):
        typ, cnt = self.header()
        if cnt is not None:
            if typ in _NP:
                dt = np.dtype(_NP[typ]).newbyteorder(self.order)
                return np.frombuffer(self.take(cnt * dt.itemsize), dtype=dt).astype(dt.newbyteorder("="))
            return [self.value(typ) for _ in range(cnt)]
        out = []
        while self.peek() != b"]":
            out.append(self.value())
        self.take(1)
        return out

    def object(self):
        typ, cnt = self.header()
        out = {}
        if cnt is not None:
            for _ in range(cnt):
                key = self.string()
                out[key] = self.value(typ)
            return out
        while self.peek() != b"}":
            key = self.string()
            out[key] = self.value()
        self.take(1)
        return out


def load(path, order=">"):
    with open(path, "rb") as f:
        return _Reader(f.read(), order).value()


def decode_array(node):
    """a JData annotated array {_ArrayType_, _ArraySize_, [_ArrayZipType_, _ArrayZipSize_, _ArrayZipData_ | _ArrayData_]}
    -> numpy array in the annotated (row-major) shape"""
    import zlib
    dtype = {"single": "<f4", "double": "<f8", "uint32": "<u4", "int32": "<i4", "uint8": "u1", "uint16": "<u2", "int16": "<i2", "uint64": "<u8"}[node["_ArrayType_"]]
    shape = [int(x) for x in np.atleast_1d(node["_ArraySize_"])]
    if "_ArrayZipData_" in node:
        raw = node["_ArrayZipData_"]
        raw = raw.tobytes() if isinstance(raw, np.ndarray) else (raw if isinstance(raw, (bytes, bytearray)) else __import__("base64").b64decode(raw))
        assert node["_ArrayZipType_"] == "zlib"
        data = np.frombuffer(zlib.decompress(raw), dtype=dtype)
    else:
        data = np.asarray(node["_ArrayData_"], dtype=dtype)
    return data.reshape(shape)
