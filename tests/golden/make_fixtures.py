"""Turn the reference's bundled datasets (data/*.rda: bzip2/gzip/xz + RDX2 XDR serialisation) into
small .npz fixtures that travel to the GPU box (where /root/reference does not exist).

Run once in the build container:  python tests/golden/make_fixtures.py
Outputs: tests/golden/{abalone,heart,wine,student}.npz
"""
import bz2
import gzip
import lzma
import os
import struct
import sys

import numpy as np

REF = "/root/reference/data"
OUT = os.path.dirname(os.path.abspath(__file__))


class Reader:
    """Minimal reader for R's XDR serialisation format (version 2/3) — enough for data.frames, numeric
    vectors/matrices, factors and dgCMatrix S4 objects."""

    def __init__(self, buf):
        self.b = buf
        self.o = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def f64(self, n):
        a = np.frombuffer(self.b, dtype=">f8", count=n, offset=self.o).astype(np.float64)
        self.o += 8 * n
        return a

    def ints(self, n):
        a = np.frombuffer(self.b, dtype=">i4", count=n, offset=self.o).astype(np.int32)
        self.o += 4 * n
        return a

    def header(self):
        assert self.b[self.o:self.o + 5] == b"RDX2\n" or self.b[self.o:self.o + 5] == b"RDX3\n", self.b[:5]
        self.o += 5
        assert self.b[self.o:self.o + 2] == b"X\n"
        self.o += 2
        version = self.i32()
        self.i32()
        self.i32()
        if version == 3:
            n = self.i32()
            self.o += n

    def item(self):
        flags = self.i32()
        typ = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        is_obj = bool(flags & 0x100)
        if typ == 254:   # NILVALUE_SXP
            return None
        if typ == 255:   # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if typ == 1:     # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if typ == 2:     # LISTSXP (pairlist)
            out = {}
            attr = self.item() if has_attr else None
            tag = self.item() if has_tag else None
            car = self.item()
            out[tag] = car
            cdr = self.item()
            if isinstance(cdr, dict):
                out.update(cdr)
            return out
        if typ == 9:     # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            s = self.b[self.o:self.o + n].decode("utf-8", "replace")
            self.o += n
            return s
        if typ in (10, 13):   # LGLSXP, INTSXP
            n = self.i32()
            v = self.ints(n)
        elif typ == 14:  # REALSXP
            n = self.i32()
            v = self.f64(n)
        elif typ == 16:  # STRSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif typ == 19:  # VECSXP
            n = self.i32()
            v = [self.item() for _ in range(n)]
        elif typ == 25:  # S4SXP
            v = "S4"
        else:
            raise NotImplementedError(f"SEXP type {typ} at offset {self.o}")
        attrs = self.item() if has_attr else None
        return {"value": v, "attr": attrs or {}, "obj": is_obj} if attrs else v


def load_rda(path):
    raw = open(path, "rb").read()
    if raw[:3] == b"BZh":
        raw = bz2.decompress(raw)
    elif raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    elif raw[:6] == b"\xfd7zXZ\x00":
        raw = lzma.decompress(raw)
    r = Reader(raw)
    r.header()
    return r.item()


def val(o):
    return o["value"] if isinstance(o, dict) and "value" in o else o


def attr(o):
    return o.get("attr", {}) if isinstance(o, dict) else {}


def named_list(o):
    names = val(attr(o)["names"])
    return dict(zip(names, val(o)))


def as_dense(o):
    """numeric matrix / data.frame of numeric columns / dgCMatrix -> (dense array or CSC triple)."""
    a = attr(o)
    v = val(o)
    if isinstance(v, str) and v == "S4":   # dgCMatrix
        dim = val(a["Dim"])
        return dict(i=val(a["i"]).astype(np.int32), p=val(a["p"]).astype(np.int32), x=val(a["x"]),
                    shape=np.array(dim, dtype=np.int64))
    if isinstance(v, list):   # data.frame
        cols = []
        for c in v:
            cv = val(c)
            cols.append(np.asarray(cv, dtype=np.float64))
        return np.stack(cols, axis=1)
    if "dim" in a:
        d = val(a["dim"])
        return np.asarray(v, dtype=np.float64).reshape(tuple(d), order="F")
    return np.asarray(v, dtype=np.float64)


def main():
    for name in ("abalone", "heart", "wine", "student"):
        top = load_rda(os.path.join(REF, name + ".rda"))
        obj = named_list(top[name])
        x = as_dense(obj["x"])
        yraw = obj["y"]
        y = val(yraw)
        if isinstance(y, list):   # data.frame response (student)
            y = np.stack([np.asarray(val(c), dtype=np.float64) for c in y], axis=1)
        else:
            ya = attr(yraw)
            if "dim" in ya:
                y = np.asarray(y, dtype=np.float64).reshape(tuple(val(ya["dim"])), order="F")
            else:
                y = np.asarray(y)   # factor codes (1-based ints) or numeric
        out = {"y": y}
        if isinstance(x, dict):
            out.update({"x_i": x["i"], "x_p": x["p"], "x_x": x["x"], "x_shape": x["shape"]})
        else:
            out["x"] = x
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        shape = x["shape"] if isinstance(x, dict) else x.shape
        print(name, "x", tuple(shape), "sparse" if isinstance(x, dict) else "dense", "y", y.shape, y.dtype)


if __name__ == "__main__":
    sys.exit(main())
