"""Minimal pure-Python reader for R's `.rda` files (bz2/gzip/xz + RDX2/RDX3, XDR serialization).

Test infrastructure only: used by `make_fixtures.py` to decode the reference's bundled
data (`/root/reference/data/*.rda`) into `.npz` golden fixtures, because R is not installed
in this image. It is NOT part of the product path.

Format notes follow R's `serialize.c` (public R-internals documentation); only the SEXP
types that occur in the reference's fixtures are handled.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np

NA_INTEGER = -(2**31)


class RObject:
    """Any R value carrying attributes (S4 slots live in attributes)."""

    __slots__ = ("value", "attr", "kind")

    def __init__(self, value, attr=None, kind=""):
        self.value = value
        self.attr = attr or {}
        self.kind = kind

    def __repr__(self):
        return f"RObject<{self.kind}>({type(self.value).__name__}, attrs={list(self.attr)})"

    def __getitem__(self, key):
        """List access by name (uses the `names` attribute)."""
        names = self.attr.get("names")
        if names is None:
            raise KeyError(key)
        names = list(unwrap(names))
        return self.value[names.index(key)]

    def slot(self, name):
        return self.attr[name]


def unwrap(x):
    return x.value if isinstance(x, RObject) else x


class _Reader:
    def __init__(self, buf: bytes):
        self.buf = buf
        self.pos = 0
        self.refs = []

    def i32(self):
        (v,) = struct.unpack_from(">i", self.buf, self.pos)
        self.pos += 4
        return v

    def raw(self, n):
        b = self.buf[self.pos:self.pos + n]
        self.pos += n
        return b

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def item(self):
        flags = self.i32()
        typ = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if typ == 254:  # NILVALUE_SXP
            return None
        if typ in (253, 242):  # GLOBALENV / EMPTYENV
            return RObject(None, kind="env")
        if typ in (252, 251):  # UNBOUNDVALUE / MISSINGARG
            return None
        if typ == 241:  # BASEENV
            return RObject(None, kind="env")
        if typ == 255:  # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if typ == 1:  # SYMSXP
            name = self.item()
            sym = RObject(name, kind="symbol")
            self.refs.append(sym)
            return sym
        if typ in (249, 250):  # NAMESPACESXP / PACKAGESXP (persistent name STRSXP)
            self.i32()  # 0
            n = self.i32()
            info = [self.item() for _ in range(n)]
            obj = RObject(info, kind="namespace")
            self.refs.append(obj)
            return obj
        if typ == 4:  # ENVSXP
            obj = RObject(None, kind="env")
            self.refs.append(obj)
            self.i32()  # locked
            self.item()  # enclos
            self.item()  # frame
            self.item()  # hashtab
            self.item()  # attrib
            return obj
        if typ == 22:  # EXTPTRSXP
            obj = RObject(None, kind="extptr")
            self.refs.append(obj)
            self.item()  # prot
            self.item()  # tag
            if has_attr:
                self.item()
            return obj
        if typ in (2, 6, 5):  # LISTSXP / LANGSXP / PROMSXP: iterate on cdr
            out = []
            attrs = None
            while True:
                if has_attr:
                    attrs = self.item()
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((unwrap(tag) if tag is not None else None, car))
                nflags = self.i32()
                ntyp = nflags & 0xFF
                if ntyp == 254:
                    break
                if ntyp not in (2, 6, 5):
                    # dotted pair tail: rewind and read as an item
                    self.pos -= 4
                    out.append((None, self.item()))
                    break
                has_attr = bool(nflags & 0x200)
                has_tag = bool(nflags & 0x400)
            return RObject(out, kind="pairlist")
        if typ == 9:  # CHARSXP
            n = self.i32()
            if n == -1:
                return None  # NA_character_
            return self.raw(n).decode("utf-8", errors="replace")
        if typ == 25:  # S4SXP: slots are attributes
            attr = self._attrs() if has_attr else {}
            return RObject(None, attr, kind="S4")
        if typ in (10, 13):  # LGLSXP / INTSXP
            n = self.length()
            v = np.frombuffer(self.raw(4 * n), dtype=">i4").astype(np.int32)
        elif typ == 14:  # REALSXP
            n = self.length()
            v = np.frombuffer(self.raw(8 * n), dtype=">f8").astype(np.float64)
        elif typ == 16:  # STRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif typ in (19, 20):  # VECSXP / EXPRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif typ == 24:  # RAWSXP
            n = self.length()
            v = self.raw(n)
        elif typ == 3:  # CLOSXP
            attr = self._attrs() if has_attr else {}
            self.item(); self.item(); self.item()
            return RObject(None, attr, kind="closure")
        elif typ == 238:  # ALTREP_SXP (serialization v3)
            info = self.item()
            state = self.item()
            attr = self.item()
            cls = unwrap(unwrap(info).__getitem__(0)[1]) if isinstance(unwrap(info), list) else None
            return RObject(state, {}, kind=f"altrep:{cls}")
        else:
            raise NotImplementedError(f"SEXP type {typ} at byte {self.pos}")
        attr = self._attrs() if has_attr else {}
        if attr:
            return RObject(v, attr, kind=str(typ))
        return v

    def _attrs(self):
        pl = self.item()
        out = {}
        if pl is None:
            return out
        for tag, val in pl.value:
            out[tag] = val
        return out


def read_rda(path: str) -> dict:
    """Return {object name: decoded value} for an `.rda` file."""
    with open(path, "rb") as fh:
        blob = fh.read()
    if blob[:3] == b"BZh":
        blob = bz2.decompress(blob)
    elif blob[:2] == b"\x1f\x8b":
        blob = gzip.decompress(blob)
    elif blob[:6] == b"\xfd7zXZ\x00":
        blob = lzma.decompress(blob)
    magic = blob[:5]
    if magic not in (b"RDX2\n", b"RDX3\n"):
        raise ValueError(f"not an RDX2/3 file: {magic!r}")
    r = _Reader(blob)
    r.pos = 5
    fmt = r.raw(2)
    if fmt != b"X\n":
        raise ValueError("only XDR serialization is supported")
    version = r.i32()
    r.i32()  # writer version
    r.i32()  # min reader version
    if version == 3:
        n = r.i32()
        r.raw(n)  # native encoding
    top = r.item()
    return {tag: val for tag, val in top.value}


def dataframe_columns(df) -> dict:
    """data.frame / data.table -> {column name: list or ndarray}."""
    names = list(unwrap(df.attr["names"]))
    cols = {}
    for name, col in zip(names, df.value):
        cols[name] = unwrap(col)
    return cols
