"""OpenFOAM file formats: dictionaries, polyMesh, vol/surface fields (ascii + binary).

This is the on-disk contract of the drop-in boundary: the solver reads
what `gmshToFoam`/`setFields` leave in a case directory and writes what `foamRun` would.

Reference artefacts that pin the layout:
  * binary volScalarField: /root/reference/case_H0.004_D0.0221_flat_R0.005_f2.0/0/alpha.water
    (`internalField   nonuniform List<scalar> \\nN\\n(` + N*8 bytes LE + `)` + `;`)
  * dictionaries: /root/reference/circularSloshingTank/system/{controlDict,fvSchemes,fvSolution}
  * 6DoF table:   /root/reference/circularSloshingTank/generate_motion.py:13-42
"""
from __future__ import annotations

import os
import re
import numpy as np

BANNER = (
    "/*--------------------------------*- C++ -*----------------------------------*\\\n"
    "  =========                 |\n"
    "  \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox\n"
    "   \\\\    /   O peration     | Website:  https://openfoam.org\n"
    "    \\\\  /    A nd           | Version:  13\n"
    "     \\\\/     M anipulation  |\n"
    "\\*---------------------------------------------------------------------------*/\n"
)
SEP = "// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n"
END = "\n// ************************************************************************* //\n"


class FoamError(Exception):
    """Raised for malformed files and for keywords this solver refuses to guess."""


# ----------------------------------------------------------------------------------------
# tokenizer / recursive-descent parser
# ----------------------------------------------------------------------------------------
_PUNCT = b"{}();"
_BINARY_LISTS = ("List<scalar>", "List<vector>", "List<tensor>", "List<symmTensor>", "List<label>")
_WS = b" \t\r\n"


class BinaryBlock:
    """Raw bytes of a binary `List<T>` body; decoded lazily by the field readers."""

    __slots__ = ("count", "data")

    def __init__(self, count, data):
        self.count = count
        self.data = data

    def __repr__(self):
        return f"BinaryBlock(n={self.count}, bytes={len(self.data)})"


class _Parser:
    def __init__(self, buf: bytes, name: str = "<buffer>"):
        self.b = buf
        self.i = 0
        self.n = len(buf)
        self.name = name
        self.binary = False
        self.label_bytes = 4
        self.scalar_bytes = 8

    # -- low level ---------------------------------------------------------------------
    def _skip(self):
        b, n = self.b, self.n
        while self.i < n:
            c = b[self.i]
            if c in _WS:
                self.i += 1
            elif c == 0x2F and self.i + 1 < n and b[self.i + 1] == 0x2F:  # //
                j = b.find(b"\n", self.i)
                self.i = n if j < 0 else j + 1
            elif c == 0x2F and self.i + 1 < n and b[self.i + 1] == 0x2A:  # /*
                j = b.find(b"*/", self.i + 2)
                if j < 0:
                    raise FoamError(f"{self.name}: unterminated comment")
                self.i = j + 2
            else:
                break

    def peek(self):
        self._skip()
        if self.i >= self.n:
            return None
        c = self.b[self.i]
        if c in _PUNCT:
            return chr(c)
        return self._word(advance=False)

    def next(self):
        self._skip()
        if self.i >= self.n:
            return None
        c = self.b[self.i]
        if c in _PUNCT:
            self.i += 1
            return chr(c)
        return self._word(advance=True)

    def _word(self, advance):
        b, n, i = self.b, self.n, self.i
        if b[i] == 0x22:  # string
            j = i + 1
            while j < n and b[j] != 0x22:
                j += 2 if b[j] == 0x5C else 1
            tok = b[i : j + 1].decode("latin-1")
            if advance:
                self.i = j + 1
            return tok
        j = i
        depth = 0
        while j < n:
            c = b[j]
            if c in _WS:
                break
            if c == 0x28:  # '(' inside a word such as div(phi,alpha)
                if j == i:
                    break
                depth += 1
            elif c == 0x29:
                if depth == 0:
                    break
                depth -= 1
            elif c in b"{};" and depth == 0:
                break
            j += 1
        tok = b[i:j].decode("latin-1")
        if advance:
            self.i = j
        return tok

    def expect(self, t):
        g = self.next()
        if g != t:
            raise FoamError(f"{self.name}: expected '{t}' got '{g}' near byte {self.i}")

    # -- grammar -----------------------------------------------------------------------
    def parse_dict_body(self, top=False):
        d = {}
        while True:
            t = self.peek()
            if t is None:
                if top:
                    return d
                raise FoamError(f"{self.name}: unexpected end of file in dictionary")
            if t == "}":
                if top:
                    raise FoamError(f"{self.name}: unbalanced '}}'")
                self.next()
                return d
            key = self.next()
            if key in ("(", ")", "{", ";"):
                raise FoamError(f"{self.name}: unexpected '{key}' where a keyword was expected")
            if key.startswith("#"):
                # directives (#include, #calc ...) are not used by the reference's cases
                raise FoamError(f"{self.name}: directive '{key}' is not supported")
            if key.startswith('"'):
                key = key  # regex key such as "pcorr.*" (kept quoted)
            t = self.peek()
            if t == "{":
                self.next()
                d[key] = self.parse_dict_body()
                continue
            vals = self.parse_stream()
            d[key] = vals[0] if len(vals) == 1 else vals
            if key == "FoamFile" and isinstance(d[key], dict):
                pass
        return d

    def parse_stream(self):
        """Values up to ';' — returns a python list of tokens / nested lists / BinaryBlock."""
        out = []
        while True:
            t = self.peek()
            if t is None:
                raise FoamError(f"{self.name}: missing ';'")
            if t == ";":
                self.next()
                return out
            if t == "{":
                # `key value { ... }` does not occur in our files
                self.next()
                out.append(self.parse_dict_body())
                continue
            out.append(self.parse_value(out))

    def parse_value(self, before=None):
        t = self.next()
        if t == "(":
            return self.parse_list()
        if t is None or t in ")};":
            raise FoamError(f"{self.name}: unexpected '{t}'")
        # N(...) or N{value}
        if _is_int(t):
            nxt = self._peek_raw_char()
            if nxt == "(":
                # counted list; binary if the stream is binary and the element type is numeric
                n = int(t)
                # numeric element types are raw bytes in a binary stream; `inGroups List<word> 1(wall)` of a
                # polyMesh/boundary file is text in both formats
                if before and isinstance(before[-1], str) and before[-1] in _BINARY_LISTS:
                    if self.binary:
                        return self._binary_list(n, before[-1])
                    # ascii numbers in bulk (numpy), not token by token: a 3 M-vector field in 2 s instead of 76 s
                    typ = before[-1][5:-1]
                    self._skip()
                    start = self.i + 1
                    if n == 0:
                        end = self.b.index(b")", start) + 1
                        a = np.zeros((0,) if _NCOMP[typ] == 1 else (0, _NCOMP[typ]), dtype=np.int64 if typ == "label" else np.float64)
                    else:
                        try:
                            a, end = _fast_ascii_list(self.b, start, n, _NCOMP[typ], label=(typ == "label"))
                        except ValueError as e:
                            raise FoamError(f"{self.name}: malformed {before[-1]} of {n} entries ({e})")
                    self.i = end
                    return a
                self.next()
                return self.parse_list()
            if nxt == "{":
                self.next()
                v = self.parse_value()
                self.expect("}")
                return ["__uniformlist__", int(t), v]
        m = re.match(r"^(\d+)$", t)
        return t

    def _peek_raw_char(self):
        self._skip()
        if self.i >= self.n:
            return None
        return chr(self.b[self.i])

    def parse_list(self):
        out = []
        while True:
            t = self.peek()
            if t is None:
                raise FoamError(f"{self.name}: unterminated list")
            if t == ")":
                self.next()
                return out
            if t == "{":
                # list of dictionaries with names: `name { ... }`
                self.next()
                body = self.parse_dict_body()
                name = out.pop() if out else None
                out.append((name, body))
                continue
            out.append(self.parse_value(out))

    def _binary_list(self, n, typ):
        width = {
            "List<scalar>": self.scalar_bytes,
            "List<vector>": 3 * self.scalar_bytes,
            "List<tensor>": 9 * self.scalar_bytes,
            "List<symmTensor>": 6 * self.scalar_bytes,
            "List<label>": self.label_bytes,
        }.get(typ)
        if width is None:
            raise FoamError(f"{self.name}: binary list of type {typ} is not supported")
        self._skip()
        if self.b[self.i] != 0x28:
            raise FoamError(f"{self.name}: expected '(' before binary data")
        start = self.i + 1
        end = start + n * width
        if end >= self.n or self.b[end] != 0x29:
            raise FoamError(f"{self.name}: binary list of {n} x {width} bytes is truncated")
        self.i = end + 1
        return BinaryBlock(n, self.b[start:end])


def _is_int(t):
    return t.isdigit()


def parse_header(buf: bytes, name="<buffer>"):
    """Return (header dict, parser positioned after the FoamFile block)."""
    p = _Parser(buf, name)
    t = p.peek()
    if t != "FoamFile":
        return {}, p
    p.next()
    p.expect("{")
    hdr = p.parse_dict_body()
    p.binary = hdr.get("format", "ascii") == "binary"
    arch = hdr.get("arch", "")
    m = re.search(r"label=(\d+)", arch)
    if m:
        p.label_bytes = int(m.group(1)) // 8
    m = re.search(r"scalar=(\d+)", arch)
    if m:
        p.scalar_bytes = int(m.group(1)) // 8
    return hdr, p


def read_dict(path):
    """Parse an OpenFOAM dictionary file into nested python dicts (values: str / list)."""
    with open(path, "rb") as f:
        buf = f.read()
    hdr, p = parse_header(buf, path)
    d = p.parse_dict_body(top=True)
    d["FoamFile"] = hdr
    return d


def parse_dict_string(s: str):
    _, p = parse_header(s.encode(), "<string>")
    return p.parse_dict_body(top=True)


# typed access with hard errors (no silent defaults: SURVEY.md §8b)
def lookup(d, key, path="", default=FoamError):
    if key in d:
        return d[key]
    # OpenFOAM regex keys ("pcorr.*"), last match wins
    hit = None
    for k, v in d.items():
        if k.startswith('"') and re.fullmatch(k.strip('"'), key):
            hit = v
    if hit is not None:
        return hit
    if default is FoamError:
        raise FoamError(f"keyword '{key}' is missing in {path}")
    return default


def to_float(v, what=""):
    if isinstance(v, list):
        # e.g. `[dims] value` or `uniform 0`
        v = v[-1]
    try:
        return float(v)
    except (TypeError, ValueError):
        raise FoamError(f"expected a number for {what}, got {v!r}")


def to_vector(v, what=""):
    if isinstance(v, list) and len(v) == 3 and not isinstance(v[0], list):
        return np.array([float(x) for x in v])
    if isinstance(v, list) and v and isinstance(v[-1], list):
        return np.array([float(x) for x in v[-1]])
    raise FoamError(f"expected a vector for {what}, got {v!r}")


def to_bool(v, what=""):
    s = str(v).lower()
    if s in ("yes", "on", "true", "y", "t", "1"):
        return True
    if s in ("no", "off", "false", "n", "f", "0", "none"):
        return False
    raise FoamError(f"expected a switch for {what}, got {v!r}")


# ----------------------------------------------------------------------------------------
# numeric lists
# ----------------------------------------------------------------------------------------
_NCOMP = {"scalar": 1, "vector": 3, "tensor": 9, "symmTensor": 6, "label": 1}


def _decode_list(val, typ, p):
    """val: BinaryBlock | nested python list -> numpy array (n, ncomp) or (n,)"""
    nc = _NCOMP[typ]
    if isinstance(val, np.ndarray):  # an ascii list the parser already decoded in bulk
        a = val.astype(np.int64 if typ == "label" else np.float64, copy=False)
    elif isinstance(val, BinaryBlock):
        if typ == "label":
            dt = "<i4" if p.label_bytes == 4 else "<i8"
        else:
            dt = "<f8" if p.scalar_bytes == 8 else "<f4"
        a = np.frombuffer(val.data, dtype=dt)
        a = a.astype(np.int64 if typ == "label" else np.float64)
    else:
        if typ == "label":
            a = np.array([int(x) for x in val], dtype=np.int64)
        elif nc == 1:
            a = np.array([float(x) for x in val], dtype=np.float64)
        else:
            a = np.array([[float(c) for c in x] for x in val], dtype=np.float64)
    if nc > 1:
        a = a.reshape(-1, nc)
    return a


def _fast_ascii_list(buf, start, n, nc, label=False):
    """Parse n entries of an ascii numeric list beginning right after '(' at buf[start]."""
    # find the closing parenthesis by counting: for nc>1 each entry has its own parens
    if nc == 1:
        end = buf.index(b")", start)
        txt = buf[start:end]
    else:
        # entries look like (a b c); strip parens wholesale
        # closing ')' of the list is the first ')' that follows another ')' (after whitespace)
        m = re.compile(rb"\)\s*\)").search(buf, start)
        if m is None:
            if n == 0:
                end = buf.index(b")", start)
                return np.zeros((0, nc)), end + 1
            raise FoamError("unterminated vector list")
        end = m.end() - 1
        txt = buf[start:end].replace(b"(", b" ").replace(b")", b" ")
    a = np.array(txt.split(), dtype=np.int64 if label else np.float64)
    if a.size != n * nc:
        raise FoamError(f"list size mismatch: header says {n} x {nc}, found {a.size} numbers")
    if nc > 1:
        a = a.reshape(n, nc)
    return a, end + 1


def _read_counted(buf, pos, p, typ, what):
    """Read `N ( ... )` (ascii or binary) of element type typ starting at/after pos."""
    m = re.compile(rb"\s*(\d+)\s*\(").match(buf, pos)
    if m is None:
        raise FoamError(f"{what}: expected 'N (' ")
    n = int(m.group(1))
    start = m.end()
    nc = _NCOMP[typ]
    if p.binary:
        if typ == "label":
            width, dt = p.label_bytes, ("<i4" if p.label_bytes == 4 else "<i8")
        else:
            width, dt = p.scalar_bytes * nc, ("<f8" if p.scalar_bytes == 8 else "<f4")
        end = start + n * width
        if buf[end : end + 1] != b")":
            raise FoamError(f"{what}: binary list truncated")
        a = np.frombuffer(buf, dtype=dt, count=n * (nc if typ != "label" else 1), offset=start)
        a = a.astype(np.int64 if typ == "label" else np.float64)
        if nc > 1:
            a = a.reshape(n, nc)
        return a, end + 1
    return _fast_ascii_list(buf, start, n, nc, label=(typ == "label"))


def _body_start(buf):
    """Offset just after the FoamFile { } block."""
    i = buf.find(b"FoamFile")
    if i < 0:
        return 0
    j = buf.index(b"}", i)
    return j + 1


def _skip_comments(buf, pos):
    n = len(buf)
    while pos < n:
        c = buf[pos : pos + 1]
        if c in b" \t\r\n":
            pos += 1
        elif buf[pos : pos + 2] == b"//":
            j = buf.find(b"\n", pos)
            pos = n if j < 0 else j + 1
        elif buf[pos : pos + 2] == b"/*":
            pos = buf.index(b"*/", pos) + 2
        else:
            break
    return pos


# ----------------------------------------------------------------------------------------
# polyMesh
# ----------------------------------------------------------------------------------------
class PolyMesh:
    """constant/polyMesh in memory.

    points (P,3) f64; face_offsets (nFaces+1,) / face_labels: CSR of point labels per face;
    owner (nFaces,), neighbour (nInternal,) ; patches: list of dict(name,type,nFaces,startFace)
    cell_zones: dict name -> label array.   All labels int32 (label=32, as the reference's
    OpenFOAM 13 build writes: arch "LSB;label=32;scalar=64").
    """

    def __init__(self, points, face_offsets, face_labels, owner, neighbour, patches, cell_zones=None):
        self.points = np.ascontiguousarray(points, dtype=np.float64)
        self.face_offsets = np.ascontiguousarray(face_offsets, dtype=np.int32)
        self.face_labels = np.ascontiguousarray(face_labels, dtype=np.int32)
        self.owner = np.ascontiguousarray(owner, dtype=np.int32)
        self.neighbour = np.ascontiguousarray(neighbour, dtype=np.int32)
        self.patches = patches
        self.cell_zones = cell_zones or {}

    @property
    def n_points(self):
        return self.points.shape[0]

    @property
    def n_faces(self):
        return self.owner.shape[0]

    @property
    def n_internal(self):
        return self.neighbour.shape[0]

    @property
    def n_cells(self):
        if not self.owner.size:
            return 0
        # the last cell may own no face at all (all four faces shared with lower-numbered cells)
        return max(int(self.owner.max()), int(self.neighbour.max()) if self.neighbour.size else -1) + 1

    def patch(self, name):
        for p in self.patches:
            if p["name"] == name:
                return p
        raise KeyError(name)

    def check(self):
        """Structural invariants OpenFOAM's checkMesh enforces on addressing (bit-exact part)."""
        nI = self.n_internal
        own, nei = self.owner, self.neighbour
        assert np.all(own[:nI] < nei), "owner < neighbour violated"
        key = own[:nI].astype(np.int64) * (self.n_cells + 1) + nei
        assert np.all(np.diff(key) > 0), "internal faces are not in upper-triangular order"
        s = nI
        for p in self.patches:
            assert p["startFace"] == s, f"patch {p['name']} startFace {p['startFace']} != {s}"
            s += p["nFaces"]
        assert s == self.n_faces
        return True


def _hdr(cls, obj, location, fmt="ascii", note=None):
    s = BANNER + "FoamFile\n{\n"
    s += f"    format      {fmt};\n"
    s += f"    class       {cls};\n"
    if fmt == "binary":
        s += '    arch        "LSB;label=32;scalar=64";\n'
    if note:
        s += f'    note        "{note}";\n'
    if location is not None:
        s += f'    location    "{location}";\n'
    s += f"    object      {obj};\n}}\n" + SEP + "\n"
    return s


def write_polymesh(case_dir, mesh: PolyMesh, binary=True, region_dir="constant/polyMesh"):
    d = os.path.join(case_dir, region_dir)
    os.makedirs(d, exist_ok=True)
    loc = region_dir
    fmt = "binary" if binary else "ascii"
    nC, nF, nI, nP = mesh.n_cells, mesh.n_faces, mesh.n_internal, mesh.n_points
    note = f"nPoints:{nP}  nCells:{nC}  nFaces:{nF}  nInternalFaces:{nI}"
    write_points(os.path.join(d, "points"), mesh.points, binary, loc)
    # faces
    with open(os.path.join(d, "faces"), "wb") as f:
        if binary:
            f.write(_hdr("faceCompactList", "faces", loc, fmt).encode())
            f.write(f"\n{nF + 1}\n(".encode())
            f.write(mesh.face_offsets.astype("<i4").tobytes())
            f.write(b")\n")
            f.write(f"\n{mesh.face_labels.size}\n(".encode())
            f.write(mesh.face_labels.astype("<i4").tobytes())
            f.write(b")\n")
        else:
            f.write(_hdr("faceList", "faces", loc, fmt).encode())
            f.write(f"\n{nF}\n(\n".encode())
            off, lab = mesh.face_offsets, mesh.face_labels
            lines = []
            for i in range(nF):
                l = lab[off[i] : off[i + 1]]
                lines.append(f"{len(l)}(" + " ".join(map(str, l)) + ")")
            f.write(("\n".join(lines) + "\n)\n").encode())
        f.write(END.encode())
    for nm, arr in (("owner", mesh.owner), ("neighbour", mesh.neighbour)):
        with open(os.path.join(d, nm), "wb") as f:
            f.write(_hdr("labelList", nm, loc, fmt, note).encode())
            f.write(f"\n{arr.size}\n(".encode())
            if binary:
                f.write(arr.astype("<i4").tobytes())
                f.write(b")\n")
            else:
                f.write(("\n" + "\n".join(map(str, arr)) + "\n)\n").encode())
            f.write(END.encode())
    with open(os.path.join(d, "boundary"), "w") as f:
        f.write(_hdr("polyBoundaryMesh", "boundary", loc, "ascii"))
        f.write(f"{len(mesh.patches)}\n(\n")
        for p in mesh.patches:
            f.write(f"    {p['name']}\n    {{\n        type            {p['type']};\n")
            if p["type"] == "wall":
                f.write("        inGroups        List<word> 1(wall);\n")
            for k in ("myProcNo", "neighbProcNo"):
                if k in p:
                    f.write(f"        {k:<15} {p[k]};\n")
            f.write(f"        nFaces          {p['nFaces']};\n        startFace       {p['startFace']};\n    }}\n")
        f.write(")\n" + END)
    if mesh.cell_zones:
        with open(os.path.join(d, "cellZones"), "wb") as f:
            f.write(_hdr("regIOobject", "cellZones", loc, fmt).encode())
            f.write(f"\n{len(mesh.cell_zones)}\n(\n".encode())
            for nm, lab in mesh.cell_zones.items():
                f.write(f"{nm}\n{{\n    type cellZone;\n    cellLabels      List<label> ".encode())
                f.write(f"\n{lab.size}\n(".encode())
                if binary:
                    f.write(np.asarray(lab).astype("<i4").tobytes())
                else:
                    f.write(("\n" + "\n".join(map(str, lab)) + "\n").encode())
                f.write(b")\n;\n}\n")
            f.write(b")\n")
            f.write(END.encode())


def write_points(path, points, binary=True, location="constant/polyMesh"):
    fmt = "binary" if binary else "ascii"
    with open(path, "wb") as f:
        f.write(_hdr("vectorField", "points", location, fmt).encode())
        f.write(f"\n{points.shape[0]}\n(".encode())
        if binary:
            f.write(np.ascontiguousarray(points, dtype="<f8").tobytes())
            f.write(b")\n")
        else:
            f.write(b"\n")
            f.write("\n".join(f"({p[0]!r} {p[1]!r} {p[2]!r})" for p in points.tolist()).encode())
            f.write(b"\n)\n")
        f.write(END.encode())


def read_points(path):
    with open(path, "rb") as f:
        buf = f.read()
    _, p = parse_header(buf, path)
    pos = _skip_comments(buf, _body_start(buf))
    a, _ = _read_counted(buf, pos, p, "vector", path)
    return a


def _read_label_list(path):
    with open(path, "rb") as f:
        buf = f.read()
    hdr, p = parse_header(buf, path)
    pos = _skip_comments(buf, _body_start(buf))
    a, _ = _read_counted(buf, pos, p, "label", path)
    return a, hdr


def read_faces(path):
    with open(path, "rb") as f:
        buf = f.read()
    hdr, p = parse_header(buf, path)
    pos = _skip_comments(buf, _body_start(buf))
    if hdr.get("class") == "faceCompactList":
        off, pos = _read_counted(buf, pos, p, "label", path)
        pos = _skip_comments(buf, pos)
        lab, pos = _read_counted(buf, pos, p, "label", path)
        return off.astype(np.int32), lab.astype(np.int32)
    # ascii faceList: N ( 3(a b c) 4(a b c d) ... )
    m = re.compile(rb"\s*(\d+)\s*\(").match(buf, pos)
    n = int(m.group(1))
    body = buf[m.end() :]
    # every face is k(l0 l1 ...): turn parens into spaces and parse one flat int stream
    end = body.rfind(b")")
    end = body.rfind(b")", 0, end) + 1 if n else 0
    flat = np.array(body[:end].replace(b"(", b" ").replace(b")", b" ").split(), dtype=np.int64)
    off = np.zeros(n + 1, dtype=np.int64)
    lab = np.empty(flat.size - n, dtype=np.int64)
    # sizes are irregular: walk (vectorised when all faces share a size)
    i = 0
    k = 0
    sizes = np.empty(n, dtype=np.int64)
    if n and (flat.size % n == 0) and np.all(flat[:: flat.size // n] == flat[0]) and flat[0] == flat.size // n - 1:
        s = int(flat[0])
        sizes[:] = s
        lab = flat.reshape(n, s + 1)[:, 1:].reshape(-1)
    else:
        for fi in range(n):
            s = int(flat[i])
            sizes[fi] = s
            lab[k : k + s] = flat[i + 1 : i + 1 + s]
            i += s + 1
            k += s
    off[1:] = np.cumsum(sizes)
    return off.astype(np.int32), lab.astype(np.int32)


def read_boundary(path):
    with open(path, "rb") as f:
        buf = f.read()
    hdr, p = parse_header(buf, path)
    v = p.parse_value()
    if not isinstance(v, list):
        raise FoamError(f"{path}: expected a list of patches")
    patches = []
    for item in v:
        if not isinstance(item, tuple):
            continue
        name, body = item
        q = {"name": name, "type": body.get("type", "patch"), "nFaces": int(body["nFaces"]), "startFace": int(body["startFace"])}
        for k in ("myProcNo", "neighbProcNo"):
            if k in body:
                q[k] = int(body[k])
        patches.append(q)
    return patches


def read_cell_zones(path):
    if not os.path.exists(path):
        return {}
    with open(path, "rb") as f:
        buf = f.read()
    hdr, p = parse_header(buf, path)
    v = p.parse_value()
    zones = {}
    for item in v:
        if isinstance(item, tuple):
            name, body = item
            cl = body.get("cellLabels")
            if isinstance(cl, list) and len(cl) >= 2:
                zones[name] = _decode_list(cl[-1], "label", p).astype(np.int32)
    return zones


def read_polymesh(case_dir, region_dir="constant/polyMesh"):
    d = os.path.join(case_dir, region_dir)
    if not os.path.isdir(d):
        raise FoamError(f"{d}: no polyMesh (run gmshToFoam or the repo's mesh generator first)")
    points = read_points(os.path.join(d, "points"))
    off, lab = read_faces(os.path.join(d, "faces"))
    owner, _ = _read_label_list(os.path.join(d, "owner"))
    neighbour, _ = _read_label_list(os.path.join(d, "neighbour"))
    patches = read_boundary(os.path.join(d, "boundary"))
    zones = read_cell_zones(os.path.join(d, "cellZones"))
    return PolyMesh(points, off, lab, owner, neighbour, patches, zones)


# ----------------------------------------------------------------------------------------
# fields
# ----------------------------------------------------------------------------------------
_CLASS_TYPE = {
    "volScalarField": "scalar",
    "volVectorField": "vector",
    "surfaceScalarField": "scalar",
    "surfaceVectorField": "vector",
    "volTensorField": "tensor",
}


class Field:
    """A vol/surface field: internal values + per-patch dictionaries.

    internal: float (uniform scalar), (ncomp,) array (uniform vector) or (n[,ncomp]) array.
    boundary: dict patch -> dict(type=..., value=array|uniform, ...raw entries)
    """

    def __init__(self, cls, name, dimensions, internal, boundary):
        self.cls = cls
        self.name = name
        self.dimensions = dimensions
        self.internal = internal
        self.boundary = boundary

    def internal_array(self, n):
        nc = _NCOMP[_CLASS_TYPE[self.cls]]
        a = np.asarray(self.internal, dtype=np.float64)
        if nc == 1:
            return np.full(n, float(a)) if a.ndim == 0 else a.reshape(n)
        if a.ndim == 1:
            return np.tile(a, (n, 1))
        return a.reshape(n, nc)


def _field_value(val, typ, p, what):
    """Decode `uniform X` / `nonuniform List<T> N (...)` token lists."""
    if not isinstance(val, list):
        raise FoamError(f"{what}: malformed field value {val!r}")
    if val[0] == "uniform":
        v = val[1]
        if isinstance(v, list):
            return np.array([float(c) for c in v])
        return float(v)
    if val[0] == "nonuniform":
        # ['nonuniform', 'List<scalar>', payload]  (ascii count is swallowed by parse_value)
        if len(val) == 3:
            return _decode_list(val[2], typ, p)
        if len(val) == 2:  # `nonuniform 0()` style
            return np.zeros((0,) if _NCOMP[typ] == 1 else (0, _NCOMP[typ]))
    raise FoamError(f"{what}: unsupported field value {val[:2]!r}")


def read_field(path):
    with open(path, "rb") as f:
        buf = f.read()
    hdr, p = parse_header(buf, path)
    cls = hdr.get("class")
    if cls not in _CLASS_TYPE:
        raise FoamError(f"{path}: unsupported field class {cls!r}")
    typ = _CLASS_TYPE[cls]
    d = p.parse_dict_body(top=True)
    dims = d.get("dimensions")
    internal = _field_value(d["internalField"], typ, p, path + ":internalField")
    boundary = {}
    for name, body in d.get("boundaryField", {}).items():
        e = dict(body)
        if "value" in e:
            e["value"] = _field_value(e["value"], typ, p, f"{path}:{name}.value")
        boundary[name] = e
    return Field(cls, hdr.get("object", os.path.basename(path)), dims, internal, boundary)


def _fmt_num(x, prec=6):
    return f"{x:.{prec}g}"


def _write_list(f, a, typ, binary, prec):
    nc = _NCOMP[typ]
    a = np.asarray(a, dtype=np.float64)
    n = a.shape[0]
    f.write(f"nonuniform List<{typ}> \n{n}\n(".encode())
    if binary:
        f.write(np.ascontiguousarray(a, dtype="<f8").tobytes())
        f.write(b")")
    else:
        f.write(b"\n")
        if nc == 1:
            f.write("\n".join(_fmt_num(x, prec) for x in a.tolist()).encode())
        else:
            f.write("\n".join("(" + " ".join(_fmt_num(c, prec) for c in r) + ")" for r in a.tolist()).encode())
        f.write(b"\n)")


def _write_value(f, v, typ, binary, prec):
    a = np.asarray(v, dtype=np.float64)
    nc = _NCOMP[typ]
    if a.ndim == 0:
        f.write(f"uniform {_fmt_num(float(a), prec)}".encode())
    elif nc > 1 and a.ndim == 1:
        f.write(("uniform (" + " ".join(_fmt_num(c, prec) for c in a.tolist()) + ")").encode())
    else:
        _write_list(f, a, typ, binary, prec)


def write_field(path, fld: Field, binary=True, precision=6, location=None):
    """Write a field the way OpenFOAM's writeFormat binary / writePrecision 6 does
    (reference: circularSloshingTank/system/controlDict:35-37)."""
    typ = _CLASS_TYPE[fld.cls]
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(_hdr(fld.cls, fld.name, location, "binary" if binary else "ascii").encode())
        dims = fld.dimensions
        if isinstance(dims, list):  # tokens of `[0 1 -2 0 0 0 0]` as the parser split them
            dims = " ".join(str(x) for x in dims)
        f.write(f"dimensions      {dims};\n\n".encode())
        f.write(b"internalField   ")
        _write_value(f, fld.internal, typ, binary, precision)
        f.write(b";\n\nboundaryField\n{\n")
        for name, e in fld.boundary.items():
            f.write(f"    {name}\n    {{\n".encode())
            for k, v in e.items():
                if k == "value":
                    continue
                if isinstance(v, (list, tuple)):
                    v = _join(v)
                f.write(f"        {k:<15} {v};\n".encode())
            if "value" in e:
                f.write(b"        value           ")
                _write_value(f, e["value"], typ, binary, precision)
                f.write(b";\n")
            f.write(b"    }\n")
        f.write(b"}\n")
        f.write(END.encode())


def _join(v):
    out = []
    for x in v:
        if isinstance(x, (list, tuple)):
            out.append("(" + _join(x) + ")")
        else:
            out.append(str(x))
    return " ".join(out)


# ----------------------------------------------------------------------------------------
# time names  (controlDict: timeFormat general; timePrecision 6)
# ----------------------------------------------------------------------------------------
def time_name(t, precision=6):
    """OpenFOAM Time::timeName with timeFormat general: C++ ostream << setprecision(p)."""
    s = f"{t:.{precision}g}"
    if "e" in s:
        m, e = s.split("e")
        if "." in m:
            m = m.rstrip("0").rstrip(".")
        sign = "-" if e.startswith("-") else "+"
        e = e.lstrip("+-").lstrip("0") or "0"
        s = f"{m}e{sign}{int(e):02d}"
    return s


def time_dirs(case_dir):
    """Numeric directories of a case, sorted by value: [(value, name)]."""
    out = []
    for nm in os.listdir(case_dir):
        if os.path.isdir(os.path.join(case_dir, nm)):
            try:
                out.append((float(nm), nm))
            except ValueError:
                pass
    return sorted(out)
