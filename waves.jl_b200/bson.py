"""Episodes in the reference's on-disk format: `FileIO.save(episode, path)` is `BSON.bson(path, s=…, a=…, t=…, y=…)`
(src/data.jl:60-62) and `Episode(path=…)` reads it back with `BSON.load` (src/data.jl:64-71, scripts/main.jl:144-150).

BSON.jl is a third-party dependency of the reference (`Project.toml:7`, no version pinned: the repository has no Manifest).  Its
file format is plain BSON (bsonspec.org) plus a "lowering" of Julia values into tagged documents:

  Vector{Float32}, Matrix{Float32}, ... (bits element type)  {tag: "array", type: <datatype>, size: [Int64...], data: <binary, column-major>}
  Vector of non-bits elements (structs, vectors)             a plain BSON array of the lowered elements
  any other struct                                           {tag: "struct", type: <datatype>, data: [lowered fields in declaration order]}
  DataType                                                   {tag: "datatype", name: [module path..., type name], params: [<datatype>...]}

This module restates that lowering for exactly the values an Episode holds (`WaveEnvState`, `TwoDim`, `Cloak`,
`AdjustableRadiiScatterers`, `Cylinders`, Float32 arrays) and writes a file `Waves.Episode(path = …)` is meant to load.  Objects
that appear several times (the same `dim.x` vector in every state) are written out each time instead of through BSON.jl's
`_backrefs` table; `BSON.load` accepts both.  PARITY UNPINNED: Julia is not installed here, so no file written by this module
has been opened by BSON.jl; what the tests pin is the BSON container (byte-level known answers from bsonspec.org), the lowering
rules above (literal expected documents) and a full read-back.  The documented alternative that avoids this module is the
converter in julia/WavesB200.jl (`episode_from_npz`), which builds the Episode inside Julia and lets BSON.jl write it.
"""
from __future__ import annotations

import struct

import numpy as np

# ---- BSON container (bsonspec.org v1.1): the element types BSON.jl emits for this data ----


class Int32(int):
    """Marks an integer to be written as BSON int32 (plain Python ints are written as int64, Julia's Int)."""


def _cstring(s: str) -> bytes:
    b = s.encode("utf-8")
    if b"\x00" in b:
        raise ValueError("BSON key contains NUL")
    return b + b"\x00"


def _element(key: str, v) -> bytes:
    k = _cstring(key)
    if isinstance(v, bool):
        return b"\x08" + k + (b"\x01" if v else b"\x00")
    if isinstance(v, float):
        return b"\x01" + k + struct.pack("<d", v)
    if isinstance(v, Int32):
        return b"\x10" + k + struct.pack("<i", int(v))
    if isinstance(v, (int, np.integer)):
        return b"\x12" + k + struct.pack("<q", int(v))
    if isinstance(v, str):
        b = v.encode("utf-8") + b"\x00"
        return b"\x02" + k + struct.pack("<i", len(b)) + b
    if v is None:
        return b"\x0a" + k
    if isinstance(v, (bytes, bytearray, memoryview)):
        b = bytes(v)
        return b"\x05" + k + struct.pack("<i", len(b)) + b"\x00" + b
    if isinstance(v, dict):
        return b"\x03" + k + encode_document(v)
    if isinstance(v, (list, tuple)):
        return b"\x04" + k + encode_document({str(i): x for i, x in enumerate(v)})
    raise TypeError(f"no BSON encoding for {type(v).__name__}")


def encode_document(d: dict) -> bytes:
    body = b"".join(_element(str(k), v) for k, v in d.items())
    return struct.pack("<i", len(body) + 5) + body + b"\x00"


def decode_document(buf: bytes, pos: int = 0, as_list: bool = False):
    """-> (value, end).  Inverse of encode_document for the element types above."""
    (n,) = struct.unpack_from("<i", buf, pos)
    end = pos + n
    if n < 5 or end > len(buf) or buf[end - 1] != 0:
        raise ValueError("malformed BSON document")
    pos += 4
    out = {}
    while pos < end - 1:
        t = buf[pos]
        z = buf.index(b"\x00", pos + 1)
        key = buf[pos + 1:z].decode("utf-8")
        pos = z + 1
        if t == 0x01:
            v = struct.unpack_from("<d", buf, pos)[0]
            pos += 8
        elif t == 0x02:
            (ln,) = struct.unpack_from("<i", buf, pos)
            v = buf[pos + 4:pos + 4 + ln - 1].decode("utf-8")
            pos += 4 + ln
        elif t in (0x03, 0x04):
            v, pos = decode_document(buf, pos, as_list=(t == 0x04))
        elif t == 0x05:
            (ln,) = struct.unpack_from("<i", buf, pos)
            v = bytes(buf[pos + 5:pos + 5 + ln])
            pos += 5 + ln
        elif t == 0x08:
            v = buf[pos] != 0
            pos += 1
        elif t == 0x0A:
            v = None
        elif t == 0x10:
            v = Int32(struct.unpack_from("<i", buf, pos)[0])
            pos += 4
        elif t == 0x12:
            v = struct.unpack_from("<q", buf, pos)[0]
            pos += 8
        else:
            raise ValueError(f"BSON element type 0x{t:02x} is not used by this format")
        out[key] = v
    if as_list:
        return [out[str(i)] for i in range(len(out))], end
    return out, end


# ---- BSON.jl lowering of the Julia values an Episode holds ----
_JL_ELTYPE = {np.dtype(np.float32): ("Core", "Float32"), np.dtype(np.float64): ("Core", "Float64"),
              np.dtype(np.int64): ("Core", "Int64"), np.dtype(np.int32): ("Core", "Int32")}
_NP_ELTYPE = {v: k for k, v in _JL_ELTYPE.items()}


def jl_datatype(*path: str, params=()) -> dict:
    return {"tag": "datatype", "params": list(params), "name": list(path)}


def jl_array(a: np.ndarray, julia_shape=None) -> dict:
    """A bits-element Julia array.  `a` is the C-order image of the Julia array, i.e. its axes are the Julia axes reversed
    (our (ny, nx) plane is Julia's (nx, ny) matrix: same bytes); `julia_shape` overrides the reversed shape."""
    a = np.ascontiguousarray(a)
    if a.dtype not in _JL_ELTYPE:
        raise TypeError(f"no Julia bits type for dtype {a.dtype}")
    shape = tuple(reversed(a.shape)) if julia_shape is None else tuple(julia_shape)
    if int(np.prod(shape)) != a.size:
        raise ValueError("julia_shape does not match the number of elements")
    return {"tag": "array", "type": jl_datatype(*_JL_ELTYPE[a.dtype]), "size": [int(s) for s in shape], "data": a.tobytes()}


def jl_struct(type_path, fields, params=()) -> dict:
    return {"tag": "struct", "type": jl_datatype(*type_path, params=params), "data": list(fields)}


def lower_cylinders(c) -> dict:
    """Cylinders(pos::(n, 2) matrix, r, c) (src/designs.jl:69-73); our pos is (n, 2) C-order, Julia's is column-major."""
    return jl_struct(("Waves", "Cylinders"), [jl_array(np.asarray(c.pos, np.float32).T, julia_shape=(len(c.r), 2)),
                                              jl_array(np.asarray(c.r, np.float32)), jl_array(np.asarray(c.c, np.float32))])


def lower_design(d, scatterers="AdjustableRadiiScatterers") -> dict:
    """Cloak(config::AbstractScatterers, core::Cylinders) (src/designs.jl:210-213) or bare Cylinders; `scatterers` names the
    wrapper type of the adjustable part (AdjustableRadiiScatterers in build_triple_ring_design_space, src/designs.jl:179-181)."""
    from .host import Cloak, Cylinders
    if isinstance(d, Cloak):
        cfg = jl_struct(("Waves", scatterers), [lower_cylinders(d.config)])
        return jl_struct(("Waves", "Cloak"), [cfg, lower_cylinders(d.core)])
    if isinstance(d, Cylinders):
        return lower_cylinders(d)
    raise TypeError(f"no lowering for design type {type(d).__name__}")


def lower_action(a, scatterers="AdjustableRadiiScatterers") -> dict:
    """rand(action_space(env)) is an AbstractScatterers around Cylinders (src/designs.jl:186-191); the Python mirror keeps the
    bare Cylinders."""
    from .host import Cylinders
    if isinstance(a, Cylinders) and scatterers:
        return jl_struct(("Waves", scatterers), [lower_cylinders(a)])
    return lower_design(a, scatterers)


def lower_state(dim, tspan, image, design, scatterers="AdjustableRadiiScatterers") -> dict:
    """WaveEnvState(dim::TwoDim, tspan, wave, design) (src/env.jl:5-10); image is (4, res_y, res_x) = Julia (res_x, res_y, 4)."""
    jdim = jl_struct(("Waves", "TwoDim"), [jl_array(np.asarray(dim.x, np.float32)), jl_array(np.asarray(dim.y, np.float32))])
    return jl_struct(("Waves", "WaveEnvState"), [jdim, jl_array(np.asarray(tspan, np.float32)), jl_array(np.asarray(image, np.float32)),
                                                 lower_design(design, scatterers)])


def episode_document(ep, dim, scatterers="AdjustableRadiiScatterers") -> dict:
    """The document BSON.bson(path, s = ep.s, a = ep.a, t = ep.t, y = ep.y) writes (src/data.jl:60-62).  y[i] is our
    (steps+1, 3) signal = Julia's (steps+1, 3) Matrix (env.signal, src/env.jl:111-114), so it is transposed to column-major."""
    return {
        "s": [lower_state(dim, s[0], s[1], s[2], scatterers) for s in ep.s],
        "a": [lower_action(a, scatterers) for a in ep.a],
        "t": [jl_array(np.asarray(t, np.float32)) for t in ep.t],
        "y": [jl_array(np.asarray(y, np.float32).T, julia_shape=np.asarray(y).shape) for y in ep.y],
    }


def save_episode(ep, dim, path: str, scatterers="AdjustableRadiiScatterers") -> None:
    with open(path, "wb") as f:
        f.write(encode_document(episode_document(ep, dim, scatterers)))


# ---- reading back (what tags[:array] / tags[:struct] of BSON.jl do, into plain Python values) ----
def raise_value(v):
    """Lowered document -> NumPy arrays (C-order images, axes reversed) / {"type": name, "fields": [...]} / lists."""
    if isinstance(v, list):
        return [raise_value(x) for x in v]
    if isinstance(v, dict) and v.get("tag") == "array":
        dt = _NP_ELTYPE[tuple(v["type"]["name"])]
        return np.frombuffer(v["data"], dtype=dt).reshape(tuple(reversed(v["size"]))).copy()
    if isinstance(v, dict) and v.get("tag") == "struct":
        return {"type": ".".join(v["type"]["name"]), "fields": [raise_value(x) for x in v["data"]]}
    if isinstance(v, dict) and "tag" not in v:
        return {k: raise_value(x) for k, x in v.items()}
    return v


def load_episode(path: str) -> dict:
    """-> {"s": [...], "a": [...], "t": [arrays], "y": [(steps+1, 3) arrays]} with structs as {"type", "fields"} dicts."""
    with open(path, "rb") as f:
        doc, _ = decode_document(f.read())
    out = raise_value(doc)
    out["y"] = [np.ascontiguousarray(y.T) for y in out["y"]]   # back to the mirror's (steps+1, 3)
    return out
