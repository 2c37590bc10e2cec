"""Episode files in the reference's BSON layout (src/data.jl:60-71) -- waves.jl_b200/bson.py.  CPU only.

What is pinned here: the BSON container against the byte-level examples of bsonspec.org, the BSON.jl lowering rules against
literal expected documents, and a full write -> read-back of an episode.  Not pinned (no Julia in this image): that BSON.jl
itself opens the file."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import waves_b200 as wb  # noqa: E402
from waves_b200 import bson  # noqa: E402


def test_container_matches_the_bson_specification_examples():
    # bsonspec.org: {"hello": "world"}
    assert bson.encode_document({"hello": "world"}) == b"\x16\x00\x00\x00\x02hello\x00\x06\x00\x00\x00world\x00\x00"
    # bsonspec.org: {"BSON": ["awesome", 5.05, 1986]} (int32 element)
    want = (b"\x31\x00\x00\x00\x04BSON\x00\x26\x00\x00\x00\x020\x00\x08\x00\x00\x00awesome\x00\x011\x00\x33\x33\x33\x33\x33\x33\x14\x40"
            b"\x102\x00\xc2\x07\x00\x00\x00\x00")
    assert bson.encode_document({"BSON": ["awesome", 5.05, bson.Int32(1986)]}) == want
    doc, end = bson.decode_document(want)
    assert end == len(want) and doc == {"BSON": ["awesome", 5.05, 1986]}
    # Julia Int -> int64, binary subtype 0, bool, null
    d = {"n": 7, "b": b"\x01\x02\x03", "t": True, "z": None, "sub": {"k": -1}}
    enc = bson.encode_document(d)
    assert enc[4:5] == b"\x12" and struct.unpack_from("<q", enc, 7)[0] == 7
    assert bson.decode_document(enc)[0] == d


def test_lowering_of_julia_arrays_and_structs():
    # Matrix{Float32}(3, 2) column-major: our C-order image has the axes reversed
    jl = np.arange(6, dtype=np.float32).reshape(2, 3)        # Julia A[i, j] = jl[j, i], size (3, 2)
    a = bson.jl_array(jl)
    assert a["tag"] == "array" and a["size"] == [3, 2] and a["type"] == {"tag": "datatype", "params": [], "name": ["Core", "Float32"]}
    assert a["data"] == jl.tobytes() and len(a["data"]) == 24
    cyl = wb.Cylinders([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]], [0.1, 0.2, 0.3], [10.0, 20.0, 30.0])
    c = bson.lower_cylinders(cyl)
    assert c["tag"] == "struct" and c["type"]["name"] == ["Waves", "Cylinders"] and len(c["data"]) == 3
    pos = c["data"][0]
    assert pos["size"] == [3, 2]
    # column-major (3, 2): all x coordinates, then all y coordinates (src/designs.jl:69-73)
    assert np.array_equal(np.frombuffer(pos["data"], np.float32), np.array([1, 3, 5, 2, 4, 6], np.float32))
    cloak = wb.Cloak(cyl, wb.Cylinders([[0.0, 0.0]], [2.0], [3100.0]))
    d = bson.lower_design(cloak)
    assert d["type"]["name"] == ["Waves", "Cloak"]
    assert d["data"][0]["type"]["name"] == ["Waves", "AdjustableRadiiScatterers"] and d["data"][0]["data"][0] == c
    assert d["data"][1]["type"]["name"] == ["Waves", "Cylinders"]


def _episode(actions=3, steps=10, res=8):
    rng = np.random.default_rng(0)
    dim = wb.TwoDim(5.0, 16)
    ds = wb.build_triple_ring_design_space()
    ep = wb.Episode()
    d = ds.rand(rng)
    t0 = np.float32(0.0)
    for _ in range(actions):
        ts = wb.build_tspan(t0, 1e-5, steps)
        act = wb.build_action_space(d, 0.25).rand(rng)
        ep.s.append((ts, rng.standard_normal((4, res, res)).astype(np.float32), d))
        ep.a.append(act)
        ep.t.append(ts)
        ep.y.append(rng.standard_normal((steps + 1, 3)).astype(np.float32))
        d = ds(d, act)
        t0 = ts[-1]
    return ep, dim


def test_episode_file_round_trip(tmp_path):
    ep, dim = _episode()
    path = str(tmp_path / "episode1.bson")
    ep.save_bson(path, dim)
    raw = open(path, "rb").read()
    assert struct.unpack_from("<i", raw, 0)[0] == len(raw) and raw[-1] == 0
    doc, _ = bson.decode_document(raw)
    assert sorted(doc) == ["a", "s", "t", "y"]                       # BSON.bson(path, s = ..., a = ..., t = ..., y = ...)
    assert doc["y"][0]["size"] == [11, 3] and doc["t"][0]["size"] == [11]
    assert doc["s"][0]["type"]["name"] == ["Waves", "WaveEnvState"]
    assert [f.get("type", {}).get("name") for f in doc["s"][0]["data"]][0] == ["Waves", "TwoDim"]
    assert doc["s"][0]["data"][2]["size"] == [8, 8, 4]               # imresize'd (res_x, res_y, 4) observation (src/env.jl:132-137)
    got = bson.load_episode(path)
    assert len(got["s"]) == len(got["a"]) == len(got["t"]) == len(got["y"]) == len(ep)
    for i in range(len(ep)):
        assert np.array_equal(got["t"][i], ep.t[i]) and np.array_equal(got["y"][i], ep.y[i])
        st = got["s"][i]
        assert st["type"] == "Waves.WaveEnvState"
        jdim, ts, img, design = st["fields"]
        assert np.array_equal(jdim["fields"][0], dim.x) and np.array_equal(jdim["fields"][1], dim.y)
        assert np.array_equal(ts, ep.s[i][0]) and np.array_equal(img, ep.s[i][1])
        cfg = design["fields"][0]["fields"][0]["fields"]             # Cloak -> AdjustableRadiiScatterers -> Cylinders -> (pos, r, c)
        assert np.array_equal(cfg[0].T, ep.s[i][2].config.pos) and np.array_equal(cfg[1], ep.s[i][2].config.r)
        assert np.array_equal(design["fields"][1]["fields"][2], ep.s[i][2].core.c)
        act = got["a"][i]
        assert act["type"] == "Waves.AdjustableRadiiScatterers" and np.array_equal(act["fields"][0]["fields"][1], ep.a[i].r)
    # y column-major on disk: the first column of the Julia matrix (the total-energy trace) is contiguous
    y0 = np.frombuffer(doc["y"][0]["data"], np.float32)
    assert np.array_equal(y0[:11], ep.y[0][:, 0])
