"""Checkpoint I/O on CPU (no GPU): the reader against the reference's own parser fixture (jlama-tests/.../safetensors/
TestParser.java:41-69), Jlama's Q4/I8 + ".qb" conventions (SafeTensorSupport.java:264-277, Weights.java:49-66,153-171), the
writer round trip, index.json shard maps and the header-length guards (SafeTensorSupport.java:59-70)."""
import json
import os
import struct

import numpy as np
import pytest

from jlama_b200 import native, synth
from jlama_b200 import safetensors_io as sio


def test_reference_parser_fixture(tmp_path):
    # TestParser.simpleTest: preamble 0x59 = 89 header bytes, one F32 [2,2] tensor, __metadata__ {"foo":"bar"}
    header = b'{"test":{"dtype":"F32","shape":[2,2],"data_offsets":[0,16]},"__metadata__":{"foo":"bar"}}'
    assert len(header) == 0x59
    blob = bytes.fromhex("5900000000000000") + header + struct.pack("<4f", 1.0, 2.0, 3.0, 4.0)
    p = tmp_path / "t.safetensors"
    p.write_bytes(blob)
    with sio.SafeTensors(str(p)) as st:
        assert st.names() == ["test"]
        assert st.info("test") == {"dtype": "F32", "dtype_code": native.F32, "shape": (2, 2), "nbytes": 16}
        t = st.get("test")
        assert t.shape == (2, 2) and t.tolist() == [[1.0, 2.0], [3.0, 4.0]]
        assert st.metadata("foo") == "bar" and st.metadata("nope") is None


def test_header_length_guards(tmp_path):
    p = tmp_path / "neg.safetensors"
    p.write_bytes(struct.pack("<q", -5) + b"{}")
    with pytest.raises(sio.SafeTensorsError, match="negative"):
        sio.SafeTensors(str(p))
    p2 = tmp_path / "big.safetensors"
    p2.write_bytes(struct.pack("<q", (1 << 30) + 1) + b"{}")
    with pytest.raises(sio.SafeTensorsError, match="exceeds"):
        sio.SafeTensors(str(p2))
    p3 = tmp_path / "off.safetensors"
    h = b'{"t":{"dtype":"F32","shape":[4],"data_offsets":[0,16]}}'
    p3.write_bytes(struct.pack("<q", len(h)) + h + b"\0" * 8)  # data shorter than the offsets claim
    with pytest.raises(sio.SafeTensorsError):
        sio.SafeTensors(str(p3))


def test_jq4_checkpoint_round_trip(tmp_path):
    cfg = synth.get_config("tiny")
    w = synth.make_weights(cfg)
    d = tmp_path / "tiny-JQ4"
    sio.save_checkpoint(str(d), w, cfg)
    raw = (d / "model.safetensors").read_bytes()
    hlen = struct.unpack("<q", raw[:8])[0]
    hdr = json.loads(raw[8:8 + hlen])
    q = hdr["model.layers.0.self_attn.q_proj.weight"]
    assert q["dtype"] == "Q4" and q["shape"] == [cfg["E"], cfg["E"]]
    assert q["data_offsets"][1] - q["data_offsets"][0] == cfg["E"] * cfg["E"] // 2  # N*K/2 bytes
    qb = hdr["model.layers.0.self_attn.q_proj.weight.qb"]
    assert qb["dtype"] == "F32" and qb["shape"] == [cfg["E"], cfg["E"] // 32]
    assert hdr["model.norm.weight"]["dtype"] == "F32"
    with sio.SafeTensors(str(d)) as st:
        assert st.majority_dtype() == native.Q4  # .qb tensors are not counted (Weights.java:52)
        offs = [hdr[n]["data_offsets"][0] for n in st.names()]
        assert offs == sorted(offs)  # TensorInfo.compareTo order
        for name, (dt, data, scales) in w.items():
            dt2, data2, scales2 = st.load(name)
            assert dt2 == dt and np.array_equal(data2, data)
            assert (scales is None and scales2 is None) or np.array_equal(scales2, scales)
    mc = sio.config_from_json(str(d / "config.json"))
    assert (mc.embedding_length, mc.hidden_length, mc.num_heads, mc.num_kv_heads, mc.num_layers, mc.vocab_size, mc.head_size) == (
        cfg["E"], cfg["H"], cfg["heads"], cfg["kv_heads"], cfg["layers"], cfg["vocab"], cfg["E"] // cfg["heads"])
    assert mc.rope_theta == cfg["rope_theta"] and abs(mc.layer_norm_eps - cfg["eps"]) < 1e-12


def test_index_json_shards_and_i8(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.standard_normal((4, 64)).astype(np.float32)
    q8 = rng.integers(-127, 128, (8, 64), dtype=np.int8)
    s8 = rng.random((8, 2), dtype=np.float32)
    bf = rng.integers(0, 65536, (2, 32), dtype=np.uint16)
    d = tmp_path / "sharded"
    d.mkdir()
    sio.write_safetensors(str(d / "model-00001-of-00002.safetensors"), {"a": (native.F32, a, None), "b": (native.BF16, bf, None)})
    sio.write_safetensors(str(d / "model-00002-of-00002.safetensors"), {"w": (native.I8, q8, None), "w.qb": (native.F32, s8, None)},
                          metadata={"format": "pt"})
    (d / "model.safetensors.index.json").write_text(json.dumps({"metadata": {}, "weight_map": {
        "a": "model-00001-of-00002.safetensors", "b": "model-00001-of-00002.safetensors",
        "w": "model-00002-of-00002.safetensors", "w.qb": "model-00002-of-00002.safetensors"}}))
    with sio.SafeTensors(str(d)) as st:
        assert sorted(st.names()) == ["a", "b", "w", "w.qb"]
        assert np.array_equal(st.get("a"), a) and np.array_equal(st.get("b"), bf)
        dt, data, scales = st.load("w")
        assert dt == native.I8 and np.array_equal(data, q8) and np.array_equal(scales, s8)
        assert st.metadata("format") == "pt"
    with pytest.raises(sio.SafeTensorsError):
        sio.SafeTensors(str(tmp_path / "missing-dir-or-file"))


def _file(tmp_path, name, header, data=b""):
    h = header if isinstance(header, bytes) else json.dumps(header).encode()
    p = tmp_path / name
    p.write_bytes(struct.pack("<q", len(h)) + h + data)
    return str(p)


@pytest.mark.parametrize("entry,why", [
    ({"dtype": "F32", "shape": [4, 4], "data_offsets": [0, 16]}, "does not match"),        # 64 bytes promised, 16 present
    ({"dtype": "Q4", "shape": [2, 64], "data_offsets": [0, 16]}, "does not match"),        # Q4 is N*K/2 = 64 bytes
    ({"dtype": "F32", "shape": [-1, 4], "data_offsets": [0, 16]}, "non-negative"),
    ({"dtype": "F32", "shape": [2, 2.5], "data_offsets": [0, 16]}, "non-negative"),
    ({"dtype": "F32", "shape": [4], "data_offsets": [-16, 0]}, "non-negative"),            # would point before the data section
    ({"dtype": "F32", "shape": [4], "data_offsets": [16, 0]}, "exceed"),
    ({"dtype": "F32", "shape": [4], "data_offsets": [0, 1 << 40]}, "exceed"),
    ({"dtype": "F32", "shape": [1 << 40, 1 << 40], "data_offsets": [0, 16]}, "overflow"),
    ({"dtype": "F32", "shape": [1, 1, 1, 1, 4], "data_offsets": [0, 16]}, "malformed"),
    ({"dtype": 7, "shape": [4], "data_offsets": [0, 16]}, "malformed"),
    ({"dtype": "F32", "shape": "4", "data_offsets": [0, 16]}, "malformed"),
    ({"dtype": "F32", "shape": [4], "data_offsets": [0]}, "malformed"),
    ("F32", "malformed"),
])
def test_header_entries_are_validated_before_anything_reads_through_them(tmp_path, entry, why):
    """Every later read (jl_st_data, the loader's row slices, the quantiser) trusts shape, dtype and data_offsets: a header whose
    numbers do not describe bytes that exist in the file is rejected when it is opened."""
    with pytest.raises(sio.SafeTensorsError, match=why):
        sio.SafeTensors(_file(tmp_path, "bad.safetensors", {"t": entry}, b"\0" * 16))


def test_unknown_dtypes_are_listed_but_not_interpreted(tmp_path):
    p = _file(tmp_path, "u.safetensors", {"x": {"dtype": "F8_E4M3", "shape": [3], "data_offsets": [0, 3]},
                                          "y": {"dtype": "F16", "shape": [2], "data_offsets": [3, 7]}}, b"abc" + b"\0" * 4)
    with sio.SafeTensors(p) as st:
        assert st.info("x")["dtype_code"] == -1 and st.info("x")["nbytes"] == 3
        assert st.info("y")["dtype"] == "F16"
        assert st.majority_dtype() == native.F32  # F16 counts as F32 (Weights.java:49-66)


def test_index_json_cannot_name_files_outside_the_checkpoint_directory(tmp_path):
    d = tmp_path / "m"
    d.mkdir()
    sio.write_safetensors(str(tmp_path / "outside.safetensors"), {"a": (native.F32, np.zeros((1, 4), np.float32), None)})
    for shard in ("../outside.safetensors", str(tmp_path / "outside.safetensors"), ""):
        (d / "model.safetensors.index.json").write_text(json.dumps({"weight_map": {"a": shard}}))
        with pytest.raises(sio.SafeTensorsError, match="shard"):
            sio.SafeTensors(str(d))


def test_mutated_headers_never_crash_the_reader(tmp_path):
    """Byte-level mutations of a valid header: the reader either rejects the file or returns tensors whose bytes are all
    inside the mapping (every byte of every listed tensor is touched)."""
    rng = np.random.default_rng(11)
    good = {"a": {"dtype": "F32", "shape": [2, 8], "data_offsets": [0, 64]},
            "w": {"dtype": "Q4", "shape": [2, 64], "data_offsets": [64, 128]},
            "w.qb": {"dtype": "F32", "shape": [2, 2], "data_offsets": [128, 144]},
            "__metadata__": {"k": "vé\\n"}}
    base = json.dumps(good).encode()
    data = bytes(range(144))
    opened = 0
    for it in range(400):
        h = bytearray(base)
        for _ in range(int(rng.integers(1, 4))):
            kind = int(rng.integers(0, 4))
            pos = int(rng.integers(0, len(h)))
            if kind == 0:
                h[pos] = int(rng.integers(0, 256))
            elif kind == 1:
                del h[pos:pos + int(rng.integers(1, 6))]
            elif kind == 2:
                h[pos:pos] = bytes(rng.integers(32, 127, size=int(rng.integers(1, 6)), dtype=np.uint8))
            else:
                tok = [b"-1", b"99999999999999999999", b"1e309", b"[", b"{", b"\\u12", b"\"", b"null", b"9223372036854775807"][int(rng.integers(0, 9))]
                h[pos:pos] = tok
        p = _file(tmp_path, "m.safetensors", bytes(h), data)
        try:
            st = sio.SafeTensors(p)
        except (sio.SafeTensorsError, UnicodeDecodeError):
            continue
        opened += 1
        for name in st.names():
            inf = st.info(name)
            i = st._index[name][0]
            if inf["nbytes"]:
                ptr = st.lib.jl_st_data(st.h, i)
                raw = np.ctypeslib.as_array(__import__("ctypes").cast(ptr, __import__("ctypes").POINTER(__import__("ctypes").c_uint8)), shape=(inf["nbytes"],))
                assert int(raw.astype(np.uint64).sum()) >= 0
        st.close()
    assert opened > 0  # some mutations only touch names / metadata and must still load


def test_config_json_mixtral_fields(tmp_path):
    cfg = synth.get_config("tiny-mixtral")
    sio.save_checkpoint(str(tmp_path / "mx"), {}, cfg)
    mc = sio.config_from_json(str(tmp_path / "mx" / "config.json"))
    assert (mc.num_experts, mc.experts_per_token) == (8, 2) and mc.num_layers == cfg["layers"]
    dense = synth.get_config("tiny")
    sio.save_checkpoint(str(tmp_path / "d"), {}, dense)
    mc = sio.config_from_json(str(tmp_path / "d" / "config.json"))
    assert (mc.num_experts, mc.experts_per_token) == (0, 0)
    (tmp_path / "broken.json").write_text("{\"hidden_size\": 12")
    with pytest.raises(sio.SafeTensorsError):
        sio.config_from_json(str(tmp_path / "broken.json"))


def test_writer_rejects_descriptions_it_cannot_write(tmp_path):
    import ctypes as C
    lib = native.load()
    a = np.zeros(4, np.float32)

    def write(ndim=1, nbytes=16, shape=(4, 0, 0, 0), name=b"t", data=a.ctypes.data):
        names = (C.c_char_p * 1)(name)
        return lib.jl_st_write(os.fsencode(str(tmp_path / "w.safetensors")), 1, names, (C.c_int * 1)(native.F32), (C.c_int * 1)(ndim),
                               (C.c_int64 * 4)(*shape), (C.c_void_p * 1)(data), (C.c_int64 * 1)(nbytes), 0, None)

    assert write() == 0
    with sio.SafeTensors(str(tmp_path / "w.safetensors")) as st:
        assert st.info("t")["shape"] == (4,)
    for kw in (dict(ndim=5), dict(ndim=-1), dict(nbytes=-16), dict(shape=(-4, 0, 0, 0)), dict(name=None), dict(data=None)):
        assert write(**kw) == native.JL_ERR_INVALID, kw
        assert b"st_write" in lib.jl_st_last_error()
    # a file this writer produced with a shape that does not match its bytes is refused by the reader, not by luck of the writer
    assert write(nbytes=8) == 0
    with pytest.raises(sio.SafeTensorsError, match="does not match"):
        sio.SafeTensors(str(tmp_path / "w.safetensors"))


def test_config_json_head_dim_and_linear_rope_scaling(tmp_path):
    """MistralConfig / LlamaConfig: `head_dim` overrides hidden_size / heads (Config.java:254); only rope_type "linear" scales the table
    (LlamaConfig.java:55-56).  MistralModel extends LlamaModel without overriding anything, so a Mistral checkpoint is this path."""
    base = {"hidden_size": 4096, "intermediate_size": 14336, "num_attention_heads": 32, "num_key_value_heads": 8, "num_hidden_layers": 32,
            "vocab_size": 32000, "max_position_embeddings": 32768, "rms_norm_eps": 1e-5, "rope_theta": 1000000.0}
    p = tmp_path / "config.json"
    p.write_text(json.dumps(dict(base, head_dim=96)))
    mc = sio.config_from_json(str(p))
    assert mc.head_size == 96 and mc.num_kv_heads == 8 and mc.rope_theta == 1000000.0 and mc.rope_scaling == 1.0
    p.write_text(json.dumps(dict(base, rope_scaling={"rope_type": "linear", "factor": 4.0})))
    assert sio.config_from_json(str(p)).rope_scaling == 4.0 and sio.config_from_json(str(p)).head_size == 128
    p.write_text(json.dumps(dict(base, rope_scaling={"rope_type": "llama3", "factor": 8.0})))
    assert sio.config_from_json(str(p)).rope_scaling == 1.0  # ignored by the reference, ignored here
    del base["num_key_value_heads"]
    p.write_text(json.dumps(base))
    assert sio.config_from_json(str(p)).num_kv_heads == 32  # multi-head attention when the key is absent
