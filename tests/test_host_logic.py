"""CPU tests of the product's host-side logic (numpy mirrors + pure host functions of the C ABI) against the
oracle, and of the C-ABI library itself: it must load without a GPU and export every declared symbol."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from jlama_b200 import native
    return native.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "jlama_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(jl_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    from jlama_b200 import native
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name


def _c_kind(decl):
    """'const float *x' -> 'ptr', 'int64_t n' -> 'i64', ... for a C parameter or return type"""
    d = decl.strip()
    if "*" in d or "[" in d:
        return "ptr"
    words = [w for w in re.split(r"\s+", d) if w not in ("const", "unsigned", "struct")]
    t = words[0]
    return {"int": "i32", "int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "float": "f32", "double": "f64", "void": "void"}[t]


def _ctypes_kind(t):
    if t is None:
        return "void"
    if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or hasattr(t, "_type_") and isinstance(t._type_, type):
        return "ptr"
    return {C.c_int: "i32", C.c_int32: "i32", C.c_int64: "i64", C.c_uint64: "u64", C.c_float: "f32", C.c_double: "f64"}[t]


def test_ctypes_signatures_agree_with_the_header_prototypes():
    """Every prototype of include/jlama_b200.h against the ctypes signature the Python binding declares for it: parameter count and,
    per parameter and return value, the machine type (pointer / int32 / int64 / uint64 / float / double).  A binding that passes a 32-bit
    int where the ABI takes int64_t, or forgets a parameter, fails here instead of corrupting an argument at run time."""
    from jlama_b200 import native
    hdr = open(os.path.join(ROOT, "include", "jlama_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    hdr = re.sub(r"typedef struct \{.*?\} \w+;", "", hdr, flags=re.S)  # struct bodies (function-pointer members) are not prototypes
    hdr = re.sub(r"^\s*#.*$", "", hdr, flags=re.M)                      # preprocessor lines
    protos = re.findall(r"([A-Za-z_][\w\s\*]*?)\b(jl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr)
    seen = set()
    for ret, name, params in protos:
        res, args = native.SIGNATURES[name]
        plist = [] if params.strip() in ("", "void") else [p for p in params.split(",")]
        assert len(plist) == len(args), (name, len(plist), len(args))
        assert _c_kind(ret + " x") == _ctypes_kind(res), (name, ret)
        for i, (p, a) in enumerate(zip(plist, args)):
            assert _c_kind(p) == _ctypes_kind(a), (name, i, p.strip(), a)
        seen.add(name)
    assert seen == set(native.SIGNATURES), seen ^ set(native.SIGNATURES)
    # the structs mirrored by hand: member counts
    raw = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "jlama_b200.h")).read(), flags=re.S)
    bodies = {name: body for body, name in re.findall(r"typedef struct \{([^{}]*)\} (\w+);", raw)}
    for cname, ctype in (("jl_sched_stats", native.SchedStats), ("jl_sched_request_info_t", native.SchedRequestInfo),
                         ("jl_sched_backend", native.SchedBackend), ("jl_dctx", native.Dctx), ("jl_model_config", native.ModelConfig)):
        n_members = 0
        for stmt in bodies[cname].split(";"):
            stmt = stmt.strip()
            if stmt:
                n_members += 1 if "(*" in stmt else len(stmt.split(","))
        assert n_members == len(ctype._fields_), (cname, n_members, len(ctype._fields_))


def test_init_fails_loudly_without_gpu_or_succeeds_on_sm100(lib):
    h = C.c_void_p()
    rc = lib.jl_init(0, C.byref(h), None)
    if rc == 0:
        lib.jl_shutdown(h)  # on the GPU box
    else:
        assert rc < 0 and not h.value
        assert b"no CPU fallback" in lib.jl_last_error(None) or b"sm_" in lib.jl_last_error(None)


def test_quantiser_mirrors_match_oracle_bit_for_bit(oracle):
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((64, 256)) * 0.02).astype(np.float32)
    x[3, :32] = 0
    x[5, 40] = np.float32(1e-45)
    x[7, 64:96] = -x[7, 64:96].max()  # ties on |v|: first index wins
    q1, s1 = oracle.quantize_q4(x)
    q2, s2 = T.quantize_q4(x)
    assert np.array_equal(q1, q2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32))
    assert np.array_equal(oracle.dequantize_q4(q1, s1), T.dequantize_q4(q2, s2))
    q1, s1 = oracle.quantize_q8_weights(x)
    q2, s2 = T.quantize_q8_weights(x)
    assert np.array_equal(q1, q2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32))
    a = rng.uniform(-1, 100, (8, 256)).astype(np.float32)
    a[2, 32:64] = 0
    q1, s1 = oracle.quantize_q8_act(a)
    q2, s2 = T.quantize_q8_activations(a)
    assert np.array_equal(q1, q2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32))
    sp = np.concatenate([a.ravel(), np.array([np.nan, np.inf, -np.inf, 1.00390625, 1.01171875], dtype=np.float32)])
    assert np.array_equal(oracle.f32_to_bf16(sp), T.float32_to_bfloat16(sp))


def test_host_functions_match_oracle(lib, oracle):
    from jlama_b200 import native
    t = np.empty((300 * 32, 2), dtype=np.float32)
    assert lib.jl_precompute_freqs_cis(64, 300, 500000.0, 1.0, native.ptr(t)) == 0
    assert np.array_equal(t, oracle.precompute_freqs_cis(64, 300, 500000.0, 1.0))
    for args in [(32, 8192, 1024, 4), (16, 131072, 512, 4), (32, 8192, 1024, 2), (2, 256, 128, 4), (4, 512, 128, 4)]:
        a, b = C.c_int(), C.c_int()
        assert lib.jl_kv_page_geometry(*args, 1 << 23, C.byref(a), C.byref(b)) == 0
        assert (a.value, b.value) == oracle.kv_page_solver(args[0], args[1], args[2], args[3])
    for shard in range(8):
        d = native.Dctx()
        assert lib.jl_dctx_build(4096, 4096, 14336, 128, 4, 32, shard, 8, 0, 1, C.byref(d)) == 0
        o = oracle.dctx(4096, 4096, 14336, 128, 4, 32, shard, 8)
        for name, _ in native.Dctx._fields_:
            assert getattr(d, name) == getattr(o, name), name
    assert lib.jl_dctx_build(4096, 4096, 14336, 128, 4, 32, 8, 8, 0, 1, C.byref(native.Dctx())) < 0


def test_synthetic_checkpoint_is_seeded_and_well_formed():
    from jlama_b200 import synth
    from jlama_b200.native import F32, Q4
    cfg = synth.get_config("tiny")
    w1, w2 = synth.make_weights(cfg), synth.make_weights(cfg)
    assert set(w1) == {n for n, *_ in synth.tensor_specs(cfg)}
    for k in w1:
        assert w1[k][0] == w2[k][0] and np.array_equal(w1[k][1], w2[k][1])
    dt, q, s = w1["model.layers.0.self_attn.q_proj.weight"]
    assert dt == Q4 and q.shape == (256, 128) and s.shape == (256, 8) and q.dtype == np.uint8
    assert w1["model.norm.weight"][0] == F32
    wd = synth.make_weights(cfg, mode="direct")
    assert wd["model.layers.1.mlp.down_proj.weight"][1].shape == (256, 256)
    # 8B weight bytes per token = 4.690 GB
    n = synth.linear_weight_count(synth.get_config("llama-3-8b"))
    assert abs(n * 0.625 / 1e9 - 4.690) < 0.01


def test_tensor_parallel_slicing_reassembles(oracle):
    from jlama_b200 import synth
    from jlama_b200.model import DistributedContext, _slice_cols, _slice_rows
    cfg = synth.get_config("small")
    w = synth.make_weights(cfg)
    t = w["model.layers.0.self_attn.o_proj.weight"]
    full = oracle.dequantize_q4(t[1], t[2])
    parts = []
    for r in range(2):
        d = DistributedContext(cfg, r, 2)
        dt, q, s = _slice_cols(t, d.attentionSegmentStart, d.attentionSegmentLength)
        parts.append(oracle.dequantize_q4(q, s))
    assert np.array_equal(np.concatenate(parts, axis=1), full)
    t = w["model.layers.0.self_attn.k_proj.weight"]
    full = oracle.dequantize_q4(t[1], t[2])
    parts = []
    for r in range(2):
        d = DistributedContext(cfg, r, 2)
        dt, q, s = _slice_rows(t, d.kvSegmentStart, d.kvSegmentLength)
        parts.append(oracle.dequantize_q4(q, s))
    assert np.array_equal(np.concatenate(parts, axis=0), full)


def test_sparsify_keeps_logical_indices():
    """Port of TestParser.testSparsify (jlama-tests/.../safetensors/TestParser.java:137-170): a column-sparse copy answers the
    logical indices of its range and refuses everything else; TensorShape.getOffset (:94-99) for the stored buffer."""
    from jlama_b200 import tensor as T
    dim = 96
    b = T.FloatBufferTensor(np.arange(dim * dim, dtype=np.float32).reshape(dim, dim))
    bt = b.sparsify(32, 20)
    assert b.data.size == dim * dim and bt.data.size == dim * 20 and bt.is_sparse()
    for row in (0, 5, dim - 1):
        for col in range(dim):
            if 32 <= col < 52:
                assert bt.get(row, col) == row * dim + col
                assert bt.data.reshape(-1)[bt.get_offset(row, col)] == row * dim + col
            else:
                with pytest.raises(IndexError):
                    bt.get(row, col)
    assert b.sparsify(0, dim) is b and bt.sparsify(0, 4) is bt  # AbstractTensor.java:152-154
    br = b.sparsify_rows(10, 7)
    assert br.get(12, 3) == 12 * dim + 3 and br.get_offset(12, 3) == 2 * dim + 3
    q = T.Q4ByteBufferTensor.from_float(np.random.default_rng(0).standard_normal((8, 128)).astype(np.float32))
    qs = q.sparsify(64, 32)
    assert qs.cols == 32 and np.array_equal(qs.to_float(), q.to_float()[:, 64:96])


def test_model_config_mirror_matches_the_c_struct(lib):
    """the ctypes mirror of jl_model_config (and the FFM StructLayout in java/.../CudaLlamaModel.java, 104 bytes) must not drift"""
    import ctypes as C
    from jlama_b200 import native
    assert lib.jl_model_config_size() == C.sizeof(native.ModelConfig) == 104
    assert native.ModelConfig.arch.offset == 100 and native.ModelConfig.rope_theta.offset == 40


def test_pure_host_functions_reject_degenerate_arguments_instead_of_faulting(lib, oracle):
    """The library never aborts the host process (the reference's natives call exit(), vector_gpu.c:96,116): zero / negative sizes in the
    pure host functions are status codes, not divisions by zero; valid random arguments agree with the oracle."""
    from jlama_b200 import native
    a, b = C.c_int(), C.c_int()
    for args in [(32, 8192, 1024, 0), (32, 8192, 1024, -4), (0, 8192, 1024, 4), (32, 0, 1024, 4), (32, 8192, 0, 4)]:
        assert lib.jl_kv_page_geometry(*args, 1 << 23, C.byref(a), C.byref(b)) == native.JL_ERR_INVALID
    assert lib.jl_kv_page_geometry(32, 8192, 1024, 4, 16, C.byref(a), C.byref(b)) == native.JL_ERR_INVALID  # page smaller than one position
    assert lib.jl_kv_page_geometry(32, 8192, 1024, 4, 1 << 23, None, C.byref(b)) == native.JL_ERR_INVALID
    rng = np.random.default_rng(3)
    for _ in range(300):
        layers, ctx_len = int(rng.integers(1, 129)), int(2 ** rng.integers(4, 18))
        kvl, esz = int(rng.integers(1, 65)) * 32, int(rng.choice([2, 4]))
        assert lib.jl_kv_page_geometry(layers, ctx_len, kvl, esz, 1 << 23, C.byref(a), C.byref(b)) == 0
        assert (a.value, b.value) == oracle.kv_page_solver(layers, ctx_len, kvl, esz)
        assert 1 <= a.value <= layers and 1 <= b.value <= ctx_len and 2 * esz * kvl * a.value * b.value <= (1 << 23)
    d = native.Dctx()
    for bad in [(4096, 4096, 14336, 0, 4, 32, 0, 8, 0, 1), (4096, 4096, 14336, 128, 0, 32, 0, 8, 0, 1), (4096, 4096, 14336, 128, 4, 32, 0, 0, 0, 1),
                (4096, 4096, 14336, 128, 4, 32, -1, 8, 0, 1), (4096, 4096, 14336, 128, 4, 32, 0, 8, 0, 0), (4096, 4096, 14336, 128, 4, 32, 0, 8, 2, 2)]:
        assert lib.jl_dctx_build(*bad, C.byref(d)) == native.JL_ERR_INVALID
    assert lib.jl_dctx_build(4096, 4096, 14336, 128, 4, 32, 0, 8, 0, 1, None) == native.JL_ERR_INVALID
    t = np.empty((4, 2), dtype=np.float32)
    for bad in [(0, 4), (3, 4), (4, 0), (-2, 4)]:
        assert lib.jl_precompute_freqs_cis(bad[0], bad[1], 10000.0, 1.0, native.ptr(t)) == native.JL_ERR_INVALID
    assert lib.jl_precompute_freqs_cis(4, 2, 10000.0, 1.0, None) == native.JL_ERR_INVALID
    assert lib.jl_model_config_size() == C.sizeof(native.ModelConfig)
    # handles that are NULL: status codes, never a fault
    assert lib.jl_sched_step(None, None) == native.JL_ERR_INVALID and lib.jl_sched_free(None) == native.JL_ERR_INVALID
    assert lib.jl_sched_submit(None, None, 0, 0, None, 0, 0, -1, C.c_float(0), 0) == -1
    assert lib.jl_st_close(None) == native.JL_ERR_INVALID and lib.jl_st_count(None) == native.JL_ERR_INVALID
    assert lib.jl_model_free(None) == native.JL_ERR_INVALID and lib.jl_model_kv_pages(None, 0) == native.JL_ERR_INVALID
    assert lib.jl_shutdown(None) == native.JL_ERR_INVALID and lib.jl_kernel_launches(None) == -1
