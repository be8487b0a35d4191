"""GPU parity tests of the TensorOperations drop-in, modelled on the reference's operator differential
tests (jlama-tests/src/test/java/com/github/tjake/jlama/tensor/operations/TestOperations.java): every
(A dtype, B dtype) pair the reference supports, its distributions (A ~ U(-1,100), W ~ U(0,1), :94-109) and
sizes (SIZE=1024, ROWS=128, BATCH=32, :46-48), compared with the CPU oracle.  All calls go through the C ABI.

Bars: integer/byte outputs (Q8/BF16/Q4 quantisers) bit-exact; float GEMMs within 1e-5 of max|C| of the
oracle's same-arithmetic restatement (the reference's own tests allow 1 % on the sum, TestOperations.java:138).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZE, ROWS, BATCH = 1024, 128, 32


def _mk(rng, rows, cols, lo=-1.0, hi=100.0):
    return rng.uniform(lo, hi, (rows, cols)).astype(np.float32)


def _tensors(T, oracle, kind, x):
    """product tensor + oracle tensor of the same values"""
    if kind == "F32":
        return T.FloatBufferTensor(x), oracle.f32(x)
    if kind == "BF16":
        t = T.BFloat16BufferTensor.from_float(x)
        return t, oracle.OTensor(oracle.BF16, t.data)
    if kind == "Q4":
        t = T.Q4ByteBufferTensor.from_float(x)
        return t, oracle.OTensor(oracle.Q4, t.data, t.scales)
    if kind == "I8W":
        t = T.Q8ByteBufferTensor.from_float(x)
        return t, oracle.OTensor(oracle.I8, t.data, t.scales)
    if kind == "I8A":
        q, s = T.quantize_q8_activations(x)
        return T.Q8ByteBufferTensor(q, s), oracle.OTensor(oracle.I8, q, s)
    raise ValueError(kind)


PAIRS = [("F32", "F32"), ("F32", "BF16"), ("F32", "Q4"), ("I8A", "Q4"), ("BF16", "BF16"), ("BF16", "Q4"),
         ("F32", "I8W"), ("I8A", "I8W")]


@pytest.mark.parametrize("akind,bkind", PAIRS)
@pytest.mark.parametrize("M", [1, 3, BATCH])
def test_batch_dot_product_pairs(cuda_ops, oracle, akind, bkind, M):
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(hash((akind, bkind, M)) & 0xFFFF)
    a, ao = _tensors(T, oracle, akind, _mk(rng, M, SIZE))
    b, bo = _tensors(T, oracle, bkind, _mk(rng, ROWS, SIZE, 0.0, 1.0))
    if bkind in ("Q4", "I8W"):
        cuda_ops.register_model_tensor(b)
    c = T.FloatBufferTensor(np.zeros((M, ROWS), dtype=np.float32))
    cuda_ops.batch_dot_product(c, a, b, 0, 0, SIZE)
    ref = oracle.batch_dot(ao, bo, 0, 0, SIZE, 0, 0, ROWS)
    scale = np.abs(ref).max()
    assert np.abs(c.data - ref).max() <= 2e-5 * scale, (akind, bkind)
    # the reference's own bar: 1 % on the sum of all outputs
    assert abs(c.data.sum(dtype=np.float64) - ref.sum(dtype=np.float64)) <= 0.01 * abs(ref.sum(dtype=np.float64))


def test_unsupported_pairs_raise(cuda_ops, oracle):
    # TestOperations.java:140 relies on UnsupportedOperationException for unsupported dtype pairs
    from jlama_b200 import native
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(1)
    a, _ = _tensors(T, oracle, "I8A", _mk(rng, 1, SIZE))
    b = T.FloatBufferTensor(_mk(rng, 8, SIZE, 0, 1))
    c = T.FloatBufferTensor(np.zeros((1, 8), dtype=np.float32))
    with pytest.raises(native.UnsupportedOperation):
        cuda_ops.batch_dot_product(c, a, b, 0, 0, SIZE)


def test_dot_product_with_offsets(cuda_ops, oracle):
    # TestOperations "offsets 512/512/512" (:156)
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(2)
    for akind, bkind in [("F32", "F32"), ("F32", "Q4"), ("I8A", "Q4")]:
        a, ao = _tensors(T, oracle, akind, _mk(rng, 1, SIZE))
        b, bo = _tensors(T, oracle, bkind, _mk(rng, 1, SIZE, 0, 1))
        got = cuda_ops.dot_product(a, b, 512, 512, 512)
        ref = oracle.batch_dot(ao, bo, 512, 512, 512, 0, 0, 1)[0, 0]
        assert abs(got - ref) <= 2e-5 * abs(ref)


def test_result_offset_and_row_chunks(cuda_ops, oracle):
    # testBatchDotProductWithResultOffset (:534-552) + chunked rows as VectorMath.pchunk issues them
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(3)
    a, ao = _tensors(T, oracle, "I8A", _mk(rng, BATCH, SIZE))
    b, bo = _tensors(T, oracle, "Q4", _mk(rng, ROWS, SIZE, 0, 1))
    cuda_ops.register_model_tensor(b)
    c1 = T.FloatBufferTensor(np.zeros((BATCH, ROWS * 2), dtype=np.float32))
    cuda_ops.batch_dot_product(c1, a, b, 0, 0, SIZE, 0, 0, ROWS)
    cuda_ops.batch_dot_product(c1, a, b, 0, 0, SIZE, ROWS, 0, ROWS)
    ref = oracle.batch_dot(ao, bo, 0, 0, SIZE, 0, 0, ROWS)
    assert np.abs(c1.data[:, :ROWS] - ref).max() <= 2e-5 * np.abs(ref).max()
    assert np.array_equal(c1.data[:, :ROWS], c1.data[:, ROWS:])
    # dotProductChunk over row chunks [32,64) and [64,128) writes result columns j (absolute)
    c2 = T.FloatBufferTensor(np.zeros((BATCH, ROWS), dtype=np.float32))
    cuda_ops.dot_product_chunk(c2, a, b, 0, SIZE, 32, 32)
    cuda_ops.dot_product_chunk(c2, a, b, 0, SIZE, 64, 64)
    assert np.all(c2.data[:, :32] == 0)
    assert np.array_equal(c2.data[:, 32:], c1.data[:, 32:ROWS])


@pytest.mark.parametrize("bkind", ["Q4", "I8W", "F32"])
def test_sparse_operands_rebase_offsets(cuda_ops, oracle, bkind):
    """Sparse (column- / row-sliced) operands behind the same logical call (TensorShape.java:94-133,
    NativeSimdTensorOperations.java:96-107): the row split of jlama-net hands the operators exactly these shapes."""
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(77)
    M, K, N = 3, 256, 64
    a, ao = _tensors(T, oracle, "F32", _mk(rng, M, K))
    b, bo = _tensors(T, oracle, bkind, _mk(rng, N, K, 0.0, 1.0))
    ref = oracle.batch_dot(ao, bo, 64, 64, 128, 0, 0, N)  # dense: columns [64, 192) of both operands
    # column-sparse a and b, logical offsets unchanged
    c = T.FloatBufferTensor(np.zeros((M, N), dtype=np.float32))
    cuda_ops.batch_dot_product(c, a.sparsify(64, 128), b.sparsify(64, 128), 64, 64, 128)
    assert np.abs(c.data - ref).max() <= 2e-5 * np.abs(ref).max()
    # row-sparse b (rows [16, 48) stored), dense result: logical rows land at their logical columns
    full = oracle.batch_dot(ao, bo, 0, 0, K, 0, 0, N)
    c = T.FloatBufferTensor(np.zeros((M, N), dtype=np.float32))
    cuda_ops.batch_dot_product(c, a, b.sparsify_rows(16, 32), 0, 0, K, 0, 16, 32)
    assert np.abs(c.data[:, 16:48] - full[:, 16:48]).max() <= 2e-5 * np.abs(full).max()
    assert not c.data[:, :16].any() and not c.data[:, 48:].any()
    # ... and a column-sparse result that stores only those 32 columns
    cs = T.FloatBufferTensor(np.zeros((M, N), dtype=np.float32)).sparsify(16, 32)
    cuda_ops.batch_dot_product(cs, a, b.sparsify_rows(16, 32), 0, 0, K, 0, 16, 32)
    assert np.abs(cs.data - full[:, 16:48]).max() <= 2e-5 * np.abs(full).max()


def test_dot_product_batch_chunk(cuda_ops, oracle):
    # testBatchChunked (:492-513)
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(4)
    a, ao = _tensors(T, oracle, "I8A", _mk(rng, 1, SIZE))
    ws = [_tensors(T, oracle, "Q4", _mk(rng, ROWS, SIZE, 0, 1)) for _ in range(2)]
    for w, _ in ws:
        cuda_ops.register_model_tensor(w)
    rs = [T.FloatBufferTensor(np.zeros((1, ROWS), dtype=np.float32)) for _ in range(2)]
    cuda_ops.dot_product_batch_chunk(rs, a, [w for w, _ in ws], 0, SIZE, 0, ROWS)
    for r, (_, wo) in zip(rs, ws):
        ref = oracle.batch_dot(ao, wo, 0, 0, SIZE, 0, 0, ROWS)
        assert np.abs(r.data - ref).max() <= 2e-5 * np.abs(ref).max()


def test_attention_style_gemm_against_host_kv_page(cuda_ops, oracle):
    # CausalSelfAttention.java:324-330: A = query row, B = KV page slice, aColOff = h*hs, bColOff = kvh*hs, K = hs
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(5)
    hs, heads, kvh, rows = 64, 8, 2, 40
    q = T.FloatBufferTensor(rng.standard_normal((1, heads * hs)).astype(np.float32))
    page = T.FloatBufferTensor(rng.standard_normal((64, kvh * hs)).astype(np.float32))
    attn = T.FloatBufferTensor(np.zeros((1, 128), dtype=np.float32))
    cuda_ops.batch_dot_product(attn, q, page, 5 * hs, 1 * hs, hs, 64, 0, rows)
    ref = q.data[:, 5 * hs:6 * hs] @ page.data[:rows, hs:2 * hs].T
    assert np.allclose(attn.data[:, 64:64 + rows], ref, rtol=1e-5, atol=1e-5)
    assert np.all(attn.data[:, :64] == 0)


def test_accumulate_maccumulate_scale_saxpy(cuda_ops, oracle):
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(6)
    a0 = _mk(rng, BATCH, SIZE)
    b0 = _mk(rng, BATCH, SIZE)
    # accumulate F32 += F32 over a sub-range, b row-broadcast and per-row (NaiveTensorOperations.java:34-46)
    for brow in (1, BATCH):
        a = T.FloatBufferTensor(a0.copy())
        cuda_ops.accumulate(a, T.FloatBufferTensor(b0[:brow].copy()), 32, 512)
        ref = oracle.accumulate(a0.copy(), oracle.f32(b0[:brow]), 32, 512)
        assert np.array_equal(a.data, ref)
    # F32 += BF16, F32 += Q4 (PanamaTensorOperations.java:2297-2325)
    bb = T.BFloat16BufferTensor.from_float(b0[:1])
    a = T.FloatBufferTensor(a0.copy())
    cuda_ops.accumulate(a, bb, 0, SIZE)
    assert np.array_equal(a.data, oracle.accumulate(a0.copy(), oracle.OTensor(oracle.BF16, bb.data), 0, SIZE))
    bq = T.Q4ByteBufferTensor.from_float(b0[:1])
    a = T.FloatBufferTensor(a0.copy())
    cuda_ops.accumulate(a, bq, 0, SIZE)
    assert np.array_equal(a.data, oracle.accumulate(a0.copy(), oracle.OTensor(oracle.Q4, bq.data, bq.scales), 0, SIZE))
    # maccumulate
    a = T.FloatBufferTensor(a0.copy())
    cuda_ops.maccumulate(a, T.FloatBufferTensor(b0.copy()), 64, 256)
    assert np.array_equal(a.data, oracle.maccumulate(a0.copy(), b0, 64, 256))
    # scale
    a = T.FloatBufferTensor(a0.copy())
    cuda_ops.scale(3.14159, a, 256, 256)
    assert np.array_equal(a.data, oracle.scale(3.14159, a0.copy(), 256, 256))
    # saxpy (scalar) and batched saxpy = P.V (PanamaTensorOperations.java:2614-2698)
    x = T.FloatBufferTensor(b0[:1].copy())
    y = T.FloatBufferTensor(a0[:1].copy())
    cuda_ops.saxpy(0.75, x, y, 128, 256, 200)
    yr = a0[:1].copy()
    oracle.saxpy(0.75, b0[:1].copy(), yr, 128, 256, 200)
    assert np.allclose(y.data, yr, rtol=1e-6, atol=1e-6)
    alpha = T.FloatBufferTensor(rng.random((1, 64)).astype(np.float32))
    xs = T.FloatBufferTensor(b0[:, :256].copy())
    y = T.FloatBufferTensor(np.zeros((1, 512), dtype=np.float32))
    cuda_ops.saxpy(alpha, xs, y, 64, 128, 64, 3, 2, 30)
    yr = np.zeros((1, 512), dtype=np.float32)
    oracle.saxpy_batch(alpha.data, xs.data, yr, 64, 128, 64, 3, 2, 30)
    assert np.array_equal(y.data, yr)


def test_quantisers_are_bit_exact(cuda_ops, oracle):
    # testQuantize / BF16 (:438-489): integer outputs must equal the reference semantics exactly
    from jlama_b200 import tensor as T
    from jlama_b200.native import BF16, I8
    rng = np.random.default_rng(7)
    x = _mk(rng, BATCH, SIZE)
    x[3, 64:96] = 0.0
    x[4, :32] = -x[4, :32]
    t = T.FloatBufferTensor(x)
    q = cuda_ops.quantize(t, I8, 0, SIZE)
    oq, os_ = oracle.quantize_q8_act(x)
    assert np.array_equal(q.data, oq) and np.array_equal(q.scales.view(np.uint32), os_.view(np.uint32))
    # sub-range
    q = cuda_ops.quantize(t, I8, 256, 512)
    assert np.array_equal(q.data[:, 256:768], oq[:, 256:768]) and np.all(q.data[:, :256] == 0)
    b = cuda_ops.quantize(t, BF16, 0, SIZE)
    assert np.array_equal(b.data, oracle.f32_to_bf16(x))
    # GPU Q4 weight quantiser is byte-identical to Q4ByteBufferTensor
    w = (rng.standard_normal((ROWS, SIZE)) * 0.02).astype(np.float32)
    w[5, 32:64] = 0
    w[6, 100] = -w[6, 96:128].__abs__().max()
    gq = cuda_ops.quantize_q4_weights(w)
    oq4, os4 = oracle.quantize_q4(w)
    assert np.array_equal(gq.data, oq4) and np.array_equal(gq.scales.view(np.uint32), os4.view(np.uint32))


def test_fused_layer_ops(cuda_ops, oracle):
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(8)
    x = rng.standard_normal((5, SIZE)).astype(np.float32) * 3
    w = (1 + 0.1 * rng.standard_normal((1, SIZE))).astype(np.float32)
    out = cuda_ops.rmsnorm(T.FloatBufferTensor(x), T.FloatBufferTensor(w), 1e-5)
    ref = oracle.rmsnorm(x, oracle.f32(w), 1e-5)
    assert np.abs(out.data - ref).max() <= 1e-6 * np.abs(ref).max()
    wb = T.BFloat16BufferTensor.from_float(w)
    out = cuda_ops.rmsnorm(T.FloatBufferTensor(x), wb, 1e-5)
    ref = oracle.rmsnorm(x, oracle.OTensor(oracle.BF16, wb.data), 1e-5)
    assert np.abs(out.data - ref).max() <= 1e-6 * np.abs(ref).max()
    s = T.FloatBufferTensor(rng.standard_normal((1, 3000)).astype(np.float32) * 4)
    sr = oracle.softmax(s.data.copy(), 0, 3000)
    cuda_ops.softmax(s, 0, 3000)
    assert np.abs(s.data - sr).max() <= 1e-5 * sr.max()
    g = rng.standard_normal((4, SIZE)).astype(np.float32) * 4
    u = rng.standard_normal((4, SIZE)).astype(np.float32)
    gt = T.FloatBufferTensor(g.copy())
    cuda_ops.silu_mul(gt, T.FloatBufferTensor(u), 0, SIZE)
    ref = oracle.silu(g) * u
    assert np.abs(gt.data - ref).max() <= 1e-6 * np.abs(ref).max()


@pytest.mark.parametrize("K,N", [(4096, 4096), (14336, 1024), (2048, 512), (512, 4096)])
def test_gemv_model_shapes_full_size(cuda_ops, oracle, K, N):
    """Llama-3-8B / 1B layer shapes (incl. a TP-8 o_proj shard K=512) at M=1 and M=4 with the model's own
    Q8 x Q4 arithmetic; oracle through the reference's C kernels when available."""
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(K + N)
    w = T.Q4ByteBufferTensor(rng.integers(0, 256, (N, K // 2), dtype=np.uint8),
                             ((0.5 + rng.random((N, K // 32))) * 0.004).astype(np.float32))
    cuda_ops.register_model_tensor(w)
    x = rng.standard_normal((4, K)).astype(np.float32)
    q, s = T.quantize_q8_activations(x)
    a = T.Q8ByteBufferTensor(q, s)
    c = T.FloatBufferTensor(np.zeros((4, N), dtype=np.float32))
    cuda_ops.batch_dot_product(c, a, w, 0, 0, K)
    ref = oracle.batch_dot(oracle.OTensor(oracle.I8, q, s), oracle.OTensor(oracle.Q4, w.data, w.scales), 0, 0, K, 0, 0, N)
    assert np.abs(c.data - ref).max() <= 2e-5 * np.abs(ref).max()
    cuda_ops.unregister_model_tensor(w)


@pytest.mark.parametrize("M,N,K", [(16, 128, 64), (128, 256, 512), (100, 384, 4096), (256, 1024, 2048)])
def test_tensor_core_gemm_matches_bf16_reference(cuda_ops, oracle, M, N, K):
    """tcgen05 prefill GEMM (BF16 operands, F32 accumulate, dequant fused into the smem fill) against a float64
    product of the SAME BF16-rounded operands (tight), and against the reference's F32 x Q4 arithmetic (1e-2 class)."""
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(M * 7 + N)
    w = T.Q4ByteBufferTensor(rng.integers(0, 256, (N, K // 2), dtype=np.uint8),
                             ((0.5 + rng.random((N, K // 32))) * 0.01 * np.where(rng.random((N, K // 32)) < 0.5, -1, 1)).astype(np.float32))
    cuda_ops.register_model_tensor(w)
    a = rng.standard_normal((M, K)).astype(np.float32)
    c = T.FloatBufferTensor(np.zeros((M, N), dtype=np.float32))
    cuda_ops.batch_dot_product_tensor_core(c, T.FloatBufferTensor(a), w, 0, 0, K)
    a16 = T.bfloat16_to_float32(T.float32_to_bfloat16(a)).astype(np.float64)
    w16 = T.bfloat16_to_float32(T.float32_to_bfloat16(w.to_float())).astype(np.float64)
    ref = a16 @ w16.T
    assert np.abs(c.data - ref).max() <= 2e-5 * np.abs(ref).max() * np.sqrt(K / 64)
    exact = oracle.batch_dot(oracle.f32(a), oracle.OTensor(oracle.Q4, w.data, w.scales), 0, 0, K, 0, 0, N)
    assert np.abs(c.data - exact).max() <= 1e-2 * np.abs(exact).max()
    # row-chunk / result-offset semantics are those of batchDotProduct
    if N >= 256:
        c2 = T.FloatBufferTensor(np.zeros((M, N), dtype=np.float32))
        cuda_ops.batch_dot_product_tensor_core(c2, T.FloatBufferTensor(a), w, 0, 0, K, 0, 128, 128)
        assert np.all(c2.data[:, :128] == 0) and np.array_equal(c2.data[:, 128:256], c.data[:, 128:256])
    cuda_ops.unregister_model_tensor(w)


@pytest.mark.parametrize("M,N,K", [(48, 128, 128), (200, 256, 4096)])
def test_tensor_core_gemm_q8_weights(cuda_ops, oracle, M, N, K):
    """the same tcgen05 GEMM with Q8_0 weights (int8 + f32 scale per 32; BASELINE config 3's format): tight against a float64
    product of the BF16-rounded operands, 1e-2 class against the exact F32 x I8 arithmetic."""
    from jlama_b200 import tensor as T
    rng = np.random.default_rng(M + N + K)
    w = T.Q8ByteBufferTensor(rng.integers(-127, 128, (N, K), dtype=np.int8), ((0.5 + rng.random((N, K // 32))) * 0.001).astype(np.float32))
    cuda_ops.register_model_tensor(w)
    a = rng.standard_normal((M, K)).astype(np.float32)
    c = T.FloatBufferTensor(np.zeros((M, N), dtype=np.float32))
    cuda_ops.batch_dot_product_tensor_core(c, T.FloatBufferTensor(a), w, 0, 0, K)
    a16 = T.bfloat16_to_float32(T.float32_to_bfloat16(a)).astype(np.float64)
    w16 = T.bfloat16_to_float32(T.float32_to_bfloat16(w.to_float())).astype(np.float64)
    ref = a16 @ w16.T
    assert np.abs(c.data - ref).max() <= 2e-5 * np.abs(ref).max() * np.sqrt(K / 64)
    exact = a.astype(np.float64) @ w.to_float().astype(np.float64).T
    assert np.abs(c.data - exact).max() <= 1e-2 * np.abs(exact).max()
    cuda_ops.unregister_model_tensor(w)


def test_layernorm_matches_reference_restatement(cuda_ops, oracle):
    """LayerNorm.forward (model/LayerNorm.java:41-67, GPT-2 family).  The reference sums sequentially in float; any other
    summation order moves the statistics by ~1e-6 relative (bar: 2e-5 of the row maximum)."""
    from jlama_b200.tensor import BFloat16BufferTensor, FloatBufferTensor, float32_to_bfloat16
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((5, 768)) * 3 + 0.7).astype(np.float32)
    w = (1 + 0.1 * rng.standard_normal((1, 768))).astype(np.float32)
    b = (0.1 * rng.standard_normal((1, 768))).astype(np.float32)
    out = cuda_ops.layernorm(FloatBufferTensor(x), FloatBufferTensor(w), FloatBufferTensor(b), 1e-5)
    ref = oracle.layernorm(x, oracle.f32(w), oracle.f32(b), 1e-5)
    assert np.abs(out.data - ref).max() <= 2e-5 * np.abs(ref).max()
    wb = float32_to_bfloat16(w)
    out2 = cuda_ops.layernorm(FloatBufferTensor(x), BFloat16BufferTensor(wb), FloatBufferTensor(b), 1e-5)
    ref2 = oracle.layernorm(x, oracle.OTensor(oracle.BF16, wb), oracle.f32(b), 1e-5)
    assert np.abs(out2.data - ref2).max() <= 2e-5 * np.abs(ref2).max()


def test_activation_functions_match_reference(cuda_ops, oracle):
    """ActivationFunction.eval (math/ActivationFunction.java:29-37): SILU and the tanh-form GELU in double, cast to float."""
    from jlama_b200.tensor import FloatBufferTensor
    rng = np.random.default_rng(6)
    x = (rng.standard_normal((3, 320)) * 4).astype(np.float32)
    for kind, ref in (("gelu", oracle.gelu), ("silu", oracle.silu)):
        t = FloatBufferTensor(x.copy())
        cuda_ops.activation(kind, t, 32, 256)
        want = x.copy()
        want[:, 32:288] = ref(x[:, 32:288])
        assert np.array_equal(t.data[:, :32], x[:, :32]) and np.array_equal(t.data[:, 288:], x[:, 288:])
        assert np.abs(t.data - want).max() <= 2e-7 * max(1.0, np.abs(want).max())
