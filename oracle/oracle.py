"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module.  The product package
``jlama_b200`` never does.

The C side (oracle/jlama_oracle.c) restates the reference arithmetic and cites
the reference file:line per function; oracle/_ref/libjlama*.so are the
reference's own C kernels compiled by oracle/Makefile.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
F32, BF16, Q4, I8 = 0, 1, 2, 3

# tensor kinds (jlama_oracle.c enums)
T_EMBED, T_OUT_NORM, T_LM_HEAD = 0, 1, 2
L_ATTN_NORM, L_Q, L_K, L_V, L_O, L_FFN_NORM, L_GATE, L_DOWN, L_UP = range(9)

HAS_F16C, HAS_AVX2 = 2, 4  # jlama-native/src/main/c/simd/vector_simd.h:13-14


class JoTensor(C.Structure):
    _fields_ = [("dtype", C.c_int), ("rows", C.c_int64), ("cols", C.c_int64),
                ("data", C.c_void_p), ("scales", C.c_void_p)]


def build(force=False):
    """Compile the restatement (and oracle/_ref when /root/reference is present)."""
    so = os.path.join(HERE, "libjlama_oracle.so")
    src = os.path.join(HERE, "jlama_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "libjlama_oracle.so"], stdout=subprocess.DEVNULL)
    ref = os.path.join(HERE, "_ref", "libjlama.so")
    if os.path.exists("/root/reference/jlama-native/src/main/c/simd/vector_simd.c") and (force or not os.path.exists(ref)):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


_lib = None


def available_cpus():
    """CPUs this process may really use: affinity mask capped by the cgroup CPU quota (what Java's
    Runtime.availableProcessors(), the input of PhysicalCoreExecutor.java:27, reports)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.999)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, (q + per - 1) // per))
            break
        except (OSError, ValueError, IndexError):
            continue
    return max(1, n)


def lib():
    global _lib
    if _lib is None:
        # worker threads must sleep, not spin, between the many small parallel regions of a decode step
        os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
        os.environ.setdefault("OMP_PROC_BIND", "false")
        if "OMP_NUM_THREADS" not in os.environ:
            os.environ["OMP_NUM_THREADS"] = str(available_cpus())
        so = build()
        L = C.CDLL(so)
        L.jo_model_create.restype = C.c_void_p
        L.jo_model_create.argtypes = [C.c_int] * 7 + [C.c_float, C.c_double, C.c_double, C.c_int]
        L.jo_model_set_tensor.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.jo_model_free.argtypes = [C.c_void_p]
        L.jo_model_reset_kv.argtypes = [C.c_void_p]
        L.jo_model_set_tp.argtypes = [C.c_void_p, C.c_int]
        L.jo_model_set_moe.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.jo_model_set_expert.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.jo_model_kv_geometry.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.jo_model_kv_row.restype = C.POINTER(C.c_float)
        L.jo_model_kv_row.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.jo_model_batch_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.jo_model_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.jo_sample_temperature.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float]
        L.jo_model_generate.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.jo_silu.restype = C.c_float
        L.jo_silu.argtypes = [C.c_float]
        L.jo_gelu.restype = C.c_float
        L.jo_gelu.argtypes = [C.c_float]
        L.jo_load_reference_kernels.argtypes = [C.c_char_p, C.c_int]
        L.jo_scale.argtypes = [C.c_float, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int]
        L.jo_saxpy.argtypes = [C.c_float, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.jo_precompute_freqs_cis.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p]
        L.jo_kv_page_solver.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.jo_rmsnorm.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def cpu_isa_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def load_reference_kernels(build=None):
    """dlopen the reference's own C kernels (oracle/_ref).  Returns a label or None.  build: None = the widest this CPU runs,
    "avx512" / "avx2" = that build only (the reference ships both code paths; their summation orders differ)."""
    fl = cpu_isa_flags()
    need512 = {"avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512_vnni"}
    cands = []
    if need512 <= fl and build in (None, "avx512"):
        # HAS_AVX2 selects the *_512 bodies (vector_simd.c:465-468)
        cands.append(("libjlama.so", HAS_F16C | HAS_AVX2, "jlama-native C kernels (AVX-512 VNNI build)"))
    if "avx2" in fl and "fma" in fl and build in (None, "avx2"):
        cands.append(("libjlama_avx2.so", HAS_F16C, "jlama-native C kernels (AVX2 build)"))
    for name, flags, label in cands:
        path = os.path.join(HERE, "_ref", name)
        if os.path.exists(path) and lib().jo_load_reference_kernels(path.encode(), flags) == 0:
            return label
    return None


def use_reference_kernels(on=True):
    lib().jo_use_reference_kernels(1 if on else 0)


# ---------------------------------------------------------------------------
# quantisers / conversions
# ---------------------------------------------------------------------------
def quantize_q4(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, cols = x.shape
    q = np.empty((rows, cols // 2), dtype=np.uint8)
    s = np.empty((rows, cols // 32), dtype=np.float32)
    lib().jo_quantize_q4(_p(x), C.c_int64(rows), C.c_int64(cols), _p(q), _p(s))
    return q, s


def dequantize_q4(q, s):
    rows, cols = q.shape[0], q.shape[1] * 2
    out = np.empty((rows, cols), dtype=np.float32)
    lib().jo_dequantize_q4(_p(q), _p(s), C.c_int64(rows), C.c_int64(cols), _p(out))
    return out


def quantize_q8_weights(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, cols = x.shape
    q = np.empty((rows, cols), dtype=np.int8)
    s = np.empty((rows, cols // 32), dtype=np.float32)
    lib().jo_quantize_q8_weights(_p(x), C.c_int64(rows), C.c_int64(cols), _p(q), _p(s))
    return q, s


def quantize_q8_act(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    rows, cols = x.shape
    q = np.empty((rows, cols), dtype=np.int8)
    s = np.empty((rows, cols // 32), dtype=np.float32)
    lib().jo_quantize_q8_act(_p(x), C.c_int64(rows), C.c_int64(cols), C.c_int64(cols), _p(q), _p(s))
    return q, s


def dequantize_q8(q, s):
    rows, cols = q.shape
    out = np.empty((rows, cols), dtype=np.float32)
    lib().jo_dequantize_q8(_p(q), _p(s), C.c_int64(rows), C.c_int64(cols), _p(out))
    return out


def f32_to_bf16(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    lib().jo_f32_to_bf16(_p(x), _p(out), C.c_int64(x.size))
    return out


def bf16_to_f32(x):
    x = np.ascontiguousarray(x, dtype=np.uint16)
    out = np.empty(x.shape, dtype=np.float32)
    lib().jo_bf16_to_f32(_p(x), _p(out), C.c_int64(x.size))
    return out


# ---------------------------------------------------------------------------
# tensors + ops
# ---------------------------------------------------------------------------
class OTensor:
    """Host tensor handed to the oracle: (dtype, data, scales)."""

    def __init__(self, dtype, data, scales=None):
        self.dtype = dtype
        self.data = np.ascontiguousarray(data)
        self.scales = None if scales is None else np.ascontiguousarray(scales, dtype=np.float32)
        self.rows = self.data.shape[0]
        self.cols = self.data.shape[1] * (2 if dtype == Q4 else 1)

    def struct(self):
        return JoTensor(self.dtype, self.rows, self.cols, self.data.ctypes.data,
                        self.scales.ctypes.data if self.scales is not None else None)

    def to_f32(self):
        if self.dtype == F32:
            return self.data.astype(np.float32)
        if self.dtype == BF16:
            return bf16_to_f32(self.data)
        if self.dtype == Q4:
            return dequantize_q4(self.data, self.scales)
        return dequantize_q8(self.data, self.scales)


def f32(x):
    return OTensor(F32, np.ascontiguousarray(x, dtype=np.float32))


def batch_dot(a, b, a_col_off, b_col_off, k, r_row_off, b_row_off, n, result=None, naive=False):
    """TensorOperations.batchDotProduct (TensorOperations.java:62-72)."""
    if result is None:
        result = np.zeros((a.rows, r_row_off + b_row_off + n), dtype=np.float32)
    sa, sb = a.struct(), b.struct()
    fn = lib().jo_batch_dot_naive if naive else lib().jo_batch_dot
    fn(_p(result), C.c_int64(result.shape[1]), C.byref(sa), C.byref(sb), a_col_off, b_col_off, k,
       r_row_off, b_row_off, n)
    return result


def ref_gemm_q8_q4(aq, as_, bq, bs, n0, n):
    m, k = aq.shape
    r = np.zeros((m, bq.shape[0]), dtype=np.float32)
    rc = lib().jo_ref_gemm_q8_q4(_p(as_), _p(aq), _p(bs), _p(bq), _p(r), m, n0, n, k, k, k, r.shape[1])
    assert rc == 0, "reference kernels not loaded"
    return r


def ref_gemm_f32_q4(a, bq, bs, n0, n):
    m, k = a.shape
    r = np.zeros((m, bq.shape[0]), dtype=np.float32)
    rc = lib().jo_ref_gemm_f32_q4(_p(a), _p(bs), _p(bq), _p(r), m, n0, n, k, k, k, r.shape[1])
    assert rc == 0
    return r


def ref_gemm_f32(a, b, n0, n):
    m, k = a.shape
    r = np.zeros((m, b.shape[0]), dtype=np.float32)
    rc = lib().jo_ref_gemm_f32(_p(a), _p(b), _p(r), m, n0, n, k, k, k, r.shape[1])
    assert rc == 0
    return r


def ref_gemm_f32_bf16(a, b_bits, n0, n):
    """reference gemm_f32_bf16 (vector_simd.h:38): F32 activations x BF16 weights (uint16 bit patterns) -> F32"""
    m, k = a.shape
    r = np.zeros((m, b_bits.shape[0]), dtype=np.float32)
    L = lib()
    L.jo_ref_gemm_f32_bf16.restype = C.c_int
    rc = L.jo_ref_gemm_f32_bf16(_p(a), _p(b_bits), _p(r), m, n0, n, k, k, k, r.shape[1])
    assert rc == 0
    return r


def accumulate(a, b, off, length):
    sb = b.struct()
    lib().jo_accumulate(_p(a), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]), C.byref(sb), off, length)
    return a


def maccumulate(a, b, off, length):
    lib().jo_maccumulate(_p(a), C.c_int64(a.shape[0]), C.c_int64(a.shape[1]), _p(b), C.c_int64(b.shape[0]),
                         C.c_int64(b.shape[1]), off, length)
    return a


def scale(f, x, off, length):
    lib().jo_scale(C.c_float(f), _p(x), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), off, length)
    return x


def saxpy(alpha, x, y, xoff, yoff, limit):
    lib().jo_saxpy(C.c_float(alpha), _p(x), _p(y), xoff, yoff, limit)
    return y


def saxpy_batch(alpha, x, y, xoff, yoff, limit, a_off, x_row_off, batch):
    lib().jo_saxpy_batch(_p(alpha), _p(x), C.c_int64(x.shape[1]), _p(y), xoff, yoff, limit, a_off, x_row_off, batch)
    return y


def softmax(x, offset, length):
    lib().jo_softmax(_p(x), offset, length)
    return x


def silu(x):
    return np.array([lib().jo_silu(float(v)) for v in np.asarray(x, dtype=np.float32).ravel()],
                    dtype=np.float32).reshape(np.shape(x))


def rmsnorm(x, w, eps, E=None, adj=0.0):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    sw = w.struct()
    E = E or x.shape[1]
    lib().jo_rmsnorm(_p(x), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), C.byref(sw), C.c_float(adj), C.c_float(eps),
                     E, 0, x.shape[1], _p(out))
    return out


def layernorm(x, w, bias, eps, E=None):
    """model/LayerNorm.java:41-67 (sequential float sums, as the reference)"""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    sw, sb = w.struct(), bias.struct()
    E = E or x.shape[1]
    L = lib()
    L.jo_layernorm.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.jo_layernorm(_p(x), C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), C.byref(sw), C.byref(sb), C.c_float(eps), E, 0, x.shape[1], _p(out))
    return out


def gelu(x):
    return np.array([lib().jo_gelu(float(v)) for v in np.asarray(x, dtype=np.float32).ravel()], dtype=np.float32).reshape(np.shape(x))


def precompute_freqs_cis(dim, end, theta, scaling=1.0):
    out = np.empty((end * (dim // 2), 2), dtype=np.float32)
    lib().jo_precompute_freqs_cis(dim, end, C.c_double(theta), C.c_double(scaling), _p(out))
    return out


def kv_page_solver(layers, ctx, kv_seg_len, dtype_size=4, max_page_bytes=1 << 23):
    a, b = C.c_int(), C.c_int()
    lib().jo_kv_page_solver(layers, ctx, kv_seg_len, dtype_size, C.c_int64(max_page_bytes), C.byref(a), C.byref(b))
    return a.value, b.value


class JoDctx(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "embeddingSegmentStart", "embeddingSegmentLength", "attentionSegmentStart", "attentionSegmentLength",
        "hiddenSegmentStart", "hiddenSegmentLength", "kvSegmentStart", "kvSegmentLength", "headStart", "headEnd",
        "groupHeadStart", "groupHeadEnd", "numberOfLayers", "layerStart", "layerEnd")]


def dctx(E, attn_len, H, head_size, group, layers, model_shard, n_model_shards, layer_shard=0, n_layer_shards=1):
    d = JoDctx()
    lib().jo_dctx_build(E, attn_len, H, head_size, group, layers, model_shard, n_model_shards, layer_shard,
                        n_layer_shards, C.byref(d))
    return d


# ---------------------------------------------------------------------------
# model
# ---------------------------------------------------------------------------
class OracleLlama:
    """CPU restatement of LlamaModel + AbstractModel.generate() at temperature 0.

    `weights` is the dict produced by jlama_b200.synth (name -> (dtype_code, data, scales)).
    """

    def __init__(self, cfg, weights, act_q8=True, tp=1):
        L = lib()
        self.cfg = cfg
        self._keep = weights
        self.h = L.jo_model_create(cfg["ctx"], cfg["E"], cfg["H"], cfg["heads"], cfg["kv_heads"], cfg["layers"],
                                   cfg["vocab"], C.c_float(cfg["eps"]), C.c_double(cfg["rope_theta"]),
                                   C.c_double(cfg.get("rope_scale", 1.0)), 1 if act_q8 else 0)
        L.jo_model_set_tp(self.h, tp)

        def put(layer, kind, name):
            if name not in weights:
                return
            dt, data, scales = weights[name]
            rows = data.shape[0]
            cols = data.shape[1] * (2 if dt == Q4 else 1)
            L.jo_model_set_tensor(self.h, layer, kind, dt, C.c_int64(rows), C.c_int64(cols), _p(data), _p(scales))

        n_exp = cfg.get("experts", 0)
        if n_exp:
            L.jo_model_set_moe(self.h, n_exp, cfg["experts_per_token"])

        def put_expert(layer, expert, which, name):
            dt, data, scales = weights[name]
            rows = data.shape[0]
            cols = data.shape[1] * (2 if dt == Q4 else 1)
            L.jo_model_set_expert(self.h, layer, expert, which, dt, C.c_int64(rows), C.c_int64(cols), _p(data), _p(scales))

        put(-1, T_EMBED, "model.embed_tokens.weight")
        put(-1, T_OUT_NORM, "model.norm.weight")
        put(-1, T_LM_HEAD, "lm_head.weight")
        for i in range(cfg["layers"]):
            b = "model.layers.%d." % i
            if n_exp:  # MixtralModel.java:88-105 tensor names
                put_expert(i, -1, 0, b + "block_sparse_moe.gate.weight")
                for e in range(n_exp):
                    for which, nm in ((0, "w1"), (1, "w2"), (2, "w3")):
                        put_expert(i, e, which, b + "block_sparse_moe.experts.%d.%s.weight" % (e, nm))
            put(i, L_ATTN_NORM, b + "input_layernorm.weight")
            put(i, L_Q, b + "self_attn.q_proj.weight")
            put(i, L_K, b + "self_attn.k_proj.weight")
            put(i, L_V, b + "self_attn.v_proj.weight")
            put(i, L_O, b + "self_attn.o_proj.weight")
            put(i, L_FFN_NORM, b + "post_attention_layernorm.weight")
            put(i, L_GATE, b + "mlp.gate_proj.weight")
            put(i, L_DOWN, b + "mlp.down_proj.weight")
            put(i, L_UP, b + "mlp.up_proj.weight")

    def reset(self):
        lib().jo_model_reset_kv(self.h)

    def kv_geometry(self):
        a, b = C.c_int(), C.c_int()
        lib().jo_model_kv_geometry(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def kv_row(self, layer, pos, which):
        p = lib().jo_model_kv_row(self.h, layer, pos, which)
        n = self.cfg["kv_heads"] * (self.cfg["E"] // self.cfg["heads"])
        return np.ctypeslib.as_array(p, shape=(n,)).copy()

    def batch_forward(self, tokens, start_pos, max_batch=256):
        tokens = np.ascontiguousarray(tokens, dtype=np.int32)
        hidden = np.empty(self.cfg["E"], dtype=np.float32)
        lib().jo_model_batch_forward(self.h, _p(tokens), len(tokens), start_pos, max_batch, _p(hidden))
        return hidden

    def sample(self, hidden):
        logits = np.empty(self.cfg["vocab"], dtype=np.float32)
        tok = lib().jo_model_sample(self.h, _p(hidden), _p(logits))
        return tok, logits

    def generate(self, prompt, n_new, max_batch=256, want_logits=True):
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        out = np.empty(n_new, dtype=np.int32)
        logits = np.empty((n_new, self.cfg["vocab"]), dtype=np.float32) if want_logits else None
        self.reset()
        n = lib().jo_model_generate(self.h, _p(prompt), len(prompt), n_new, max_batch, _p(out), _p(logits))
        return out[:n], logits

    def close(self):
        if self.h:
            lib().jo_model_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sample_temperature(logits, temperature, uniform):
    """AbstractModel.sample (:475-489) for temperature != 0 on a logits row; returns the drawn token."""
    work = np.ascontiguousarray(logits, dtype=np.float32).copy()
    return int(lib().jo_sample_temperature(_p(work), len(work), C.c_float(temperature), C.c_float(uniform)))


def num_threads():
    return lib().jo_num_threads()


def set_num_threads(n):
    lib().jo_set_num_threads(int(n))


class OracleGPT2:
    """core/model/gpt2/GPT2Model.java:54-129 restated for F32 checkpoints (test infrastructure): embedding wte + wpe (:54-69); per layer
    LayerNorm (model/LayerNorm.java:41-67, the C restatement) -> q/k/v = ln . W^T + bias (CausalSelfAttention.java:161-192) ->
    attention without rotary embedding (:199-356: scores * 1/sqrt(headSize), VectorMath.softMax, P.V) -> (att . Wo^T + bias) + x
    (:363-380, TransformerBlock.java:185) -> LayerNorm -> gelu(ln . Wfc^T + bias) (MLPBlock.java:117-141, ActivationFunction.java:29-37)
    -> (h . Wproj^T + bias) + xb (:144-160, TransformerBlock.java:203); logits = ln_f(x) . wte^T (:112-128).  Dot products are numpy
    float32 (another summation order than any reference kernel: 1e-6 class, the F32 tolerance is 1e-3)."""

    def __init__(self, cfg, weights):
        self.cfg = cfg
        g = lambda n: np.asarray(weights[n][1], dtype=np.float32)  # noqa: E731
        self.wte, self.wpe = g("wte.weight"), g("wpe.weight")
        self.lnf = (OTensor(F32, g("ln_f.weight").reshape(1, -1)), OTensor(F32, g("ln_f.bias").reshape(1, -1)))
        self.layers = []
        for i in range(cfg["layers"]):
            b = "h.%d." % i
            wq, wk, wv = np.split(np.ascontiguousarray(g(b + "attn.c_attn.weight").T), 3, axis=0)
            bq, bk, bv = np.split(g(b + "attn.c_attn.bias").reshape(-1), 3)
            self.layers.append(dict(
                ln1=(OTensor(F32, g(b + "ln_1.weight").reshape(1, -1)), OTensor(F32, g(b + "ln_1.bias").reshape(1, -1))),
                ln2=(OTensor(F32, g(b + "ln_2.weight").reshape(1, -1)), OTensor(F32, g(b + "ln_2.bias").reshape(1, -1))),
                wq=wq, wk=wk, wv=wv, bq=bq, bk=bk, bv=bv,
                wo=np.ascontiguousarray(g(b + "attn.c_proj.weight").T), bo=g(b + "attn.c_proj.bias").reshape(-1),
                wfc=np.ascontiguousarray(g(b + "mlp.c_fc.weight").T), bfc=g(b + "mlp.c_fc.bias").reshape(-1),
                wpr=np.ascontiguousarray(g(b + "mlp.c_proj.weight").T), bpr=g(b + "mlp.c_proj.bias").reshape(-1)))
        self.reset()

    def reset(self):
        self.k = [[] for _ in self.layers]
        self.v = [[] for _ in self.layers]

    @staticmethod
    def _gelu(x):
        xd = x.astype(np.float64)
        return (0.5 * xd * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (xd + 0.044715 * xd ** 3)))).astype(np.float32)

    def forward(self, token, pos):
        c = self.cfg
        E, nh = c["E"], c["heads"]
        hs = E // nh
        scale = np.float32(1.0 / np.sqrt(float(hs)))
        x = (self.wte[token] + self.wpe[pos]).astype(np.float32)
        for li, L in enumerate(self.layers):
            ln = layernorm(x[None, :], L["ln1"][0], L["ln1"][1], c["eps"])[0]
            q = (L["wq"] @ ln + L["bq"]).astype(np.float32)
            self.k[li].append((L["wk"] @ ln + L["bk"]).astype(np.float32))
            self.v[li].append((L["wv"] @ ln + L["bv"]).astype(np.float32))
            K, V = np.stack(self.k[li]), np.stack(self.v[li])
            att = np.empty(E, dtype=np.float32)
            for h in range(nh):
                sl = slice(h * hs, (h + 1) * hs)
                s = ((K[:, sl] @ q[sl]).astype(np.float32) * scale).astype(np.float32)
                s = np.ascontiguousarray(s)
                softmax(s, 0, len(s))
                att[sl] = (s @ V[:, sl]).astype(np.float32)
            xb = ((L["wo"] @ att + L["bo"]).astype(np.float32) + x).astype(np.float32)
            ln2 = layernorm(xb[None, :], L["ln2"][0], L["ln2"][1], c["eps"])[0]
            hdn = self._gelu((L["wfc"] @ ln2 + L["bfc"]).astype(np.float32))
            x = ((L["wpr"] @ hdn + L["bpr"]).astype(np.float32) + xb).astype(np.float32)
        return x

    def logits(self, x):
        ln = layernorm(x[None, :], self.lnf[0], self.lnf[1], self.cfg["eps"])[0]
        return (self.wte @ ln).astype(np.float32)

    def generate(self, prompt, n_new):
        """greedy; returns (tokens, logits per step)"""
        self.reset()
        x = None
        for p, tok in enumerate(prompt):
            x = self.forward(int(tok), p)
        toks, lgs = [], []
        pos = len(prompt)
        for _ in range(n_new):
            lg = self.logits(x)
            tok = int(np.argmax(lg))  # first maximum, like AbstractModel.sample's strict '>' scan
            toks.append(tok)
            lgs.append(lg)
            x = self.forward(tok, pos)
            pos += 1
        return toks, lgs

    def close(self):
        pass
