"""Model configs (public HF config.json values) and seeded synthetic checkpoints.

There are no checkpoints or network here, so benchmarks and parity tests run on synthetic weights of
the real architecture:

  mode "quantize": W ~ N(0, 0.02^2) f32 (o_proj/down_proj: 0.02/sqrt(2L); embedding N(0,1)), norm weights
                   1 + N(0, 0.1^2), then quantised with the reference quantiser semantics
                   (tensor.quantize_q4 / quantize_q8_weights).
  mode "direct"  : the packed Q4 bytes and the f32 block scales are drawn directly (uniform nibbles,
                   scales of magnitude ~0.02/4.6 with random sign) -- every byte pattern is a valid JQ4
                   tensor, and an 8B model is generated in seconds instead of minutes.

Weights are returned as  name -> (dtype_code, data, scales)  with Jlama's tensor names
(LlamaModel.java:72,115-141,152-156).
"""
import numpy as np

from .native import BF16, F32, I8, Q4
from .tensor import quantize_q4, quantize_q8_weights

CONFIGS = {
    # tiny configs for parity tests (the oracle finishes in seconds)
    "tiny": dict(ctx=256, E=256, H=512, heads=8, kv_heads=4, layers=2, vocab=512, eps=1e-5, rope_theta=10000.0),
    "tiny-mha": dict(ctx=128, E=128, H=256, heads=4, kv_heads=4, layers=2, vocab=320, eps=1e-5, rope_theta=10000.0),
    "small": dict(ctx=512, E=512, H=1536, heads=8, kv_heads=2, layers=4, vocab=2048, eps=1e-5, rope_theta=500000.0),
    "small-hs128": dict(ctx=512, E=1024, H=2048, heads=8, kv_heads=2, layers=2, vocab=1024, eps=1e-5, rope_theta=500000.0),
    # 8 KV heads so that the jlama-net model-shard split reaches 8 ranks (JlamaService.java:65-68 caps shards at KV heads)
    "llama-tp8-test": dict(ctx=512, E=1024, H=3584, heads=16, kv_heads=8, layers=3, vocab=4096, eps=1e-5, rope_theta=500000.0),
    # mixture of experts (Mixtral layout: 8 experts, top-2) at test size
    "tiny-mixtral": dict(ctx=256, E=256, H=512, heads=8, kv_heads=4, layers=2, vocab=512, eps=1e-5, rope_theta=10000.0, experts=8, experts_per_token=2),
    "mixtral-tp8-test": dict(ctx=256, E=1024, H=1024, heads=16, kv_heads=8, layers=2, vocab=1024, eps=1e-5, rope_theta=1000000.0, experts=8, experts_per_token=2),
    "small-mixtral": dict(ctx=512, E=1024, H=2048, heads=8, kv_heads=2, layers=3, vocab=1024, eps=1e-5, rope_theta=1000000.0, experts=4, experts_per_token=2),
    # BASELINE.json configs (public HF config.json dims)
    "llama-3.2-1b": dict(ctx=131072, E=2048, H=8192, heads=32, kv_heads=8, layers=16, vocab=128256, eps=1e-5,
                         rope_theta=500000.0, tied=True),
    "llama-3-8b": dict(ctx=8192, E=4096, H=14336, heads=32, kv_heads=8, layers=32, vocab=128256, eps=1e-5,
                       rope_theta=500000.0),
    "mixtral-8x7b": dict(ctx=32768, E=4096, H=14336, heads=32, kv_heads=8, layers=32, vocab=32000, eps=1e-5, rope_theta=1000000.0,
                         experts=8, experts_per_token=2),
}

SEED0 = 0x4A4C414D41  # "JLAMA"


def get_config(name, **over):
    cfg = dict(CONFIGS[name])
    cfg.setdefault("tied", False)
    cfg["name"] = name
    cfg.update(over)
    return cfg


def _direct_q4(rng, rows, cols, std):
    q = rng.integers(0, 256, size=(rows, cols // 2), dtype=np.uint8)
    # uniform nibbles in [-8, 7] have std ~4.61 -> |scale| ~ std/4.61, jittered, random sign
    mag = (std / 4.61) * (0.5 + rng.random((rows, cols // 32), dtype=np.float32))
    sign = np.where(rng.random((rows, cols // 32), dtype=np.float32) < 0.5, np.float32(-1), np.float32(1))
    return q, (mag * sign).astype(np.float32)


def _direct_q8(rng, rows, cols, std):
    q = rng.integers(-127, 128, size=(rows, cols), dtype=np.int8)
    mag = (std / 73.3) * (0.5 + rng.random((rows, cols // 32), dtype=np.float32))
    return q, mag.astype(np.float32)


def make_tensor(seed, rows, cols, wdtype, mode, std=0.02, q4_fn=None):
    """q4_fn: optional drop-in for tensor.quantize_q4 (same bytes, faster): the GPU weight quantiser
    (CudaTensorOperations.quantize_q4_weights) in bench.py, the oracle's C quantiser in tests."""
    rng = np.random.default_rng(seed)
    if wdtype == Q4:
        if mode == "direct":
            q, s = _direct_q4(rng, rows, cols, std)
        else:
            q, s = (q4_fn or quantize_q4)(rng.standard_normal((rows, cols), dtype=np.float32) * np.float32(std))
        return (Q4, q, s)
    if wdtype == I8:
        if mode == "direct":
            q, s = _direct_q8(rng, rows, cols, std)
        else:
            q, s = quantize_q8_weights(rng.standard_normal((rows, cols), dtype=np.float32) * np.float32(std))
        return (I8, q, s)
    w = rng.standard_normal((rows, cols), dtype=np.float32) * np.float32(std)
    if wdtype == BF16:
        from .tensor import float32_to_bfloat16
        return (BF16, float32_to_bfloat16(w), None)
    return (F32, w, None)


def tensor_specs(cfg):
    """(name, rows, cols, kind) for every tensor of a Llama checkpoint; kind in {'linear','norm','embed','head'}."""
    E, H, V = cfg["E"], cfg["H"], cfg["vocab"]
    hs = E // cfg["heads"]
    kvl = cfg["kv_heads"] * hs
    specs = [("model.embed_tokens.weight", V, E, "embed")]
    for i in range(cfg["layers"]):
        b = "model.layers.%d." % i
        specs += [
            (b + "input_layernorm.weight", 1, E, "norm"),
            (b + "self_attn.q_proj.weight", E, E, "linear"),
            (b + "self_attn.k_proj.weight", kvl, E, "linear"),
            (b + "self_attn.v_proj.weight", kvl, E, "linear"),
            (b + "self_attn.o_proj.weight", E, E, "linear"),
            (b + "post_attention_layernorm.weight", 1, E, "norm"),
        ]
        if cfg.get("experts"):  # MixtralModel.java:88-105
            specs.append((b + "block_sparse_moe.gate.weight", cfg["experts"], E, "linear"))
            for e in range(cfg["experts"]):
                p = b + "block_sparse_moe.experts.%d." % e
                specs += [(p + "w1.weight", H, E, "linear"), (p + "w2.weight", E, H, "linear"), (p + "w3.weight", H, E, "linear")]
        else:
            specs += [
                (b + "mlp.gate_proj.weight", H, E, "linear"),
                (b + "mlp.down_proj.weight", E, H, "linear"),
                (b + "mlp.up_proj.weight", H, E, "linear"),
            ]
    specs.append(("model.norm.weight", 1, E, "norm"))
    if not cfg.get("tied"):
        specs.append(("lm_head.weight", V, E, "head"))
    return specs


def tensor_seed(cfg, name):
    import zlib
    return (SEED0 + zlib.crc32(name.encode())) & 0x7FFFFFFFFFFFFFFF


def make_one(cfg, name, rows, cols, kind, wdtype=Q4, mode="quantize", embed_dtype=None, q4_fn=None):
    seed = tensor_seed(cfg, name)
    if kind == "norm":
        rng = np.random.default_rng(seed)
        return (F32, (1.0 + 0.1 * rng.standard_normal((rows, cols), dtype=np.float32)).astype(np.float32), None)
    if kind == "embed":
        # Jlama's quantiser also quantises the embedding table (skip pattern is only "norm",
        # QuantizeCommand.java:36-38), so a JQ4 checkpoint has a Q4 embedding (LlamaModel.java:91-97)
        dt = wdtype if embed_dtype is None else embed_dtype
        return make_tensor(seed, rows, cols, dt, mode, std=1.0, q4_fn=q4_fn)
    # residual-branch output projections are down-scaled by 1/sqrt(2L) (GPT-2 / Llama style init) so that the
    # synthetic network is residual-dominated and well-conditioned like a trained one; with every matrix at
    # std 0.02 a random transformer amplifies 1e-7 summation-order differences into O(1) logit changes.
    std = 0.02
    if name.endswith("o_proj.weight") or name.endswith("down_proj.weight") or name.endswith(".w2.weight"):
        std = 0.02 / float(np.sqrt(2.0 * cfg["layers"]))
    return make_tensor(seed, rows, cols, wdtype, mode, std=std, q4_fn=q4_fn)


def make_weights(cfg, wdtype=Q4, mode="quantize", embed_dtype=None, threads=None, q4_fn=None):
    """Full synthetic checkpoint in host memory (tensors generated in parallel; each has its own seed)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    specs = tensor_specs(cfg)
    threads = threads or min(32, os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=threads) as ex:
        vals = list(ex.map(lambda sp: make_one(cfg, sp[0], sp[1], sp[2], sp[3], wdtype, mode, embed_dtype, q4_fn), specs))
    return {sp[0]: v for sp, v in zip(specs, vals)}


def lazy_weights(cfg, wdtype=Q4, mode="quantize", embed_dtype=None, q4_fn=None):
    """get(name) -> tensor generated on demand (same seeds and bytes as make_weights).  A rank of a tensor- / expert-parallel job
    only asks for what it holds, so a Mixtral-8x7B shard never materialises the other ranks' 29 GB."""
    specs = {sp[0]: sp for sp in tensor_specs(cfg)}

    def get(name):
        sp = specs.get(name)
        return None if sp is None else make_one(cfg, sp[0], sp[1], sp[2], sp[3], wdtype, mode, embed_dtype, q4_fn)
    return get


def linear_weight_count(cfg):
    """Parameters streamed per decoded token: all linear layers + lm_head."""
    E, H, V = cfg["E"], cfg["H"], cfg["vocab"]
    hs = E // cfg["heads"]
    kvl = cfg["kv_heads"] * hs
    if cfg.get("experts"):  # per token: attention + router + the selected experts
        per_layer = E * E * 2 + kvl * E * 2 + cfg["experts"] * E + cfg["experts_per_token"] * H * E * 3
    else:
        per_layer = E * E * 2 + kvl * E * 2 + H * E * 3
    return per_layer * cfg["layers"] + V * E


def random_prompt(cfg, n, seed=1234):
    return np.random.default_rng(seed).integers(0, cfg["vocab"], size=n, dtype=np.int32)


# ---- GPT-2 family (BASELINE config 1: "GPT-2-small F32") ------------------------------------------------------------------------
GPT2_CONFIGS = {
    "gpt2-tiny": dict(ctx=64, E=128, H=512, heads=4, layers=2, vocab=256, eps=1e-5),
    "gpt2-small": dict(ctx=1024, E=768, H=3072, heads=12, layers=12, vocab=50257, eps=1e-5),  # gpt2 (124M) dims
}


def get_gpt2_config(name):
    cfg = dict(GPT2_CONFIGS[name])
    cfg["name"], cfg["kv_heads"] = name, cfg["heads"]
    return cfg


def make_gpt2_weights(cfg):
    """Synthetic F32 GPT-2 checkpoint in the file layout the reference loads (GPT2Model.java:54-129): Conv1D weights are [in, out]."""
    E, H, V, L = cfg["E"], cfg["H"], cfg["vocab"], cfg["layers"]
    out = {}

    def t(name, shape, std, mean=0.0):
        rng = np.random.default_rng(tensor_seed(cfg, name))
        out[name] = (F32, (mean + std * rng.standard_normal(shape, dtype=np.float32)).astype(np.float32), None)

    t("wte.weight", (V, E), 0.5)
    t("wpe.weight", (cfg["ctx"], E), 0.1)
    res = 0.02 / float(np.sqrt(2.0 * L))
    for i in range(L):
        b = "h.%d." % i
        t(b + "ln_1.weight", (E,), 0.1, 1.0), t(b + "ln_1.bias", (E,), 0.05)
        t(b + "attn.c_attn.weight", (E, 3 * E), 0.02), t(b + "attn.c_attn.bias", (3 * E,), 0.01)
        t(b + "attn.c_proj.weight", (E, E), res), t(b + "attn.c_proj.bias", (E,), 0.01)
        t(b + "ln_2.weight", (E,), 0.1, 1.0), t(b + "ln_2.bias", (E,), 0.05)
        t(b + "mlp.c_fc.weight", (E, H), 0.02), t(b + "mlp.c_fc.bias", (H,), 0.01)
        t(b + "mlp.c_proj.weight", (H, E), res), t(b + "mlp.c_proj.bias", (E,), 0.01)
    t("ln_f.weight", (E,), 0.1, 1.0), t("ln_f.bias", (E,), 0.05)
    return out
