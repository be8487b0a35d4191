"""Decode at long contexts ("one decode point at position 2048 (and 8000) with KV bytes in the numerator"):
Llama-3-8B JQ4, one session, F32 KV pages; the prompt goes through the tensor-core prefill path, then `n` tokens are decoded in the
device-resident loop (persistent kernel, flat attention with context splits) and timed with CUDA events by the library.
   python tools/long_context_bench.py [--positions 2048,8000] [--steps 32]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from jlama_b200 import native, synth  # noqa: E402
from jlama_b200.model import LlamaModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--positions", default="2048,8000")
    ap.add_argument("--steps", type=int, default=32)
    a = ap.parse_args()
    cfg = synth.get_config(a.model)
    ctx = native.Context(0)
    peak, _ = bench.measured_peaks()
    w = synth.make_weights(cfg, mode="direct")
    hs = cfg["E"] // cfg["heads"]
    for p in [int(x) for x in a.positions.split(",")]:
        p = min(p, cfg["ctx"] - a.steps - 16)
        m = LlamaModel(ctx, cfg, w, max_context=p + a.steps + 16, prefill_tensor_core=1, max_batch=2048)
        prompt = synth.random_prompt(cfg, p, seed=7)
        m.batch_forward(prompt, 0)
        first, _ = m.sample(want_logits=False)
        m.decode_resident(first, p, 8)
        m.decode_resident(first, p + 8, a.steps)
        ms, _ = m.last_timing()
        step = ms / a.steps / 1e3
        kv = cfg["layers"] * 2 * cfg["kv_heads"] * hs * 4 * (p + 8 + a.steps / 2.0 + 1)
        wb = m.weight_bytes()
        print(json.dumps({"position": p + 8, "ms_per_step": 1e3 * step, "tokens_per_s": 1.0 / step, "weight_bytes": wb, "kv_bytes_per_token": kv,
                          "frac_of_hbm_peak_weights_plus_kv": (wb + kv) / 1e9 / step / peak,
                          "kv_GBps_over_short_context_step": None}), flush=True)
        m.close()
    ctx.close()


if __name__ == "__main__":
    main()
