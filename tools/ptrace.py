"""Phase timeline of the persistent decode kernel (CTA 0's globaltimer stamps): per layer the time from "dependency met" to
"phase done" of QKV / attention / o_proj / gate+up / down, the waits in between, lm_head and the whole token.
Usage (GPU box): JL_PD_TRACE=1 python tools/ptrace.py [--model llama-3-8b] [--prompt 32] [--json]"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

os.environ.setdefault("JL_PD_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jlama_b200 import native, synth  # noqa: E402
from jlama_b200.model import LlamaModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--prompt", type=int, default=32)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args()
    cfg = synth.get_config(a.model)
    ctx = native.Context(0)
    w = synth.make_weights(cfg, wdtype=native.Q4, mode="direct")
    m = LlamaModel(ctx, cfg, w, max_context=max(512, a.prompt + a.steps + 8))
    assert m.decode_mode(1) == 3, "persistent kernel not active"
    prompt = synth.random_prompt(cfg, a.prompt)
    m.batch_forward(prompt, 0)
    first, _ = m.sample(want_logits=False)
    toks = m.decode_resident(first, a.prompt, a.steps)
    tot_ms, _ = m.last_timing()
    L = cfg["layers"]
    buf = np.zeros(L * 16 + 32, dtype=np.uint64)
    n = ctx.lib.jl_model_debug_trace(m.h, buf.ctypes.data_as(C.c_void_p), buf.size)
    assert n > 0, ctx.lib.jl_last_error(ctx.h)
    t = buf.astype(np.int64)
    t0 = int(t[0])
    per = t[1:1 + L * 8].reshape(L, 8)
    lm0, end = int(t[1 + L * 8]), int(t[1 + L * 8 + 1])
    names = ["qkv", "attn(cta0)", "wait att", "o_proj", "wait o", "gate+up", "wait gu", "down+wait"]
    # stamps: 0 qkv dep met, 1 qkv done, 2 attention done, 3 o dep met, 4 o done, 5 gu dep met, 6 gu done, 7 down dep met
    seg = np.zeros((L, 8))
    for l in range(L):
        s = per[l]
        nxt = per[l + 1][0] if l + 1 < L else lm0
        seg[l] = [s[1] - s[0], (s[2] - s[1]) if s[2] else 0, s[3] - (s[2] if s[2] else s[1]), s[4] - s[3], s[5] - s[4], s[6] - s[5], s[7] - s[6],
                  nxt - s[7]]
    seg /= 1e3
    mean = seg[1:].mean(axis=0) if L > 1 else seg[0]
    out = {"model": a.model, "position": a.prompt + a.steps - 1, "token_us": (end - t0) / 1e3, "event_ms_per_token": tot_ms / a.steps,
           "layer_us": float(mean.sum()), "lm_head_us": (end - lm0) / 1e3, "phases_us": {k: float(v) for k, v in zip(names, mean)}}
    if a.json:
        print(json.dumps(out), flush=True)
    else:
        print("# %s persistent decode, position %d: token %.1f us (events: %.3f ms/token), layer %.2f us, lm_head+argmax %.1f us" % (
            a.model, out["position"], out["token_us"], out["event_ms_per_token"], out["layer_us"], out["lm_head_us"]))
        for k, v in zip(names, mean):
            print("  %-12s %7.2f us" % (k, v))
    m.close()
    ctx.close()


if __name__ == "__main__":
    main()
