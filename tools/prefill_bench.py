"""Prefill throughput of the tensor-core path at the BASELINE config-3 prompt length (diagnostic, run on the GPU box):
   python tools/prefill_bench.py [--model llama-3-8b] [--tokens 2048] [--batch 256,512]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jlama_b200 import native, synth  # noqa: E402
from jlama_b200.model import LlamaModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--tokens", type=int, default=2048)
    ap.add_argument("--batch", default="256,512,1024,2048")
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()
    cfg = synth.get_config(args.model)
    ctx = native.Context(0)
    w = synth.make_weights(cfg, mode="direct")
    prompt = synth.random_prompt(cfg, args.tokens, seed=99)
    for mb in [int(x) for x in args.batch.split(",")]:
        m = LlamaModel(ctx, cfg, w, max_context=args.tokens + 64, prefill_tensor_core=1, max_batch=mb)
        m.batch_forward(prompt, 0)  # warm-up: allocates the KV pages of the whole prompt, loads every kernel
        best = None
        for _ in range(args.repeat):
            m.reset_session(0)
            ctx.sync()
            l0 = ctx.kernel_launches()
            t0 = time.perf_counter()
            m.batch_forward(prompt, 0)
            ctx.sync()
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
        dt = best
        flops = 2.0 * (synth.linear_weight_count(cfg) - cfg["vocab"] * cfg["E"]) * args.tokens
        print("%s prefill %d tokens, chunks of %d: %.1f ms  %.0f tokens/s  %.1f linear TFLOP/s  (%d launches)" % (
            cfg["name"], args.tokens, mb, dt * 1e3, args.tokens / dt, flops / dt / 1e12, ctx.kernel_launches() - l0), flush=True)
        m.close()
    ctx.close()


if __name__ == "__main__":
    main()
