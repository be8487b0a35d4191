"""Serving throughput through the session scheduler (csrc/jl_sched.cu): R requests with mixed prompt / output lengths over S session
slots, continuous batching vs the same requests run one after the other through generate().  NOT RUN YET: written after the round's GPU
budget was spent; no number in DESIGN.md comes from it.

    python tools/sched_bench.py [--model llama-3-8b] [--slots 8] [--requests 32] [--prefill-budget 512] [--temperature 0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jlama_b200 import native, synth  # noqa: E402
from jlama_b200.model import LlamaModel  # noqa: E402
from jlama_b200.scheduler import SessionScheduler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--slots", type=int, default=8)
    ap.add_argument("--requests", type=int, default=32)
    ap.add_argument("--prefill-budget", type=int, default=512)
    ap.add_argument("--temperature", type=float, default=0.0)
    ap.add_argument("--max-context", type=int, default=1024)
    args = ap.parse_args()
    cfg = synth.get_config(args.model)
    rng = np.random.default_rng(7)
    work = [(synth.random_prompt(cfg, int(rng.integers(16, 384)), seed=700 + i), int(rng.integers(16, 128))) for i in range(args.requests)]
    ctx = native.Context(0)
    m = LlamaModel(ctx, cfg, synth.make_weights(cfg, mode="direct"), max_sessions=args.slots, max_context=args.max_context, prefill_tensor_core=1)
    m.generate(work[0][0], 4)  # warm-up: kernel load, graphs, pages of session 0
    out = {"model": cfg["name"], "slots": args.slots, "requests": args.requests, "prefill_budget": args.prefill_budget,
           "generated_tokens": int(sum(n for _, n in work)), "prompt_tokens": int(sum(len(p) for p, _ in work))}
    with SessionScheduler(m, prefill_tokens_per_step=args.prefill_budget) as sched:
        t0 = time.perf_counter()
        ids = [sched.submit(p, n, temperature=args.temperature, seed=i) for i, (p, n) in enumerate(work)]
        st = sched.run()
        dt = time.perf_counter() - t0
        infos = [sched.info(r) for r in ids]
        out["scheduler"] = {"seconds": dt, "generated_tokens_per_s": out["generated_tokens"] / dt, "decode_rows_per_call": st.decode_rows / max(1, st.decode_calls),
                            "mean_queue_ms": float(np.mean([i.queue_ms for i in infos])), "mean_prompt_ms": float(np.mean([i.prompt_ms for i in infos])),
                            "mean_ms_per_generated_token": float(np.mean([i.generate_ms / max(1, i.n_generated - 1) for i in infos]))}
    t0 = time.perf_counter()
    for p, n in work:
        m.generate(p, n)
    dt = time.perf_counter() - t0
    out["one_by_one"] = {"seconds": dt, "generated_tokens_per_s": out["generated_tokens"] / dt}
    print(json.dumps(out), flush=True)
    m.close()
    ctx.close()


if __name__ == "__main__":
    main()
