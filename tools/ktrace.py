"""Device-side timeline of one decoded token: per traced launch (quantised GEMVs + decode attention) the start of the
first CTA, the end of the last CTA and the end of CTA 0's prologue (globaltimer), i.e. kernel durations AND the gaps
between them inside the replayed CUDA graph.  Usage: python tools/ktrace.py [--model llama-3-8b] [--flags N]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jlama_b200 import native, synth  # noqa: E402
from jlama_b200.model import LlamaModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama-3-8b")
    ap.add_argument("--flags", type=int, default=0)
    ap.add_argument("--prompt", type=int, default=32)
    ap.add_argument("--layers", type=int, default=2, help="layers printed in full")
    ap.add_argument("--json", action="store_true", help="print one JSON summary line instead of the table (used by bench.py)")
    a = ap.parse_args()
    cfg = synth.get_config(a.model)
    ctx = native.Context(0)
    w = synth.make_weights(cfg, wdtype=native.Q4, mode="direct")
    cap = 8192
    ctx.check(ctx.lib.jl_debug_ktrace(ctx.h, cap))
    m = LlamaModel(ctx, cfg, w, max_context=512, flags=a.flags)
    prompt = synth.random_prompt(cfg, a.prompt)
    m.batch_forward(prompt, 0)
    first, _ = m.sample(want_logits=False)
    buf = np.zeros((cap, 16), dtype=np.uint64)
    n0 = ctx.lib.jl_debug_ktrace_read(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), cap)
    toks = m.decode_resident(first, a.prompt, 8)  # captures the graph (slots n0..), replays it
    ctx.check(ctx.lib.jl_debug_ktrace_clear(ctx.h))
    m.decode_resident(int(toks[-1]), a.prompt + 8, 1)
    tot_ms, _ = m.last_timing()
    n = ctx.lib.jl_debug_ktrace_read(ctx.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), cap)
    rows = [buf[i] for i in range(n0, n) if buf[i][1] > 0]
    rows.sort(key=lambda r: int(r[0]))
    t0 = int(rows[0][0])
    if a.json:
        import json
        gemv_us = sum((int(r[1]) - int(r[0])) / 1e3 for r in rows if (int(r[3]) & 0xF00) == 0x100)
        attn_us = sum((int(r[1]) - int(r[0])) / 1e3 for r in rows if (int(r[3]) & 0xF00) == 0xA00)
        print(json.dumps({"model": a.model, "position": a.prompt + 8, "traced_launches": len(rows), "step_ms_events": tot_ms,
                          "gemv_launches": sum(1 for r in rows if (int(r[3]) & 0xF00) == 0x100), "gemv_us": gemv_us,
                          "attention_us": attn_us, "span_us": (int(rows[-1][1]) - t0) / 1e3}), flush=True)
        m.close()
        ctx.close()
        return
    print("# %s flags=%d JL_PF=%s: %d traced launches, event-timed step %.3f ms, trace span %.3f ms" % (
        a.model, a.flags, os.environ.get("JL_PF", "0"), len(rows), tot_ms, (int(rows[-1][1]) - t0) / 1e6))
    print("# kind                        start_us   dur_us  prologue_us  gap_before_us")
    per_layer = 5
    prev_end = t0
    agg = {}
    for i, r in enumerate(rows):
        s, e, mid, tag = int(r[0]), int(r[1]), int(r[2]), int(r[3])
        if tag & 0xF00 == 0xA00:
            kind = "attention splits=%d" % (tag >> 16)
        else:
            kind = "gemv epi=%d pro=%d rows=%d K=%d" % (tag & 0xF, (tag >> 4) & 0xF, (tag >> 16) & 0xFFFFFF, tag >> 40)
        dur, pro, gap = (e - s) / 1e3, (mid - s) / 1e3 if mid else float("nan"), (s - prev_end) / 1e3
        if i < per_layer * a.layers or i >= len(rows) - 2:
            print("%-28s %9.2f %8.2f %11.2f %13.2f" % (kind, (s - t0) / 1e3, dur, pro, gap))
        g = agg.setdefault(kind, [0, 0.0, 0.0, 0.0, np.zeros(16), np.zeros(16)])
        g[0] += 1
        g[1] += dur
        g[2] += pro if mid else 0.0
        g[3] += gap
        for j in range(4, 16):
            if int(r[j]):
                g[4][j] += (int(r[j]) - s) / 1e3
                g[5][j] += 1
        prev_end = e
    print("# averages over the step")
    for k, g in agg.items():
        print("%-28s n=%3d  dur %7.2f  prologue %6.2f  gap_before %6.2f" % (k, g[0], g[1] / g[0], g[2] / g[0], g[3] / g[0]))
        st = ["s%d=%.2f" % (j, g[4][j] / g[5][j]) for j in range(4, 16) if g[5][j]]
        if st:
            print("    CTA0 stamps (us after first CTA start): " + " ".join(st))
    m.close()
    ctx.close()


if __name__ == "__main__":
    main()
