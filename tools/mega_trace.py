"""Phase trace of one megakernel decode step (diagnostic): where does a CTA spend its time per op?"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jlama_b200 import native, synth  # noqa: E402
from jlama_b200.model import LlamaModel  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
    cfg = synth.get_config(name)
    w = synth.make_weights(cfg, mode="direct")
    ctx = native.Context(0)
    m = LlamaModel(ctx, cfg, w, max_context=512, flags=native.MODEL_MEGA)
    prompt = synth.random_prompt(cfg, 32)
    m.reset_session(0)
    m.batch_forward(prompt, 0)
    first, _ = m.sample(want_logits=False)
    if not os.environ.get('JL_MEGA_DBG'):
        m.decode_resident(first, 32, 8)
    n_ops = cfg["layers"] * 4 + 1
    buf = np.zeros((3, n_ops, 8), dtype=np.int64)
    ctx.check(ctx.lib.jl_model_debug_trace(m.h, 0, int(first), 40, native.ptr(buf), buf.size))
    ghz = 1.965
    names = ["qkv", "o", "gateup", "down"]
    for c, cname in enumerate(("cta0", "ctaMid", "ctaLast")):
        t = buf[c].astype(np.float64) / (ghz * 1e3)  # us
        t0 = t[0, 0]
        print("== %s: total %.1f us" % (cname, t[n_ops - 1, 5] - t0))
        agg = {}
        for op in range(n_ops):
            nm = names[op % 4] if op < n_ops - 1 else "lm_head"
            a = agg.setdefault(nm, np.zeros(5))
            st = t[op]
            if st[2] == 0:  # no rows
                continue
            a += np.array([st[1] - st[0], st[2] - st[1], st[3] - st[2], st[4] - st[3], st[5] - st[4]])
        print("%-8s %10s %10s %10s %10s %10s   (us summed over layers)" % ("op", "attention", "dep-wait", "prologue", "stages", "signal"))
        for nm, a in agg.items():
            print("%-8s %10.1f %10.1f %10.1f %10.1f %10.1f" % (nm, *a))
        # first layers in detail
        for op in range(0, 8):
            st = t[op] - t0
            print("  op %2d %-7s " % (op, names[op % 4]) + " ".join("%8.1f" % x for x in st[:6]))
    m.close()
    ctx.close()


if __name__ == "__main__":
    main()
