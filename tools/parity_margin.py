"""How decisive is the reference's own arg-max on the bench prompt?  (CPU only; diagnostic for the in-run parity field)
Prints, for the first tokens after the 32-token bench prompt at the bench's 8B shapes and weights, the top logits of the two CPU
implementations of the same arithmetic (the reference's AVX-512 C kernels and the plain-C port) and the gap between the top two
relative to the largest logit -- to be read next to parity.max_logit_rel_err."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from jlama_b200 import synth  # noqa: E402
from oracle import oracle as o  # noqa: E402


def main():
    model = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    cfg = synth.get_config(model)
    w = bench.make_model_weights(cfg, "quantize", q4_fn=o.quantize_q4)
    prompt = synth.random_prompt(cfg, 32)
    out = {}
    for ref in (True, False):
        label = o.load_reference_kernels()
        o.use_reference_kernels(ref and label is not None)
        o.set_num_threads(o.available_cpus())
        m = o.OracleLlama(cfg, w, act_q8=True)
        toks, logits = m.generate(prompt, n)
        m.close()
        out[ref] = (toks, logits)
        for i in range(n):
            lg = np.asarray(logits[i])
            top = np.argsort(lg)[-3:][::-1]
            mx = np.abs(lg).max()
            print("%-22s step %d: token %6d  top3 %s  logits %s  gap(top1-top2)/max|logit| = %.3e" % (
                "reference C kernels" if ref else "plain-C port", i, toks[i], top.tolist(), np.round(lg[top], 4).tolist(), (lg[top[0]] - lg[top[1]]) / mx), flush=True)
    for i in range(n):
        a, b = np.asarray(out[True][1][i]), np.asarray(out[False][1][i])
        print("step %d: the two CPU implementations differ by %.3e of max|logit| (tokens %d / %d)" % (i, np.abs(a - b).max() / np.abs(b).max(), out[True][0][i], out[False][0][i]))
        if out[True][0][i] != out[False][0][i]:
            break


if __name__ == "__main__":
    main()
