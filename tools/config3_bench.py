"""BASELINE config 3 alone (Llama-3-8B Q8_0, 8 sessions, prefill 2048 / decode 128): the bench's config3 section as its own
process, for A/B runs of kernel variants through environment switches (JL_G8_KSPLIT, JL_ATTN_FLAT)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from jlama_b200 import native, synth  # noqa: E402


def main():
    cfg = synth.get_config("llama-3-8b")
    ctx = native.Context(0)
    peak, _ = bench.measured_peaks()
    sessions = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    prompt = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    r = bench.config3_workload(ctx, cfg, peak, sessions=sessions, prompt_tokens=prompt)
    r.pop("workload", None)
    print(json.dumps({"env": {k: os.environ.get(k) for k in ("JL_G8_KSPLIT", "JL_ATTN_FLAT")}, **r}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
