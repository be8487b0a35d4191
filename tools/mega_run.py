"""Runs a handful of megakernel decode steps (target for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jlama_b200 import native, synth  # noqa: E402
from jlama_b200.model import LlamaModel  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "llama-3-8b"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = synth.get_config(name)
w = synth.make_weights(cfg, mode="direct")
ctx = native.Context(0)
flags = native.MODEL_MEGA
if os.environ.get('JL_NO_MEGA'):
    flags = 0
if os.environ.get('JL_PDL'):
    flags |= native.MODEL_PDL
m = LlamaModel(ctx, cfg, w, max_context=512, flags=flags)
prompt = synth.random_prompt(cfg, 32)
m.reset_session(0)
m.batch_forward(prompt, 0)
first, _ = m.sample(want_logits=False)
out = m.decode_resident(first, 32, steps)
print("tokens", list(out), "ms/step", m.last_timing()[0] / steps)
m.close()
ctx.close()
