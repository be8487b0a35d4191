"""Micro-benchmark of the tcgen05 prefill GEMM at Llama-3-8B layer shapes: TFLOP/s vs the measured BF16 peak."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jlama_b200 import native  # noqa: E402

SHAPES = [("q/o 8B", 4096, 4096), ("gate 8B", 14336, 4096), ("down 8B", 4096, 14336), ("lm_head 8B", 128256, 4096)]


def main():
    peak = 1454.3
    if os.path.exists("MEASURED_PEAKS.json"):
        peak = json.load(open("MEASURED_PEAKS.json")).get("bf16_tflops_sustained", peak)
    ctx = native.Context(0)
    rng = np.random.default_rng(0)
    for name, n, k in SHAPES:
        q = rng.integers(0, 256, (n, k // 2), dtype=np.uint8)
        s = ((0.5 + rng.random((n, k // 32))) * 0.004).astype(np.float32)
        tid = ctx.lib.jl_register_tensor(ctx.h, native.Q4, n, k, native.ptr(q), native.ptr(s))
        for t in (128, 256, 1024, 2048):
            us = C.c_double()
            ctx.check(ctx.lib.jl_debug_gemm_tc_bench(ctx.h, tid, t, 20, C.byref(us)))
            tf = 2.0 * t * n * k / (us.value * 1e-6) / 1e12
            print("%-11s N=%6d K=%5d T=%4d  %9.1f us  %7.1f TFLOP/s  %.3f of measured sustained BF16 peak" % (name, n, k, t, us.value, tf, tf / peak), flush=True)
        ctx.check(ctx.lib.jl_unregister_tensor(ctx.h, tid))
    ctx.close()


if __name__ == "__main__":
    main()
