"""Micro-benchmark of the quantised GEMV kernel at Llama-3-8B / 1B layer shapes (diagnostic, run on the GPU box):
   python tools/gemv_bench.py        -> achieved GB/s per shape, with and without PDL overlap
CAVEAT: these are back-to-back launches in a stream, and stream launches are quantised to about
2.05 us on this stack, so 5-15 us kernels report multiples of 2.05 us.  Use tools/ktrace.py (timeline inside the replayed
CUDA graph) to compare kernel variants."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jlama_b200 import native  # noqa: E402

SHAPES = [("qkv 8B", 6144, 4096), ("o 8B", 4096, 4096), ("gate 8B", 14336, 4096), ("down 8B", 4096, 14336),
          ("lm_head 8B", 128256, 4096), ("qkv 1B", 3072, 2048), ("down 1B", 2048, 8192)]


def main():
    batch = "--batch" in sys.argv  # M = 1, 2, 4, 8 without PDL: the 2..8-row tensor-core kernel (jl_gemm8.cu) against M = 1
    quick = "--quick" in sys.argv  # M=1, no PDL, 8B shapes only: for comparing JL_GEMV_CFG / JL_GEMV_AHEAD variants
    tag = "quick"
    peak = 6563.9
    if os.path.exists("MEASURED_PEAKS.json"):
        peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    ctx = native.Context(0)
    rng = np.random.default_rng(0)
    rows_out = []
    for name, n, k in (SHAPES[:5] if quick else SHAPES):
        reps = max(1, int(400e6 // (n * k * 0.625)) + 1)  # > 3x L2 of distinct weights
        if n * reps > 400000:
            reps = max(1, 400000 // n)
        q = rng.integers(0, 256, (n * reps, k // 2), dtype=np.uint8)
        s = ((0.5 + rng.random((n * reps, k // 32))) * 0.004).astype(np.float32)
        tid = ctx.lib.jl_register_tensor(ctx.h, native.Q4, n * reps, k, native.ptr(q), native.ptr(s))
        assert tid > 0
        for m in ((1,) if quick else ((1, 2, 4, 8) if batch else (1, 4))):
            for mode, mname in ((0, "q8"), (1, "norm+q8"), (2, "f32")):
                if mode == 2 and m * k * 4 > 190 * 1024:
                    continue
                if quick and mode == 2 and n < 100000:
                    continue
                for pdl in ((0,) if quick or batch else (0, 1)):
                    us = C.c_double()
                    ctx.check(ctx.lib.jl_debug_gemv_bench(ctx.h, tid, n, m, mode, 200, pdl, C.byref(us)))
                    gbs = n * k * 0.625 / 1e9 / (us.value * 1e-6)
                    rows_out.append((name, n, k, m, mname, pdl, us.value, gbs, gbs / peak))
                    print(("[%s] " % tag if quick else "") +
                          "%-11s N=%6d K=%5d M=%d %-8s pdl=%d  %8.2f us  %7.1f GB/s  %.3f of measured peak" % rows_out[-1], flush=True)
        ctx.check(ctx.lib.jl_unregister_tensor(ctx.h, tid))
    ctx.close()


if __name__ == "__main__":
    main()
