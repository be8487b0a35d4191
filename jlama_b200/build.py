"""Builds jlama_b200/libjlama_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

The library is plain CUDA runtime + a C ABI (include/jlama_b200.h); no torch, no pybind.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libjlama_b200.so")
OBJ = os.path.join(HERE, "build")
SOURCES = ["jl_runtime.cu", "jl_gemv.cu", "jl_gemm_tc.cu", "jl_elementwise.cu", "jl_attention.cu", "jl_model.cu", "jl_comm.cu", "jl_pdecode.cu", "jl_safetensors.cu", "jl_gemm8.cu", "jl_attn_prefill.cu", "jl_sched.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "jlama_b200.h"))
    return hdrs


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    hdr_time = max(os.path.getmtime(h) for h in _deps())
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s.replace(".cu", ".o"))
        objs.append(obj)
        if force or _newer(src, obj) or hdr_time > os.path.getmtime(obj):
            extra = ["-Xptxas", "-v"] if verbose else []
            jobs.append([NVCC] + FLAGS + extra + ["-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return cmd, r.returncode, r.stdout

    failed = False
    with ThreadPoolExecutor(max_workers=max(1, min(8, len(jobs) or 1))) as ex:
        for cmd, rc, out in ex.map(run, jobs):
            if rc != 0 or verbose:
                sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
            failed |= rc != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if jobs or not os.path.exists(OUT):
        cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-ldl", "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
