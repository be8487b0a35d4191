"""The checkpoint reader / loader (jlama_b200/csrc/jl_safetensors.cu) parses files it did not write: build it as plain host C++ with
AddressSanitizer + UBSan and run tests/native/safetensors_sanitize.cpp -- every shard of well-formed Llama and Mixtral checkpoints at
tp 1 / 2 / 4 with a jl_register_tensor stub that reads every byte it is handed, then thousands of mutated headers and config.json
texts.  An out-of-bounds read of a single byte, an overflow or a misaligned typed load ends the run with a non-zero exit code."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_checkpoint_reader_under_address_and_ub_sanitizers(tmp_path):
    # jl_safetensors.cu includes "jl_common.cuh" (device helpers it does not use): give the copy a host-only stand-in
    shutil.copy(os.path.join(ROOT, "jlama_b200", "csrc", "jl_safetensors.cu"), tmp_path / "jl_safetensors.cu")
    shutil.copy(os.path.join(ROOT, "tests", "native", "fake_jl_common.cuh"), tmp_path / "jl_common.cuh")
    shutil.copy(os.path.join(ROOT, "tests", "native", "safetensors_sanitize.cpp"), tmp_path / "safetensors_sanitize.cpp")
    exe = str(tmp_path / "st_asan")
    cmd = ["g++", "-std=c++17", "-x", "c++", "-g", "-O1", "-w", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", os.path.join(ROOT, "include"),
           "safetensors_sanitize.cpp", "-o", exe]
    b = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=str(tmp_path))
    if b.returncode != 0 and ("cannot find" in b.stdout or "unrecognized" in b.stdout):
        pytest.skip("this g++ has no sanitizer runtime: " + b.stdout[-200:])
    assert b.returncode == 0, b.stdout[-3000:]
    work = tmp_path / "work"
    work.mkdir()
    r = subprocess.run([exe, str(work)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1", UBSAN_OPTIONS="halt_on_error=1"))
    assert r.returncode == 0 and r.stdout.strip().splitlines()[-1].startswith("ok:"), r.stdout[-4000:]
    assert "ERROR: AddressSanitizer" not in r.stdout and "runtime error" not in r.stdout
