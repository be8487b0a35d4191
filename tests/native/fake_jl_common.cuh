// Stand-in for jlama_b200/csrc/jl_common.cuh when jl_safetensors.cu is built as plain host C++ for the sanitizer harness
// (tests/test_safetensors_sanitizers.py copies jl_safetensors.cu next to this file under the name jl_common.cuh expects).
// jl_safetensors.cu needs nothing device-side: the public header, the standard containers and jl_set_error.
#pragma once
#include <stdint.h>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "jlama_b200.h"

int jl_set_error(jl_ctx *ctx, int code, const char *fmt, ...);
