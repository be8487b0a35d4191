"""Host-side mirror of Jlama's checkpoint I/O (core/safetensors/SafeTensorSupport.java, Weights.java, SafeTensorIndex.java)
on top of the C ABI (csrc/jl_safetensors.cu): the safetensors container with Jlama's dtype strings "Q4" / "I8" and the
"<name>.qb" f32 block-scale tensors, index.json shard maps, the offline quantiser (jlama quantize) and the loader."""
import ctypes as C
import os

import numpy as np

from . import native
from .native import BF16, F32, I8, Q4

F16 = 4
_NP = {F32: np.float32, BF16: np.uint16, Q4: np.uint8, I8: np.int8, F16: np.float16}
_NAMES = {F32: "F32", BF16: "BF16", Q4: "Q4", I8: "I8", F16: "F16"}


class SafeTensorsError(ValueError):
    pass


class SafeTensors:
    """SafeTensorSupport.loadWeights / readWeights: a file, or a model directory (single file or index.json + shards)."""

    def __init__(self, path):
        self.lib = native.load()
        h = C.c_void_p()
        rc = self.lib.jl_st_open(os.fsencode(path), C.byref(h))
        if rc != 0:
            raise SafeTensorsError(self.lib.jl_st_last_error().decode())
        self.h = h
        self._index = {}
        for i in range(self.lib.jl_st_count(self.h)):
            name, dt, nd, nb = C.c_char_p(), C.c_int(), C.c_int(), C.c_int64()
            shape = (C.c_int64 * 4)()
            self.lib.jl_st_info(self.h, i, C.byref(name), C.byref(dt), C.byref(nd), shape, C.byref(nb))
            self._index[name.value.decode()] = (i, dt.value, tuple(shape[k] for k in range(nd.value)), nb.value)

    def names(self):
        """tensor names in data-offset order (TensorInfo.compareTo)"""
        return list(self._index)

    def info(self, name):
        i, dt, shape, nbytes = self._index[name]
        return {"dtype": _NAMES.get(dt, "?"), "dtype_code": dt, "shape": shape, "nbytes": nbytes}

    def metadata(self, key):
        v = self.lib.jl_st_metadata(self.h, key.encode())
        return v.decode() if v is not None else None

    def majority_dtype(self):
        return self.lib.jl_st_majority_dtype(self.h)

    def get(self, name):
        """numpy copy of the raw tensor (Q4: packed bytes [rows, cols/2]; BF16: uint16 bit patterns)"""
        i, dt, shape, nbytes = self._index[name]
        ptr = self.lib.jl_st_data(self.h, i)
        raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,)).copy() if nbytes else np.zeros(0, np.uint8)
        a = raw.view(_NP[dt])
        if dt == Q4:
            return a.reshape(shape[0], shape[1] // 2) if len(shape) == 2 else a
        return a.reshape(shape)

    def load(self, name):
        """Weights.load: (dtype_code, data, scales) with the ".qb" block scales paired (Weights.java:153-171)"""
        i, dt, shape, _ = self._index[name]
        scales = self.get(name + ".qb") if dt in (Q4, I8) else None
        return (dt, self.get(name), scales)

    def close(self):
        if self.h:
            self.lib.jl_st_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def write_safetensors(path, tensors, metadata=None):
    """tensors: ordered mapping name -> (dtype_code, ndarray, logical_shape or None).  Quantised tensors are written as
    "Q4"/"I8" with their logical shape; the caller adds the "<name>.qb" F32 entries (save_checkpoint does)."""
    lib = native.load()
    n = len(tensors)
    names = (C.c_char_p * n)(*[k.encode() for k in tensors])
    dts = (C.c_int * n)()
    nds = (C.c_int * n)()
    shp = (C.c_int64 * (4 * n))()
    ptrs = (C.c_void_p * n)()
    nbs = (C.c_int64 * n)()
    keep = []
    for i, (k, (dt, arr, shape)) in enumerate(tensors.items()):
        arr = np.ascontiguousarray(arr)
        keep.append(arr)
        shape = tuple(shape) if shape is not None else arr.shape
        dts[i], nds[i] = dt, len(shape)
        for j, d in enumerate(shape):
            shp[4 * i + j] = d
        ptrs[i] = arr.ctypes.data
        nbs[i] = arr.nbytes
    meta = metadata or {}
    kv = (C.c_char_p * (2 * len(meta)))(*[x.encode() for pair in meta.items() for x in pair]) if meta else None
    rc = lib.jl_st_write(os.fsencode(path), n, names, dts, nds, shp, ptrs, nbs, len(meta), kv)
    if rc != 0:
        raise SafeTensorsError(lib.jl_st_last_error().decode())


def save_checkpoint(dirpath, weights, config=None):
    """A synth-style weight dict (name -> (dtype_code, data, scales)) as a Jlama checkpoint directory: model.safetensors with
    "<name>.qb" scale tensors after each quantised tensor (SafeTensorSupport.java:264-277) + config.json."""
    import json
    os.makedirs(dirpath, exist_ok=True)
    out = {}
    for name, (dt, data, scales) in weights.items():
        if dt == Q4:
            out[name] = (Q4, data, (data.shape[0], data.shape[1] * 2))
            out[name + ".qb"] = (F32, scales, None)
        elif dt == I8:
            out[name] = (I8, data, None)
            out[name + ".qb"] = (F32, scales, None)
        else:
            out[name] = (dt, data, None)
    write_safetensors(os.path.join(dirpath, "model.safetensors"), out)
    if config is not None:
        hf = {"architectures": ["LlamaForCausalLM"], "hidden_size": config["E"], "intermediate_size": config["H"],
              "num_attention_heads": config["heads"], "num_key_value_heads": config["kv_heads"], "num_hidden_layers": config["layers"],
              "vocab_size": config["vocab"], "max_position_embeddings": config["ctx"], "rms_norm_eps": config["eps"],
              "rope_theta": config["rope_theta"], "tie_word_embeddings": bool(config.get("tied"))}
        if config.get("experts"):  # MixtralConfig.java: num_local_experts / num_experts_per_tok
            hf.update({"architectures": ["MixtralForCausalLM"], "num_local_experts": config["experts"],
                       "num_experts_per_tok": config["experts_per_token"]})
        with open(os.path.join(dirpath, "config.json"), "w") as f:
            json.dump(hf, f)


def quantize_model(ctx, src_dir, dst_dir, qtype=Q4, skip=None, drop=None):
    """`jlama quantize` (SafeTensorSupport.quantizeModel :215-332) with the block quantisers on the GPU."""
    ctx.check(ctx.lib.jl_quantize_model(ctx.h, os.fsencode(src_dir), os.fsencode(dst_dir), qtype,
                                        None if skip is None else ",".join(skip).encode(), None if drop is None else ",".join(drop).encode()))
    return dst_dir


def config_from_json(path):
    mc = native.ModelConfig()
    lib = native.load()
    if lib.jl_config_from_json(os.fsencode(path), C.byref(mc)) != 0:
        raise SafeTensorsError(lib.jl_st_last_error().decode())
    return mc
