"""CudaTensorOperations: the host-side mirror of Jlama's operator plug-in interface
(core/tensor/operations/TensorOperations.java:25-161), method for method, forwarding to the C ABI
exactly like NativeGPUTensorOperations forwards to libjlamagpu
(jlama-native/.../operations/NativeGPUTensorOperations.java:216-335).

Same names, argument meaning and error behaviour as the reference (unsupported dtype pairs raise
UnsupportedOperation, shape violations raise ValueError/JlamaNativeError); results land in the
caller's host tensors.  There is no CPU delegate: every method runs a CUDA kernel.
"""
import ctypes as C

import numpy as np

from . import native
from .native import BF16, F32, I8, Q4, ptr
from .tensor import (AbstractTensor, BFloat16BufferTensor, FloatBufferTensor, Q4ByteBufferTensor,
                     Q8ByteBufferTensor)


class CudaTensorOperations:
    def __init__(self, ctx=None, device=0):
        self.ctx = ctx or native.Context(device)
        self._own = ctx is None
        self.lib = self.ctx.lib
        self._registered = {}  # AbstractTensor.getUid() -> device tensor id (NativeGPUTensorOperations.java:104-151)

    # -- TensorOperations.java:28-39 ---------------------------------------------------------------
    def name(self):
        return "CUDA sm_100a Operations"

    def parallel_split_size(self):
        return 1  # NativeGPUTensorOperations.java:99-101: one call per GEMM, no host-side chunking

    def preferred_working_quantized_type(self):
        return I8

    def register_model_tensor(self, t):
        if t.uid in self._registered:
            return self._registered[t.uid]
        tid = self.lib.jl_register_tensor(self.ctx.h, t.dtype, t.rows, t.cols, ptr(t.data), ptr(t.scales))
        if tid < 0:
            self.ctx.check(native.JL_ERR_OOM if b"memory" in (self.lib.jl_last_error(self.ctx.h) or b"") else native.JL_ERR_INVALID)
        self._registered[t.uid] = tid
        return tid

    def unregister_model_tensor(self, t):
        tid = self._registered.pop(t.uid, None)
        if tid is not None:
            self.ctx.check(self.lib.jl_unregister_tensor(self.ctx.h, tid))

    # -- dot products ------------------------------------------------------------------------------
    def dot_product(self, a, b, aoffset=0, boffset=0, limit=None):
        limit = a.cols - aoffset if limit is None else limit
        r = FloatBufferTensor(np.zeros((1, 1), dtype=np.float32))
        self.batch_dot_product(r, a, b, aoffset, boffset, limit, 0, 0, 1)
        return float(r.data[0, 0])

    def batch_dot_product(self, result, a, b, a_column_offset, b_column_offset, column_limit, r_row_offset=0,
                          b_row_offset=0, row_chunk_size=None):
        """result[i, j + rRowOffset] = sum_t a[i, aOff+t] * b[j, bOff+t] for j in [bRowOffset, bRowOffset+N)
        (TensorOperations.java:62-72; Panama/Native index semantics, PanamaTensorOperations.java:848)."""
        if row_chunk_size is None:
            row_chunk_size = b.rows
        if not isinstance(result, FloatBufferTensor):
            raise native.UnsupportedOperation(native.JL_ERR_UNSUPPORTED, "result must be F32")
        if a.rows != result.rows:
            raise ValueError("BAD M")  # PanamaTensorOperations.java:107
        if not (r_row_offset == 0 or r_row_offset >= b_row_offset):
            raise ValueError("Result offset must be >= b row offset")  # :108
        h = self.ctx.h
        # sparse operands store a slice of their logical shape: rebase every offset onto the stored extent exactly like
        # NativeSimdTensorOperations.java:96-107 (aOffset = at.getOffset(0, aColumnOffset), bOffset relative to b's stored
        # columns, rOffset = result.sparseColumnOffset - b.sparseRowOffset - rRowOffset, adjBRowOffset = bRowOffset - b.sparseRowOffset)
        if a.is_sparse() or b.is_sparse() or result.is_sparse():
            r_off = result.sparse_col_off - b.sparse_row_off - r_row_offset
            return self._batch_dot_product_raw(result, a, b, a_column_offset - a.sparse_col_off, b_column_offset - b.sparse_col_off,
                                               column_limit, r_off, b_row_offset - b.sparse_row_off, row_chunk_size)
        return self._batch_dot_product_raw(result, a, b, a_column_offset, b_column_offset, column_limit, -r_row_offset, b_row_offset,
                                           row_chunk_size)

    def _batch_dot_product_raw(self, result, a, b, a_column_offset, b_column_offset, column_limit, roffset, b_row_offset, row_chunk_size):
        """offsets relative to the STORED buffers; `roffset` is subtracted from the output column index (vector_simd.c:344)"""
        h = self.ctx.h
        if b.uid in self._registered:
            rc = self.lib.jl_gemm(h, a.dtype, ptr(a.data), ptr(a.scales), a_column_offset, a.cols,
                                  self._registered[b.uid], b_column_offset, ptr(result.data), roffset,
                                  a.rows, b_row_offset, row_chunk_size, column_limit, result.cols)
        else:
            if b.dtype in (Q4, I8):
                # unregistered quantised B: register on the fly (the reference would take its CPU delegate,
                # NativeGPUTensorOperations.java:321-334; there is none here)
                self.register_model_tensor(b)
                return self._batch_dot_product_raw(result, a, b, a_column_offset, b_column_offset, column_limit,
                                                   roffset, b_row_offset, row_chunk_size)
            if a.dtype == I8:
                raise native.UnsupportedOperation(native.JL_ERR_UNSUPPORTED, "I8 x %d" % b.dtype)
            rc = self.lib.jl_gemm_host(h, a.dtype, ptr(a.data), a_column_offset, a.cols, b.dtype, ptr(b.data),
                                       b_column_offset, b.cols, ptr(result.data), roffset, a.rows, b_row_offset,
                                       row_chunk_size, column_limit, result.cols)
        self.ctx.check(rc)

    def batch_dot_product_tensor_core(self, result, a, b, a_column_offset, b_column_offset, column_limit, r_row_offset=0,
                                      b_row_offset=0, row_chunk_size=None):
        """batchDotProduct on tcgen05 (BF16 operands, F32 accumulate) for prefill-shaped calls; B must be registered Q4."""
        if row_chunk_size is None:
            row_chunk_size = b.rows
        self.ctx.check(self.lib.jl_gemm_tc(self.ctx.h, a.dtype, ptr(a.data), a_column_offset, a.cols, self._registered[b.uid],
                                           b_column_offset, ptr(result.data), -r_row_offset, a.rows, b_row_offset, row_chunk_size,
                                           column_limit, result.cols))

    def dot_product_chunk(self, result, a, b, column_offset, column_limit, row_offset, row_chunk_size):
        # TensorOperations.java:74-84
        self.batch_dot_product(result, a, b, column_offset, column_offset, column_limit, 0, row_offset, row_chunk_size)

    def dot_product_batch_chunk(self, results, a, bs, offset, limit, chunk_start, chunk_size):
        # TensorOperations.java:86-99
        if len(results) != len(bs):
            raise ValueError("result.length == b.length")
        if all(b.uid in self._registered for b in bs) and all(isinstance(r, FloatBufferTensor) for r in results):
            n = len(bs)
            ids = (C.c_int64 * n)(*[self._registered[b.uid] for b in bs])
            rp = (C.c_void_p * n)(*[r.data.ctypes.data for r in results])
            self.ctx.check(self.lib.jl_gemm_batch(self.ctx.h, n, a.dtype, ptr(a.data), ptr(a.scales), offset, a.cols, ids,
                                                  offset, rp, 0, a.rows, chunk_start, chunk_size, limit, results[0].cols))
            return
        for r, b in zip(results, bs):
            self.dot_product_chunk(r, a, b, offset, limit, chunk_start, chunk_size)

    # -- element-wise ------------------------------------------------------------------------------
    def accumulate(self, a, b, offset, length):
        if not isinstance(a, FloatBufferTensor):
            raise native.UnsupportedOperation(native.JL_ERR_UNSUPPORTED, "accumulate: a must be F32")
        self.ctx.check(self.lib.jl_accumulate(self.ctx.h, ptr(a.data), a.rows, a.cols, b.dtype, ptr(b.data), ptr(b.scales),
                                              b.rows, b.cols, offset, length))

    def maccumulate(self, a, b, offset, length):
        if not (isinstance(a, FloatBufferTensor) and isinstance(b, FloatBufferTensor)):
            raise native.UnsupportedOperation(native.JL_ERR_UNSUPPORTED, "maccumulate: F32 only")
        self.ctx.check(self.lib.jl_maccumulate(self.ctx.h, ptr(a.data), a.rows, a.cols, ptr(b.data), b.rows, b.cols, offset,
                                               length))

    def saxpy(self, alpha, x, y, xoffset, yoffset, limit, a_offset=None, x_row_offset=None, batch_size=None):
        """Both overloads: saxpy(float alpha, x, y, ...) and the batched
        saxpy(AbstractTensor alpha, x, y, xoffset, yoffset, limit, aOffset, xRowOffset, batchSize)."""
        h = self.ctx.h
        if isinstance(alpha, AbstractTensor):
            if y.rows != 1:
                raise ValueError("y must have one row")  # TensorOperations.java:131
            self.ctx.check(self.lib.jl_saxpy_batch(h, ptr(alpha.data), ptr(x.data), x.cols, ptr(y.data), xoffset, yoffset,
                                                   limit, a_offset, x_row_offset, batch_size))
        else:
            if x.rows != 1 or y.rows != 1:
                raise ValueError("saxpy needs single-row tensors")  # NaiveTensorOperations.java:107
            al = np.array([alpha], dtype=np.float32)
            self.ctx.check(self.lib.jl_saxpy_batch(h, ptr(al), ptr(x.data), x.cols, ptr(y.data), xoffset, yoffset, limit,
                                                   0, 0, 1))

    def scale(self, factor, x, offset, length):
        self.ctx.check(self.lib.jl_scale(self.ctx.h, C.c_float(factor), ptr(x.data), x.rows, x.cols, offset, length))

    def quantize(self, t, qtype, offset, length):
        """TensorOperations.quantize (:145-149) with the Panama semantics (PanamaTensorOperations.java:1598-1622):
        same dtype -> returned as is; F32 -> I8 / BF16 on the GPU; caller owns the result."""
        if t.dtype == qtype:
            return t
        if not isinstance(t, FloatBufferTensor):
            raise native.UnsupportedOperation(native.JL_ERR_UNSUPPORTED, "quantize from dtype %d" % t.dtype)
        if qtype == I8:
            q = np.zeros((t.rows, t.cols), dtype=np.int8)
            s = np.zeros((t.rows, t.cols // 32), dtype=np.float32)
            self.ctx.check(self.lib.jl_quantize_q8(self.ctx.h, ptr(t.data), t.rows, t.cols, offset, length, ptr(q), ptr(s)))
            return Q8ByteBufferTensor(q, s)
        if qtype == BF16:
            out = np.zeros((t.rows, t.cols), dtype=np.uint16)
            self.ctx.check(self.lib.jl_quantize_bf16(self.ctx.h, ptr(t.data), t.rows, t.cols, offset, length, ptr(out)))
            return BFloat16BufferTensor(out)
        raise native.UnsupportedOperation(native.JL_ERR_UNSUPPORTED, "quantize to dtype %d" % qtype)

    # -- fused layer-level entry points (outside the reference interface) -------------
    def rmsnorm(self, x, weights, eps, embedding_length=None, weight_adjustment=0.0, offset=0, length=None):
        length = x.cols - offset if length is None else length
        out = np.zeros_like(x.data)
        self.ctx.check(self.lib.jl_rmsnorm(self.ctx.h, ptr(x.data), x.rows, x.cols, weights.dtype, ptr(weights.data),
                                           C.c_float(weight_adjustment), C.c_float(eps), embedding_length or x.cols, offset,
                                           length, ptr(out)))
        return FloatBufferTensor(out)

    def layernorm(self, x, weights, bias, eps, embedding_length=None, offset=0, length=None):
        """LayerNorm.forward (model/LayerNorm.java:41-67)"""
        length = x.cols - offset if length is None else length
        out = np.zeros_like(x.data)
        self.ctx.check(self.lib.jl_layernorm(self.ctx.h, ptr(x.data), x.rows, x.cols, weights.dtype, ptr(weights.data), bias.dtype,
                                             ptr(bias.data), C.c_float(eps), embedding_length or x.cols, offset, length, ptr(out)))
        return FloatBufferTensor(out)

    def activation(self, kind, x, offset=0, length=None):
        """ActivationFunction.eval in place: kind in {"silu", "gelu", "tanh"} (math/ActivationFunction.java:29-37)"""
        length = x.cols - offset if length is None else length
        code = {"silu": 0, "gelu": 1, "gelu_pytorch_tanh": 1, "tanh": 2}[kind]
        self.ctx.check(self.lib.jl_activation(self.ctx.h, code, ptr(x.data), x.rows, x.cols, offset, length))

    def softmax(self, x, offset, length):
        self.ctx.check(self.lib.jl_softmax(self.ctx.h, ptr(x.data), offset, length))

    def silu_mul(self, gate, up, offset, length):
        self.ctx.check(self.lib.jl_silu_mul(self.ctx.h, ptr(gate.data), ptr(up.data), gate.rows, gate.cols, offset, length))

    def quantize_q4_weights(self, x):
        x = np.ascontiguousarray(x, dtype=np.float32)
        q = np.zeros((x.shape[0], x.shape[1] // 2), dtype=np.uint8)
        s = np.zeros((x.shape[0], x.shape[1] // 32), dtype=np.float32)
        self.ctx.check(self.lib.jl_quantize_q4_weights(self.ctx.h, ptr(x), x.shape[0], x.shape[1], ptr(q), ptr(s)))
        return Q4ByteBufferTensor(q, s)

    def close(self):
        if self._own:
            self.ctx.close()
