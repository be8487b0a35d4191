/*
 * Reference-side binding for libjlama_b200.so (include/jlama_b200.h): a TensorOperations implementation that forwards
 * every method to the C ABI through the Java 22 Foreign Function & Memory API -- the same mechanism jlama-native uses
 * for libjlama / libjlamagpu (jlama-native/src/main/java22/.../NativeSimd.java:58-59).
 *
 * SOURCE ONLY: this image has no JDK, so the class is not compiled or tested here; INTEGRATION.md explains how a
 * maintainer wires it in.  The Python mirror jlama_b200/ops.py makes exactly the same calls and IS tested
 * (tests/test_gpu_ops.py).
 */
package com.github.tjake.jlama.tensor.operations;

import com.github.tjake.jlama.safetensors.DType;
import com.github.tjake.jlama.tensor.AbstractTensor;
import com.github.tjake.jlama.tensor.FloatBufferTensor;
import com.github.tjake.jlama.tensor.Q4ByteBufferTensor;
import com.github.tjake.jlama.tensor.Q8ByteBufferTensor;
import com.google.common.base.Preconditions;
import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import java.util.concurrent.ConcurrentHashMap;

import static java.lang.foreign.ValueLayout.*;

public final class CudaTensorOperations implements TensorOperations {
    private static final int JL_F32 = 0, JL_BF16 = 1, JL_Q4 = 2, JL_I8 = 3;
    private static final int JL_ERR_INVALID = -1, JL_ERR_UNSUPPORTED = -4;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB;
    static {
        System.loadLibrary("jlama_b200"); // or JarSupport.maybeLoadLibrary("jlama_b200") like NativeSimd
        LIB = SymbolLookup.loaderLookup().or(LINKER.defaultLookup());
    }

    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
    }

    private static final MethodHandle jl_init = h("jl_init", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle jl_last_error = h("jl_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle jl_register_tensor =
        h("jl_register_tensor", FunctionDescriptor.of(JAVA_LONG, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS));
    private static final MethodHandle jl_gemm = h("jl_gemm", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS,
        JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT));
    private static final MethodHandle jl_gemm_host = h("jl_gemm_host", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS,
        JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT));
    private static final MethodHandle jl_accumulate = h("jl_accumulate", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT,
        JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT));
    private static final MethodHandle jl_maccumulate = h("jl_maccumulate", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT,
        JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT));
    private static final MethodHandle jl_scale =
        h("jl_scale", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_FLOAT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT));
    private static final MethodHandle jl_saxpy_batch = h("jl_saxpy_batch", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS,
        JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT));
    private static final MethodHandle jl_quantize_q8 = h("jl_quantize_q8",
        FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));

    private final MemorySegment ctx;
    /** AbstractTensor.getUid() -> device tensor id (NativeGPUTensorOperations.java:104-151 keeps the same map). */
    private final ConcurrentHashMap<String, Long> registered = new ConcurrentHashMap<>();

    /** Public no-arg constructor: TensorOperationsProvider instantiates back-ends reflectively and expects a throw when
     *  the device or the library is unusable (TensorOperationsProvider.java:57-73). */
    public CudaTensorOperations() {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ADDRESS), info = a.allocate(JAVA_LONG, 4);
            int rc = (int) jl_init.invokeExact(0, out, info);
            if (rc != 0) throw new UnsupportedOperationException("jl_init failed: " + rc + " (no B200 / CUDA driver?)");
            this.ctx = out.get(ADDRESS, 0);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    @Override public String name() { return "CUDA sm_100a Operations"; }
    @Override public int parallelSplitSize() { return 1; } // one call per GEMM, no host-side chunking (NativeGPUTensorOperations.java:99-101)
    @Override public DType preferredWorkingQuantizedType() { return DType.I8; }

    private void check(int rc) {
        if (rc == 0) return;
        String msg;
        try {
            msg = ((MemorySegment) jl_last_error.invokeExact(ctx)).reinterpret(4096).getString(0);
        } catch (Throwable t) {
            msg = "error " + rc;
        }
        if (rc == JL_ERR_UNSUPPORTED) throw new UnsupportedOperationException(msg);
        if (rc == JL_ERR_INVALID) throw new IllegalArgumentException(msg);
        throw new RuntimeException(msg);
    }

    private static int code(DType t) {
        return switch (t) { case F32 -> JL_F32; case BF16 -> JL_BF16; case Q4 -> JL_Q4; case I8 -> JL_I8;
                            default -> throw new UnsupportedOperationException(t.name()); };
    }
    private static MemorySegment scales(AbstractTensor t) {
        if (t instanceof Q4ByteBufferTensor q) return q.getBlockF().getMemorySegment();
        if (t instanceof Q8ByteBufferTensor q) return q.getBlockF().getMemorySegment();
        return MemorySegment.NULL;
    }

    @Override public void registerModelTensor(AbstractTensor t) {
        registered.computeIfAbsent(t.getUid(), k -> {
            try {
                // a sparse (row- or column-sliced) weight owns only its slice: register the stored extent
                // (TensorShape.java:119-133), offsets are rebased per call below
                long id = (long) jl_register_tensor.invokeExact(ctx, code(t.dType()), (long) t.shape().sparseRowLength(),
                                                               (long) t.shape().sparseColumnLength(), t.getMemorySegment(), scales(t));
                if (id < 0) throw new OutOfMemoryError("jl_register_tensor"); // no CPU delegate: fail loudly
                return id;
            } catch (Error | RuntimeException e) { throw e; } catch (Throwable e) { throw new RuntimeException(e); }
        });
    }

    @Override public void batchDotProduct(AbstractTensor result, AbstractTensor a, AbstractTensor b, int aColumnOffset,
                                          int bColumnOffset, int columnLimit, int rRowOffset, int bRowOffset, int rowChunkSize) {
        Preconditions.checkArgument(a.shape().first() == result.shape().first(), "BAD M");            // PanamaTensorOperations.java:107
        Preconditions.checkArgument(rRowOffset == 0 || rRowOffset >= bRowOffset, "Result offset must be >= b row offset"); // :108
        if (!(result instanceof FloatBufferTensor)) throw new UnsupportedOperationException("result must be F32");
        try {
            Long id = registered.get(b.getUid());
            int rc;
            if (id == null && (b.dType() == DType.Q4 || b.dType() == DType.I8)) { registerModelTensor(b); id = registered.get(b.getUid()); }
            // Offsets rebased onto the stored (sparse) extents exactly like NativeSimdTensorOperations.java:96-107:
            //   aOffset = at.getOffset(0, aColumnOffset), bOffset = bt.getOffset(bt.sparseRowOffset, bColumnOffset),
            //   rOffset = result.sparseColumnOffset - bt.sparseRowOffset - rRowOffset, adjBRowOffset = bRowOffset - bt.sparseRowOffset
            final int aOff = a.getOffset(0, aColumnOffset);
            final int bOff = bColumnOffset - b.shape().sparseColumnOffset();
            final int rOff = result.shape().sparseColumnOffset() - b.shape().sparseRowOffset() - rRowOffset;
            final int bRow = bRowOffset - b.shape().sparseRowOffset();
            if (id != null) {
                rc = (int) jl_gemm.invokeExact(ctx, code(a.dType()), a.getMemorySegment(), scales(a), aOff, a.getStride(),
                                               (long) id, bOff, result.getMemorySegment(), rOff, a.shape().first(),
                                               bRow, rowChunkSize, columnLimit, result.getStride());
            } else {
                rc = (int) jl_gemm_host.invokeExact(ctx, code(a.dType()), a.getMemorySegment(), aOff, a.getStride(),
                                                    code(b.dType()), b.getMemorySegment(), bOff, b.getStride(),
                                                    result.getMemorySegment(), rOff, a.shape().first(), bRow,
                                                    rowChunkSize, columnLimit, result.getStride());
            }
            check(rc);
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    @Override public void accumulate(AbstractTensor a, AbstractTensor b, int offset, int length) {
        try {
            check((int) jl_accumulate.invokeExact(ctx, a.getMemorySegment(), a.shape().first(), a.getStride(), code(b.dType()),
                                                  b.getMemorySegment(), scales(b), b.shape().first(), b.getStride(), offset, length));
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    @Override public void maccumulate(AbstractTensor a, AbstractTensor b, int offset, int length) {
        try {
            check((int) jl_maccumulate.invokeExact(ctx, a.getMemorySegment(), a.shape().first(), a.getStride(), b.getMemorySegment(),
                                                   b.shape().first(), b.getStride(), offset, length));
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    @Override public void saxpy(float alpha, AbstractTensor x, AbstractTensor y, int xoffset, int yoffset, int limit) {
        try (Arena ar = Arena.ofConfined()) {
            MemorySegment al = ar.allocate(JAVA_FLOAT); al.set(JAVA_FLOAT, 0, alpha);
            check((int) jl_saxpy_batch.invokeExact(ctx, al, x.getMemorySegment(), x.getStride(), y.getMemorySegment(), xoffset, yoffset,
                                                   limit, 0, 0, 1));
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    @Override public void saxpy(AbstractTensor alpha, AbstractTensor x, AbstractTensor y, int xoffset, int yoffset, int limit,
                                int aOffset, int xRowOffset, int batchSize) {
        try {
            check((int) jl_saxpy_batch.invokeExact(ctx, alpha.getMemorySegment(), x.getMemorySegment(), x.getStride(),
                                                   y.getMemorySegment(), xoffset, yoffset, limit, aOffset, xRowOffset, batchSize));
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    @Override public void scale(float factor, AbstractTensor x, int offset, int length) {
        try {
            check((int) jl_scale.invokeExact(ctx, factor, x.getMemorySegment(), x.shape().first(), x.getStride(), offset, length));
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    @Override public AbstractTensor quantize(AbstractTensor t, DType qtype, int offset, int length) {
        if (t.dType() == qtype) return t;                               // PanamaTensorOperations.java:1598-1622
        if (qtype != DType.I8 || t.dType() != DType.F32) return TensorOperations.super.quantize(t, qtype, offset, length);
        Q8ByteBufferTensor q = new Q8ByteBufferTensor(t.shape());       // caller closes (PanamaTensorOperations.java:1686-1687)
        try {
            check((int) jl_quantize_q8.invokeExact(ctx, t.getMemorySegment(), t.shape().first(), t.getStride(), offset, length,
                                                   q.getMemorySegment(), q.getBlockF().getMemorySegment()));
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable th) { throw new RuntimeException(th); }
        return q;
    }
}
