/*
 * Reference-side binding for the MODEL-LEVEL entry points of libjlama_b200.so (include/jlama_b200.h, jl_model_*): a LlamaModel whose
 * layers live on the GPU.  It keeps AbstractModel's public call shape -- batchForward / sample / generate
 * (core/model/AbstractModel.java:295-329,443-491,516-646) -- and replaces the per-layer Java objects that LlamaModel builds in
 * loadTransformerBlockWeights (core/model/llama/LlamaModel.java:117-143: RMSNorm, CausalSelfAttention, MLPBlock per layer) by one
 * device-resident model: weights are registered once, KV pages, activations and the sampled token stay in HBM, and a decode step is
 * one kernel launch.  Row (f2) of SURVEY.md section 8.
 *
 * SOURCE ONLY: this image has no JDK, so the class is not compiled or tested here.  jlama_b200/model.py (LlamaModel) makes exactly the
 * same calls in the same order and IS tested (tests/test_gpu_model.py, tests/test_gpu_checkpoint.py); INTEGRATION.md section 3
 * shows where a maintainer plugs this class in.
 */
package com.github.tjake.jlama.model;

import com.github.tjake.jlama.safetensors.Config;
import com.github.tjake.jlama.safetensors.DType;
import com.github.tjake.jlama.safetensors.WeightLoader;
import com.github.tjake.jlama.tensor.AbstractTensor;
import com.github.tjake.jlama.tensor.Q4ByteBufferTensor;
import com.github.tjake.jlama.tensor.Q8ByteBufferTensor;
import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import java.util.ArrayList;
import java.util.List;

import static java.lang.foreign.ValueLayout.*;

public final class CudaLlamaModel implements AutoCloseable {
    // include/jlama_b200.h
    private static final int JL_F32 = 0, JL_BF16 = 1, JL_Q4 = 2, JL_I8 = 3;
    private static final int T_EMBED = 0, T_OUT_NORM = 1, T_LM_HEAD = 2;
    private static final int L_ATTN_NORM = 0, L_Q = 1, L_K = 2, L_V = 3, L_O = 4, L_FFN_NORM = 5, L_GATE = 6, L_DOWN = 7, L_UP = 8;

    /** jl_model_config, field for field (all 4-byte except the two doubles; natural alignment like the C struct). */
    private static final StructLayout CONFIG = MemoryLayout.structLayout(
        JAVA_INT.withName("context_length"), JAVA_INT.withName("embedding_length"), JAVA_INT.withName("hidden_length"),
        JAVA_INT.withName("num_heads"), JAVA_INT.withName("num_kv_heads"), JAVA_INT.withName("num_layers"),
        JAVA_INT.withName("vocab_size"), JAVA_INT.withName("head_size"), JAVA_FLOAT.withName("layer_norm_eps"),
        MemoryLayout.paddingLayout(4), JAVA_DOUBLE.withName("rope_theta"), JAVA_DOUBLE.withName("rope_scaling"),
        JAVA_INT.withName("working_qtype"), JAVA_INT.withName("kv_dtype"), JAVA_INT.withName("max_batch"),
        JAVA_INT.withName("max_sessions"), JAVA_INT.withName("max_context"), JAVA_INT.withName("tp_rank"), JAVA_INT.withName("tp_size"),
        JAVA_INT.withName("prefill_tensor_core"), JAVA_INT.withName("flags"), JAVA_INT.withName("num_experts"),
        JAVA_INT.withName("experts_per_token"), JAVA_INT.withName("arch") /* 0 = Llama / Mixtral blocks, 1 = GPT-2 blocks */);

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB;
    static {
        System.loadLibrary("jlama_b200");
        LIB = SymbolLookup.loaderLookup().or(LINKER.defaultLookup());
    }
    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
    }
    private static final MethodHandle jl_init = h("jl_init", FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle jl_last_error = h("jl_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle jl_register_tensor =
        h("jl_register_tensor", FunctionDescriptor.of(JAVA_LONG, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS, ADDRESS));
    private static final MethodHandle jl_model_create = h("jl_model_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle jl_model_set_tensor =
        h("jl_model_set_tensor", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_LONG));
    private static final MethodHandle jl_model_set_expert_tensor =
        h("jl_model_set_expert_tensor", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_LONG));
    private static final MethodHandle jl_model_finalize = h("jl_model_finalize", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle jl_model_free = h("jl_model_free", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle jl_model_reset_session = h("jl_model_reset_session", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    private static final MethodHandle jl_model_batch_forward =
        h("jl_model_batch_forward", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));
    private static final MethodHandle jl_model_sample =
        h("jl_model_sample", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_FLOAT, JAVA_FLOAT, ADDRESS, ADDRESS));
    private static final MethodHandle jl_model_decode =
        h("jl_model_decode", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle jl_model_generate =
        h("jl_model_generate", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle jl_model_kv_save = h("jl_model_kv_save", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle jl_model_kv_load = h("jl_model_kv_load", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));

    private final Arena arena = Arena.ofShared();
    private final MemorySegment ctx, model;
    private final Config c;
    private final List<Long> tensorIds = new ArrayList<>();

    /**
     * Mirrors LlamaModel's constructor chain (llama/LlamaModel.java:68-115): config + weights in, a model ready for batchForward out.
     * `weights` is the same WeightLoader the reference hands to loadInputWeights / loadTransformerBlockWeights / loadOutputWeights;
     * the DistributedContext in `c.dctx()` selects this rank's rows / columns exactly as Weights.getLoadOffsets (:99-117) does.
     */
    public CudaLlamaModel(Config c, WeightLoader weights, DType workingQType, DType kvDType, int maxSessions, int tpRank, int tpSize) {
        this.c = c;
        try {
            MemorySegment out = arena.allocate(ADDRESS), info = arena.allocate(JAVA_LONG, 4);
            check((int) jl_init.invokeExact(tpRank /* one GPU per process */, out, info), MemorySegment.NULL);
            this.ctx = out.get(ADDRESS, 0);
            MemorySegment cfg = arena.allocate(CONFIG);
            cfg.set(JAVA_INT, off("context_length"), c.contextLength);
            cfg.set(JAVA_INT, off("embedding_length"), c.embeddingLength);
            cfg.set(JAVA_INT, off("hidden_length"), c.hiddenLength);
            cfg.set(JAVA_INT, off("num_heads"), c.numberOfHeads);
            cfg.set(JAVA_INT, off("num_kv_heads"), c.numberOfKeyValueHeads);
            cfg.set(JAVA_INT, off("num_layers"), c.numberOfLayers);
            cfg.set(JAVA_INT, off("vocab_size"), c.vocabularySize);
            cfg.set(JAVA_INT, off("head_size"), c.headSize);
            cfg.set(JAVA_FLOAT, off("layer_norm_eps"), c.layerNormEps);
            cfg.set(JAVA_DOUBLE, off("rope_theta"), c.ropeFreqsTheta);
            cfg.set(JAVA_DOUBLE, off("rope_scaling"), c.ropeScalingFactor == null ? 1.0 : c.ropeScalingFactor);
            cfg.set(JAVA_INT, off("working_qtype"), code(workingQType));
            cfg.set(JAVA_INT, off("kv_dtype"), code(kvDType));
            cfg.set(JAVA_INT, off("max_batch"), 256);               // jlama.max_batch_size (AbstractModel.java:57)
            cfg.set(JAVA_INT, off("max_sessions"), maxSessions);
            cfg.set(JAVA_INT, off("tp_rank"), tpRank);
            cfg.set(JAVA_INT, off("tp_size"), tpSize);
            MemorySegment mo = arena.allocate(ADDRESS);
            check((int) jl_model_create.invokeExact(ctx, cfg, mo), ctx);
            this.model = mo.get(ADDRESS, 0);

            // loadInputWeights / loadOutputWeights (LlamaModel.java:88-97,145-156)
            bind(-1, T_EMBED, weights.load("model.embed_tokens.weight"));
            bind(-1, T_OUT_NORM, weights.load("model.norm.weight"));
            if (weights.isWeightPresent("lm_head.weight")) bind(-1, T_LM_HEAD, weights.load("lm_head.weight")); // else tied to the embedding
            // loadTransformerBlockWeights (:117-143): the same names, the same row / column splits
            for (int i = c.dctx().layerStart; i < c.dctx().layerEnd; i++) {
                String b = "model.layers." + i + ".";
                bind(i, L_ATTN_NORM, weights.load(b + "input_layernorm.weight"));
                bind(i, L_Q, weights.load(b + "self_attn.q_proj.weight", c.dctx(), true, false));
                bind(i, L_K, weights.load(b + "self_attn.k_proj.weight", c.dctx(), true, false));
                bind(i, L_V, weights.load(b + "self_attn.v_proj.weight", c.dctx(), true, false));
                bind(i, L_O, weights.load(b + "self_attn.o_proj.weight", c.dctx(), false, true));
                bind(i, L_FFN_NORM, weights.load(b + "post_attention_layernorm.weight"));
                bind(i, L_GATE, weights.load(b + "mlp.gate_proj.weight", c.dctx(), true, false));
                bind(i, L_DOWN, weights.load(b + "mlp.down_proj.weight", c.dctx(), false, true));
                bind(i, L_UP, weights.load(b + "mlp.up_proj.weight", c.dctx(), true, false));
            }
            check((int) jl_model_finalize.invokeExact(model), ctx);
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    private static long off(String f) { return CONFIG.byteOffset(MemoryLayout.PathElement.groupElement(f)); }
    private static int code(DType t) {
        return switch (t) { case F32 -> JL_F32; case BF16 -> JL_BF16; case Q4 -> JL_Q4; case I8 -> JL_I8;
                            default -> throw new UnsupportedOperationException(t.name()); };
    }
    private static MemorySegment scales(AbstractTensor t) {
        if (t instanceof Q4ByteBufferTensor q) return q.getBlockF().getMemorySegment();
        if (t instanceof Q8ByteBufferTensor q) return q.getBlockF().getMemorySegment();
        return MemorySegment.NULL;
    }
    private void check(int rc, MemorySegment ctxOrNull) throws Throwable {
        if (rc >= 0) return;
        String msg = ctxOrNull.equals(MemorySegment.NULL) ? "" : ((MemorySegment) jl_last_error.invokeExact(ctxOrNull)).reinterpret(1024).getString(0);
        if (rc == -4) throw new UnsupportedOperationException(msg);
        if (rc == -1) throw new IllegalArgumentException(msg);
        throw new RuntimeException("libjlama_b200: " + rc + " " + msg);
    }
    /** registerModelTensor + jl_model_set_tensor: the stored (sparse) extent of the tensor is what this rank holds. */
    private void bind(int layer, int slot, AbstractTensor t) throws Throwable {
        long id = (long) jl_register_tensor.invokeExact(ctx, code(t.dType()), (long) t.shape().sparseRowLength(),
                                                       (long) t.shape().sparseColumnLength(), t.getMemorySegment(), scales(t));
        if (id < 0) throw new OutOfMemoryError("jl_register_tensor");
        tensorIds.add(id);
        check((int) jl_model_set_tensor.invokeExact(model, layer, slot, id), ctx);
    }

    /** AbstractModel.batchForward(int[] tokens, int startPos, KvBuffer kv) (:295-312): the session index stands for the KvBuffer. */
    public void batchForward(int session, int[] tokens, int startPos) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment tk = a.allocateFrom(JAVA_INT, tokens);
            check((int) jl_model_batch_forward.invokeExact(model, session, tk, tokens.length, startPos), ctx);
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** AbstractModel.sample (:443-491) on the last forwarded row: arg-max at temperature 0, else the reference's prefix-sum rule. */
    public int sample(int session, float temperature, float uniform, float[] logitsOrNull) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment tok = a.allocate(JAVA_INT);
            MemorySegment lg = logitsOrNull == null ? MemorySegment.NULL : a.allocate(JAVA_FLOAT, logitsOrNull.length);
            check((int) jl_model_sample.invokeExact(model, session, temperature, uniform, tok, lg), ctx);
            if (logitsOrNull != null) MemorySegment.copy(lg, JAVA_FLOAT, 0, logitsOrNull, 0, logitsOrNull.length);
            return tok.get(JAVA_INT, 0);
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** forward(token, position) + sample for n concurrent sessions in one launch (the reference's batch: KvBufferCache.java:58-60). */
    public int[] decode(int[] sessions, int[] tokens, int[] positions) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment s = a.allocateFrom(JAVA_INT, sessions), t = a.allocateFrom(JAVA_INT, tokens), p = a.allocateFrom(JAVA_INT, positions);
            MemorySegment nxt = a.allocate(JAVA_INT, tokens.length);
            check((int) jl_model_decode.invokeExact(model, tokens.length, s, t, p, nxt, MemorySegment.NULL), ctx);
            return nxt.toArray(JAVA_INT);
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** AbstractModel.generate (:516-646) at temperature 0 over token ids; timingsMs = {prompt, generate} like Generator.Response. */
    public int[] generate(int session, int[] prompt, int nNew, double[] timingsMs) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment pr = a.allocateFrom(JAVA_INT, prompt), out = a.allocate(JAVA_INT, nNew), tm = a.allocate(JAVA_DOUBLE, 2);
            check((int) jl_model_generate.invokeExact(model, session, pr, prompt.length, nNew, out, MemorySegment.NULL, tm), ctx);
            if (timingsMs != null) { timingsMs[0] = tm.getAtIndex(JAVA_DOUBLE, 0); timingsMs[1] = tm.getAtIndex(JAVA_DOUBLE, 1); }
            return out.toArray(JAVA_INT);
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** KvBufferCache.KvBufferPage persistence (KvBufferCache.java:121-176): <dir>/<session>-L<l>C<c>.page, the reference's own files. */
    public int saveKv(int session, String workingDirectory, java.util.UUID sessionId) {
        try (Arena a = Arena.ofConfined()) {
            int n = (int) jl_model_kv_save.invokeExact(model, session, a.allocateFrom(workingDirectory), a.allocateFrom(sessionId.toString()));
            check(n, ctx);
            return n;
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }
    public int loadKv(int session, String workingDirectory, java.util.UUID sessionId) {
        try (Arena a = Arena.ofConfined()) {
            int n = (int) jl_model_kv_load.invokeExact(model, session, a.allocateFrom(workingDirectory), a.allocateFrom(sessionId.toString()));
            check(n, ctx);
            return n;
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** jl_model* for the classes of this package that take the model as an argument (CudaSessionScheduler). */
    MemorySegment handle() { return model; }

    public void resetSession(int session) {
        try { check((int) jl_model_reset_session.invokeExact(model, session), ctx); }
        catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    @Override public void close() {
        try { int rc = (int) jl_model_free.invokeExact(model); } catch (Throwable ignored) { }
        arena.close();
    }
}
