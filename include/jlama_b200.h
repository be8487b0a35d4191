/*
 * jlama_b200.h -- C ABI of libjlama_b200.so, the Blackwell (sm_100a) back-end for
 * tjake/Jlama's quantized forward pass.
 *
 * This is the drop-in boundary: plain C types only (pointers, ints, sizes), the
 * conventions jextract/Panama-FFI binds (the reference's existing natives:
 * jlama-native/src/main/c/simd/vector_simd.h:22-39 and
 * jlama-native/src/main/c/gpu/vector_gpu.h:7-19).  Each entry point cites the
 * reference interface it replaces.  Reference paths are relative to
 * /root/reference; "core/" = jlama-core/src/main/java/com/github/tjake/jlama/,
 * "native/java/" = jlama-native/src/main/java/com/github/tjake/jlama/tensor/operations/.
 *
 * Conventions
 *  - every function returns int status (JL_OK or a negative JL_ERR_*) or an
 *    int64 handle with -1 = failure (vector_gpu.h:10-15 convention).  Nothing
 *    aborts the process (the reference GPU natives call exit(),
 *    vector_gpu.c:96,116 -- deliberately not reproduced).
 *  - there is no CPU fallback: with no usable sm_100 device jl_init fails.
 *  - thread-safe: calls on one jl_ctx are serialised internally; the entry points
 *    of one jl_model serialise on the model's own lock and run on its own stream
 *    (request threads may call jl_model_generate concurrently, one session each).
 *  - dtype codes: Jlama DType (core/safetensors/DType.java) subset.
 */
#ifndef JLAMA_B200_H
#define JLAMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JL_OK 0
#define JL_ERR_INVALID (-1)     /* IllegalArgumentException analogue (Guava Preconditions in the reference) */
#define JL_ERR_CUDA (-2)        /* CUDA runtime error; see jl_last_error */
#define JL_ERR_OOM (-3)         /* device memory exhausted (NativeGPUTensorOperations.java:145-150 "limitReached") */
#define JL_ERR_UNSUPPORTED (-4) /* UnsupportedOperationException analogue (TestOperations.java:140 relies on it) */
#define JL_ERR_NCCL (-5)

/* dtypes */
#define JL_F32 0
#define JL_BF16 1
#define JL_Q4 2 /* core/tensor/Q4ByteBufferTensor.java:34-259: 32-element blocks, byte j = q[j] | q[j+16]<<4, separate f32 scales [rows, cols/32] */
#define JL_I8 3 /* core/tensor/Q8ByteBufferTensor.java:37-223: int8 + f32 scale per 32 */

typedef struct jl_ctx jl_ctx;
typedef struct jl_model jl_model;

/* ---- lifecycle: replaces init_gpu (vector_gpu.h:7, vector_gpu.c:227-239) -------------- */
/* info[0]=free bytes, [1]=total bytes, [2]=SM count, [3]=compute capability*10 (100 for B200). */
int jl_init(int device, jl_ctx **out, int64_t *info /* nullable, int64[4] */);
int jl_shutdown(jl_ctx *ctx);
/* never NULL.  errno style: the text of the calling thread's most recent failing call (any entry point of this library), or, if this
 * thread has not failed yet, the most recent failure of any thread on `ctx`.  Valid until the calling thread's next failing call. */
const char *jl_last_error(jl_ctx *ctx);
const char *jl_version(void);
int jl_sync(jl_ctx *ctx);
/* number of kernels this library launched since jl_init (bench.py's gpu_launches claim) */
int64_t jl_kernel_launches(jl_ctx *ctx);

/* kernel timeline: every GEMV / decode-attention launch made while tracing is on gets a slot of 16 words
 * {first CTA start ns, last CTA end ns, CTA 0 after its prologue ns, tag, CTA 0 start ns, 11 stage stamps}; tools/ktrace.py */
int jl_debug_ktrace(jl_ctx *ctx, int capacity);
int jl_debug_ktrace_clear(jl_ctx *ctx);
int jl_debug_ktrace_read(jl_ctx *ctx, uint64_t *out, int max_slots);
/* diagnostic: time the quantised GEMV kernel on device-resident operands.  `b_id` is a registered weight
 * [rows, k]; each of `iters` launches handles `n` consecutive rows starting at a rotating offset so that the
 * stream of weights is larger than L2; activations [m, k] f32 live in HBM.  mode: 0 = Q8-quantising prologue +
 * store, 1 = RMSNorm + Q8 prologue + store, 2 = f32 activations + store, 3 = Q8 prologue + residual epilogue.
 * Returns the average microseconds per launch measured with CUDA events around the whole sequence. */
int jl_debug_gemv_bench(jl_ctx *ctx, int64_t b_id, int n, int m, int mode, int iters, int use_pdl, double *avg_us);

/* diagnostic: average microseconds of one tcgen05 prefill GEMM launch C[t, rows] = A_bf16[t, k] * W^T on
 * device-resident operands (W = registered Q4 tensor [rows, k]). */
int jl_debug_gemm_tc_bench(jl_ctx *ctx, int64_t b_id, int t, int iters, double *avg_us);

/* ---- TensorOperations.registerModelTensor (core/tensor/operations/TensorOperations.java:39;
 *      NativeGPUTensorOperations.java:104-151 -> register_tensor, vector_gpu.h:10) ------------
 * Copies a weight (and its Q4/I8 block scales) to HBM once; returns a tensor id, -1 on failure.
 * `rows` x `cols` is the logical shape; Q4 data is rows*cols/2 bytes, scales rows*cols/32 floats.
 * Unlike the reference (int size => < 2 GiB) sizes are 64-bit. */
int64_t jl_register_tensor(jl_ctx *ctx, int dtype, int64_t rows, int64_t cols, const void *data, const float *scales);
int jl_unregister_tensor(jl_ctx *ctx, int64_t id);

/* ---- TensorOperations.batchDotProduct (TensorOperations.java:62-72) with a registered B ----
 * Drop-in for gpu_gemm (vector_gpu.h:17) / gemm_q8_q4, gemm_f32_q4, gemm_f32, gemm_f32_bf16
 * (vector_simd.h:22-39).  A and R are HOST buffers (the Java MemorySegments):
 *   r[ldc*i + j - roffset] = sum_{t<k} A[i, a_col_off+t] * B[j, b_col_off+t],  i in [0,m), j in [n0,n0+n)
 * a_dtype in {JL_F32, JL_BF16, JL_I8}; for JL_I8 `a_scales` are the Q8 block scales [m, lda/32]
 * (produced by jl_quantize_q8 or by Jlama itself).  lda/ldc in elements.
 * Offsets follow NativeSimdTensorOperations.java:96-107 (roffset is subtracted from the output index,
 * vector_simd.c:344). */
int jl_gemm(jl_ctx *ctx, int a_dtype, const void *a, const float *a_scales, int a_col_off, int lda, int64_t b_id,
            int b_col_off, float *r, int roffset, int m, int n0, int n, int k, int ldc);
/* dotProductBatchChunk (TensorOperations.java:86-99; gemm_*_batch, vector_simd.h:24-39): one A, several B/R. */
int jl_gemm_batch(jl_ctx *ctx, int batch_num, int a_dtype, const void *a, const float *a_scales, int a_col_off, int lda,
                  const int64_t *b_ids, int b_col_off, float *const *r, int roffset, int m, int n0, int n, int k,
                  int ldc);
/* batchDotProduct on the tcgen05 tensor cores for m >= 16 rows of F32/BF16 activations against a registered Q4 or Q8_0 (JL_I8)
 * weight (prefill shape).  Same index semantics as jl_gemm; operands are rounded to BF16 (accumulation in F32), so
 * the result matches the reference's F32 x Q4 GEMM (PanamaTensorOperations.java:289-548) to ~1e-3 of max|C|, not bit-wise.
 * Requires k %% 64 == 0, n %% 128 == 0, n0 %% 128 == 0, b_col_off %% 64 == 0 (JL_ERR_UNSUPPORTED otherwise). */
int jl_gemm_tc(jl_ctx *ctx, int a_dtype, const void *a, int a_col_off, int lda, int64_t b_id, int b_col_off, float *r,
               int roffset, int m, int n0, int n, int k, int ldc);
/* batchDotProduct where B is a HOST tensor too (attention scores against a KV page held by Java,
 * CausalSelfAttention.java:324-330).  b_dtype in {JL_F32, JL_BF16}. */
int jl_gemm_host(jl_ctx *ctx, int a_dtype, const void *a, int a_col_off, int lda, int b_dtype, const void *b,
                 int b_col_off, int ldb, float *r, int roffset, int m, int n0, int n, int k, int ldc);

/* ---- remaining TensorOperations methods on HOST f32 buffers (TensorOperations.java:101-149) ----
 * All run as CUDA kernels (upload, kernel, read back); there is no CPU path. */
/* a[r, off:off+len] += b[(r or 0), off:off+len]; b_dtype in {JL_F32, JL_BF16, JL_Q4 (+b_scales)}
 * (PanamaTensorOperations.java:2151-2325) */
int jl_accumulate(jl_ctx *ctx, float *a, int a_rows, int lda, int b_dtype, const void *b, const float *b_scales,
                  int b_rows, int ldb, int offset, int length);
/* a *= b (PanamaTensorOperations.java:2062-2097) */
int jl_maccumulate(jl_ctx *ctx, float *a, int a_rows, int lda, const float *b, int b_rows, int ldb, int offset,
                   int length);
/* x[r, off:off+len] *= factor (PanamaTensorOperations.java:2474-2514) */
int jl_scale(jl_ctx *ctx, float factor, float *x, int rows, int ldx, int offset, int length);
/* y[yoff+i] += alpha * x[xoff+i] (PanamaTensorOperations.java:2567-2611) */
int jl_saxpy(jl_ctx *ctx, float alpha, const float *x, float *y, int xoffset, int yoffset, int limit);
/* y[0, yoff+i] += sum_r alpha[aoff+r] * x[xrow+r, xoff+i]  (TensorOperations.java:122-137,
 * PanamaTensorOperations.java:2614-2698) */
int jl_saxpy_batch(jl_ctx *ctx, const float *alpha, const float *x, int ldx, float *y, int xoffset, int yoffset,
                   int limit, int a_offset, int x_row_offset, int batch);
/* quantize(t, I8, offset, length) (TensorOperations.java:145-149; PanamaTensorOperations.java:1684-1774):
 * d = max/127, q = (byte)(x*(127/max) + 0.5f) truncating; writes q[rows, ldx] (only [offset, offset+length))
 * and scales[rows, ldx/32]. */
int jl_quantize_q8(jl_ctx *ctx, const float *x, int rows, int ldx, int offset, int length, int8_t *q, float *scales);
/* quantize(t, BF16, ...) : FloatConversions.float32ToBFloat16 RNE (core/math/FloatConversions.java:35-61) */
int jl_quantize_bf16(jl_ctx *ctx, const float *x, int rows, int ldx, int offset, int length, uint16_t *out);
/* AbstractTensor.quantize(Q4) weight quantiser on the GPU (Q4ByteBufferTensor.java:66-120);
 * byte-identical to the reference's files. */
int jl_quantize_q4_weights(jl_ctx *ctx, const float *x, int64_t rows, int64_t cols, uint8_t *q, float *scales);
/* AbstractTensor.quantize(I8) weight quantiser on the GPU (Q8ByteBufferTensor.java:47-90): iscale = 127/max,
 * q = (byte)Math.round(x * iscale) -- round half up, unlike the truncating activation quantiser above. */
int jl_quantize_q8_weights(jl_ctx *ctx, const float *x, int64_t rows, int64_t cols, int8_t *q, float *scales);

/* ---- layer-level fused entry points on HOST buffers --------------------------------------
 * The reference does these as scalar Java loops outside TensorOperations (SURVEY 7 "hard parts");
 * a device-resident layer subclass calls them instead. */
/* RMSNorm.forward (core/model/RMSNorm.java:34-56) */
int jl_rmsnorm(jl_ctx *ctx, const float *x, int rows, int ldx, int w_dtype, const void *w, float weight_adjustment,
               float eps, int embedding_length, int offset, int length, float *out);
/* LayerNorm.forward (core/model/LayerNorm.java:41-67; GPT-2 family): statistics over [offset, offset+length) divided by
 * embedding_length, out = (x - mean) * invStddev * w + bias.  Weights / bias F32 or BF16. */
int jl_layernorm(jl_ctx *ctx, const float *x, int rows, int ldx, int w_dtype, const void *w, int b_dtype, const void *bias, float eps,
                 int embedding_length, int offset, int length, float *out);
/* ActivationFunction.eval in place (core/math/ActivationFunction.java:29-37): type 0 SILU, 1 GELU (tanh form, also
 * GELU_PYTORCH_TANH), 2 TANH; double-precision math cast to float like the reference. */
#define JL_ACT_SILU 0
#define JL_ACT_GELU 1
#define JL_ACT_TANH 2
int jl_activation(jl_ctx *ctx, int type, float *x, int rows, int ld, int offset, int length);
/* VectorMath.softMax (core/math/VectorMath.java:69-90) */
int jl_softmax(jl_ctx *ctx, float *x, int offset, int length);
/* MLPBlock activation loop + maccumulate (core/model/MLPBlock.java:132-141): gate = silu(gate) * up */
int jl_silu_mul(jl_ctx *ctx, float *gate, const float *up, int rows, int ld, int offset, int length);
/* VectorMath.precomputeFreqsCis (core/math/VectorMath.java:148-165): out[end*dim/2][2] (cos,sin).
 * Host-side double libm, identical recipe to the reference. */
int jl_precompute_freqs_cis(int dim, int end, double theta, double scaling_factor, float *out);

/* ---- DistributedContext / KvBufferCache host logic (pure functions) -----------------------
 * core/model/DistributedContext.java:60-98 */
typedef struct {
    int embeddingSegmentStart, embeddingSegmentLength;
    int attentionSegmentStart, attentionSegmentLength;
    int hiddenSegmentStart, hiddenSegmentLength;
    int kvSegmentStart, kvSegmentLength;
    int headStart, headEnd, groupHeadStart, groupHeadEnd;
    int numberOfLayers, layerStart, layerEnd;
} jl_dctx;
int jl_dctx_build(int embedding_length, int attention_length, int hidden_length, int head_size, int head_group_size,
                  int num_layers, int model_shard, int num_model_shards, int layer_shard, int num_layer_shards,
                  jl_dctx *out);
/* KvBufferCache.computePageSize (core/tensor/KvBufferCache.java:224-280) */
int jl_kv_page_geometry(int num_layers, int context_length, int kv_segment_length, int dtype_size,
                        int64_t max_page_bytes, int *layers_per_page, int *ctx_per_page);

/* ---- device-resident model: AbstractModel / LlamaModel mirror -----------------------------
 * core/model/AbstractModel.java:100-183,295-329,443-491,516-646; core/model/llama/LlamaModel.java:68-184.
 * The host driver (C++) replays generate()'s exact call sequence with every layer op as a CUDA
 * kernel on device-resident activations, paged KV in HBM, CUDA-graphed decode. */
typedef struct {
    int context_length, embedding_length, hidden_length;
    int num_heads, num_kv_heads, num_layers, vocab_size, head_size;
    float layer_norm_eps;
    double rope_theta;   /* LlamaConfig.java:27-57 */
    double rope_scaling; /* only rope_type "linear" is honoured by the reference (:55-56) */
    int working_qtype;   /* JL_I8 (default, Q8 activations), JL_F32, JL_BF16 (AbstractModel.java:119-176) */
    int kv_dtype;        /* JL_F32 (reference default workingDType) or JL_BF16 */
    int max_batch;       /* jlama.max_batch_size, 256 (AbstractModel.java:57,304) */
    int max_sessions;    /* concurrent KvBuffers (KvBufferCache.java:58-60) */
    int max_context;     /* positions to reserve pages for (<= context_length); 0 = context_length */
    int tp_rank, tp_size; /* jlama-net model shard (DistributedContext modelShard/numModelShards) */
    int prefill_tensor_core; /* 1: prompt chunks of >= 16 rows use the tcgen05 BF16 GEMM path; 0: exact-integer SIMT path */
    int flags;           /* JL_MODEL_* */
    int num_experts;       /* 0 = dense MLP; > 0: Mixtral-style mixture of experts (core/model/MoEBlock.java) */
    int experts_per_token; /* top-k (MixtralConfig numberOfExpertsPerToken) */
    int arch;              /* JL_ARCH_LLAMA (Llama / Mixtral blocks) or JL_ARCH_GPT2 (core/model/gpt2/GPT2Model.java) */
} jl_model_config;

#define JL_ARCH_LLAMA 0
/* GPT-2 blocks: learned position embeddings (wte + wpe, GPT2Model.java:54-69), LayerNorm with bias (LayerNorm.java:41-67), biased
 * q/k/v/o and MLP projections (CausalSelfAttention.java:184-192,380; MLPBlock.java:126-128,160), GELU, no RoPE, lm_head tied to wte.
 * Slots: JL_T_EMBED = wte, JL_T_OUT_NORM = ln_f.weight; JL_L_ATTN_NORM / JL_L_FFN_NORM = ln_1 / ln_2 weights, JL_L_Q/K/V = the three
 * row blocks of c_attn.weight^T, JL_L_O = attn c_proj^T, JL_L_GATE = c_fc^T, JL_L_DOWN = mlp c_proj^T (JL_L_UP stays empty);
 * biases and wpe through jl_model_set_aux_tensor.  Single rank, dense (non-MoE). */
#define JL_ARCH_GPT2 1
/* jl_model_set_aux_tensor `which`: layer < 0 */
#define JL_AUX_POS_EMBED 0     /* wpe.weight [context_length, E] */
#define JL_AUX_OUT_NORM_BIAS 1 /* ln_f.bias [1, E] */
/* layer >= 0 */
#define JL_AUX_ATTN_NORM_BIAS 0 /* ln_1.bias */
#define JL_AUX_Q_BIAS 1         /* c_attn.bias split in three (GPT2Model.java:79) */
#define JL_AUX_K_BIAS 2
#define JL_AUX_V_BIAS 3
#define JL_AUX_O_BIAS 4         /* attn.c_proj.bias */
#define JL_AUX_FFN_NORM_BIAS 5  /* ln_2.bias */
#define JL_AUX_FC_BIAS 6        /* mlp.c_fc.bias [1, H] */
#define JL_AUX_PROJ_BIAS 7      /* mlp.c_proj.bias */

#define JL_MODEL_NO_GRAPH 1 /* launch decode kernels eagerly instead of through a CUDA graph */
#define JL_MODEL_NO_PDL 2   /* (default) no programmatic dependent launch between the decode kernels */
#define JL_MODEL_NO_PERSISTENT 4 /* decode through the CUDA graph of per-op kernels instead of the persistent decode kernel */
#define JL_MODEL_PDL 16     /* opt in: programmatic dependent launch between the per-op decode kernels (measured: no gain) */

/* tensor slots */
#define JL_T_EMBED 0
#define JL_T_OUT_NORM 1
#define JL_T_LM_HEAD 2
#define JL_L_ATTN_NORM 0
#define JL_L_Q 1
#define JL_L_K 2
#define JL_L_V 3
#define JL_L_O 4
#define JL_L_FFN_NORM 5
#define JL_L_GATE 6
#define JL_L_DOWN 7
#define JL_L_UP 8

/* sizeof(jl_model_config) of this build: bindings that mirror the struct by hand (ctypes, FFM StructLayout) compare it with theirs */
int jl_model_config_size(void);
int jl_model_create(jl_ctx *ctx, const jl_model_config *cfg, jl_model **out);
/* layer < 0: global slot (JL_T_*); else per-layer slot (JL_L_*).  The tensor must already hold this
 * rank's shard (rows for q/k/v/gate/up, columns for o/down: LlamaModel.java:120-133,
 * Weights.getLoadOffsets core/safetensors/Weights.java:99-117). */
int jl_model_set_tensor(jl_model *m, int layer, int slot, int64_t tensor_id);
/* Mixture-of-experts layers (core/model/mixtral/MixtralModel.java:88-105): expert < 0 binds the router
 * ("block_sparse_moe.gate.weight" [num_experts, E]); otherwise which = 0 w1 (gate_proj [H, E]), 1 w2 (down_proj [E, H]),
 * 2 w3 (up_proj [H, E]) of that expert.  The dense JL_L_GATE / JL_L_UP / JL_L_DOWN slots stay empty for MoE models. */
int jl_model_set_expert_tensor(jl_model *m, int layer, int expert, int which, int64_t tensor_id);
/* GPT-2 family: bias vectors ([1, n] F32/BF16) and the position-embedding table, see JL_AUX_* */
int jl_model_set_aux_tensor(jl_model *m, int layer, int which, int64_t tensor_id);
/* allocate scratch + KV page pool, build RoPE table, capture graphs */
int jl_model_finalize(jl_model *m);
int jl_model_free(jl_model *m);
/* KvBuffer lifecycle (KvBufferCache.getKvBuffer / close) */
int jl_model_reset_session(jl_model *m, int session);
/* AbstractModel.batchForward (:295-312): prompt tokens (HOST int32) in chunks of max_batch. */
int jl_model_batch_forward(jl_model *m, int session, const int32_t *tokens, int n, int start_pos);
/* AbstractModel.sample (:443-491) on the last forwarded row of `session`: final norm, lm_head, argmax
 * (temperature 0) or softmax sampling with `uniform`.  logits_out (HOST, [vocab]) may be NULL. */
int jl_model_sample(jl_model *m, int session, float temperature, float uniform, int32_t *token_out, float *logits_out);
/* One decode step for `n` sessions at once (the reference's "batch" = concurrent sessions):
 * forward(token_i, position_i) + sample at temperature 0, one CUDA-graph launch.
 * tokens/positions/next_tokens are HOST int32[n]; logits_out HOST [n, vocab] or NULL. */
int jl_model_decode(jl_model *m, int n, const int32_t *sessions, const int32_t *tokens, const int32_t *positions,
                    int32_t *next_tokens, float *logits_out);
/* jl_model_decode with per-row sampling: a row whose temperature is not 0 draws its token with AbstractModel.sample's rule (:475-489:
 * exp((l - max) / T) in double, float running sum against `uniform`) from its logits row; rows at temperature 0 keep the arg-max.
 * temperatures / uniforms: HOST float[n].  logits_out receives the raw logits. */
int jl_model_decode_sample(jl_model *m, int n, const int32_t *sessions, const int32_t *tokens, const int32_t *positions,
                           const float *temperatures, const float *uniforms, int32_t *next_tokens, float *logits_out);
/* generate() with a temperature: uniforms[i] (HOST float[n_new]) stands for the i-th ThreadLocalRandom.nextFloat() of the loop */
int jl_model_generate_sample(jl_model *m, int session, const int32_t *prompt, int n_prompt, int n_new, float temperature,
                             const float *uniforms, int32_t *out_tokens, double *timings_ms);
/* AbstractModel.generate (:516-646) at temperature 0 over token ids: returns n_new tokens (the first is
 * sampled from the prompt's last row).  timings_ms (nullable double[2]) = {prompt, generate} wall ms like
 * Generator.Response.  logits_out HOST [n_new, vocab] or NULL. */
int jl_model_generate(jl_model *m, int session, const int32_t *prompt, int n_prompt, int n_new, int32_t *out_tokens,
                      float *logits_out, double *timings_ms);
/* device-side decode loop without per-token host round trips: feeds each step's argmax into the next
 * step on the GPU (tokens stay in HBM); copies the n_new tokens back at the end.  Used for `value`
 * in bench.py (inputs resident in HBM). */
int jl_model_decode_resident(jl_model *m, int session, int32_t first_token, int start_pos, int n_new,
                             int32_t *out_tokens);
/* test hooks: copy a K/V row (f32) or the hidden rows of the last batch_forward chunk to HOST */
int jl_model_read_kv(jl_model *m, int session, int layer, int position, int which /*0=K,1=V*/, float *out);
/* KvBufferCache.KvBufferPage persistence (core/tensor/KvBufferCache.java:121-176): write / read the allocated KV pages of a session
 * as  <dir>/<session_name>-L<layerPage>C<contextPage>.page  (raw little-endian page bytes, the reference's file naming and layout).
 * Return the number of pages written / read (>= 0) or a JL_ERR_*.  After a load the caller continues at the position it saved. */
int jl_model_kv_save(jl_model *m, int session, const char *dir, const char *session_name);
int jl_model_kv_load(jl_model *m, int session, const char *dir, const char *session_name);
/* Host spill of an idle session (SURVEY 8f.3 "device page pool, host spill/restore"): offload copies the allocated KV pages of `session`
 * to host memory, frees their HBM and leaves the slot empty; returns a handle (> 0) or a negative JL_ERR_*.  restore zeroes `session`
 * (any slot), brings the pages of `handle` back and forgets the handle; the caller continues at the position it stopped at.  discard
 * drops a handle that will not be restored.  kv_pages = allocated pages of a session. */
int64_t jl_model_kv_offload(jl_model *m, int session);
int jl_model_kv_restore(jl_model *m, int session, int64_t handle);
int jl_model_kv_discard(jl_model *m, int64_t handle);
int jl_model_kv_pages(jl_model *m, int session);
int jl_model_read_hidden(jl_model *m, int session, float *out /* [embedding_length] last row */);
/* test hook: copy `n` floats of an internal activation buffer of the LAST forward/decode call to HOST (row 0 first).
 * which: 0 = x (hidden after the last layer), 1 = xb (after attention + residual), 2 = q, 3 = k, 4 = v (raw projections),
 * 5 = att (attention output), 6 = h (silu(gate) * up), 7 = logits. */
int jl_model_debug_read(jl_model *m, int which, float *out, int64_t n);
/* diagnostic: CTA 0's phase stamps (globaltimer ns) of the last persistent decode launch (tools/ptrace.py); phase tracing
 * is enabled with JL_PD_TRACE=1 in the environment when the model is created.  Returns the number of words written. */
int jl_model_debug_trace(jl_model *m, uint64_t *out, int64_t out_words);
/* how jl_model_decode executes for n sessions: 3 = persistent decode kernel (one cooperative launch per token, n == 1),
 * 1 = CUDA graph of per-op kernels, 0 = eager */
int jl_model_decode_mode(jl_model *m, int n);
/* per-token algorithmic bytes of the decode weight stream on this rank (roofline numerator) */
int64_t jl_model_weight_bytes(jl_model *m);
/* event-timed duration (ms) of the last jl_model_decode / decode_resident GPU work, and of its
 * dominant kernel class (the quantised GEMV), measured with CUDA events on the model stream */
int jl_model_last_timing(jl_model *m, double *total_ms, double *gemv_ms);

/* ---- JQ4 / JQ8 checkpoints: safetensors with Jlama's dtype strings "Q4" / "I8" + "<name>.qb" f32 block scales ------------
 * core/safetensors/SafeTensorSupport.java:54-102,215-332; Weights.java:49-66,99-179; SafeTensorIndex.java:59-236.
 * Files are mapped whole (64-bit), so the reference's <= 2 GiB mmap splits (SafeTensorIndex.java:121-236) are not needed. */
typedef struct jl_st jl_st;
#define JL_ST_F16 4 /* dtype code reported for F16 tensors (read-only; bound to a model as F32, Weights.java:137-152) */
/* path: a .safetensors file, or a directory with model.safetensors or model.safetensors.index.json + shards */
int jl_st_open(const char *path, jl_st **out);
int jl_st_close(jl_st *st);
const char *jl_st_last_error(void);
int jl_st_count(jl_st *st);
int jl_st_find(jl_st *st, const char *name); /* index or -1 */
/* tensors are ordered by data offset (TensorInfo.compareTo); dtype -1 = a dtype string this library does not know */
int jl_st_info(jl_st *st, int i, const char **name, int *dtype, int *ndim, int64_t *shape4, int64_t *nbytes);
const void *jl_st_data(jl_st *st, int i); /* pointer into the read-only mapping */
const char *jl_st_metadata(jl_st *st, const char *key); /* "__metadata__" entry or NULL */
int jl_st_majority_dtype(jl_st *st); /* Weights.findDType: ".qb" tensors not counted, F16 counts as F32 */
/* one file: 8-byte LE header length, JSON header (data_offsets relative to the data section), raw bytes in the given order */
int jl_st_write(const char *path, int n, const char *const *names, const int *dtypes, const int *ndims, const int64_t *shapes4,
                const void *const *data, const int64_t *nbytes, int n_meta, const char *const *meta_kv);
/* SafeTensorSupport.quantizeModel with the block quantisers on the GPU (byte-identical files): 2-D tensors whose name contains
 * none of the comma-separated `skip_csv` substrings (NULL = "norm", QuantizeCommand.java:36-38) and whose first dim is not 1
 * (AbstractTensor.java:284) become qtype (JL_Q4 / JL_I8) + "<name>.qb"; names starting with a `drop_csv` prefix are omitted. */
int jl_quantize_model(jl_ctx *ctx, const char *src_dir, const char *dst_dir, int qtype, const char *skip_csv, const char *drop_csv);
/* config.json -> jl_model_config (safetensors/Config.java:253-287, llama/LlamaConfig.java:27-57); working_qtype I8, F32 KV, tp 1 */
int jl_config_from_json(const char *config_json_path, jl_model_config *cfg);
/* LlamaModel.loadInputWeights / loadTransformerBlockWeights / loadOutputWeights (llama/LlamaModel.java:68-156): registers this
 * rank's shard of every tensor of the checkpoint (rows per Weights.getLoadOffsets :99-117, columns per AbstractTensor.sparsify)
 * and binds it; call jl_model_finalize afterwards.  ids_out receives the registered tensor ids (for jl_unregister_tensor). */
int jl_model_load_safetensors(jl_model *m, jl_ctx *ctx, jl_st *st, int64_t *ids_out, int ids_cap, int *n_ids);
/* this model's DistributedContext and shard count */
int jl_model_tp_layout(jl_model *m, jl_dctx *out, int *tp_size);

/* ---- concurrent sessions: request queue + iteration-level batching (csrc/jl_sched.cu) ------------------------------------
 * The reference runs one thread per request, each calling AbstractModel.generate (core/model/AbstractModel.java:516-646) on its own
 * KvBuffer (core/tensor/KvBufferCache.java:58-60; jlama-net/.../openai/OpenAIChatService.java:64-74,107-160).  Here requests queue in
 * front of the batched decode step: every jl_sched_step admits queued requests into free session slots (FIFO), forwards prompt chunks,
 * runs ONE decode step for all generating requests (rows of different lengths side by side, at most the model's rows-per-call per
 * backend call, each row greedy or sampled at its own temperature) and retires finished requests so that their slots are reused by the
 * next step.  When no slot is free, the least
 * recently finished kept session is spilled to host memory (KvBufferCache's pages are file-backed in the reference, :121-176; here the
 * idle session's pages leave HBM) and restored into any free slot when its follow-up arrives.  Under tensor parallelism every rank runs the same scheduler on the same request stream (it is deterministic).
 * Free the scheduler before its model. */
typedef struct jl_sched jl_sched;
/* request states */
#define JL_SCHED_QUEUED 0
#define JL_SCHED_PREFILL 1
#define JL_SCHED_DECODING 2
#define JL_SCHED_FINISHED 3
#define JL_SCHED_FAILED 4
/* Generator.FinishReason (core/model/functions/Generator.java) + the scheduler's own exits */
#define JL_FINISH_NONE 0
#define JL_FINISH_MAX_TOKENS 1 /* max_new tokens produced, or the reserved context is full (generate :538,:590) */
#define JL_FINISH_STOP_TOKEN 2 /* c.eosTokens.contains(next) (:604-608) */
#define JL_FINISH_CANCELLED 3
#define JL_FINISH_ERROR 4      /* a backend call failed for this request; jl_sched_last_error */
/* submit flags */
#define JL_SCHED_KEEP_SESSION 1 /* keep the session slot and its KV after the request finishes, for a continuation (the reference's
                                   session header: the next generate() on that UUID starts at kvmem.getCurrentContextPosition(), :533) */
/* the four model calls the policy is written against; all return JL_OK or a JL_ERR_* */
typedef struct {
    int (*reset_session)(void *user, int session);                                                   /* jl_model_reset_session */
    int (*batch_forward)(void *user, int session, const int32_t *tokens, int n, int start_pos);       /* jl_model_batch_forward */
    int (*sample)(void *user, int session, float temperature, float uniform, int32_t *token_out);     /* jl_model_sample */
    int (*decode)(void *user, int n, const int32_t *sessions, const int32_t *tokens, const int32_t *positions,
                  const float *temperatures, const float *uniforms, int32_t *next_tokens);            /* jl_model_decode_sample */
    /* optional, all three or none: host spill of kept sessions (jl_model_kv_offload / kv_restore / kv_discard) */
    int (*offload)(void *user, int session, int64_t *handle_out);
    int (*restore)(void *user, int session, int64_t handle);
    int (*discard)(void *user, int64_t handle);
} jl_sched_backend;
typedef struct {
    int admitted, prefill_tokens, decode_rows, decode_calls, finished; /* what this step (or run) did */
    int active, queued;                                                 /* afterwards */
    int spilled;                                                        /* kept sessions moved to host memory to free their slot */
} jl_sched_stats;
typedef struct {
    int state, finish_reason, session /* -1: none, or spilled */, start_pos, n_prompt, n_prefilled, n_generated, next_position;
    int spilled; /* 1: this finished request's kept KV lives in host memory; its continuation restores it into a free slot */
    int64_t submit_step, first_token_step, finish_step; /* scheduler step counters: queueing delay and time to first token in steps */
    /* host-clock milliseconds: waiting for a slot; admission -> first token (Generator.Response promptTimeMs, AbstractModel.java:561-568);
     * first token -> finish (generateTimeMs, :589,:623).  A phase still running is measured up to the call. */
    double queue_ms, prompt_ms, generate_ms;
} jl_sched_request_info_t;
/* max_active: session slots to use (<= the model's max_sessions; 0 = all).  prefill_tokens_per_step: prompt tokens forwarded per step over
 * all admitted requests (chunked prefill, bounds the latency a long prompt adds to the running requests' decode steps); 0 = no bound. */
int jl_sched_create(jl_model *m, int max_active, int prefill_tokens_per_step, jl_sched **out);
/* the same policy over caller-supplied model calls (CPU tests drive it with the oracle; a host with its own model object plugs in here) */
int jl_sched_create_backend(const jl_sched_backend *be, void *user, int n_sessions, int max_rows_per_decode, int max_context,
                            int prefill_tokens_per_step, jl_sched **out);
int jl_sched_free(jl_sched *s);
const char *jl_sched_last_error(jl_sched *s);
/* Queue a request: prompt token ids, at most max_new generated tokens (the first is sampled from the prompt's last row and, like the
 * reference's, not stop-checked), optional stop tokens.  continue_request >= 0: a finished JL_SCHED_KEEP_SESSION request whose session
 * this request appends to.  temperature 0 = arg-max; otherwise every token is drawn with the reference's rule (AbstractModel.java:475-489)
 * from a per-request uniform stream seeded with `seed` (the reference draws ThreadLocalRandom.nextFloat(), :576,:594): the k-th token of
 * a request uses its k-th draw however the steps were batched, so (seed, prompt) reproduces.  Returns the request id (> 0) or -1.
 * Callable from any thread, also while a step runs. */
int64_t jl_sched_submit(jl_sched *s, const int32_t *prompt, int n_prompt, int max_new, const int32_t *stop_tokens, int n_stop, int flags,
                        int64_t continue_request, float temperature, uint64_t seed);
int jl_sched_cancel(jl_sched *s, int64_t request); /* takes effect at the next step boundary */
/* one scheduling iteration; returns the first backend error of the step (the affected requests are JL_SCHED_FAILED, the others go on) */
int jl_sched_step(jl_sched *s, jl_sched_stats *stats /* nullable */);
/* step until nothing is queued or running (max_steps <= 0: no limit); totals are summed over the steps */
int jl_sched_run(jl_sched *s, int max_steps, jl_sched_stats *totals /* nullable */);
/* tokens generated so far (streaming read), state and finish reason; copies min(cap, n) tokens, *n_tokens = n */
int jl_sched_result(jl_sched *s, int64_t request, int32_t *tokens, int cap, int *n_tokens, int *state, int *finish_reason);
int jl_sched_request_info(jl_sched *s, int64_t request, jl_sched_request_info_t *info);
/* forget a finished / failed request; frees the session slot it kept */
int jl_sched_release(jl_sched *s, int64_t request);
int jl_sched_counts(jl_sched *s, int *queued, int *active, int *free_slots);

/* ---- jlama-net replacement: NCCL over NVLink instead of gRPC -------------------------------
 * JlamaService.combine (jlama-net/src/main/java/com/github/tjake/jlama/net/grpc/JlamaService.java:300-359)
 * == all-reduce SUM of [M,E] f32.  One process per GPU; rank 0 creates the id, the launcher
 * (torch.distributed / any store) broadcasts it. */
int jl_comm_unique_id(jl_ctx *ctx, uint8_t *id128 /* 128 bytes */);
int jl_comm_init(jl_ctx *ctx, const uint8_t *id128, int rank, int world);
/* HOST-buffer all-reduce for tests of the collective itself */
int jl_comm_allreduce_f32(jl_ctx *ctx, float *host_buf, int64_t count);
int jl_comm_destroy(jl_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* JLAMA_B200_H */
