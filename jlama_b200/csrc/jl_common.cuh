// Internal declarations shared by the translation units of libjlama_b200.so.
// sm_100a only.  No torch types anywhere in this library.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/jlama_b200.h"

#define QBLOCK 32

struct DevTensor {
    int dtype = 0;
    int64_t rows = 0, cols = 0;
    void *data = nullptr;    // f32 / bf16 / packed nibbles / int8
    float *scales = nullptr; // [rows, cols/32] for Q4 / I8
    size_t bytes = 0;
    int64_t id = 0;          // registry id (0 for views / temporaries)
    int refs = 0;            // models this tensor is bound to (jl_model_set_tensor); unregister refuses while > 0
};

struct jl_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr; // op-level API stream
    std::mutex mu;
    std::mutex err_mu;      // guards last_error only: model entry points report errors without holding `mu`
    std::string last_error; // the most recent error of any thread; jl_last_error prefers the calling thread's own (jl_runtime.cu)
    std::unordered_map<int64_t, DevTensor> tensors;
    std::vector<struct jl_model *> models; // live models (freed by jl_shutdown before the tensors they point at)
    int64_t next_id = 1;
    long long launches = 0;
    // op-level scratch (grown on demand)
    void *scratch[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratch_bytes[4] = {0, 0, 0, 0};
    // kernel timeline diagnostics (jl_debug_ktrace): device [cap][4] globaltimer stamps, one slot per traced launch
    unsigned long long *ktrace = nullptr;
    int ktrace_cap = 0, ktrace_n = 0;
    // comm
    void *nccl_comm = nullptr;
    int rank = 0, world = 1;
};

int jl_set_error(jl_ctx *ctx, int code, const char *fmt, ...);
#define JL_CUDA_CHECK(ctx, expr)                                                                      \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            return jl_set_error(ctx, _e == cudaErrorMemoryAllocation ? JL_ERR_OOM : JL_ERR_CUDA,      \
                                "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

void *jl_scratch(jl_ctx *ctx, int slot, size_t bytes); // nullptr on OOM
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remember what was set per (device, kernel)
// so that a second context on another GPU of the same process configures its own copy.
#define JL_MAX_DEVICES 64
template <typename K>
static inline cudaError_t jl_ensure_dyn_smem(K kernel, int device, size_t bytes, size_t (&configured)[JL_MAX_DEVICES]) {
    if (bytes <= 40 * 1024) return cudaSuccess; // static shared memory counts against the 48 KB default too
    const int d = device >= 0 && device < JL_MAX_DEVICES ? device : 0;
    if (bytes <= configured[d]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) configured[d] = bytes;
    return e;
}
// called when the timeline buffer is replaced: captured graphs hold the old slot pointers
void jl_models_invalidate_graphs(jl_ctx *ctx);

// ---------------------------------------------------------------------------------------------
// Quantised GEMV / small-M GEMM (jl_gemv.cu)
// ---------------------------------------------------------------------------------------------
enum GemvPrologue {
    PRO_Q8_GLOBAL = 0,  // activations already int8 + scales in global memory
    PRO_F32_QUANT,      // f32 activations -> Q8 (PanamaTensorOperations.quantizeQ8 semantics) in the prologue
    PRO_RMSNORM_QUANT,  // f32 hidden row -> RMSNorm (RMSNorm.java:34-56) -> Q8
    PRO_F32,            // f32 activations used as-is (F32 x W path)
    PRO_RMSNORM_F32,    // RMSNorm then f32 activations
    PRO_BF16_GLOBAL,    // bf16 activations in global -> f32
};
enum GemvEpilogue {
    EPI_STORE = 0,    // out[m, out_off + row] = acc
    EPI_ADD_RESIDUAL, // out[m, row] = acc + residual[m, row]
    EPI_SILU_MUL,     // seg0 = gate, seg1 = up: out[m,row] = silu(g) * u   (MLPBlock.java:132-141)
};

struct GemvSeg {
    const void *w;       // weight rows, leading dim = ldw elements
    const float *ws;     // block scales (Q4/I8) with leading dim ldw/32
    float *out;          // f32 output [M, out_ld]
    int rows;            // rows in this segment
    int out_ld;
    int out_off;         // added to the row index on store (already includes -roffset)
};

struct GemvParams {
    GemvSeg seg[3];
    int nseg;
    int w_dtype;             // JL_Q4 / JL_I8 / JL_BF16 / JL_F32
    int ldw;                 // weight leading dimension in elements
    int w_col_off;           // first weight column (elements, multiple of 32 for quantised)
    int K;                   // reduction length
    int M;                   // activation rows (1..GEMV_MAX_M)
    // activations
    const void *a;           // f32 / int8 / bf16 [M, lda]
    const float *a_scales;   // for PRO_Q8_GLOBAL: [M, lda/32]
    int lda;
    int a_col_off;
    // norm prologue
    const void *norm_w;      // f32 or bf16 [E]
    int norm_w_dtype;
    float norm_adj, norm_eps;
    int norm_E;
    double norm_inv_E;       // filled by jl_launch_gemv
    // epilogue
    const float *residual;   // [M, res_ld]
    int res_ld;
    int row0;                // first global row handled by this launch (n0)
    int total_rows;          // rows handled by this launch
    unsigned long long *trace; // diagnostics slot (jl_debug_ktrace) or nullptr
    // mixture-of-experts indirection (MoEBlock.java:98-137): when sel != nullptr the weights of segment i are
    // w_tab[i][*sel] / ws_tab[i][*sel] -- the expert index is only known on the device (router top-k of this row)
    const int *sel;
    const void *const *w_tab[2];
    const float *const *ws_tab[2];
};
__device__ __forceinline__ void gemv_seg_base(const GemvParams &p, int seg, const void *&w, const float *&ws) {
    if (p.sel) {
        const int e = __ldg(p.sel);
        w = p.w_tab[seg][e], ws = p.ws_tab[seg][e];
    } else {
        w = p.seg[seg].w, ws = p.seg[seg].ws;
    }
}

#define GEMV_MAX_M 8

// Launch on `stream`.  use_pdl: launch with the programmatic-stream-serialization attribute.
int jl_launch_gemv(jl_ctx *ctx, cudaStream_t stream, const GemvParams &p, int prologue, int epilogue, bool use_pdl);
int jl_gemv_max_m(int w_dtype, int prologue, int K);
// batched decode (2..8 rows): integer block products on the tensor cores (jl_gemm8.cu)
bool jl_gemm8_supported(const GemvParams &p, int prologue, int epilogue);
int jl_launch_gemm8(jl_ctx *ctx, cudaStream_t stream, const GemvParams &p, int prologue, int epilogue); // largest M chunk (1/2/4/8) that fits shared memory, 0 if none

// ---------------------------------------------------------------------------------------------
// Element-wise / normalisation / sampling kernels (jl_elementwise.cu)
// ---------------------------------------------------------------------------------------------
int jl_launch_accumulate(jl_ctx *ctx, cudaStream_t s, float *a, int a_rows, int lda, int b_dtype, const void *b,
                         const float *b_scales, int b_rows, int ldb, int offset, int length);
int jl_launch_maccumulate(jl_ctx *ctx, cudaStream_t s, float *a, int a_rows, int lda, const float *b, int b_rows, int ldb,
                          int offset, int length);
int jl_launch_scale(jl_ctx *ctx, cudaStream_t s, float f, float *x, int rows, int ldx, int offset, int length);
int jl_launch_saxpy_batch(jl_ctx *ctx, cudaStream_t s, const float *alpha, const float *x, int ldx, float *y, int xoffset,
                          int yoffset, int limit, int a_offset, int x_row_offset, int batch);
int jl_launch_quantize_q8(jl_ctx *ctx, cudaStream_t s, const float *x, int rows, int ldx, int offset, int length,
                          int8_t *q, float *scales);
int jl_launch_quantize_bf16(jl_ctx *ctx, cudaStream_t s, const float *x, int rows, int ldx, int offset, int length,
                            uint16_t *out);
int jl_launch_quantize_q4w(jl_ctx *ctx, cudaStream_t s, const float *x, int64_t rows, int64_t cols, uint8_t *q,
                           float *scales);
int jl_launch_quantize_q8w(jl_ctx *ctx, cudaStream_t s, const float *x, int64_t rows, int64_t cols, int8_t *q, float *scales);
int jl_launch_rmsnorm(jl_ctx *ctx, cudaStream_t s, const float *x, int rows, int ldx, int w_dtype, const void *w, float adj,
                      float eps, int E, int offset, int length, float *out);
int jl_launch_layernorm(jl_ctx *ctx, cudaStream_t s, const float *x, int rows, int ldx, int w_dtype, const void *w, int b_dtype,
                        const void *bias, float eps, int E, int offset, int length, float *out);
int jl_launch_activation(jl_ctx *ctx, cudaStream_t s, int type, float *x, int rows, int ld, int offset, int length);
// MoE router: softmax over the expert logits of each row (VectorMath.softMax) + top-k by the reference's replace-the-minimum
// scan (MoEBlock.java:151-168); sel[row * k + i] = i-th selected expert
int jl_launch_moe_route(jl_ctx *ctx, cudaStream_t s, float *logits, int rows, int n_experts, int k, int32_t *sel);
int jl_launch_softmax(jl_ctx *ctx, cudaStream_t s, float *x, int offset, int length);
int jl_launch_silu_mul(jl_ctx *ctx, cudaStream_t s, float *gate, const float *up, int rows, int ld, int offset, int length);
// fused producers of the BF16 operand of the tensor-core prefill GEMMs
int jl_launch_rmsnorm_bf16(jl_ctx *ctx, cudaStream_t s, const float *x, int rows, int ldx, int w_dtype, const void *w, float adj, float eps,
                           int E, uint16_t *out, int ldo);
int jl_launch_silu_mul_bf16(jl_ctx *ctx, cudaStream_t s, const float *gate, const float *up, int rows, int ld, int length, uint16_t *out, int ldo);
// embedding rows -> f32 hidden (LlamaModel.java:68-100); tokens on device
int jl_launch_embed(jl_ctx *ctx, cudaStream_t s, const DevTensor &wte, const int32_t *tokens, int n, float *out, int E);
int jl_launch_pos_embed_add(jl_ctx *ctx, cudaStream_t s, float *x, int rows, int E, const DevTensor &wpe, const int32_t *positions);
// argmax with strict '>' (lowest index wins; AbstractModel.java:455-469): two-stage
int jl_launch_argmax(jl_ctx *ctx, cudaStream_t s, const float *logits, int rows, int vocab, int ld, int32_t *out_tokens,
                     void *scratch);
size_t jl_argmax_scratch_bytes(int rows);

// ---------------------------------------------------------------------------------------------
// Attention over the paged KV cache (jl_attention.cu)
// ---------------------------------------------------------------------------------------------
struct KvLayout {
    // Page = [layers_per_page, 2, ctx_per_page, kv_len] (KvBufferCache.java:102); page table per session.
    int layers_per_page, ctx_per_page, kv_len; // kv_len = this rank's kvSegmentLength
    int n_ctx_pages;                            // page-table stride per (session, layer_page)
    int n_layer_pages;
    int kv_dtype;                               // JL_F32 or JL_BF16
    void *const *page_table;                    // device: [sessions][n_layer_pages][n_ctx_pages] -> page base
};

struct AttnParams {
    KvLayout kv;
    int layer;
    int heads, kv_heads, head_size; // this rank's heads
    int head0_global;               // first global head on this rank (for the RoPE table quirk)
    int kv_head0_global;
    const float *q;  // [rows, q_ld]  raw projections (pre-RoPE)
    const float *k;  // [rows, kv_ld]
    const float *v;  // [rows, kv_ld]
    int q_ld, kv_ld;
    float *out;      // [rows, q_ld] attention output (this rank's heads)
    const float *rope; // [(positions) * hs/2][2]
    int rows;                  // query rows in this launch
    const int32_t *sessions;   // device [rows] session of each row
    const int32_t *positions;  // device [rows] absolute position of each row
    float scale;
    // split-K workspace
    float *ws;       // [rows, heads, splits, hs + 2]
    int splits;
};

// Writes rotated K and V of every row into the pages and rotates q in place (CausalSelfAttention.java:199-311).
int jl_launch_rope_kv_append(jl_ctx *ctx, cudaStream_t s, const AttnParams &p, float *q_inplace, bool use_pdl);
// Causal attention of each row against positions [0, pos] of its session (CausalSelfAttention.java:314-356).
int jl_launch_paged_attention(jl_ctx *ctx, cudaStream_t s, const AttnParams &p, int max_pos, bool use_pdl);
// tiled prefill attention for one session's chunk of consecutive positions (jl_attn_prefill.cu)
bool jl_prefill_attention_supported(const AttnParams &p);
int jl_launch_prefill_attention(jl_ctx *ctx, cudaStream_t s, const AttnParams &p, int session, int pos0);
// Decode steps (every row a different session): RoPE + KV append + attention fused in one kernel; q/k/v are the raw
// projections.  done_cnt: [rows * kv_heads] zero-initialised split-arrival counters (self-resetting).
// max_pos >= 0: the largest position among the rows (lets batched decode take the flat task); -1: tiled task only
int jl_launch_fused_decode_attention(jl_ctx *ctx, cudaStream_t s, const AttnParams &p, unsigned *done_cnt, bool use_pdl, int max_pos = -1);

// ---------------------------------------------------------------------------------------------
// Larger-M GEMM paths for prefill (jl_gemm.cu)
// ---------------------------------------------------------------------------------------------
// tcgen05 tensor-core GEMM (jl_gemm_tc.cu): C[T, out_col_off+n] (+residual) = A_bf16[T,K] * dequant(W[n, w_col_off+k])^T
int jl_launch_gemm_tc(jl_ctx *ctx, cudaStream_t stream, const uint16_t *a_bf16, int lda, int T, const DevTensor &W, int n_rows,
                      int w_col_off, int K, float *out, int ldc, int out_col_off, const float *residual, int res_ld);

// ---------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ uint4 ldg_nc_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_nc_f32(const float *p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
// Loads that should stay in L2 across decode steps (norm weights: 16 KB per layer, re-read every token but otherwise
// flushed by the 4.7 GB weight stream in between): L2 evict_last policy.
__device__ __forceinline__ unsigned long long l2_evict_last_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_keep_u4(const void *p, unsigned long long pol) {
    uint4 r;
    asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
// The decode weight stream (4.7 GB per token through a 126 MB L2) is tagged evict_first, so that it recycles its own
// lines instead of flushing what the step re-reads: activations, KV pages, RoPE and norm tables -- and kernel code.
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const void *p, unsigned long long pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float ldg_stream_f32(const float *p, unsigned long long pol) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define KTRACE_WORDS 16 // per slot: first CTA start, last CTA end, CTA 0 after prologue, tag, CTA 0 start, 11 free stamps
__device__ __forceinline__ void ktrace_begin(unsigned long long *t, unsigned long long tag) {
    if (t && threadIdx.x == 0) {
        const unsigned long long now = globaltimer_ns();
        atomicMin(&t[0], now);
        if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) t[3] = tag, t[4] = now;
    }
}
// extra stamp `idx` (5..15) by thread 0 of CTA 0, ordered after `dep` has been computed
__device__ __forceinline__ void ktrace_stamp(unsigned long long *t, int idx, float dep) {
    if (t && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now) : "f"(dep) : "memory");
        t[idx] = now;
    }
}
__device__ __forceinline__ void ktrace_mid(unsigned long long *t) {
    if (t && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) t[2] = globaltimer_ns();
}
__device__ __forceinline__ void ktrace_end(unsigned long long *t) {
    if (t && threadIdx.x == 0) atomicMax(&t[1], globaltimer_ns());
}
inline unsigned long long *jl_ktrace_slot(jl_ctx *ctx) {
    if (!ctx->ktrace || ctx->ktrace_n >= ctx->ktrace_cap) return nullptr;
    return ctx->ktrace + (size_t)KTRACE_WORDS * ctx->ktrace_n++;
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float bf16_bits_to_f32(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }

// jafama-equivalent activation in double, cast to float (ActivationFunction.java:29-37)
__device__ __forceinline__ float silu_ref(float x) { return (float)((double)x * (1.0 / (1.0 + exp(-(double)x)))); }

template <typename K, typename... Args>
static inline cudaError_t jl_launch_kernel(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                           Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}
