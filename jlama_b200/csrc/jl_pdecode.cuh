// Persistent decode kernel: shared declarations (jl_pdecode.cu implements, jl_model.cu launches).
#pragma once
#include "jl_common.cuh"

#ifndef PD_THREADS
#define PD_THREADS 512
#endif
#define PD_MAX_TP 8
#define PD_SYNC_WORDS 32 // u64 words: [0] epoch, [1] status (0 ok, else the phase id that timed out), [8..] barrier counters

struct PdLayer {
    const uint8_t *w[7]; // q, k, v, o, gate, up, down (Q4 nibbles or int8, row-major, this rank's shard)
    const float *s[7];   // block scales
    const void *attn_norm, *ffn_norm;
    int attn_norm_dt, ffn_norm_dt;
};
enum { PW_Q = 0, PW_K, PW_V, PW_O, PW_GATE, PW_UP, PW_DOWN };

// barrier counters (cumulative; the target of use number u is u * arrivals)
enum { PC_EMBED = 8, PC_QKV, PC_ATT, PC_O, PC_GU, PC_DOWN, PC_LM, PC_FINAL };

struct PdParams {
    int layers, E, H, attn_seg, kv_seg, heads, kv_heads, head_size; // H, attn_seg, kv_seg, heads, kv_heads: this rank's shard
    int vocab, vocab_rows, vocab0; // total vocabulary, lm_head rows held by this rank, first of them
    int head0_global, kv_head0_global;
    float eps, attn_scale;
    const PdLayer *lw; // unused (the layer table travels in constant memory)
    int embed_dt;
    const void *embed_w;
    const float *embed_s;
    const void *out_norm;
    int out_norm_dt;
    const uint8_t *lm_w; // this rank's rows of the lm_head (or tied embedding) in the linear-weight dtype
    const float *lm_s;
    float *x, *xb, *q, *k, *v, *att, *h, *logits, *attn_ws;
    const float *rope;
    KvLayout kv;
    int32_t *tokens, *positions, *next;
    const int32_t *sessions;
    int32_t *hist, *counter;
    int hist_cap, resident;
    unsigned long long *sync;         // [PD_SYNC_WORDS]
    unsigned long long *argmax_slots; // [grid] packed (ordered logit bits << 32 | ~index) per CTA
    unsigned *att_done;               // [kv_heads] split arrival counters (self-resetting)
    int splits;
    int ntok, split_cap;              // tokens per launch (resident loop), upper bound of the context splits
    // tensor parallel exchange over NVLink peer memory (world > 1): LL lines {v0, tag, v1, tag}
    int world, rank;
    uint4 *ll_o[PD_MAX_TP]; // ll_o[d] = rank d's receive buffer for o_proj partials: [world][E/2] lines
    uint4 *ll_d[PD_MAX_TP]; // same for down_proj partials
    uint4 *ll_a[PD_MAX_TP]; // per-rank arg-max candidates: [world] lines {lo, tag, hi, tag}
    int want_logits;        // world > 1: broadcast this rank's logits slice to every rank's logits buffer
    float *logits_peer[PD_MAX_TP];
    unsigned long long *trace; // optional diagnostics: CTA 0 stamps [layers * 8 + 8] globaltimer values
};

bool jl_pdecode_supported(const PdParams &p, int w_dtype, int grid);
// one decoded token; `stream` order serialises consecutive tokens.  The model constants (dims, buffers, per-layer weight
// pointers) live in __constant__ memory and are uploaded when `owner` (the model) differs from the previous launch's owner on
// this device; splits / resident / want_logits travel as kernel arguments.
int jl_launch_pdecode(jl_ctx *ctx, cudaStream_t stream, const PdParams &p, const PdLayer *layers_host, const void *owner, int w_dtype);
// the owner is going away (or its constants changed): the next launch re-uploads
void jl_pdecode_forget(int device, const void *owner);
