// Persistent decode megakernel: shared declarations (jl_mega.cu implements, jl_model.cu launches).
#pragma once
#include "jl_common.cuh"

#define MEGA_MAX_M 4

struct MegaLayer {
    const uint8_t *w[7]; // q, k, v, o, gate, up, down  (packed Q4 nibbles, row pitch K/2 bytes)
    const float *s[7];   // block scales, row pitch K/32 floats
    const void *attn_norm, *ffn_norm;
    int attn_norm_dt, ffn_norm_dt;
};
enum { MW_Q = 0, MW_K, MW_V, MW_O, MW_GATE, MW_UP, MW_DOWN };

struct MegaParams {
    int layers, E, H, attn_seg, kv_seg, heads, kv_heads, head_size, vocab;
    int head0_global, kv_head0_global;
    int M; // rows (concurrent sessions) in this step
    float eps, attn_scale;
    const MegaLayer *lw; // device array [layers]
    // embedding / head
    int embed_dt;
    const void *embed_w;
    const float *embed_s;
    const void *out_norm;
    int out_norm_dt;
    const uint8_t *lm_w;
    const float *lm_s;
    // global scratch (L2 resident)
    float *x, *xb, *q, *k, *v, *att, *h, *logits, *attn_ws;
    const float *rope;
    KvLayout kv;
    int32_t *tokens, *positions, *next;
    const int32_t *sessions;
    int32_t *hist, *counter;
    int hist_cap, resident;
    unsigned int *sync;      // [layers*5 + 3] completion counters, zeroed before launch
    unsigned long long *argmax_slots; // [MEGA_MAX_M][grid] packed (ordered logit bits, ~index)
    int splits;              // attention context splits per (row, kv head)
    unsigned int *att_done;  // [layers][M*kv_heads] split arrival counters (zeroed before launch)
    int grid;                // CTAs (= SM count) the schedule was built for
    int dbg;                 // diagnostics: 1 = consumers skip the math, 2 = skip dependencies/prologues/attention
    int l2_ahead;            // 1: the helper warp prefetches the next ops' weights into L2
    int uarea_bytes;         // filled by the launcher
    long long *trace;        // optional [3 CTAs][n_ops][8] clock64 stamps (diagnostics), or nullptr
};

size_t jl_mega_sync_words(int layers);
// true when the model shape is supported by the megakernel
bool jl_mega_supported(const MegaParams &p);
int jl_launch_mega(jl_ctx *ctx, cudaStream_t stream, const MegaParams &p);
