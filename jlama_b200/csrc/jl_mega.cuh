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

// Host-precomputed schedule: every CTA's stages in execution order.  A stage = up to 4 bulk copies.
// Records are 96 bytes: MegaCopy[4] followed by MegaMeta, contiguous per CTA so that the producer can
// pull batches of records into shared memory with one bulk copy.
struct MegaCopy {
    unsigned long long src;
    uint32_t bytes; // 0 = unused
    uint32_t dst;   // byte offset inside the ring slot
};
struct MegaMeta {
    uint32_t total_bytes;
    int32_t row0;
    int16_t nrows, R, wpr, seg;
    uint8_t pair, type, wpr_shift, r_shift; // wpr and R are powers of two
    int32_t op, K;
    int32_t pad2;
};
static_assert(sizeof(MegaMeta) == 32 && sizeof(MegaCopy) == 16, "schedule records are read with 128-bit loads");

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
    const unsigned char *records; // [total stages][96]: MegaCopy[4] | MegaMeta
    int grid;                // CTAs (= SM count) the schedule was built for
    int dbg;                 // diagnostics: 1 = consumers skip the math, 2 = skip dependencies/prologues/attention
    int l2_ahead;            // 1: prefetch the next batches of stages into L2 while the ring is full
    const int *cta_first;    // [grid + 1] first stage of each CTA
    const int *op_first;     // [grid][n_ops + 1] first stage of each op within the CTA's range
    long long *trace;        // optional [3 CTAs][n_ops][8] clock64 stamps (diagnostics), or nullptr
};

size_t jl_mega_sync_words(int layers);
// Build the static schedule for a grid of G CTAs (host).  `shape` needs the model dims; `layers` are host copies.
void jl_mega_build_table(const MegaParams &shape, const MegaLayer *layers, int G, std::vector<unsigned char> &records,
                         std::vector<int> &cta_first, std::vector<int> &op_first);
// true when the model shape is supported by the megakernel
bool jl_mega_supported(const MegaParams &p);
int jl_launch_mega(jl_ctx *ctx, cudaStream_t stream, const MegaParams &p);
