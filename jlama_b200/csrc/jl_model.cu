// Device-resident host driver: the C++ mirror of AbstractModel / LlamaModel / TransformerBlock /
// CausalSelfAttention / MLPBlock (core/model/AbstractModel.java:295-329,443-491,516-646;
// core/model/llama/LlamaModel.java:68-184; core/model/TransformerBlock.java:158-215;
// core/model/CausalSelfAttention.java:145-385; core/model/MLPBlock.java:106-166).
// It replays generate()'s exact call sequence; every layer op is a CUDA kernel on activations that
// never leave HBM, KV lives in device pages shaped like KvBufferCache's, and the per-token decode
// step is captured once into a CUDA graph with programmatic dependent launch between kernels.
#include "jl_common.cuh"
#include <string>
#include "jl_pdecode.cuh"

#include <chrono>
#include <map>
#include <math.h>
#include <new>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int jl_comm_allreduce_dev(jl_ctx *ctx, cudaStream_t stream, float *buf, size_t count);
int jl_comm_allgather_dev(jl_ctx *ctx, cudaStream_t stream, const void *send, void *recv, size_t bytes_per_rank);

struct jl_model {
    jl_ctx *ctx = nullptr;
    // serialises the entry points that use the model stream and its staging buffers: the reference calls generate() from one thread per
    // request (OpenAIChatService.java:107-160); recursive because generate() is built from the other entry points
    std::recursive_mutex mu;
    jl_model_config cfg;
    jl_dctx d;
    cudaStream_t stream = nullptr;
    bool finalized = false;
    DevTensor g[3];
    bool g_set[3] = {false, false, false};
    std::vector<DevTensor> l; // [layers][9]
    std::vector<char> l_set;
    std::vector<int64_t> bound_ids; // registry ids whose refs this model holds (one entry per set_tensor call)
    int group = 1, attn_seg = 0, kv_seg = 0, h_seg = 0, heads_local = 0, kv_heads_local = 0;
    // rope
    float *rope = nullptr;
    // KV pages
    KvLayout kv;
    void **page_table_dev = nullptr;
    std::vector<void *> page_table_host;
    size_t page_bytes = 0;
    int max_context = 0;
    // scratch
    int maxB = 0;
    float *x = nullptr, *xb = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *att = nullptr, *hbuf = nullptr,
          *partial = nullptr, *logits = nullptr, *attn_ws = nullptr, *last_hidden = nullptr, *ln = nullptr, *hbuf2 = nullptr;
    uint16_t *abf = nullptr; // bf16 activations for the tensor-core prefill GEMMs
    bool tc_ok = false;
    int max_splits = 32;
    int32_t *d_tokens = nullptr, *d_positions = nullptr, *d_sessions = nullptr, *d_next = nullptr, *d_hist = nullptr,
            *d_counter = nullptr;
    void *argmax_scratch = nullptr;
    int32_t *h_pinned = nullptr; // pinned staging: tokens | positions | sessions | next
    int hist_cap = 0;
    // graphs keyed by (n, splits, resident)
    std::map<long long, cudaGraphExec_t> graphs;
    std::map<long long, long long> graph_launches; // kernels per graph replay
    // eager-mode event timing of the GEMV launches
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    double last_total_ms = 0, last_gemv_ms = 0;
    bool timing_valid = false;
    int64_t weight_bytes = 0;
    unsigned *fda_done = nullptr;
    // mixture of experts (MoEBlock.java): router + per-expert w1 / w2 / w3, device pointer tables for the indirect GEMV launches
    // GPT-2 family (cfg.arch == JL_ARCH_GPT2): biases per layer [layers][8], wpe and ln_f.bias
    std::vector<DevTensor> aux;
    std::vector<char> aux_set;
    DevTensor gaux[2];
    bool gaux_set[2] = {false, false};
    int n_exp = 0, exp_k = 0;
    std::vector<DevTensor> moe_gate; // [layers]
    std::vector<DevTensor> moe_w;    // [layers][n_exp][3]
    std::vector<char> moe_set;       // [layers][1 + n_exp * 3]
    void **moe_wtab = nullptr;       // device [layers][3][n_exp] weight pointers
    float **moe_stab = nullptr;      // device [layers][3][n_exp] scale pointers
    float *moe_logits = nullptr;     // [maxB][n_exp]
    int32_t *moe_sel = nullptr;      // [maxB][exp_k]
    // persistent decode kernel (jl_pdecode.cu)
    bool pd_ok = false;
    int pd_wdtype = JL_Q4;
    std::vector<PdLayer> pd_layers; // host copy; uploaded to constant memory by jl_launch_pdecode
    unsigned long long *pd_sync = nullptr, *pd_slots = nullptr, *pd_trace = nullptr;
    unsigned *pd_att_done = nullptr;
    int pd_vocab0 = 0, pd_vocab_rows = 0;
    // tensor-parallel exchange buffers: local allocations + every rank's copy opened through CUDA IPC
    uint4 *ll_o = nullptr, *ll_d = nullptr, *ll_a = nullptr;
    // sessions spilled to host memory (jl_model_kv_offload): handle -> (layer page, context page, page bytes) per allocated page
    struct KvSpill {
        std::vector<std::pair<int, int>> pages;
        std::vector<char> bytes;
    };
    std::map<int64_t, KvSpill> spills;
    int64_t next_spill = 1;
    uint4 *peer_ll_o[PD_MAX_TP] = {}, *peer_ll_d[PD_MAX_TP] = {}, *peer_ll_a[PD_MAX_TP] = {};
    float *peer_logits[PD_MAX_TP] = {};
    std::vector<void *> ipc_opened;
};

#define M_CHECK(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != JL_OK) return _rc; \
    } while (0)

// Programmatic dependent launch between the decode kernels: opt in with JL_MODEL_PDL or JL_PDL=1.  Measured on the 8B
// decode step (tools/ktrace.py): launch gaps turn into ~0.6 us overlaps, but the latency-bound attention kernel runs
// ~4 us longer, so the step is no faster (1.76-1.83 ms with, 1.80 ms without).  Every PDL-launched kernel only issues
// read-only weight loads before griddepcontrol.wait.
static bool use_pdl(const jl_model *m) {
    static int env = -1;
    if (env < 0) {
        const char *e = getenv("JL_PDL");
        env = e ? (atoi(e) ? 1 : 0) : 2;
    }
    if (m->cfg.flags & JL_MODEL_NO_PDL) return false;
    if (env != 2) return env == 1;
    return (m->cfg.flags & JL_MODEL_PDL) != 0;
}
static bool use_graph(const jl_model *m) { return !(m->cfg.flags & JL_MODEL_NO_GRAPH); }
static bool use_pd(const jl_model *m) {
    static int env = -1;
    if (env < 0) {
        const char *e = getenv("JL_PERSISTENT");
        env = e ? (atoi(e) ? 1 : 0) : 1;
    }
    return m->pd_ok && env == 1 && !(m->cfg.flags & (JL_MODEL_NO_PERSISTENT | JL_MODEL_NO_GRAPH));
}

extern "C" int jl_model_config_size(void) { return (int)sizeof(jl_model_config); }

extern "C" int jl_model_create(jl_ctx *ctx, const jl_model_config *cfg, jl_model **out) {
    if (!ctx || !cfg || !out) return JL_ERR_INVALID;
    *out = nullptr;
    const jl_model_config &c = *cfg;
    if (c.embedding_length <= 0 || c.num_heads <= 0 || c.num_kv_heads <= 0 || c.num_layers <= 0 || c.vocab_size <= 0 ||
        c.hidden_length <= 0 || c.context_length <= 0 || c.num_heads % c.num_kv_heads)
        return jl_set_error(ctx, JL_ERR_INVALID, "model_create: bad config");
    jl_model *m = new jl_model();
    m->ctx = ctx;
    m->cfg = c;
    if (m->cfg.head_size <= 0) m->cfg.head_size = c.embedding_length / c.num_heads; // Config.java:254
    if (m->cfg.max_batch <= 0) m->cfg.max_batch = 256;
    if (m->cfg.max_sessions <= 0) m->cfg.max_sessions = 1;
    if (m->cfg.tp_size <= 0) m->cfg.tp_size = 1;
    if (m->cfg.rope_scaling == 0.0) m->cfg.rope_scaling = 1.0;
    if (m->cfg.working_qtype != JL_I8 && m->cfg.working_qtype != JL_F32)
        return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "model_create: working_qtype must be JL_I8 or JL_F32"), delete m, JL_ERR_UNSUPPORTED;
    if (m->cfg.kv_dtype != JL_F32 && m->cfg.kv_dtype != JL_BF16)
        return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "model_create: kv_dtype must be JL_F32 or JL_BF16"), delete m, JL_ERR_UNSUPPORTED;
    const int hs = m->cfg.head_size;
    m->group = c.num_heads / c.num_kv_heads;
    // jlama-net limits model shards to the number of KV heads (JlamaService.java:65-68)
    if (m->cfg.tp_size > c.num_kv_heads || c.num_kv_heads % m->cfg.tp_size)
        return jl_set_error(ctx, JL_ERR_INVALID, "model_create: tp_size must divide num_kv_heads"), delete m, JL_ERR_INVALID;
    int rc = jl_dctx_build(c.embedding_length, c.num_heads * hs, c.hidden_length, hs, m->group, c.num_layers, m->cfg.tp_rank,
                           m->cfg.tp_size, 0, 1, &m->d);
    if (rc) return delete m, jl_set_error(ctx, rc, "model_create: bad shard config");
    m->attn_seg = m->d.attentionSegmentLength;
    m->kv_seg = m->d.kvSegmentLength;
    m->h_seg = m->d.hiddenSegmentLength;
    m->heads_local = m->attn_seg / hs;
    m->kv_heads_local = m->kv_seg / hs;
    m->l.resize((size_t)c.num_layers * 9);
    m->l_set.assign((size_t)c.num_layers * 9, 0);
    if (c.arch == JL_ARCH_GPT2) {
        if (m->cfg.tp_size != 1 || c.num_experts > 0 || c.num_heads != c.num_kv_heads)
            return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "model_create: GPT-2 blocks are single-rank, dense, multi-head"), delete m, JL_ERR_UNSUPPORTED;
        m->aux.resize((size_t)c.num_layers * 8);
        m->aux_set.assign((size_t)c.num_layers * 8, 0);
    } else if (c.arch != JL_ARCH_LLAMA)
        return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "model_create: unknown arch %d", c.arch), delete m, JL_ERR_UNSUPPORTED;
    if (c.num_experts > 0) {
        if (c.num_experts > 64 || c.experts_per_token < 1 || c.experts_per_token > c.num_experts || (c.num_experts % m->cfg.tp_size))
            return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "model_create: %d experts / top-%d (max 64 experts, single rank)", c.num_experts,
                                c.experts_per_token),
                   delete m, JL_ERR_UNSUPPORTED;
        m->n_exp = c.num_experts, m->exp_k = c.experts_per_token;
        m->moe_gate.resize(c.num_layers);
        m->moe_w.resize((size_t)c.num_layers * m->n_exp * 3);
        m->moe_set.assign((size_t)c.num_layers * (1 + m->n_exp * 3), 0);
    }
    m->max_context = c.max_context > 0 && c.max_context < c.context_length ? c.max_context : c.context_length;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->models.push_back(m);
    }
    *out = m;
    return JL_OK;
}

extern "C" int jl_model_set_tensor(jl_model *m, int layer, int slot, int64_t tensor_id) {
    if (!m) return JL_ERR_INVALID;
    jl_ctx *ctx = m->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto it = ctx->tensors.find(tensor_id);
    if (it == ctx->tensors.end()) return jl_set_error(ctx, JL_ERR_INVALID, "model_set_tensor: unknown tensor id");
    if (m->finalized) return jl_set_error(ctx, JL_ERR_INVALID, "model_set_tensor: model already finalized");
    DevTensor &t = it->second;
    auto bind = [&](DevTensor &slot_t, bool was_set) {
        if (was_set && slot_t.id) {
            auto old = ctx->tensors.find(slot_t.id);
            if (old != ctx->tensors.end() && old->second.refs > 0) old->second.refs--;
            for (size_t i = 0; i < m->bound_ids.size(); i++)
                if (m->bound_ids[i] == slot_t.id) {
                    m->bound_ids.erase(m->bound_ids.begin() + i);
                    break;
                }
        }
        t.refs++;
        m->bound_ids.push_back(t.id);
        slot_t = t;
    };
    const jl_model_config &c = m->cfg;
    const int E = c.embedding_length;
    auto expect = [&](int64_t rows, int64_t cols) {
        return (t.rows == rows && t.cols == cols)
                   ? JL_OK
                   : jl_set_error(ctx, JL_ERR_INVALID, "model_set_tensor: layer %d slot %d expects [%lld,%lld], got [%lld,%lld]",
                                  layer, slot, (long long)rows, (long long)cols, (long long)t.rows, (long long)t.cols);
    };
    if (layer < 0) {
        if (slot < 0 || slot > 2) return jl_set_error(ctx, JL_ERR_INVALID, "bad global slot");
        if (slot == JL_T_OUT_NORM) {
            M_CHECK(expect(1, E));
            if (t.dtype != JL_F32 && t.dtype != JL_BF16) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "norm weights must be F32/BF16");
        } else
            M_CHECK(expect(c.vocab_size, E));
        bind(m->g[slot], m->g_set[slot]);
        m->g_set[slot] = true;
        return JL_OK;
    }
    if (layer >= c.num_layers || slot < 0 || slot > 8) return jl_set_error(ctx, JL_ERR_INVALID, "bad layer slot");
    switch (slot) {
        case JL_L_ATTN_NORM:
        case JL_L_FFN_NORM:
            M_CHECK(expect(1, E));
            if (t.dtype != JL_F32 && t.dtype != JL_BF16) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "norm weights must be F32/BF16");
            break;
        case JL_L_Q: M_CHECK(expect(m->attn_seg, E)); break;
        case JL_L_K:
        case JL_L_V: M_CHECK(expect(m->kv_seg, E)); break;
        case JL_L_O: M_CHECK(expect(E, m->attn_seg)); break;
        case JL_L_GATE:
        case JL_L_UP: M_CHECK(expect(m->h_seg, E)); break;
        case JL_L_DOWN: M_CHECK(expect(E, m->h_seg)); break;
    }
    bind(m->l[(size_t)layer * 9 + slot], m->l_set[(size_t)layer * 9 + slot] != 0);
    m->l_set[(size_t)layer * 9 + slot] = 1;
    return JL_OK;
}

extern "C" int jl_model_set_expert_tensor(jl_model *m, int layer, int expert, int which, int64_t tensor_id) {
    if (!m) return JL_ERR_INVALID;
    jl_ctx *ctx = m->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (m->n_exp <= 0) return jl_set_error(ctx, JL_ERR_INVALID, "model_set_expert_tensor: the model has no experts");
    if (m->finalized) return jl_set_error(ctx, JL_ERR_INVALID, "model_set_expert_tensor: model already finalized");
    auto it = ctx->tensors.find(tensor_id);
    if (it == ctx->tensors.end()) return jl_set_error(ctx, JL_ERR_INVALID, "model_set_expert_tensor: unknown tensor id");
    if (layer < 0 || layer >= m->cfg.num_layers || expert >= m->n_exp || which < 0 || which > 2)
        return jl_set_error(ctx, JL_ERR_INVALID, "model_set_expert_tensor: bad layer / expert / slot");
    DevTensor &t = it->second;
    const int E = m->cfg.embedding_length, H = m->cfg.hidden_length;
    const int64_t rows = expert < 0 ? m->n_exp : (which == 1 ? E : H), cols = expert < 0 ? E : (which == 1 ? H : E);
    if (t.rows != rows || t.cols != cols)
        return jl_set_error(ctx, JL_ERR_INVALID, "model_set_expert_tensor: expects [%lld,%lld], got [%lld,%lld]", (long long)rows, (long long)cols,
                            (long long)t.rows, (long long)t.cols);
    t.refs++;
    m->bound_ids.push_back(t.id);
    if (expert < 0) {
        m->moe_gate[layer] = t;
        m->moe_set[(size_t)layer * (1 + m->n_exp * 3)] = 1;
    } else {
        m->moe_w[((size_t)layer * m->n_exp + expert) * 3 + which] = t;
        m->moe_set[(size_t)layer * (1 + m->n_exp * 3) + 1 + expert * 3 + which] = 1;
    }
    return JL_OK;
}

extern "C" int jl_model_set_aux_tensor(jl_model *m, int layer, int which, int64_t tensor_id) {
    if (!m) return JL_ERR_INVALID;
    jl_ctx *ctx = m->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    const jl_model_config &c = m->cfg;
    if (c.arch != JL_ARCH_GPT2) return jl_set_error(ctx, JL_ERR_INVALID, "model_set_aux_tensor: only GPT-2 models take bias / position tensors");
    if (m->finalized) return jl_set_error(ctx, JL_ERR_INVALID, "model_set_aux_tensor: model already finalized");
    auto it = ctx->tensors.find(tensor_id);
    if (it == ctx->tensors.end()) return jl_set_error(ctx, JL_ERR_INVALID, "model_set_aux_tensor: unknown tensor id");
    DevTensor &t = it->second;
    if (t.dtype != JL_F32 && t.dtype != JL_BF16) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "model_set_aux_tensor: F32 / BF16 only");
    const int E = c.embedding_length;
    int64_t rows = 1, cols = E;
    if (layer < 0) {
        if (which != JL_AUX_POS_EMBED && which != JL_AUX_OUT_NORM_BIAS) return jl_set_error(ctx, JL_ERR_INVALID, "model_set_aux_tensor: bad global slot");
        if (which == JL_AUX_POS_EMBED) rows = c.context_length;
    } else {
        if (layer >= c.num_layers || which < 0 || which > 7) return jl_set_error(ctx, JL_ERR_INVALID, "model_set_aux_tensor: bad layer slot");
        if (which == JL_AUX_K_BIAS || which == JL_AUX_V_BIAS) cols = m->kv_seg;
        else if (which == JL_AUX_Q_BIAS) cols = m->attn_seg;
        else if (which == JL_AUX_FC_BIAS) cols = m->h_seg;
    }
    if (t.rows != rows || t.cols != cols)
        return jl_set_error(ctx, JL_ERR_INVALID, "model_set_aux_tensor: layer %d slot %d expects [%lld,%lld], got [%lld,%lld]", layer, which,
                            (long long)rows, (long long)cols, (long long)t.rows, (long long)t.cols);
    t.refs++;
    m->bound_ids.push_back(t.id);
    if (layer < 0) m->gaux[which] = t, m->gaux_set[which] = true;
    else m->aux[(size_t)layer * 8 + which] = t, m->aux_set[(size_t)layer * 8 + which] = 1;
    return JL_OK;
}

static int dev_alloc(jl_ctx *ctx, void **p, size_t bytes) {
    if (cudaMalloc(p, bytes ? bytes : 16) != cudaSuccess) {
        cudaGetLastError();
        return jl_set_error(ctx, JL_ERR_OOM, "out of device memory (%zu bytes)", bytes);
    }
    return JL_OK;
}

extern "C" int jl_model_finalize(jl_model *m) {
    if (!m) return JL_ERR_INVALID;
    jl_ctx *ctx = m->ctx;
    if (m->finalized) return JL_OK;
    const jl_model_config &c = m->cfg;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    for (int i = 0; i < 2; i++)
        if (!m->g_set[i]) return jl_set_error(ctx, JL_ERR_INVALID, "model_finalize: global tensor %d missing", i);
    for (size_t i = 0; i < m->l_set.size(); i++) {
        const size_t slot = i % 9;
        if (m->n_exp > 0 && (slot == JL_L_GATE || slot == JL_L_DOWN || slot == JL_L_UP)) continue; // experts instead of a dense MLP
        if (c.arch == JL_ARCH_GPT2 && slot == JL_L_UP) continue;                                      // c_fc -> GELU -> c_proj: no up projection
        if (!m->l_set[i]) return jl_set_error(ctx, JL_ERR_INVALID, "model_finalize: layer %zu slot %zu missing", i / 9, slot);
    }
    // expert parallelism (BASELINE config 5: "expert FFN GEMMs sharded one-per-GPU"): rank r holds the experts e with
    // e % tp_size == r (whole, not row-split); the router is replicated.  Tensors of other ranks' experts stay unset.
    for (size_t i = 0; i < m->moe_set.size(); i++) {
        const size_t per = 1 + (size_t)m->n_exp * 3, slot = i % per;
        const bool mine = slot == 0 || (int)((slot - 1) / 3) % c.tp_size == c.tp_rank;
        if (mine && !m->moe_set[i]) return jl_set_error(ctx, JL_ERR_INVALID, "model_finalize: layer %zu expert tensor %zu missing", i / per, slot);
        if (!mine && m->moe_set[i]) return jl_set_error(ctx, JL_ERR_INVALID, "model_finalize: layer %zu expert tensor %zu belongs to another rank", i / per, slot);
    }
    // the fused QKV and gate+up launches decode all their segments with one weight dtype: a checkpoint that mixes
    // precisions inside a fused group (e.g. Q in Q4, K/V left in BF16) must be rejected, not mis-read
    for (int L = 0; L < c.num_layers; L++) {
        const DevTensor *lw = &m->l[(size_t)L * 9];
        if (lw[JL_L_K].dtype != lw[JL_L_Q].dtype || lw[JL_L_V].dtype != lw[JL_L_Q].dtype)
            return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "model_finalize: layer %d q/k/v weights must share one dtype (got %d/%d/%d)", L,
                                lw[JL_L_Q].dtype, lw[JL_L_K].dtype, lw[JL_L_V].dtype);
        if (m->n_exp == 0 && c.arch != JL_ARCH_GPT2 && lw[JL_L_UP].dtype != lw[JL_L_GATE].dtype)
            return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "model_finalize: layer %d gate/up weights must share one dtype (got %d/%d)", L,
                                lw[JL_L_GATE].dtype, lw[JL_L_UP].dtype);
    }
    if (c.arch == JL_ARCH_GPT2) {
        for (size_t i = 0; i < m->aux_set.size(); i++)
            if (!m->aux_set[i]) return jl_set_error(ctx, JL_ERR_INVALID, "model_finalize: layer %zu bias slot %zu missing", i / 8, i % 8);
        if (!m->gaux_set[0] || !m->gaux_set[1]) return jl_set_error(ctx, JL_ERR_INVALID, "model_finalize: wpe / ln_f.bias missing");
    }
    const int E = c.embedding_length, hs = c.head_size;
    JL_CUDA_CHECK(ctx, cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    // RoPE table (Config.java:271-276) padded by 2*kv_heads positions: head h reads position pos+2*kvh
    // (CausalSelfAttention.java:260-268); the reference would throw past the end of its table.
    {
        const int positions = c.context_length + 2 * c.num_kv_heads;
        std::vector<float> tbl((size_t)positions * (hs / 2) * 2);
        if (c.arch == JL_ARCH_GPT2) // no rotary embedding: the identity rotation (cos 1, sin 0) leaves q and k bit for bit
            for (size_t i = 0; i < tbl.size(); i += 2) tbl[i] = 1.0f, tbl[i + 1] = 0.0f;
        else jl_precompute_freqs_cis(hs, positions, c.rope_theta, c.rope_scaling, tbl.data());
        M_CHECK(dev_alloc(ctx, (void **)&m->rope, tbl.size() * 4));
        JL_CUDA_CHECK(ctx, cudaMemcpy(m->rope, tbl.data(), tbl.size() * 4, cudaMemcpyHostToDevice));
    }
    // KV geometry (KvBufferCache.java:99-112,224-280)
    const int kv_esz = c.kv_dtype == JL_F32 ? 4 : 2;
    int lpp, cpp;
    if (jl_kv_page_geometry(c.num_layers, c.context_length, m->kv_seg, kv_esz, 1 << 23, &lpp, &cpp))
        return jl_set_error(ctx, JL_ERR_INVALID, "kv page geometry");
    m->kv.layers_per_page = lpp;
    m->kv.ctx_per_page = cpp;
    m->kv.kv_len = m->kv_seg;
    m->kv.n_layer_pages = (c.num_layers + lpp - 1) / lpp;
    m->kv.n_ctx_pages = (m->max_context + cpp - 1) / cpp;
    m->kv.kv_dtype = c.kv_dtype;
    m->page_bytes = (size_t)lpp * 2 * cpp * m->kv_seg * kv_esz;
    const size_t nent = (size_t)c.max_sessions * m->kv.n_layer_pages * m->kv.n_ctx_pages;
    m->page_table_host.assign(nent, nullptr);
    M_CHECK(dev_alloc(ctx, (void **)&m->page_table_dev, nent * sizeof(void *)));
    JL_CUDA_CHECK(ctx, cudaMemset(m->page_table_dev, 0, nent * sizeof(void *)));
    m->kv.page_table = m->page_table_dev;
    // scratch
    m->maxB = c.max_batch > c.max_sessions ? c.max_batch : c.max_sessions;
    const size_t B = m->maxB;
    M_CHECK(dev_alloc(ctx, (void **)&m->x, B * E * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->xb, B * E * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->q, B * m->attn_seg * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->k, B * m->kv_seg * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->v, B * m->kv_seg * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->att, B * m->attn_seg * 4));
    // experts are held whole (expert parallelism), so their hidden activations need the full hidden length
    M_CHECK(dev_alloc(ctx, (void **)&m->hbuf, B * (size_t)(m->n_exp > 0 ? c.hidden_length : m->h_seg) * 4));
    if (m->n_exp > 0) {
        M_CHECK(dev_alloc(ctx, (void **)&m->moe_logits, B * m->n_exp * 4));
        M_CHECK(dev_alloc(ctx, (void **)&m->moe_sel, B * m->exp_k * 4));
        const size_t ne = (size_t)c.num_layers * 3 * m->n_exp;
        std::vector<void *> wt(ne);
        std::vector<float *> st(ne);
        int wd0 = m->moe_w[(size_t)c.tp_rank * 3].dtype; // expert tp_rank is the first one this rank holds
        for (int L = 0; L < c.num_layers; L++)
            for (int w = 0; w < 3; w++)
                for (int e = 0; e < m->n_exp; e++) {
                    const DevTensor &t = m->moe_w[((size_t)L * m->n_exp + e) * 3 + w];
                    if (e % c.tp_size != c.tp_rank) { // another rank's expert: null table entries, the GEMV contributes nothing
                        wt[((size_t)L * 3 + w) * m->n_exp + e] = nullptr;
                        st[((size_t)L * 3 + w) * m->n_exp + e] = nullptr;
                        continue;
                    }
                    if (t.dtype != wd0 || (wd0 != JL_Q4 && wd0 != JL_I8))
                        return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "model_finalize: expert weights must share one quantised dtype (Q4 or I8)");
                    wt[((size_t)L * 3 + w) * m->n_exp + e] = t.data;
                    st[((size_t)L * 3 + w) * m->n_exp + e] = t.scales;
                }
        M_CHECK(dev_alloc(ctx, (void **)&m->moe_wtab, ne * sizeof(void *)));
        M_CHECK(dev_alloc(ctx, (void **)&m->moe_stab, ne * sizeof(float *)));
        JL_CUDA_CHECK(ctx, cudaMemcpy(m->moe_wtab, wt.data(), ne * sizeof(void *), cudaMemcpyHostToDevice));
        JL_CUDA_CHECK(ctx, cudaMemcpy(m->moe_stab, st.data(), ne * sizeof(float *), cudaMemcpyHostToDevice));
    }
    M_CHECK(dev_alloc(ctx, (void **)&m->partial, B * E * 4));
    if (c.arch == JL_ARCH_GPT2) M_CHECK(dev_alloc(ctx, (void **)&m->ln, B * E * 4)); // LayerNorm output (its own tensor in the reference too)
    if (c.prefill_tensor_core && c.arch == JL_ARCH_LLAMA) {
        size_t kmax = E > m->h_seg ? E : m->h_seg;
        if ((size_t)m->attn_seg > kmax) kmax = m->attn_seg;
        M_CHECK(dev_alloc(ctx, (void **)&m->ln, B * E * 4));
        M_CHECK(dev_alloc(ctx, (void **)&m->hbuf2, B * m->h_seg * 4));
        M_CHECK(dev_alloc(ctx, (void **)&m->abf, B * kmax * 2));
        bool ok = m->n_exp == 0 && c.arch == JL_ARCH_LLAMA && c.tp_size == 1 && (E % 128) == 0 && (m->h_seg % 128) == 0 && (m->attn_seg % 128) == 0 &&
                  (m->kv_seg % 128) == 0;
        for (int L = 0; L < c.num_layers && ok; L++)
            for (int sl : {JL_L_Q, JL_L_K, JL_L_V, JL_L_O, JL_L_GATE, JL_L_DOWN, JL_L_UP}) {
                const int dt = m->l[(size_t)L * 9 + sl].dtype;
                ok = ok && (dt == JL_Q4 || dt == JL_I8); // both block formats are dequantised into the BF16 weight tile
            }
        m->tc_ok = ok;
    }
    M_CHECK(dev_alloc(ctx, (void **)&m->logits, (size_t)c.max_sessions * c.vocab_size * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->last_hidden, (size_t)c.max_sessions * E * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->attn_ws, B * m->heads_local * m->max_splits * (hs + 2) * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->d_tokens, B * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->d_positions, B * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->d_sessions, B * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->d_next, B * 4));
    m->hist_cap = 1 << 16;
    M_CHECK(dev_alloc(ctx, (void **)&m->d_hist, (size_t)m->hist_cap * 4));
    M_CHECK(dev_alloc(ctx, (void **)&m->d_counter, 4));
    M_CHECK(dev_alloc(ctx, &m->argmax_scratch, jl_argmax_scratch_bytes(c.max_sessions)));
    M_CHECK(dev_alloc(ctx, (void **)&m->fda_done, (size_t)m->maxB * m->kv_heads_local * sizeof(unsigned)));
    JL_CUDA_CHECK(ctx, cudaMemset(m->fda_done, 0, (size_t)m->maxB * m->kv_heads_local * sizeof(unsigned)));
    JL_CUDA_CHECK(ctx, cudaMallocHost((void **)&m->h_pinned, B * 4 * 4));
    JL_CUDA_CHECK(ctx, cudaEventCreate(&m->ev_begin));
    JL_CUDA_CHECK(ctx, cudaEventCreate(&m->ev_end));
    // algorithmic decode bytes per token on this rank: every linear weight once + lm_head
    auto tb = [](const DevTensor &t) { return (int64_t)t.bytes; };
    int64_t wb = 0;
    for (int L = 0; L < c.num_layers; L++)
        for (int s : {JL_L_Q, JL_L_K, JL_L_V, JL_L_O, JL_L_GATE, JL_L_DOWN, JL_L_UP})
            if (m->l_set[(size_t)L * 9 + s]) wb += tb(m->l[(size_t)L * 9 + s]);
    // mixture of experts: the router plus the experts_per_token selected experts are streamed per token
    for (int L = 0; L < c.num_layers && m->n_exp > 0; L++) {
        wb += tb(m->moe_gate[L]);
        // expert parallel: on average experts_per_token / tp_size of them live on this rank
        for (int w = 0; w < 3; w++) wb += (int64_t)m->exp_k * tb(m->moe_w[((size_t)L * m->n_exp + c.tp_rank) * 3 + w]) / c.tp_size;
    }
    wb += tb(m->g_set[JL_T_LM_HEAD] ? m->g[JL_T_LM_HEAD] : m->g[JL_T_EMBED]);
    m->weight_bytes = wb;
    // ---- persistent decode kernel eligibility: Q8 activations, every linear weight (and the lm_head) in ONE quantised dtype ----
    {
        const DevTensor &head = m->g_set[JL_T_LM_HEAD] ? m->g[JL_T_LM_HEAD] : m->g[JL_T_EMBED];
        const int wd = m->l[JL_L_Q].dtype;
        bool ok = m->n_exp == 0 && c.arch == JL_ARCH_LLAMA && c.working_qtype == JL_I8 && (wd == JL_Q4 || wd == JL_I8) && head.dtype == wd;
        for (int L = 0; L < c.num_layers && ok; L++)
            for (int sl : {JL_L_Q, JL_L_K, JL_L_V, JL_L_O, JL_L_GATE, JL_L_DOWN, JL_L_UP}) ok = ok && m->l[(size_t)L * 9 + sl].dtype == wd;
        // lm_head rows of this rank (vocabulary-sharded under tensor parallelism: SURVEY 8e; the reference computes the
        // full GEMV on the coordinator, AbstractModel.java:444-449)
        const int W = c.tp_size, V = c.vocab_size;
        m->pd_vocab0 = (int)(((long long)V * c.tp_rank) / W);
        m->pd_vocab_rows = (int)(((long long)V * (c.tp_rank + 1)) / W) - m->pd_vocab0;
        if (ok && W > 1 && !ctx->nccl_comm) ok = false; // the IPC handles travel over the communicator
        if (ok) {
            PdParams probe = {};
            probe.layers = c.num_layers, probe.E = E, probe.H = m->h_seg, probe.attn_seg = m->attn_seg, probe.kv_seg = m->kv_seg;
            probe.heads = m->heads_local, probe.kv_heads = m->kv_heads_local, probe.head_size = hs, probe.world = W;
            ok = jl_pdecode_supported(probe, wd, ctx->sm_count);
            // flat decode attention: at most 1024 positions per (kv head, split) task
            int cap = ctx->sm_count / (m->kv_heads_local > 0 ? m->kv_heads_local : 1);
            if (cap > m->max_splits) cap = m->max_splits;
            if (cap < 1 || (m->max_context + cap - 1) / cap > 1024 - 32) ok = false;
        }
        if (ok) {
            std::vector<PdLayer> pl(c.num_layers);
            const int order[7] = {JL_L_Q, JL_L_K, JL_L_V, JL_L_O, JL_L_GATE, JL_L_UP, JL_L_DOWN};
            for (int L = 0; L < c.num_layers; L++) {
                for (int i = 0; i < 7; i++) {
                    pl[L].w[i] = (const uint8_t *)m->l[(size_t)L * 9 + order[i]].data;
                    pl[L].s[i] = m->l[(size_t)L * 9 + order[i]].scales;
                }
                pl[L].attn_norm = m->l[(size_t)L * 9 + JL_L_ATTN_NORM].data;
                pl[L].attn_norm_dt = m->l[(size_t)L * 9 + JL_L_ATTN_NORM].dtype;
                pl[L].ffn_norm = m->l[(size_t)L * 9 + JL_L_FFN_NORM].data;
                pl[L].ffn_norm_dt = m->l[(size_t)L * 9 + JL_L_FFN_NORM].dtype;
            }
            m->pd_layers = pl;
            M_CHECK(dev_alloc(ctx, (void **)&m->pd_sync, PD_SYNC_WORDS * 8));
            JL_CUDA_CHECK(ctx, cudaMemset(m->pd_sync, 0, PD_SYNC_WORDS * 8));
            M_CHECK(dev_alloc(ctx, (void **)&m->pd_slots, (size_t)ctx->sm_count * 8));
            M_CHECK(dev_alloc(ctx, (void **)&m->pd_att_done, (size_t)m->kv_heads_local * sizeof(unsigned)));
            JL_CUDA_CHECK(ctx, cudaMemset(m->pd_att_done, 0, (size_t)m->kv_heads_local * sizeof(unsigned)));
            m->pd_wdtype = wd;
            if (W > 1) {
                // LL receive buffers of this rank and the peers' copies (cudaIpc: one process per GPU)
                const size_t lines = (size_t)W * E / 2 + 8;
                M_CHECK(dev_alloc(ctx, (void **)&m->ll_o, lines * 16));
                M_CHECK(dev_alloc(ctx, (void **)&m->ll_d, lines * 16));
                M_CHECK(dev_alloc(ctx, (void **)&m->ll_a, (size_t)PD_MAX_TP * 16));
                JL_CUDA_CHECK(ctx, cudaMemset(m->ll_o, 0, lines * 16));
                JL_CUDA_CHECK(ctx, cudaMemset(m->ll_d, 0, lines * 16));
                JL_CUDA_CHECK(ctx, cudaMemset(m->ll_a, 0, (size_t)PD_MAX_TP * 16));
                void *local[4] = {m->ll_o, m->ll_d, m->ll_a, m->logits};
                cudaIpcMemHandle_t mine[4];
                for (int i = 0; i < 4; i++) JL_CUDA_CHECK(ctx, cudaIpcGetMemHandle(&mine[i], local[i]));
                cudaIpcMemHandle_t *d_send = nullptr, *d_recv = nullptr;
                M_CHECK(dev_alloc(ctx, (void **)&d_send, sizeof mine));
                M_CHECK(dev_alloc(ctx, (void **)&d_recv, sizeof mine * W));
                JL_CUDA_CHECK(ctx, cudaMemcpy(d_send, mine, sizeof mine, cudaMemcpyHostToDevice));
                int rc = jl_comm_allgather_dev(ctx, m->stream, d_send, d_recv, sizeof mine);
                if (rc == JL_OK && cudaStreamSynchronize(m->stream) != cudaSuccess) rc = jl_set_error(ctx, JL_ERR_CUDA, "handle exchange failed");
                std::vector<cudaIpcMemHandle_t> all((size_t)4 * W);
                if (rc == JL_OK) JL_CUDA_CHECK(ctx, cudaMemcpy(all.data(), d_recv, sizeof mine * W, cudaMemcpyDeviceToHost));
                cudaFree(d_send);
                cudaFree(d_recv);
                if (rc != JL_OK) return rc;
                for (int r = 0; r < W; r++) {
                    void *ptrs[4];
                    for (int i = 0; i < 4; i++) {
                        if (r == c.tp_rank) {
                            ptrs[i] = local[i];
                        } else {
                            cudaError_t e = cudaIpcOpenMemHandle(&ptrs[i], all[(size_t)r * 4 + i], cudaIpcMemLazyEnablePeerAccess);
                            if (e != cudaSuccess)
                                return jl_set_error(ctx, JL_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d): %s (peer access over NVLink is required)", r,
                                                    cudaGetErrorString(e));
                            m->ipc_opened.push_back(ptrs[i]);
                        }
                    }
                    m->peer_ll_o[r] = (uint4 *)ptrs[0], m->peer_ll_d[r] = (uint4 *)ptrs[1], m->peer_ll_a[r] = (uint4 *)ptrs[2];
                    m->peer_logits[r] = (float *)ptrs[3];
                }
            }
            if (getenv("JL_PD_TRACE")) {
                M_CHECK(dev_alloc(ctx, (void **)&m->pd_trace, ((size_t)c.num_layers * 16 + 32) * 8));
                JL_CUDA_CHECK(ctx, cudaMemset(m->pd_trace, 0, ((size_t)c.num_layers * 16 + 32) * 8));
            }
            m->pd_ok = true;
            // per-token stream of this rank: the lm_head is vocabulary-sharded on the persistent path
            if (W > 1) m->weight_bytes += -tb(head) + (int64_t)((double)tb(head) * m->pd_vocab_rows / V);
        }
    }
    m->finalized = true;
    return JL_OK;
}

extern "C" int64_t jl_model_weight_bytes(jl_model *m) { return m ? m->weight_bytes : -1; }

int jl_model_num_experts(jl_model *m) { return m ? m->n_exp : 0; } // jl_model_load_safetensors

// limits the session scheduler (jl_sched.cu) plans against: {max_sessions, reserved context, max_batch, rows per decode call}
void jl_model_limits(jl_model *m, int out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0;
    if (!m || !m->finalized) return;
    out[0] = m->cfg.max_sessions, out[1] = m->max_context, out[2] = m->cfg.max_batch;
    out[3] = m->cfg.max_sessions < GEMV_MAX_M ? m->cfg.max_sessions : GEMV_MAX_M;
}

// The timeline buffer was replaced (jl_debug_ktrace): drop every captured decode graph, they are re-captured on demand.
void jl_models_invalidate_graphs(jl_ctx *ctx) {
    for (jl_model *m : ctx->models) {
        if (m->stream) cudaStreamSynchronize(m->stream);
        for (auto &kv : m->graphs) cudaGraphExecDestroy(kv.second);
        m->graphs.clear();
        m->graph_launches.clear();
    }
}

extern "C" int jl_model_free(jl_model *m) {
    if (!m) return JL_ERR_INVALID;
    jl_ctx *ctx = m->ctx;
    cudaSetDevice(ctx->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        for (int64_t id : m->bound_ids) {
            auto it = ctx->tensors.find(id);
            if (it != ctx->tensors.end() && it->second.refs > 0) it->second.refs--;
        }
        for (size_t i = 0; i < ctx->models.size(); i++)
            if (ctx->models[i] == m) {
                ctx->models.erase(ctx->models.begin() + i);
                break;
            }
    }
    jl_pdecode_forget(ctx->device, m);
    for (auto &kv : m->graphs) cudaGraphExecDestroy(kv.second);
    for (void *p : m->ipc_opened) cudaIpcCloseMemHandle(p);
    for (void *p : m->page_table_host)
        if (p) cudaFree(p);
    void *bufs[] = {m->ln, m->hbuf2, m->abf, m->rope, m->page_table_dev, m->x, m->xb, m->q, m->k, m->v, m->att, m->hbuf, m->partial, m->logits,
                    m->last_hidden, m->attn_ws, m->d_tokens, m->d_positions, m->d_sessions, m->d_next, m->d_hist, m->d_counter,
                    m->argmax_scratch, m->fda_done, m->moe_wtab, m->moe_stab, m->moe_logits, m->moe_sel, m->pd_sync, m->pd_slots, m->pd_att_done, m->pd_trace, m->ll_o, m->ll_d, m->ll_a};
    for (void *p : bufs)
        if (p) cudaFree(p);
    if (m->h_pinned) cudaFreeHost(m->h_pinned);
    for (auto e : m->ev_pool) cudaEventDestroy(e);
    if (m->ev_begin) cudaEventDestroy(m->ev_begin);
    if (m->ev_end) cudaEventDestroy(m->ev_end);
    if (m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return JL_OK;
}

extern "C" int jl_model_reset_session(jl_model *m, int session) {
    if (!m || !m->finalized || session < 0 || session >= m->cfg.max_sessions) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
    // pages stay allocated (a KvBuffer keeps its pages until close); contents are dead once positions restart.
    // Zero them so a fresh session starts from the reference's zero-initialised page (TensorCache.java:105).
    const size_t per = (size_t)m->kv.n_layer_pages * m->kv.n_ctx_pages;
    for (size_t i = 0; i < per; i++) {
        void *p = m->page_table_host[(size_t)session * per + i];
        if (p) JL_CUDA_CHECK(ctx, cudaMemsetAsync(p, 0, m->page_bytes, m->stream));
    }
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
    return JL_OK;
}

// make sure pages for positions [p0, p1] of `session` exist (KvBufferCache.java:307-318 lazy pages)
static int ensure_pages(jl_model *m, int session, int p0, int p1) {
    jl_ctx *ctx = m->ctx;
    if (p1 >= m->max_context)
        return jl_set_error(ctx, JL_ERR_INVALID, "position %d exceeds reserved context %d", p1, m->max_context);
    bool changed = false;
    for (int cp = p0 / m->kv.ctx_per_page; cp <= p1 / m->kv.ctx_per_page; cp++)
        for (int lp = 0; lp < m->kv.n_layer_pages; lp++) {
            size_t idx = ((size_t)session * m->kv.n_layer_pages + lp) * m->kv.n_ctx_pages + cp;
            if (!m->page_table_host[idx]) {
                void *p = nullptr;
                M_CHECK(dev_alloc(ctx, &p, m->page_bytes));
                JL_CUDA_CHECK(ctx, cudaMemsetAsync(p, 0, m->page_bytes, m->stream));
                m->page_table_host[idx] = p;
                JL_CUDA_CHECK(ctx, cudaMemcpyAsync(&m->page_table_dev[idx], &m->page_table_host[idx], sizeof(void *),
                                                   cudaMemcpyHostToDevice, m->stream));
                changed = true;
            }
        }
    if (changed) JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream)); // host staging of the pointer must outlive the copy
    return JL_OK;
}

static cudaEvent_t next_event(jl_model *m) {
    if (m->ev_used == m->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        m->ev_pool.push_back(e);
    }
    return m->ev_pool[m->ev_used++];
}

// one quantised/dense GEMM call site; M rows of `a` against W; splits M into GEMV_MAX_M chunks.
// `timed`: bracket with events (eager mode only) for the roofline numerator.
static int run_gemm(jl_model *m, GemvParams p, int prologue, int epilogue, int M, size_t a_row_bytes, bool timed,
                    bool allow_pdl = true) {
    jl_ctx *ctx = m->ctx;
    if (timed) cudaEventRecord(next_event(m), m->stream);
    // rows per launch: as many as the staged activations allow (F32 activations at K = 14336 fit 2 rows, Q8 all 8)
    const int chunk = jl_gemv_max_m(p.w_dtype, prologue, p.K);
    if (chunk < 1) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "gemm: K=%d does not fit shared memory", p.K);
    for (int m0 = 0; m0 < M; m0 += chunk) {
        GemvParams c = p;
        c.M = M - m0 < chunk ? M - m0 : chunk;
        c.a = (const char *)p.a + (size_t)m0 * a_row_bytes;
        for (int s = 0; s < c.nseg; s++) c.seg[s].out = p.seg[s].out + (size_t)m0 * p.seg[s].out_ld;
        if (c.residual) c.residual = p.residual + (size_t)m0 * p.res_ld;
        M_CHECK(jl_launch_gemv(ctx, m->stream, c, prologue, epilogue, allow_pdl && use_pdl(m)));
    }
    if (timed) cudaEventRecord(next_event(m), m->stream);
    return JL_OK;
}

static void set_w(GemvParams &p, int seg, const DevTensor &t, float *out, int out_ld) {
    p.seg[seg].w = t.data;
    p.seg[seg].ws = t.scales;
    p.seg[seg].out = out;
    p.seg[seg].rows = (int)t.rows;
    p.seg[seg].out_ld = out_ld;
    p.seg[seg].out_off = 0;
}

// AbstractModel.forward (:314-329) over M rows whose tokens/positions/sessions are already on the device.
static int forward_rows_gpt2(jl_model *m, int M, int max_pos, int splits, bool distinct_sessions);

static int forward_rows(jl_model *m, int M, int max_pos, int splits, bool timed, bool distinct_sessions = false) {
    if (m->cfg.arch == JL_ARCH_GPT2) return forward_rows_gpt2(m, M, max_pos, splits, distinct_sessions);
    jl_ctx *ctx = m->ctx;
    const jl_model_config &c = m->cfg;
    const int E = c.embedding_length, hs = c.head_size;
    const bool q8 = c.working_qtype == JL_I8;
    const bool pdl = use_pdl(m);
    M_CHECK(jl_launch_embed(ctx, m->stream, m->g[JL_T_EMBED], m->d_tokens, M, m->x, E));
    for (int L = 0; L < c.num_layers; L++) {
        const DevTensor *lw = &m->l[(size_t)L * 9];
        // Q8 activations only pair with Q4/I8 weights (AbstractModel.java:119-176)
        auto act_q = [&](const DevTensor &w) { return q8 && (w.dtype == JL_Q4 || w.dtype == JL_I8); };
        // ---- pre-norm + quantise + QKV (TransformerBlock.java:170-175, CausalSelfAttention.java:161-171) ----
        {
            GemvParams p = {};
            p.nseg = 3;
            set_w(p, 0, lw[JL_L_Q], m->q, m->attn_seg);
            set_w(p, 1, lw[JL_L_K], m->k, m->kv_seg);
            set_w(p, 2, lw[JL_L_V], m->v, m->kv_seg);
            // k and v outputs are indexed by their own local row: out col = local
            p.w_dtype = lw[JL_L_Q].dtype;
            p.ldw = E;
            p.K = E;
            p.a = m->x;
            p.lda = E;
            p.norm_w = lw[JL_L_ATTN_NORM].data;
            p.norm_w_dtype = lw[JL_L_ATTN_NORM].dtype;
            p.norm_adj = 0.0f;
            p.norm_eps = c.layer_norm_eps;
            p.norm_E = E;
            p.total_rows = m->attn_seg + 2 * m->kv_seg;
            M_CHECK(run_gemm(m, p, act_q(lw[JL_L_Q]) ? PRO_RMSNORM_QUANT : PRO_RMSNORM_F32, EPI_STORE, M, (size_t)E * 4, timed));
        }
        // ---- KV append + RoPE, attention (CausalSelfAttention.java:199-356) ----
        AttnParams ap = {};
        ap.kv = m->kv;
        ap.layer = L;
        ap.heads = m->heads_local;
        ap.kv_heads = m->kv_heads_local;
        ap.head_size = hs;
        ap.head0_global = m->d.headStart;
        ap.kv_head0_global = m->d.groupHeadStart;
        ap.q = m->q;
        ap.k = m->k;
        ap.v = m->v;
        ap.q_ld = m->attn_seg;
        ap.kv_ld = m->kv_seg;
        ap.out = m->att;
        ap.rope = m->rope;
        ap.rows = M;
        ap.sessions = m->d_sessions;
        ap.positions = m->d_positions;
        ap.scale = (float)(1.0 / sqrt((double)hs)); // CausalSelfAttention.java:134
        ap.ws = m->attn_ws;
        ap.splits = splits;
        if (distinct_sessions) {
            // decode: one fused kernel (RoPE + KV append + attention); `splits` = about one per 64 positions
            // the launch may be replayed from a graph for every context of its split bucket (run_decode): choose the kernel by the
            // largest position the bucket can see, not by the current one
            const int bound = splits >= m->max_splits ? m->max_context - 1 : splits * 64 * M - 1;
            M_CHECK(jl_launch_fused_decode_attention(ctx, m->stream, ap, m->fda_done, pdl, bound));
        } else {
            M_CHECK(jl_launch_rope_kv_append(ctx, m->stream, ap, m->q, pdl));
            M_CHECK(jl_launch_paged_attention(ctx, m->stream, ap, max_pos, pdl));
        }
        // ---- o_proj (+ reducer) + residual (CausalSelfAttention.java:363-378, TransformerBlock.java:185) ----
        {
            GemvParams p = {};
            p.nseg = 1;
            const bool tp = c.tp_size > 1;
            set_w(p, 0, lw[JL_L_O], tp ? m->partial : m->xb, E);
            p.w_dtype = lw[JL_L_O].dtype;
            p.ldw = m->attn_seg;
            p.K = m->attn_seg;
            p.a = m->att;
            p.lda = m->attn_seg;
            p.residual = m->x;
            p.res_ld = E;
            p.total_rows = E;
            // never an early launch for o_proj: its weight requests would flood the memory system under the latency-bound
            // attention kernel
            M_CHECK(run_gemm(m, p, act_q(lw[JL_L_O]) ? PRO_F32_QUANT : PRO_F32, tp ? EPI_STORE : EPI_ADD_RESIDUAL, M,
                             (size_t)m->attn_seg * 4, timed, false));
            if (tp) {
                M_CHECK(jl_comm_allreduce_dev(ctx, m->stream, m->partial, (size_t)M * E));
                // xb = reduced + x
                JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->xb, m->partial, (size_t)M * E * 4, cudaMemcpyDeviceToDevice, m->stream));
                M_CHECK(jl_launch_accumulate(ctx, m->stream, m->xb, M, E, JL_F32, m->x, nullptr, M, E, 0, E));
            }
        }
        if (m->n_exp > 0) {
            // ---- mixture of experts (MoEBlock.java:73-149), one row at a time: router GEMV -> softmax + top-k on the device ->
            // per selected expert gate/up (SiLU*up fused) and down_proj through the expert pointer tables; the expert results
            // are summed UNWEIGHTED in selection order (:139-143), then the residual is added (TransformerBlock.java:203).
            // Row b of the result is the sum over row b's experts (the reference's copy of expert 0's result into row 0 for
            // every batch row, MoEBlock.java:141, is a bug we do not reproduce; with one row per call they coincide).
            // Expert parallel (tp_size > 1): every rank routes identically (replicated router), a selected expert held by another
            // rank has a null table entry and its launches only write zeros / pass the running sum through (gemv_absent); the
            // per-rank sums then meet in one all-reduce.  With two experts per token that sum is exact in any order.
            const int NE = m->n_exp, KE = m->exp_k, H = c.hidden_length;
            void **wt = m->moe_wtab + (size_t)L * 3 * NE;
            float **st = m->moe_stab + (size_t)L * 3 * NE;
            const size_t own = ((size_t)L * NE + c.tp_rank) * 3; // first expert held by this rank: shapes / dtype of every expert
            const bool aq = act_q(m->moe_w[own]);
            for (int b = 0; b < M; b++) {
                {
                    GemvParams p = {};
                    p.nseg = 1;
                    set_w(p, 0, m->moe_gate[L], m->moe_logits + (size_t)b * NE, NE);
                    p.w_dtype = m->moe_gate[L].dtype, p.ldw = E, p.K = E;
                    p.a = m->xb + (size_t)b * E, p.lda = E;
                    p.norm_w = lw[JL_L_FFN_NORM].data, p.norm_w_dtype = lw[JL_L_FFN_NORM].dtype, p.norm_eps = c.layer_norm_eps, p.norm_E = E;
                    p.total_rows = NE;
                    M_CHECK(run_gemm(m, p, act_q(m->moe_gate[L]) ? PRO_RMSNORM_QUANT : PRO_RMSNORM_F32, EPI_STORE, 1, (size_t)E * 4, false, false));
                }
                M_CHECK(jl_launch_moe_route(ctx, m->stream, m->moe_logits + (size_t)b * NE, 1, NE, KE, m->moe_sel + (size_t)b * KE));
                for (int i = 0; i < KE; i++) {
                    const DevTensor &proto = m->moe_w[own];
                    GemvParams p = {};
                    p.nseg = 2;
                    set_w(p, 0, proto, m->hbuf, H);
                    set_w(p, 1, proto, m->hbuf, H);
                    p.sel = m->moe_sel + (size_t)b * KE + i;
                    p.w_tab[0] = (const void *const *)(wt + 0 * NE), p.ws_tab[0] = (const float *const *)(st + 0 * NE); // w1
                    p.w_tab[1] = (const void *const *)(wt + 2 * NE), p.ws_tab[1] = (const float *const *)(st + 2 * NE); // w3
                    p.w_dtype = proto.dtype, p.ldw = E, p.K = E;
                    p.a = m->xb + (size_t)b * E, p.lda = E;
                    p.norm_w = lw[JL_L_FFN_NORM].data, p.norm_w_dtype = lw[JL_L_FFN_NORM].dtype, p.norm_eps = c.layer_norm_eps, p.norm_E = E;
                    p.total_rows = H;
                    // (no programmatic early launch: the selection is read before any dependency wait)
                    M_CHECK(run_gemm(m, p, aq ? PRO_RMSNORM_QUANT : PRO_RMSNORM_F32, EPI_SILU_MUL, 1, (size_t)E * 4, false, false));
                    GemvParams d = {};
                    d.nseg = 1;
                    const DevTensor &dproto = m->moe_w[own + 1];
                    set_w(d, 0, dproto, m->x + (size_t)b * E, E);
                    d.sel = p.sel;
                    d.w_tab[0] = (const void *const *)(wt + 1 * NE), d.ws_tab[0] = (const float *const *)(st + 1 * NE); // w2
                    d.w_dtype = dproto.dtype, d.ldw = H, d.K = H;
                    d.a = m->hbuf, d.lda = H;
                    d.residual = m->x + (size_t)b * E, d.res_ld = E; // expert i > 0 accumulates onto the sum so far
                    d.total_rows = E;
                    M_CHECK(run_gemm(m, d, aq ? PRO_F32_QUANT : PRO_F32, i == 0 ? EPI_STORE : EPI_ADD_RESIDUAL, 1, (size_t)H * 4, false, false));
                }
            }
            if (c.tp_size > 1) M_CHECK(jl_comm_allreduce_dev(ctx, m->stream, m->x, (size_t)M * E));
            M_CHECK(jl_launch_accumulate(ctx, m->stream, m->x, M, E, JL_F32, m->xb, nullptr, M, E, 0, E)); // + residual
            continue;
        }
        // ---- pre-FF norm + quantise + gate/up + SiLU*up (TransformerBlock.java:187-196, MLPBlock.java:117-141) ----
        {
            GemvParams p = {};
            p.nseg = 2;
            set_w(p, 0, lw[JL_L_GATE], m->hbuf, m->h_seg);
            set_w(p, 1, lw[JL_L_UP], m->hbuf, m->h_seg);
            p.w_dtype = lw[JL_L_GATE].dtype;
            p.ldw = E;
            p.K = E;
            p.a = m->xb;
            p.lda = E;
            p.norm_w = lw[JL_L_FFN_NORM].data;
            p.norm_w_dtype = lw[JL_L_FFN_NORM].dtype;
            p.norm_eps = c.layer_norm_eps;
            p.norm_E = E;
            p.total_rows = m->h_seg;
            M_CHECK(run_gemm(m, p, act_q(lw[JL_L_GATE]) ? PRO_RMSNORM_QUANT : PRO_RMSNORM_F32, EPI_SILU_MUL, M, (size_t)E * 4, timed));
        }
        // ---- down_proj (+ reducer) + residual (MLPBlock.java:144-160, TransformerBlock.java:203) ----
        {
            GemvParams p = {};
            p.nseg = 1;
            const bool tp = c.tp_size > 1;
            set_w(p, 0, lw[JL_L_DOWN], tp ? m->partial : m->x, E);
            p.w_dtype = lw[JL_L_DOWN].dtype;
            p.ldw = m->h_seg;
            p.K = m->h_seg;
            p.a = m->hbuf;
            p.lda = m->h_seg;
            p.residual = m->xb;
            p.res_ld = E;
            p.total_rows = E;
            M_CHECK(run_gemm(m, p, act_q(lw[JL_L_DOWN]) ? PRO_F32_QUANT : PRO_F32, tp ? EPI_STORE : EPI_ADD_RESIDUAL, M,
                             (size_t)m->h_seg * 4, timed));
            if (tp) {
                M_CHECK(jl_comm_allreduce_dev(ctx, m->stream, m->partial, (size_t)M * E));
                JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->x, m->partial, (size_t)M * E * 4, cudaMemcpyDeviceToDevice, m->stream));
                M_CHECK(jl_launch_accumulate(ctx, m->stream, m->x, M, E, JL_F32, m->xb, nullptr, M, E, 0, E));
            }
        }
    }
    return JL_OK;
}

// Prefill chunk on the tensor cores: the same op sequence as forward_rows with every weight GEMM on tcgen05
// (BF16 operands, F32 accumulate; jl_gemm_tc.cu).  Norms, RoPE, attention, SiLU stay f32.
static int forward_rows_tc(jl_model *m, int M, int max_pos, int splits, int session, int pos0) {
    jl_ctx *ctx = m->ctx;
    const jl_model_config &c = m->cfg;
    const int E = c.embedding_length, hs = c.head_size, H = m->h_seg;
    M_CHECK(jl_launch_embed(ctx, m->stream, m->g[JL_T_EMBED], m->d_tokens, M, m->x, E));
    for (int L = 0; L < c.num_layers; L++) {
        const DevTensor *lw = &m->l[(size_t)L * 9];
        M_CHECK(jl_launch_rmsnorm_bf16(ctx, m->stream, m->x, M, E, lw[JL_L_ATTN_NORM].dtype, lw[JL_L_ATTN_NORM].data, 0.0f, c.layer_norm_eps, E,
                                       m->abf, E));
        M_CHECK(jl_launch_gemm_tc(ctx, m->stream, m->abf, E, M, lw[JL_L_Q], m->attn_seg, 0, E, m->q, m->attn_seg, 0, nullptr, 0));
        M_CHECK(jl_launch_gemm_tc(ctx, m->stream, m->abf, E, M, lw[JL_L_K], m->kv_seg, 0, E, m->k, m->kv_seg, 0, nullptr, 0));
        M_CHECK(jl_launch_gemm_tc(ctx, m->stream, m->abf, E, M, lw[JL_L_V], m->kv_seg, 0, E, m->v, m->kv_seg, 0, nullptr, 0));
        AttnParams ap = {};
        ap.kv = m->kv, ap.layer = L, ap.heads = m->heads_local, ap.kv_heads = m->kv_heads_local, ap.head_size = hs;
        ap.head0_global = m->d.headStart, ap.kv_head0_global = m->d.groupHeadStart;
        ap.q = m->q, ap.k = m->k, ap.v = m->v, ap.q_ld = m->attn_seg, ap.kv_ld = m->kv_seg, ap.out = m->att, ap.rope = m->rope;
        ap.rows = M, ap.sessions = m->d_sessions, ap.positions = m->d_positions;
        ap.scale = (float)(1.0 / sqrt((double)hs)), ap.ws = m->attn_ws, ap.splits = splits;
        M_CHECK(jl_launch_rope_kv_append(ctx, m->stream, ap, m->q, false));
        // rows = consecutive positions pos0.. of one session: tiled tensor-core attention (jl_attn_prefill.cu)
        if (jl_prefill_attention_supported(ap)) M_CHECK(jl_launch_prefill_attention(ctx, m->stream, ap, session, pos0));
        else M_CHECK(jl_launch_paged_attention(ctx, m->stream, ap, max_pos, false));
        M_CHECK(jl_launch_quantize_bf16(ctx, m->stream, m->att, M, m->attn_seg, 0, m->attn_seg, m->abf));
        M_CHECK(jl_launch_gemm_tc(ctx, m->stream, m->abf, m->attn_seg, M, lw[JL_L_O], E, 0, m->attn_seg, m->xb, E, 0, m->x, E));
        M_CHECK(jl_launch_rmsnorm_bf16(ctx, m->stream, m->xb, M, E, lw[JL_L_FFN_NORM].dtype, lw[JL_L_FFN_NORM].data, 0.0f, c.layer_norm_eps, E,
                                       m->abf, E));
        M_CHECK(jl_launch_gemm_tc(ctx, m->stream, m->abf, E, M, lw[JL_L_GATE], H, 0, E, m->hbuf, H, 0, nullptr, 0));
        M_CHECK(jl_launch_gemm_tc(ctx, m->stream, m->abf, E, M, lw[JL_L_UP], H, 0, E, m->hbuf2, H, 0, nullptr, 0));
        M_CHECK(jl_launch_silu_mul_bf16(ctx, m->stream, m->hbuf, m->hbuf2, M, H, H, m->abf, H));
        M_CHECK(jl_launch_gemm_tc(ctx, m->stream, m->abf, H, M, lw[JL_L_DOWN], E, 0, H, m->x, E, 0, m->xb, E));
    }
    return JL_OK;
}

// GPT-2 blocks (core/model/gpt2/GPT2Model.java:54-129) on the same per-op kernels: x = wte[token] + wpe[position]; per layer
//   ln1 = LayerNorm(x) -> q,k,v = ln1 . W^T + bias -> attention (no rotary: identity table) -> xb = (att . Wo^T + bo) + x
//   ln2 = LayerNorm(xb) -> h = gelu(ln2 . Wfc^T + bfc) -> x = (h . Wproj^T + bproj) + xb
// in the order TransformerBlock.forward / CausalSelfAttention.forward / MLPBlock.forward apply them (bias after the reducer,
// residual last: CausalSelfAttention.java:363-380, MLPBlock.java:126-160, TransformerBlock.java:185,203).
static int forward_rows_gpt2(jl_model *m, int M, int max_pos, int splits, bool distinct_sessions) {
    jl_ctx *ctx = m->ctx;
    const jl_model_config &c = m->cfg;
    const int E = c.embedding_length, hs = c.head_size, H = m->h_seg;
    const bool q8 = c.working_qtype == JL_I8;
    auto act_q = [&](const DevTensor &w) { return q8 && (w.dtype == JL_Q4 || w.dtype == JL_I8); };
    auto bias = [&](float *a, int lda, const DevTensor &b, int n) {
        return jl_launch_accumulate(ctx, m->stream, a, M, lda, b.dtype, b.data, nullptr, 1, (int)b.cols, 0, n);
    };
    M_CHECK(jl_launch_embed(ctx, m->stream, m->g[JL_T_EMBED], m->d_tokens, M, m->x, E));
    M_CHECK(jl_launch_pos_embed_add(ctx, m->stream, m->x, M, E, m->gaux[JL_AUX_POS_EMBED], m->d_positions));
    for (int L = 0; L < c.num_layers; L++) {
        const DevTensor *lw = &m->l[(size_t)L * 9];
        const DevTensor *ax = &m->aux[(size_t)L * 8];
        M_CHECK(jl_launch_layernorm(ctx, m->stream, m->x, M, E, lw[JL_L_ATTN_NORM].dtype, lw[JL_L_ATTN_NORM].data, ax[JL_AUX_ATTN_NORM_BIAS].dtype,
                                    ax[JL_AUX_ATTN_NORM_BIAS].data, c.layer_norm_eps, E, 0, E, m->ln));
        {
            GemvParams p = {};
            p.nseg = 3;
            set_w(p, 0, lw[JL_L_Q], m->q, m->attn_seg);
            set_w(p, 1, lw[JL_L_K], m->k, m->kv_seg);
            set_w(p, 2, lw[JL_L_V], m->v, m->kv_seg);
            p.w_dtype = lw[JL_L_Q].dtype, p.ldw = E, p.K = E, p.a = m->ln, p.lda = E;
            p.total_rows = m->attn_seg + 2 * m->kv_seg;
            M_CHECK(run_gemm(m, p, act_q(lw[JL_L_Q]) ? PRO_F32_QUANT : PRO_F32, EPI_STORE, M, (size_t)E * 4, false, false));
        }
        M_CHECK(bias(m->q, m->attn_seg, ax[JL_AUX_Q_BIAS], m->attn_seg));
        M_CHECK(bias(m->k, m->kv_seg, ax[JL_AUX_K_BIAS], m->kv_seg));
        M_CHECK(bias(m->v, m->kv_seg, ax[JL_AUX_V_BIAS], m->kv_seg));
        AttnParams ap = {};
        ap.kv = m->kv, ap.layer = L, ap.heads = m->heads_local, ap.kv_heads = m->kv_heads_local, ap.head_size = hs;
        ap.head0_global = m->d.headStart, ap.kv_head0_global = m->d.groupHeadStart;
        ap.q = m->q, ap.k = m->k, ap.v = m->v, ap.q_ld = m->attn_seg, ap.kv_ld = m->kv_seg, ap.out = m->att, ap.rope = m->rope;
        ap.rows = M, ap.sessions = m->d_sessions, ap.positions = m->d_positions;
        ap.scale = (float)(1.0 / sqrt((double)hs)), ap.ws = m->attn_ws, ap.splits = splits;
        if (distinct_sessions) {
            const int bound = splits >= m->max_splits ? m->max_context - 1 : splits * 64 * M - 1;
            M_CHECK(jl_launch_fused_decode_attention(ctx, m->stream, ap, m->fda_done, false, bound));
        } else {
            M_CHECK(jl_launch_rope_kv_append(ctx, m->stream, ap, m->q, false));
            M_CHECK(jl_launch_paged_attention(ctx, m->stream, ap, max_pos, false));
        }
        {
            GemvParams p = {};
            p.nseg = 1;
            set_w(p, 0, lw[JL_L_O], m->xb, E);
            p.w_dtype = lw[JL_L_O].dtype, p.ldw = m->attn_seg, p.K = m->attn_seg, p.a = m->att, p.lda = m->attn_seg, p.total_rows = E;
            M_CHECK(run_gemm(m, p, act_q(lw[JL_L_O]) ? PRO_F32_QUANT : PRO_F32, EPI_STORE, M, (size_t)m->attn_seg * 4, false, false));
        }
        M_CHECK(bias(m->xb, E, ax[JL_AUX_O_BIAS], E));
        M_CHECK(jl_launch_accumulate(ctx, m->stream, m->xb, M, E, JL_F32, m->x, nullptr, M, E, 0, E)); // + residual
        M_CHECK(jl_launch_layernorm(ctx, m->stream, m->xb, M, E, lw[JL_L_FFN_NORM].dtype, lw[JL_L_FFN_NORM].data, ax[JL_AUX_FFN_NORM_BIAS].dtype,
                                    ax[JL_AUX_FFN_NORM_BIAS].data, c.layer_norm_eps, E, 0, E, m->ln));
        {
            GemvParams p = {};
            p.nseg = 1;
            set_w(p, 0, lw[JL_L_GATE], m->hbuf, H);
            p.w_dtype = lw[JL_L_GATE].dtype, p.ldw = E, p.K = E, p.a = m->ln, p.lda = E, p.total_rows = H;
            M_CHECK(run_gemm(m, p, act_q(lw[JL_L_GATE]) ? PRO_F32_QUANT : PRO_F32, EPI_STORE, M, (size_t)E * 4, false, false));
        }
        M_CHECK(bias(m->hbuf, H, ax[JL_AUX_FC_BIAS], H));
        M_CHECK(jl_launch_activation(ctx, m->stream, JL_ACT_GELU, m->hbuf, M, H, 0, H));
        {
            GemvParams p = {};
            p.nseg = 1;
            set_w(p, 0, lw[JL_L_DOWN], m->x, E);
            p.w_dtype = lw[JL_L_DOWN].dtype, p.ldw = H, p.K = H, p.a = m->hbuf, p.lda = H, p.total_rows = E;
            M_CHECK(run_gemm(m, p, act_q(lw[JL_L_DOWN]) ? PRO_F32_QUANT : PRO_F32, EPI_STORE, M, (size_t)H * 4, false, false));
        }
        M_CHECK(bias(m->x, E, ax[JL_AUX_PROJ_BIAS], E));
        M_CHECK(jl_launch_accumulate(ctx, m->stream, m->x, M, E, JL_F32, m->xb, nullptr, M, E, 0, E)); // + residual
    }
    return JL_OK;
}

// AbstractModel.sample (:443-473) for `n` hidden rows [n, E] -> logits [n, vocab] -> argmax tokens
static int sample_rows(jl_model *m, const float *hidden, int n, bool timed) {
    jl_ctx *ctx = m->ctx;
    const jl_model_config &c = m->cfg;
    const int E = c.embedding_length;
    const DevTensor &w = m->g_set[JL_T_LM_HEAD] ? m->g[JL_T_LM_HEAD] : m->g[JL_T_EMBED];
    GemvParams p = {};
    p.nseg = 1;
    set_w(p, 0, w, m->logits, c.vocab_size);
    p.w_dtype = w.dtype;
    p.ldw = E;
    p.K = E;
    p.a = hidden;
    p.lda = E;
    if (c.arch == JL_ARCH_GPT2) { // ln_f is a LayerNorm with bias (GPT2Model.java:112-128), the logits weights are wte
        M_CHECK(jl_launch_layernorm(ctx, m->stream, hidden, n, E, m->g[JL_T_OUT_NORM].dtype, m->g[JL_T_OUT_NORM].data, m->gaux[JL_AUX_OUT_NORM_BIAS].dtype,
                                    m->gaux[JL_AUX_OUT_NORM_BIAS].data, c.layer_norm_eps, E, 0, E, m->ln));
        p.a = m->ln;
        p.total_rows = c.vocab_size;
        M_CHECK(run_gemm(m, p, PRO_F32, EPI_STORE, n, (size_t)E * 4, timed));
        M_CHECK(jl_launch_argmax(ctx, m->stream, m->logits, n, c.vocab_size, c.vocab_size, m->d_next, m->argmax_scratch));
        return JL_OK;
    }
    p.norm_w = m->g[JL_T_OUT_NORM].data;
    p.norm_w_dtype = m->g[JL_T_OUT_NORM].dtype;
    p.norm_eps = c.layer_norm_eps;
    p.norm_E = E;
    p.total_rows = c.vocab_size;
    // lm_head uses un-quantised F32 activations (AbstractModel.java:444-449)
    M_CHECK(run_gemm(m, p, PRO_RMSNORM_F32, EPI_STORE, n, (size_t)E * 4, timed));
    M_CHECK(jl_launch_argmax(ctx, m->stream, m->logits, n, c.vocab_size, c.vocab_size, m->d_next, m->argmax_scratch));
    return JL_OK;
}

static int pick_splits(const jl_model *m, int max_pos, int rows) {
    // enough CTAs to cover the SMs, at most one split per 128 positions
    int want = (max_pos + 1 + 127) / 128;
    int fill = (m->ctx->sm_count * 2) / (m->kv_heads_local * rows > 0 ? m->kv_heads_local * rows : 1);
    if (fill < 1) fill = 1;
    int s = want < fill ? want : fill;
    if (s > m->max_splits) s = m->max_splits;
    if (s < 1) s = 1;
    // bucket to powers of two so few graphs are needed
    int b = 1;
    while (b < s) b <<= 1;
    return b > m->max_splits ? m->max_splits : b;
}

extern "C" int jl_model_batch_forward(jl_model *m, int session, const int32_t *tokens, int n, int start_pos) {
    if (!m || !m->finalized || !tokens || n <= 0 || session < 0 || session >= m->cfg.max_sessions || start_pos < 0)
        return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    const int E = m->cfg.embedding_length;
    M_CHECK(ensure_pages(m, session, start_pos, start_pos + n - 1));
    for (int i = 0; i < n; i += m->cfg.max_batch) { // AbstractModel.java:304
        const int cnt = n - i < m->cfg.max_batch ? n - i : m->cfg.max_batch;
        int32_t *hp = m->h_pinned;
        for (int j = 0; j < cnt; j++) {
            hp[j] = tokens[i + j];
            hp[m->maxB + j] = start_pos + i + j;
            hp[2 * m->maxB + j] = session;
        }
        JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_tokens, hp, (size_t)cnt * 4, cudaMemcpyHostToDevice, m->stream));
        JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_positions, hp + m->maxB, (size_t)cnt * 4, cudaMemcpyHostToDevice, m->stream));
        JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_sessions, hp + 2 * m->maxB, (size_t)cnt * 4, cudaMemcpyHostToDevice, m->stream));
        const int max_pos = start_pos + i + cnt - 1;
        if (m->tc_ok && cnt >= 16) M_CHECK(forward_rows_tc(m, cnt, max_pos, pick_splits(m, max_pos, cnt), session, start_pos + i));
        else M_CHECK(forward_rows(m, cnt, max_pos, pick_splits(m, max_pos, cnt), false));
        // keep the last row for sample()
        JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->last_hidden + (size_t)session * E, m->x + (size_t)(cnt - 1) * E, (size_t)E * 4,
                                           cudaMemcpyDeviceToDevice, m->stream));
        JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream)); // pinned staging is reused by the next chunk
    }
    return JL_OK;
}

__global__ void __launch_bounds__(256) softmax_sample_kernel(float *logits, int vocab, float temperature, float uniform,
                                                             int32_t *out);

extern "C" int jl_model_sample(jl_model *m, int session, float temperature, float uniform, int32_t *token_out,
                               float *logits_out) {
    if (!m || !m->finalized || !token_out || session < 0 || session >= m->cfg.max_sessions) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    const int E = m->cfg.embedding_length, V = m->cfg.vocab_size;
    M_CHECK(sample_rows(m, m->last_hidden + (size_t)session * E, 1, false));
    if (temperature != 0.0f) {
        // AbstractModel.java:475-487: exp((l - max)/T), float prefix sum against the uniform sample
        softmax_sample_kernel<<<1, 256, 0, m->stream>>>(m->logits, V, temperature, uniform, m->d_next);
        ctx->launches++;
        JL_CUDA_CHECK(ctx, cudaGetLastError());
    }
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->h_pinned + 3 * m->maxB, m->d_next, 4, cudaMemcpyDeviceToHost, m->stream));
    if (logits_out && temperature == 0.0f)
        JL_CUDA_CHECK(ctx, cudaMemcpyAsync(logits_out, m->logits, (size_t)V * 4, cudaMemcpyDeviceToHost, m->stream));
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
    *token_out = m->h_pinned[3 * m->maxB];
    return JL_OK;
}

// exp((l-max)/T) then the first index whose running sum reaches `uniform` (AbstractModel.java:475-489).
// Sequential prefix order is part of the semantics, so one thread walks the normalised values.
__global__ void __launch_bounds__(256) softmax_sample_kernel(float *logits, int vocab, float temperature, float uniform,
                                                             int32_t *out) {
    __shared__ float red[32];
    __shared__ float bc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float mx = -INFINITY;
    for (int i = tid; i < vocab; i += blockDim.x) mx = fmaxf(mx, logits[i]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        float t = red[0];
        for (int i = 1; i < nw; i++) t = fmaxf(t, red[i]);
        bc = t;
    }
    __syncthreads();
    mx = bc;
    float sum = 0.0f;
    for (int i = tid; i < vocab; i += blockDim.x) {
        const float e = (float)exp(((double)logits[i] - (double)mx) / (double)temperature);
        logits[i] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        float t = 0.0f;
        for (int i = 0; i < nw; i++) t += red[i];
        float acc = 0.0f;
        int pick = vocab - 1;
        for (int i = 0; i < vocab; i++) {
            acc += logits[i] / t;
            if (acc >= uniform) {
                pick = i;
                break;
            }
        }
        *out = pick;
    }
}

// device-side bookkeeping at the end of a resident decode step: feed the sampled token back, advance
// the position, append to the history ring.
__global__ void advance_kernel(int32_t *tokens, int32_t *positions, const int32_t *next, int32_t *hist, int32_t *counter,
                               int n, int cap) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = threadIdx.x;
    const int cnt = *counter;
    if (i < n) {
        tokens[i] = next[i];
        positions[i] += 1;
        if (cnt * n + i < cap) hist[cnt * n + i] = next[i];
    }
    __syncthreads();
    if (i == 0) *counter = cnt + 1;
}

static int decode_body(jl_model *m, int n, int max_pos, int splits, bool resident, bool timed) {
    jl_ctx *ctx = m->ctx;
    M_CHECK(forward_rows(m, n, max_pos, splits, timed, true));
    M_CHECK(sample_rows(m, m->x, n, timed));
    if (resident) {
        JL_CUDA_CHECK(ctx, jl_launch_kernel(advance_kernel, dim3(1), dim3(256), 0, m->stream, false, m->d_tokens, m->d_positions,
                                            (const int32_t *)m->d_next, m->d_hist, m->d_counter, n, m->hist_cap));
        ctx->launches++;
    }
    return JL_OK;
}

static void fill_pd(jl_model *m, int max_pos, bool resident, bool want_logits, PdParams &p) {
    const jl_model_config &c = m->cfg;
    p = PdParams();
    p.layers = c.num_layers, p.E = c.embedding_length, p.H = m->h_seg, p.attn_seg = m->attn_seg, p.kv_seg = m->kv_seg;
    p.heads = m->heads_local, p.kv_heads = m->kv_heads_local, p.head_size = c.head_size;
    p.vocab = c.vocab_size, p.vocab_rows = m->pd_vocab_rows, p.vocab0 = m->pd_vocab0;
    p.head0_global = m->d.headStart, p.kv_head0_global = m->d.groupHeadStart;
    p.eps = c.layer_norm_eps;
    p.attn_scale = (float)(1.0 / sqrt((double)c.head_size)); // CausalSelfAttention.java:134
    p.lw = nullptr;
    p.embed_dt = m->g[JL_T_EMBED].dtype, p.embed_w = m->g[JL_T_EMBED].data, p.embed_s = m->g[JL_T_EMBED].scales;
    p.out_norm = m->g[JL_T_OUT_NORM].data, p.out_norm_dt = m->g[JL_T_OUT_NORM].dtype;
    const DevTensor &head = m->g_set[JL_T_LM_HEAD] ? m->g[JL_T_LM_HEAD] : m->g[JL_T_EMBED];
    const size_t rb = (size_t)c.embedding_length / 32 * (head.dtype == JL_Q4 ? 16 : 32);
    p.lm_w = (const uint8_t *)head.data + (size_t)m->pd_vocab0 * rb;
    p.lm_s = head.scales + (size_t)m->pd_vocab0 * (c.embedding_length / 32);
    p.x = m->x, p.xb = m->xb, p.q = m->q, p.k = m->k, p.v = m->v, p.att = m->att, p.h = m->hbuf, p.logits = m->logits;
    p.attn_ws = m->attn_ws;
    p.rope = m->rope;
    p.kv = m->kv;
    p.tokens = m->d_tokens, p.positions = m->d_positions, p.next = m->d_next, p.sessions = m->d_sessions;
    p.hist = m->d_hist, p.counter = m->d_counter, p.hist_cap = m->hist_cap, p.resident = resident ? 1 : 0;
    p.sync = m->pd_sync, p.argmax_slots = m->pd_slots, p.att_done = m->pd_att_done;
    // one context split per 64 positions, bounded by the CTAs available for (kv head, split) tasks
    int s = (max_pos + 1 + 63) / 64;
    const int cap = m->ctx->sm_count / m->kv_heads_local;
    if (s > cap) s = cap;
    if (s > m->max_splits) s = m->max_splits;
    if (s < 1) s = 1;
    p.splits = s;
    p.split_cap = cap < m->max_splits ? cap : m->max_splits;
    p.ntok = 1;
    p.world = c.tp_size, p.rank = c.tp_rank;
    for (int r = 0; r < c.tp_size && r < PD_MAX_TP; r++)
        p.ll_o[r] = m->peer_ll_o[r], p.ll_d[r] = m->peer_ll_d[r], p.ll_a[r] = m->peer_ll_a[r], p.logits_peer[r] = m->peer_logits[r];
    p.want_logits = want_logits && c.tp_size > 1 ? 1 : 0;
    p.trace = m->pd_trace;
}

// after a synchronise: did the last persistent launches drain on a timeout?  Resets the protocol state if so.
static int pd_check(jl_model *m) {
    if (!m->pd_ok) return JL_OK;
    unsigned long long st[2] = {0, 0};
    if (cudaMemcpy(st, m->pd_sync, sizeof st, cudaMemcpyDeviceToHost) != cudaSuccess) return jl_set_error(m->ctx, JL_ERR_CUDA, "decode status read failed");
    if (st[1] == 0) return JL_OK;
    cudaMemset(m->pd_sync, 0, PD_SYNC_WORDS * 8);
    cudaMemset(m->pd_att_done, 0, (size_t)m->kv_heads_local * sizeof(unsigned));
    return jl_set_error(m->ctx, JL_ERR_CUDA, "persistent decode kernel timed out waiting on counter %llu (token %llu); state reset", st[1], st[0]);
}

// run the decode body through the persistent kernel (one session), a cached CUDA graph, or eagerly
static int run_decode(jl_model *m, int n, int max_pos, bool resident, bool want_logits = false, int ntok = 1) {
    jl_ctx *ctx = m->ctx;
    if (n == 1 && use_pd(m)) {
        PdParams pp;
        fill_pd(m, max_pos, resident, want_logits, pp);
        pp.ntok = resident ? ntok : 1;
        return jl_launch_pdecode(ctx, m->stream, pp, m->pd_layers.data(), m, m->pd_wdtype);
    }
    // fused decode attention: one split per 64 positions (per 64 * n with n sessions in the step: the rows already fill the grid),
    // bucketed so that few graphs are captured
    int splits = (max_pos + 1 + 63) / 64;
    if (n > 1) splits = (splits + n - 1) / n;
    {
        static const int buckets[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32};
        int b = 32;
        for (int v : buckets)
            if (v >= splits) {
                b = v;
                break;
            }
        splits = b > m->max_splits ? m->max_splits : b;
    }
    if (!use_graph(m)) {
        m->ev_used = 0;
        int rc = decode_body(m, n, max_pos, splits, resident, true);
        m->timing_valid = rc == JL_OK;
        return rc;
    }
    const long long key = ((long long)n << 32) | ((long long)splits << 1) | (resident ? 1 : 0);
    auto it = m->graphs.find(key);
    if (it == m->graphs.end()) {
        // First decode of this (sessions, resident) kind: capture the graph of EVERY context-split bucket now, so that a
        // generation never pays a capture + instantiate (~8 ms) when its context crosses a bucket boundary later.
        // Capturing records launches only; nothing executes and no KV page is touched.
        static const int buckets[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32};
        for (int b : buckets) {
            const int sp = b > m->max_splits ? m->max_splits : b;
            const long long k2 = ((long long)n << 32) | ((long long)sp << 1) | (resident ? 1 : 0);
            if (m->graphs.count(k2)) continue;
            cudaGraph_t graph = nullptr;
            const long long before = ctx->launches;
            const int before_trace = ctx->ktrace_n;
            JL_CUDA_CHECK(ctx, cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
            int rc = decode_body(m, n, max_pos, sp, resident, false);
            cudaError_t ce = cudaStreamEndCapture(m->stream, &graph);
            const long long per_graph = ctx->launches - before;
            ctx->launches = before;
            if (sp != splits) ctx->ktrace_n = before_trace; // timeline slots only for the graph that is about to run
            if (rc != JL_OK) {
                if (graph) cudaGraphDestroy(graph);
                return rc;
            }
            if (ce != cudaSuccess) return jl_set_error(ctx, JL_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
            cudaGraphExec_t exec = nullptr;
            ce = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            if (ce != cudaSuccess) return jl_set_error(ctx, JL_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce));
            m->graphs[k2] = exec;
            m->graph_launches[k2] = per_graph;
        }
        it = m->graphs.find(key);
        if (it == m->graphs.end()) return jl_set_error(ctx, JL_ERR_INVALID, "decode: no graph for %d splits", splits);
    }
    JL_CUDA_CHECK(ctx, cudaGraphLaunch(it->second, m->stream));
    ctx->launches += m->graph_launches[key];
    return JL_OK;
}

// temperatures / uniforms: nullable [n]; a row with temperature != 0 is sampled with the reference's prefix-sum rule
// (AbstractModel.java:475-489, softmax_sample_kernel) from its logits row instead of taking the arg-max
static int decode_impl(jl_model *m, int n, const int32_t *sessions, const int32_t *tokens, const int32_t *positions,
                       const float *temperatures, const float *uniforms, int32_t *next_tokens, float *logits_out) {
    if (!m || !m->finalized || n <= 0 || n > m->cfg.max_sessions || n > GEMV_MAX_M || !sessions || !tokens || !positions ||
        !next_tokens)
        return m ? jl_set_error(m->ctx, JL_ERR_INVALID, "decode: bad arguments (n=%d, max %d)", n, GEMV_MAX_M) : JL_ERR_INVALID;
    bool sampled = false;
    for (int i = 0; temperatures && i < n; i++) sampled = sampled || temperatures[i] != 0.0f;
    if (sampled && !uniforms) return jl_set_error(m->ctx, JL_ERR_INVALID, "decode: temperatures without uniform samples");
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    int max_pos = 0;
    int32_t *hp = m->h_pinned;
    for (int i = 0; i < n; i++) {
        if (sessions[i] < 0 || sessions[i] >= m->cfg.max_sessions || positions[i] < 0)
            return jl_set_error(ctx, JL_ERR_INVALID, "decode: bad session/position");
        for (int j = 0; j < i; j++) // two rows of one session would append to and read the same KV positions concurrently
            if (sessions[j] == sessions[i]) return jl_set_error(ctx, JL_ERR_INVALID, "decode: session %d appears twice in one step", sessions[i]);
        M_CHECK(ensure_pages(m, sessions[i], positions[i], positions[i]));
        if (positions[i] > max_pos) max_pos = positions[i];
        hp[i] = tokens[i];
        hp[m->maxB + i] = positions[i];
        hp[2 * m->maxB + i] = sessions[i];
    }
    static const bool dbg_timing = getenv("JL_DEBUG_TIMING") != nullptr;
    auto c0 = std::chrono::steady_clock::now();
    JL_CUDA_CHECK(ctx, cudaEventRecord(m->ev_begin, m->stream));
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_tokens, hp, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_positions, hp + m->maxB, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_sessions, hp + 2 * m->maxB, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
    auto c1 = std::chrono::steady_clock::now();
    M_CHECK(run_decode(m, n, max_pos, false, logits_out != nullptr || sampled));
    auto c2 = std::chrono::steady_clock::now();
    if (logits_out) // before the sampling kernels turn the sampled rows into exponentials in place
        JL_CUDA_CHECK(ctx, cudaMemcpyAsync(logits_out, m->logits, (size_t)n * m->cfg.vocab_size * 4, cudaMemcpyDeviceToHost, m->stream));
    for (int i = 0; sampled && i < n; i++) {
        if (temperatures[i] == 0.0f) continue; // :471-473 short-circuit: the arg-max of the step stands
        softmax_sample_kernel<<<1, 256, 0, m->stream>>>(m->logits + (size_t)i * m->cfg.vocab_size, m->cfg.vocab_size, temperatures[i], uniforms[i],
                                                        m->d_next + i);
        ctx->launches++;
        JL_CUDA_CHECK(ctx, cudaGetLastError());
    }
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(hp + 3 * m->maxB, m->d_next, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
    JL_CUDA_CHECK(ctx, cudaEventRecord(m->ev_end, m->stream));
    auto c3 = std::chrono::steady_clock::now();
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
    auto c4 = std::chrono::steady_clock::now();
    if (dbg_timing) {
        static int calls = 0;
        if ((++calls % 32) == 0) {
            auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
            float gms = 0;
            cudaEventElapsedTime(&gms, m->ev_begin, m->ev_end);
            fprintf(stderr, "[decode] h2d-enqueue %.1f us, launch %.1f us, d2h-enqueue %.1f us, sync-wait %.1f us, gpu(events) %.1f us\n",
                    us(c0, c1), us(c1, c2), us(c2, c3), us(c3, c4), gms * 1000.0);
        }
    }
    if (n == 1 && use_pd(m)) M_CHECK(pd_check(m));
    for (int i = 0; i < n; i++) next_tokens[i] = hp[3 * m->maxB + i];
    float ms = 0;
    cudaEventElapsedTime(&ms, m->ev_begin, m->ev_end);
    m->last_total_ms = ms;
    if (!use_graph(m) && m->timing_valid) {
        double g = 0;
        for (size_t i = 0; i + 1 < m->ev_used; i += 2) {
            float t = 0;
            cudaEventElapsedTime(&t, m->ev_pool[i], m->ev_pool[i + 1]);
            g += t;
        }
        m->last_gemv_ms = g;
    }
    return JL_OK;
}

extern "C" int jl_model_decode(jl_model *m, int n, const int32_t *sessions, const int32_t *tokens, const int32_t *positions,
                               int32_t *next_tokens, float *logits_out) {
    return decode_impl(m, n, sessions, tokens, positions, nullptr, nullptr, next_tokens, logits_out);
}

extern "C" int jl_model_decode_sample(jl_model *m, int n, const int32_t *sessions, const int32_t *tokens, const int32_t *positions,
                                      const float *temperatures, const float *uniforms, int32_t *next_tokens, float *logits_out) {
    return decode_impl(m, n, sessions, tokens, positions, temperatures, uniforms, next_tokens, logits_out);
}

extern "C" int jl_model_decode_resident(jl_model *m, int session, int32_t first_token, int start_pos, int n_new,
                                        int32_t *out_tokens) {
    if (!m || !m->finalized || session < 0 || session >= m->cfg.max_sessions || n_new <= 0 || !out_tokens || start_pos < 0)
        return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    if (n_new > m->hist_cap) return jl_set_error(ctx, JL_ERR_INVALID, "decode_resident: n_new too large");
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    M_CHECK(ensure_pages(m, session, start_pos, start_pos + n_new - 1));
    int32_t *hp = m->h_pinned;
    hp[0] = first_token;
    hp[m->maxB] = start_pos;
    hp[2 * m->maxB] = session;
    hp[3 * m->maxB] = 0;
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_tokens, hp, 4, cudaMemcpyHostToDevice, m->stream));
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_positions, hp + m->maxB, 4, cudaMemcpyHostToDevice, m->stream));
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_sessions, hp + 2 * m->maxB, 4, cudaMemcpyHostToDevice, m->stream));
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(m->d_counter, hp + 3 * m->maxB, 4, cudaMemcpyHostToDevice, m->stream));
    JL_CUDA_CHECK(ctx, cudaEventRecord(m->ev_begin, m->stream));
    double gemv = 0;
    // persistent kernel: up to PD_TOKENS_PER_LAUNCH tokens per cooperative launch (the fed-back token crosses a grid barrier
    // instead of a kernel boundary); other paths: one graph replay per token
    static const int tpl_env = getenv("JL_PD_TOKENS") ? atoi(getenv("JL_PD_TOKENS")) : 16;
    const int tpl = use_pd(m) && tpl_env > 1 ? tpl_env : 1;
    for (int i = 0; i < n_new; i += tpl) {
        const int nt = n_new - i < tpl ? n_new - i : tpl;
        M_CHECK(run_decode(m, 1, start_pos + i, true, false, nt));
        if (!use_graph(m)) {
            // eager profiling mode: collect the per-GEMV event pairs of this step
            JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
            for (size_t e = 0; e + 1 < m->ev_used; e += 2) {
                float t = 0;
                cudaEventElapsedTime(&t, m->ev_pool[e], m->ev_pool[e + 1]);
                gemv += t;
            }
        }
    }
    JL_CUDA_CHECK(ctx, cudaEventRecord(m->ev_end, m->stream));
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(out_tokens, m->d_hist, (size_t)n_new * 4, cudaMemcpyDeviceToHost, m->stream));
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
    if (use_pd(m)) M_CHECK(pd_check(m));
    float ms = 0;
    cudaEventElapsedTime(&ms, m->ev_begin, m->ev_end);
    m->last_total_ms = ms;
    m->last_gemv_ms = gemv;
    return JL_OK;
}

// diagnostic: CTA 0's phase stamps (globaltimer ns) of the last persistent decode launch; needs JL_PD_TRACE=1 at finalize.
// out[0] = kernel start, out[1 + L*8 + k]: k 0 qkv dependency met, 1 qkv done, 2 attention done (attention CTAs only), 3 o_proj dependency
// met, 4 o_proj done, 5 gate/up dependency met, 6 gate/up done, 7 down dependency met; then lm_head dependency met, token published.
extern "C" int jl_model_debug_trace(jl_model *m, uint64_t *out, int64_t out_words) {
    if (!m || !m->finalized || !out) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    if (!m->pd_trace) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "phase tracing is off (set JL_PD_TRACE=1 before creating the model)");
    const size_t words = (size_t)m->cfg.num_layers * 16 + 32;
    if ((size_t)out_words < words) return jl_set_error(ctx, JL_ERR_INVALID, "trace buffer too small (%zu words needed)", words);
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
    JL_CUDA_CHECK(ctx, cudaMemcpy(out, m->pd_trace, words * 8, cudaMemcpyDeviceToHost));
    return (int)words;
}

extern "C" int jl_model_decode_mode(jl_model *m, int n) {
    if (!m || !m->finalized) return JL_ERR_INVALID;
    if (n == 1 && use_pd(m)) return 3;
    return use_graph(m) ? 1 : 0;
}

extern "C" int jl_model_last_timing(jl_model *m, double *total_ms, double *gemv_ms) {
    if (!m) return JL_ERR_INVALID;
    if (total_ms) *total_ms = m->last_total_ms;
    if (gemv_ms) *gemv_ms = m->last_gemv_ms;
    return JL_OK;
}

extern "C" int jl_model_generate(jl_model *m, int session, const int32_t *prompt, int n_prompt, int n_new, int32_t *out_tokens,
                                 float *logits_out, double *timings_ms) {
    if (!m || !m->finalized || !prompt || n_prompt <= 0 || n_new <= 0 || !out_tokens) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    const int V = m->cfg.vocab_size;
    auto t0 = std::chrono::steady_clock::now();
    M_CHECK(jl_model_reset_session(m, session));
    M_CHECK(jl_model_batch_forward(m, session, prompt, n_prompt, 0));
    int32_t next = 0;
    M_CHECK(jl_model_sample(m, session, 0.0f, 0.0f, &next, logits_out));
    auto t1 = std::chrono::steady_clock::now();
    out_tokens[0] = next;
    for (int i = 1; i < n_new; i++) {
        const int32_t pos = n_prompt + i - 1;
        int32_t nx = 0;
        M_CHECK(jl_model_decode(m, 1, &session, &next, &pos, &nx, logits_out ? logits_out + (size_t)i * V : nullptr));
        next = nx;
        out_tokens[i] = next;
    }
    auto t2 = std::chrono::steady_clock::now();
    if (timings_ms) {
        timings_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        timings_ms[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
    }
    return JL_OK;
}

// AbstractModel.generate (:516-646) with a temperature: uniforms[i] plays ThreadLocalRandom.current().nextFloat() of the i-th sample()
// call (:576, :594), so a caller that passes the same stream gets the same tokens.
extern "C" int jl_model_generate_sample(jl_model *m, int session, const int32_t *prompt, int n_prompt, int n_new, float temperature,
                                        const float *uniforms, int32_t *out_tokens, double *timings_ms) {
    if (!m || !m->finalized || !prompt || n_prompt <= 0 || n_new <= 0 || !out_tokens || (temperature != 0.0f && !uniforms)) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    auto t0 = std::chrono::steady_clock::now();
    M_CHECK(jl_model_reset_session(m, session));
    M_CHECK(jl_model_batch_forward(m, session, prompt, n_prompt, 0));
    int32_t next = 0;
    M_CHECK(jl_model_sample(m, session, temperature, temperature != 0.0f ? uniforms[0] : 0.0f, &next, nullptr));
    auto t1 = std::chrono::steady_clock::now();
    out_tokens[0] = next;
    for (int i = 1; i < n_new; i++) {
        const int32_t pos = n_prompt + i - 1;
        const float u = temperature != 0.0f ? uniforms[i] : 0.0f;
        int32_t nx = 0;
        M_CHECK(decode_impl(m, 1, &session, &next, &pos, &temperature, &u, &nx, nullptr));
        next = nx;
        out_tokens[i] = next;
    }
    auto t2 = std::chrono::steady_clock::now();
    if (timings_ms) {
        timings_ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
        timings_ms[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
    }
    return JL_OK;
}

// ---- KV persistence (KvBufferCache.KvBufferPage, core/tensor/KvBufferCache.java:121-176): the reference backs every page of a
// non-ephemeral session with the file  <workingDirectory>/<session>-L<layerPage>C<contextPage>.page  holding the raw little-endian
// page ([layersPerPage][2][contextPerPage][kvLength] of the working KV dtype).  Here pages live in HBM; save writes the allocated
// pages of a session in exactly that naming and byte layout, load maps such files back (allocating the pages) so a session can be
// resumed by this library or by the reference.  Returns the number of pages written / read, or a negative status.
extern "C" int jl_model_kv_save(jl_model *m, int session, const char *dir, const char *session_name) {
    if (!m || !m->finalized || !dir || !session_name || session < 0 || session >= m->cfg.max_sessions) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
    std::vector<char> host(m->page_bytes);
    int n = 0;
    for (int lp = 0; lp < m->kv.n_layer_pages; lp++)
        for (int cp = 0; cp < m->kv.n_ctx_pages; cp++) {
            const void *page = m->page_table_host[((size_t)session * m->kv.n_layer_pages + lp) * m->kv.n_ctx_pages + cp];
            if (!page) continue;
            JL_CUDA_CHECK(ctx, cudaMemcpy(host.data(), page, m->page_bytes, cudaMemcpyDeviceToHost));
            const std::string path = std::string(dir) + "/" + session_name + "-L" + std::to_string(lp) + "C" + std::to_string(cp) + ".page";
            FILE *f = fopen(path.c_str(), "wb");
            if (!f) return jl_set_error(ctx, JL_ERR_INVALID, "kv_save: cannot open %s", path.c_str());
            const size_t w = fwrite(host.data(), 1, m->page_bytes, f);
            if (fclose(f) != 0 || w != m->page_bytes) return jl_set_error(ctx, JL_ERR_INVALID, "kv_save: short write to %s", path.c_str());
            n++;
        }
    return n;
}

extern "C" int jl_model_kv_load(jl_model *m, int session, const char *dir, const char *session_name) {
    if (!m || !m->finalized || !dir || !session_name || session < 0 || session >= m->cfg.max_sessions) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    std::vector<char> host(m->page_bytes);
    int n = 0;
    for (int cp = 0; cp < m->kv.n_ctx_pages; cp++)
        for (int lp = 0; lp < m->kv.n_layer_pages; lp++) {
            const std::string path = std::string(dir) + "/" + session_name + "-L" + std::to_string(lp) + "C" + std::to_string(cp) + ".page";
            FILE *f = fopen(path.c_str(), "rb");
            if (!f) continue;
            const size_t r = fread(host.data(), 1, m->page_bytes, f);
            const bool more = fgetc(f) != EOF;
            fclose(f);
            // KvBufferPage re-sizes a file of the wrong length (:148); a page of another geometry cannot be this session's
            if (r != m->page_bytes || more)
                return jl_set_error(ctx, JL_ERR_INVALID, "kv_load: %s is not a %zu-byte page of this model's geometry", path.c_str(), m->page_bytes);
            M_CHECK(ensure_pages(m, session, cp * m->kv.ctx_per_page, cp * m->kv.ctx_per_page)); // allocates every layer page of cp
            void *page = m->page_table_host[((size_t)session * m->kv.n_layer_pages + lp) * m->kv.n_ctx_pages + cp];
            JL_CUDA_CHECK(ctx, cudaMemcpy(page, host.data(), m->page_bytes, cudaMemcpyHostToDevice));
            n++;
        }
    return n;
}

// ---- host spill of an idle session -----------------------------------------------------------------------------------------------
// The reference's pages are memory-mapped files, so an idle session costs no RAM (KvBufferCache.java:121-176).  Here an idle session
// would pin HBM; offload moves its pages to host memory and returns the slot to the pool, restore brings them back into any slot.
static int set_page(jl_model *m, size_t idx, void *p) {
    jl_ctx *ctx = m->ctx;
    m->page_table_host[idx] = p;
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(&m->page_table_dev[idx], &m->page_table_host[idx], sizeof(void *), cudaMemcpyHostToDevice, m->stream));
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream)); // the staged pointer must outlive the copy
    return JL_OK;
}

extern "C" int jl_model_kv_pages(jl_model *m, int session) {
    if (!m || !m->finalized || session < 0 || session >= m->cfg.max_sessions) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    const size_t per = (size_t)m->kv.n_layer_pages * m->kv.n_ctx_pages;
    int n = 0;
    for (size_t i = 0; i < per; i++) n += m->page_table_host[(size_t)session * per + i] != nullptr;
    return n;
}

extern "C" int64_t jl_model_kv_offload(jl_model *m, int session) {
    if (!m || !m->finalized || session < 0 || session >= m->cfg.max_sessions) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
    jl_model::KvSpill sp;
    for (int lp = 0; lp < m->kv.n_layer_pages; lp++)
        for (int cp = 0; cp < m->kv.n_ctx_pages; cp++)
            if (m->page_table_host[((size_t)session * m->kv.n_layer_pages + lp) * m->kv.n_ctx_pages + cp]) sp.pages.emplace_back(lp, cp);
    try {
        sp.bytes.resize(sp.pages.size() * m->page_bytes);
    } catch (const std::bad_alloc &) {
        return jl_set_error(ctx, JL_ERR_OOM, "kv_offload: out of host memory for %zu pages", sp.pages.size());
    }
    // copy everything first, free afterwards: a failure leaves the session as it was
    for (size_t i = 0; i < sp.pages.size(); i++) {
        const size_t idx = ((size_t)session * m->kv.n_layer_pages + sp.pages[i].first) * m->kv.n_ctx_pages + sp.pages[i].second;
        JL_CUDA_CHECK(ctx, cudaMemcpy(sp.bytes.data() + i * m->page_bytes, m->page_table_host[idx], m->page_bytes, cudaMemcpyDeviceToHost));
    }
    for (size_t i = 0; i < sp.pages.size(); i++) {
        const size_t idx = ((size_t)session * m->kv.n_layer_pages + sp.pages[i].first) * m->kv.n_ctx_pages + sp.pages[i].second;
        void *page = m->page_table_host[idx];
        M_CHECK(set_page(m, idx, nullptr));
        cudaFree(page);
    }
    const int64_t h = m->next_spill++;
    m->spills.emplace(h, std::move(sp));
    return h;
}

extern "C" int jl_model_kv_restore(jl_model *m, int session, int64_t handle) {
    if (!m || !m->finalized || session < 0 || session >= m->cfg.max_sessions) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    auto it = m->spills.find(handle);
    if (it == m->spills.end()) return jl_set_error(ctx, JL_ERR_INVALID, "kv_restore: unknown spill handle %lld", (long long)handle);
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    M_CHECK(jl_model_reset_session(m, session)); // pages the slot still holds from its previous user start from zero
    const jl_model::KvSpill &sp = it->second;
    for (size_t i = 0; i < sp.pages.size(); i++) {
        const size_t idx = ((size_t)session * m->kv.n_layer_pages + sp.pages[i].first) * m->kv.n_ctx_pages + sp.pages[i].second;
        if (!m->page_table_host[idx]) {
            void *p = nullptr;
            M_CHECK(dev_alloc(ctx, &p, m->page_bytes));
            M_CHECK(set_page(m, idx, p));
        }
        JL_CUDA_CHECK(ctx, cudaMemcpy(m->page_table_host[idx], sp.bytes.data() + i * m->page_bytes, m->page_bytes, cudaMemcpyHostToDevice));
    }
    m->spills.erase(it);
    return JL_OK;
}

extern "C" int jl_model_kv_discard(jl_model *m, int64_t handle) {
    if (!m) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    return m->spills.erase(handle) ? JL_OK : jl_set_error(m->ctx, JL_ERR_INVALID, "kv_discard: unknown spill handle %lld", (long long)handle);
}

extern "C" int jl_model_read_kv(jl_model *m, int session, int layer, int position, int which, float *out) {
    if (!m || !m->finalized || !out || session < 0 || session >= m->cfg.max_sessions || layer < 0 ||
        layer >= m->cfg.num_layers || position < 0 || position >= m->max_context || (which != 0 && which != 1))
        return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    const KvLayout &kv = m->kv;
    const int lp = layer / kv.layers_per_page, rl = layer % kv.layers_per_page;
    const int cp = position / kv.ctx_per_page, rc = position % kv.ctx_per_page;
    const char *base = (const char *)m->page_table_host[((size_t)session * kv.n_layer_pages + lp) * kv.n_ctx_pages + cp];
    if (!base) return jl_set_error(ctx, JL_ERR_INVALID, "read_kv: page not allocated");
    const size_t elem = (((size_t)rl * 2 + which) * kv.ctx_per_page + rc) * kv.kv_len;
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
    if (kv.kv_dtype == JL_F32) {
        JL_CUDA_CHECK(ctx, cudaMemcpy(out, base + elem * 4, (size_t)kv.kv_len * 4, cudaMemcpyDeviceToHost));
    } else {
        std::vector<uint16_t> tmp(kv.kv_len);
        JL_CUDA_CHECK(ctx, cudaMemcpy(tmp.data(), base + elem * 2, (size_t)kv.kv_len * 2, cudaMemcpyDeviceToHost));
        for (int i = 0; i < kv.kv_len; i++) {
            uint32_t u = ((uint32_t)tmp[i]) << 16;
            memcpy(&out[i], &u, 4);
        }
    }
    return JL_OK;
}

// DistributedContext of this model (which rows / columns of every tensor this rank holds) and its shard count
extern "C" int jl_model_tp_layout(jl_model *m, jl_dctx *out, int *tp_size) {
    if (!m || !out) return JL_ERR_INVALID;
    *out = m->d;
    if (tp_size) *tp_size = m->cfg.tp_size;
    return JL_OK;
}

extern "C" int jl_model_debug_read(jl_model *m, int which, float *out, int64_t n) {
    if (!m || !m->finalized || !out || n <= 0) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    const float *src[8] = {m->x, m->xb, m->q, m->k, m->v, m->att, m->hbuf, m->logits};
    const size_t B = m->maxB;
    const size_t cap[8] = {B * m->cfg.embedding_length, B * m->cfg.embedding_length, B * m->attn_seg, B * m->kv_seg, B * m->kv_seg,
                           B * m->attn_seg, B * m->h_seg, (size_t)m->cfg.max_sessions * m->cfg.vocab_size};
    if (which < 0 || which > 7 || (size_t)n > cap[which]) return jl_set_error(ctx, JL_ERR_INVALID, "debug_read: bad buffer / size");
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
    JL_CUDA_CHECK(ctx, cudaMemcpy(out, src[which], (size_t)n * 4, cudaMemcpyDeviceToHost));
    return JL_OK;
}

extern "C" int jl_model_read_hidden(jl_model *m, int session, float *out) {
    if (!m || !m->finalized || !out || session < 0 || session >= m->cfg.max_sessions) return JL_ERR_INVALID;
    std::lock_guard<std::recursive_mutex> model_lock(m->mu);
    jl_ctx *ctx = m->ctx;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(m->stream));
    JL_CUDA_CHECK(ctx, cudaMemcpy(out, m->last_hidden + (size_t)session * m->cfg.embedding_length,
                                  (size_t)m->cfg.embedding_length * 4, cudaMemcpyDeviceToHost));
    return JL_OK;
}
