// Sanitizer harness for the checkpoint reader / loader (jlama_b200/csrc/jl_safetensors.cu is host code: a JSON parser, mmap, slicing).
// Built with -fsanitize=address,undefined by tests/test_safetensors_sanitizers.py.
//   1. a well-formed JQ4 checkpoint (Llama layout and Mixtral layout) written with jl_st_write is loaded through
//      jl_model_load_safetensors for every rank of tp = 1, 2, 4; the stub jl_register_tensor reads EVERY byte of every slice it is
//      handed (data and block scales), so a slice that leaves the mapping or the temporary slice buffers is an ASAN report, and
//      checksums them against the rows / columns the rank must receive (LlamaModel.java:120-133 row and column splits);
//   2. thousands of header mutations: jl_st_open either rejects the file or returns tensors whose bytes are all readable, and the
//      loader either fails cleanly or reads in bounds.
#include "jl_safetensors.cu"

#include <random>
#include <stdarg.h>

#define REQUIRE(cond)                                                                   \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            fprintf(stderr, "REQUIRE failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            abort();                                                                    \
        }                                                                               \
    } while (0)

// ---- the rest of the library, as far as the loader calls it ------------------------------------------------------------------------
struct jl_ctx {
    int dummy;
};
struct jl_model {
    jl_dctx d;
    int tp, experts, layers;
    long sets;
};
int jl_set_error(jl_ctx *, int code, const char *, ...) { return code; }
struct Registered {
    int dtype;
    int64_t rows, cols;
    uint64_t data_sum, scale_sum;
};
static std::vector<Registered> g_registered; // what the loader handed to jl_register_tensor, in order
static uint64_t fnv(const uint8_t *p, size_t n, uint64_t h = 1469598103934665603ull) {
    for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}
static volatile uint64_t g_sink;
static uint64_t g_bytes_read;
extern "C" int64_t jl_register_tensor(jl_ctx *, int dtype, int64_t rows, int64_t cols, const void *data, const float *scales) {
    if (dtype < JL_F32 || dtype > JL_I8 || rows <= 0 || cols <= 0 || !data) return -1;
    const size_t row = dtype == JL_F32 ? (size_t)cols * 4 : dtype == JL_BF16 ? (size_t)cols * 2 : dtype == JL_Q4 ? (size_t)cols / 2 : (size_t)cols;
    uint64_t acc = 0;
    const uint8_t *p = (const uint8_t *)data;
    for (size_t i = 0; i < row * (size_t)rows; i++) acc += p[i];
    if (dtype == JL_Q4 || dtype == JL_I8) {
        if (!scales) return -1;
        const uint8_t *sp = (const uint8_t *)scales; // like the library's cudaMemcpy: bytes, whatever the alignment inside the mapping
        for (size_t i = 0; i < (size_t)rows * (size_t)(cols / 32) * 4; i++) acc += sp[i];
        g_bytes_read += (size_t)rows * (size_t)(cols / 32) * 4;
    }
    g_sink = acc;
    g_bytes_read += row * (size_t)rows;
    Registered r;
    r.dtype = dtype, r.rows = rows, r.cols = cols;
    r.data_sum = fnv(p, row * (size_t)rows);
    r.scale_sum = (dtype == JL_Q4 || dtype == JL_I8) ? fnv((const uint8_t *)scales, (size_t)rows * (size_t)(cols / 32) * 4) : 0;
    g_registered.push_back(r);
    static int64_t next = 1;
    return next++;
}
extern "C" int jl_quantize_q4_weights(jl_ctx *, const float *, int64_t, int64_t, uint8_t *, float *) { return JL_ERR_UNSUPPORTED; }
extern "C" int jl_quantize_q8_weights(jl_ctx *, const float *, int64_t, int64_t, int8_t *, float *) { return JL_ERR_UNSUPPORTED; }
extern "C" int jl_model_set_tensor(jl_model *m, int, int, int64_t id) { return m->sets++, id > 0 ? JL_OK : JL_ERR_INVALID; }
extern "C" int jl_model_set_expert_tensor(jl_model *m, int, int, int, int64_t id) { return m->sets++, id > 0 ? JL_OK : JL_ERR_INVALID; }
extern "C" int jl_model_tp_layout(jl_model *m, jl_dctx *out, int *tp) {
    *out = m->d, *tp = m->tp;
    return JL_OK;
}
int jl_model_num_experts(jl_model *m) { return m->experts; }

// ---- a tiny checkpoint in Jlama's layout --------------------------------------------------------------------------------------------
static const int E = 128, H = 256, HEADS = 4, KVH = 4, HS = 32, LAYERS = 2, VOCAB = 96; // tp = 4 still cuts 32-column shards
struct Blob {
    std::vector<std::string> names;
    std::vector<int> dtypes, ndims;
    std::vector<int64_t> shapes;
    std::vector<std::vector<uint8_t>> data;
    void add(const std::string &n, int dt, int64_t r, int64_t c, std::mt19937 &rng) {
        const size_t bytes = dt == JL_F32 ? (size_t)r * c * 4 : dt == JL_Q4 ? (size_t)r * c / 2 : (size_t)r * c;
        std::vector<uint8_t> b(bytes);
        for (auto &x : b) x = (uint8_t)rng();
        if (dt == JL_F32)
            for (size_t i = 0; i + 3 < bytes; i += 4) b[i + 3] = 0x3f; // finite floats
        names.push_back(n), dtypes.push_back(dt), ndims.push_back(2), data.push_back(std::move(b));
        shapes.insert(shapes.end(), {r, c, 0, 0});
    }
    void quant(const std::string &n, int64_t r, int64_t c, std::mt19937 &rng) {
        add(n, JL_Q4, r, c, rng);
        add(n + ".qb", JL_F32, r, c / 32, rng);
    }
    int write(const std::string &path) {
        std::vector<const char *> np;
        std::vector<const void *> dp;
        std::vector<int64_t> nb;
        for (size_t i = 0; i < names.size(); i++) np.push_back(names[i].c_str()), dp.push_back(data[i].data()), nb.push_back((int64_t)data[i].size());
        const char *meta[] = {"format", "pt"};
        return jl_st_write(path.c_str(), (int)names.size(), np.data(), dtypes.data(), ndims.data(), shapes.data(), dp.data(), nb.data(), 1, meta);
    }
};

static Blob make_checkpoint(bool moe, std::mt19937 &rng) {
    Blob b;
    b.quant("model.embed_tokens.weight", VOCAB, E, rng);
    for (int l = 0; l < LAYERS; l++) {
        const std::string p = "model.layers." + std::to_string(l) + ".";
        b.add(p + "input_layernorm.weight", JL_F32, 1, E, rng);
        b.quant(p + "self_attn.q_proj.weight", E, E, rng);
        b.quant(p + "self_attn.k_proj.weight", KVH * HS, E, rng);
        b.quant(p + "self_attn.v_proj.weight", KVH * HS, E, rng);
        b.quant(p + "self_attn.o_proj.weight", E, E, rng);
        b.add(p + "post_attention_layernorm.weight", JL_F32, 1, E, rng);
        if (moe) {
            b.quant(p + "block_sparse_moe.gate.weight", 4, E, rng);
            for (int e = 0; e < 4; e++) {
                const std::string q = p + "block_sparse_moe.experts." + std::to_string(e) + ".";
                b.quant(q + "w1.weight", H, E, rng), b.quant(q + "w2.weight", E, H, rng), b.quant(q + "w3.weight", H, E, rng);
            }
        } else {
            b.quant(p + "mlp.gate_proj.weight", H, E, rng), b.quant(p + "mlp.down_proj.weight", E, H, rng), b.quant(p + "mlp.up_proj.weight", H, E, rng);
        }
    }
    b.add("model.norm.weight", JL_F32, 1, E, rng);
    b.quant("lm_head.weight", VOCAB, E, rng);
    return b;
}

// what rank `rank` of `tp` must receive for tensor `name`: rows [r0, r0+nr) x columns [c0, c0+nc) of the stored matrix, checksummed the
// same way (Q4: nibble bytes of the column range + the f32 block scales of that range)
static Registered expect(const Blob &b, const std::string &name, int64_t r0, int64_t nr, int64_t c0, int64_t nc) {
    size_t i = 0;
    while (b.names[i] != name) i++;
    const int dt = b.dtypes[i];
    const int64_t R = b.shapes[i * 4], C = b.shapes[i * 4 + 1];
    if (nr < 0) r0 = 0, nr = R;
    if (nc < 0) c0 = 0, nc = C;
    Registered e;
    e.dtype = dt, e.rows = nr, e.cols = nc, e.scale_sum = 0;
    const double bpe = dt == JL_F32 ? 4 : (dt == JL_Q4 ? 0.5 : 1);
    uint64_t h = 1469598103934665603ull;
    for (int64_t r = r0; r < r0 + nr; r++) h = fnv(b.data[i].data() + (size_t)(r * C * bpe) + (size_t)(c0 * bpe), (size_t)(nc * bpe), h);
    e.data_sum = h;
    if (dt == JL_Q4) {
        const std::vector<uint8_t> &qb = b.data[i + 1]; // "<name>.qb" follows its tensor
        h = 1469598103934665603ull;
        for (int64_t r = r0; r < r0 + nr; r++) h = fnv(qb.data() + ((size_t)r * (size_t)(C / 32) + (size_t)(c0 / 32)) * 4, (size_t)(nc / 32) * 4, h);
        e.scale_sum = h;
    }
    return e;
}

// the order and the slices of LlamaModel.loadInputWeights / loadTransformerBlockWeights / loadOutputWeights (llama/LlamaModel.java:68-156,
// rows for q/k/v/gate/up, columns for o/down) and MixtralModel.java:88-105 with expert e on rank e % tp
static std::vector<Registered> expected_registrations(const Blob &b, const jl_model &m, bool moe) {
    std::vector<Registered> v;
    const bool sh = m.tp > 1;
    const int rank = sh ? m.d.attentionSegmentStart / m.d.attentionSegmentLength : 0;
    const int64_t ar0 = sh ? m.d.attentionSegmentStart : 0, ar = sh ? m.d.attentionSegmentLength : -1;
    const int64_t kr0 = sh ? m.d.kvSegmentStart : 0, kr = sh ? m.d.kvSegmentLength : -1;
    const int64_t hr0 = sh ? m.d.hiddenSegmentStart : 0, hr = sh ? m.d.hiddenSegmentLength : -1;
    v.push_back(expect(b, "model.embed_tokens.weight", 0, -1, 0, -1));
    v.push_back(expect(b, "model.norm.weight", 0, -1, 0, -1));
    v.push_back(expect(b, "lm_head.weight", 0, -1, 0, -1));
    for (int l = 0; l < LAYERS; l++) {
        const std::string p = "model.layers." + std::to_string(l) + ".";
        v.push_back(expect(b, p + "input_layernorm.weight", 0, -1, 0, -1));
        v.push_back(expect(b, p + "self_attn.q_proj.weight", ar0, ar, 0, -1));
        v.push_back(expect(b, p + "self_attn.k_proj.weight", kr0, kr, 0, -1));
        v.push_back(expect(b, p + "self_attn.v_proj.weight", kr0, kr, 0, -1));
        v.push_back(expect(b, p + "self_attn.o_proj.weight", 0, -1, ar0, ar));
        v.push_back(expect(b, p + "post_attention_layernorm.weight", 0, -1, 0, -1));
        if (moe) {
            v.push_back(expect(b, p + "block_sparse_moe.gate.weight", 0, -1, 0, -1));
            for (int e = 0; e < 4; e++) {
                if (e % m.tp != rank) continue;
                const std::string q = p + "block_sparse_moe.experts." + std::to_string(e) + ".";
                for (const char *w : {"w1.weight", "w2.weight", "w3.weight"}) v.push_back(expect(b, q + w, 0, -1, 0, -1));
            }
        } else {
            v.push_back(expect(b, p + "mlp.gate_proj.weight", hr0, hr, 0, -1));
            v.push_back(expect(b, p + "mlp.down_proj.weight", 0, -1, hr0, hr));
            v.push_back(expect(b, p + "mlp.up_proj.weight", hr0, hr, 0, -1));
        }
    }
    return v;
}

static jl_model shard(int rank, int tp, bool moe) {
    jl_model m = {};
    m.tp = tp, m.experts = moe ? 4 : 0, m.layers = LAYERS;
    m.d.numberOfLayers = LAYERS, m.d.layerEnd = LAYERS;
    m.d.attentionSegmentLength = HEADS * HS / tp, m.d.attentionSegmentStart = rank * m.d.attentionSegmentLength;
    m.d.kvSegmentLength = KVH * HS / tp, m.d.kvSegmentStart = rank * m.d.kvSegmentLength;
    m.d.hiddenSegmentLength = H / tp, m.d.hiddenSegmentStart = rank * m.d.hiddenSegmentLength;
    m.d.embeddingSegmentLength = E / tp, m.d.embeddingSegmentStart = rank * m.d.embeddingSegmentLength;
    return m;
}

static void touch_all(jl_st *st) {
    uint64_t acc = 0;
    for (int i = 0; i < jl_st_count(st); i++) {
        const char *name = nullptr;
        int dt = 0, nd = 0;
        int64_t shape[4], nb = 0;
        REQUIRE(jl_st_info(st, i, &name, &dt, &nd, shape, &nb) == JL_OK);
        REQUIRE(jl_st_find(st, name) >= 0);
        const uint8_t *p = (const uint8_t *)jl_st_data(st, i);
        for (int64_t k = 0; k < nb; k++) acc += p[k];
    }
    jl_st_majority_dtype(st);
    jl_st_metadata(st, "format");
    g_sink = acc;
}

int main(int argc, char **argv) {
    REQUIRE(argc == 2);
    const std::string dir = argv[1];
    std::mt19937 rng(2024);
    jl_ctx ctx = {};
    // ---- 1. well-formed checkpoints, every shard of tp 1 / 2 / 4 ------------------------------------------------------------------------
    long slices = 0;
    for (int moe = 0; moe < 2; moe++) {
        Blob b = make_checkpoint(moe != 0, rng);
        const std::string path = dir + (moe ? "/moe.safetensors" : "/dense.safetensors");
        REQUIRE(b.write(path) == JL_OK);
        jl_st *st = nullptr;
        REQUIRE(jl_st_open(path.c_str(), &st) == JL_OK);
        REQUIRE(jl_st_count(st) == (int)b.names.size());
        touch_all(st);
        for (int tp : {1, 2, 4})
            for (int rank = 0; rank < tp; rank++) {
                jl_model m = shard(rank, tp, moe != 0);
                int64_t ids[256];
                int n = 0;
                g_registered.clear();
                REQUIRE(jl_model_load_safetensors(&m, &ctx, st, ids, 256, &n) == JL_OK);
                REQUIRE(n == m.sets && n > 0);
                // the bytes each rank received are exactly its rows / columns of the stored tensors
                const std::vector<Registered> want = expected_registrations(b, m, moe != 0);
                REQUIRE(want.size() == g_registered.size());
                for (size_t i = 0; i < want.size(); i++) {
                    const Registered &g = g_registered[i], &w = want[i];
                    REQUIRE(g.dtype == w.dtype && g.rows == w.rows && g.cols == w.cols);
                    REQUIRE(g.data_sum == w.data_sum && g.scale_sum == w.scale_sum);
                }
                slices += n;
            }
        REQUIRE(jl_st_close(st) == JL_OK);
    }
    // ---- 2. header mutations -----------------------------------------------------------------------------------------------------------------
    Blob b = make_checkpoint(false, rng);
    const std::string good = dir + "/good.safetensors", bad = dir + "/mut.safetensors";
    REQUIRE(b.write(good) == JL_OK);
    std::string raw;
    REQUIRE(read_text(good, raw));
    int64_t hlen;
    memcpy(&hlen, raw.data(), 8);
    const std::string header = raw.substr(8, (size_t)hlen), payload = raw.substr(8 + (size_t)hlen);
    const char *inserts[] = {"-1", "99999999999999999999", "1e309", "[", "{", "\\u12", "\"", "null", "9223372036854775807", "0", ",", "}", "]"};
    int opened = 0, loaded = 0;
    for (int it = 0; it < 4000; it++) {
        std::string h = header;
        const int edits = 1 + (int)(rng() % 3);
        for (int e = 0; e < edits && !h.empty(); e++) {
            const size_t pos = rng() % h.size();
            switch (rng() % 4) {
                case 0: h[pos] = (char)rng(); break;
                case 1: h.erase(pos, 1 + rng() % 6); break;
                case 2: h.insert(pos, 1 + rng() % 4, (char)(32 + rng() % 95)); break;
                default: h.insert(pos, inserts[rng() % 13]);
            }
        }
        int64_t n = (int64_t)h.size();
        if (it % 97 == 0) n = (int64_t)(rng() % 3 ? -(int64_t)(rng() % 1000) : (1LL << 40)); // the length field itself
        FILE *f = fopen(bad.c_str(), "wb");
        REQUIRE(f);
        fwrite(&n, 8, 1, f), fwrite(h.data(), 1, h.size(), f), fwrite(payload.data(), 1, it % 5 ? payload.size() : payload.size() / 2, f);
        fclose(f);
        jl_st *st = nullptr;
        if (jl_st_open(bad.c_str(), &st) != JL_OK) {
            REQUIRE(st == nullptr && jl_st_last_error()[0] != 0);
            continue;
        }
        opened++;
        touch_all(st);
        jl_model m = shard((int)(rng() % 2), 2, false);
        int n_ids = 0;
        if (jl_model_load_safetensors(&m, &ctx, st, nullptr, 0, &n_ids) == JL_OK) loaded++;
        jl_st_close(st);
    }
    REQUIRE(opened > 0);
    // ---- 3. config.json reader on mutated text -------------------------------------------------------------------------------------------------
    const std::string cfg = "{\"hidden_size\": 64, \"intermediate_size\": 128, \"num_attention_heads\": 4, \"num_key_value_heads\": 2, \"num_hidden_layers\": 2,"
                            " \"vocab_size\": 96, \"max_position_embeddings\": 256, \"rms_norm_eps\": 1e-5, \"rope_theta\": 10000.0,"
                            " \"rope_scaling\": {\"rope_type\": \"linear\", \"factor\": 2.0}, \"num_local_experts\": 4, \"num_experts_per_tok\": 2}";
    int cfg_ok = 0;
    for (int it = 0; it < 2000; it++) {
        std::string h = cfg;
        if (it)
            for (int e = 0; e < 1 + (int)(rng() % 3) && !h.empty(); e++) {
                const size_t pos = rng() % h.size();
                if (rng() % 2) h[pos] = (char)rng();
                else h.insert(pos, inserts[rng() % 13]);
            }
        FILE *f = fopen((dir + "/config.json").c_str(), "wb");
        fwrite(h.data(), 1, h.size(), f);
        fclose(f);
        jl_model_config mc;
        if (jl_config_from_json((dir + "/config.json").c_str(), &mc) == JL_OK) {
            cfg_ok++;
            if (it == 0) REQUIRE(mc.embedding_length == 64 && mc.num_experts == 4 && mc.experts_per_token == 2 && mc.rope_scaling == 2.0 && mc.head_size == 16);
        }
    }
    REQUIRE(cfg_ok > 0);
    printf("ok: %ld slices registered in bounds (%llu bytes read), %d / 4000 mutated headers opened, %d loaded, %d / 2000 configs accepted\n", slices,
           (unsigned long long)g_bytes_read, opened, loaded, cfg_ok);
    return 0;
}
