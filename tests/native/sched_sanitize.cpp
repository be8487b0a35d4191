// Sanitizer harness for the session scheduler (jlama_b200/csrc/jl_sched.cu is plain host C++): the policy code is compiled with
// -fsanitize=address,undefined and again with -fsanitize=thread, and driven by several threads at once -- producers submitting fresh
// requests, kept sessions and follow-up turns, a canceller, a poller streaming results, a releaser -- while the main thread steps.
// The toy backend keeps a token history per session slot (the KV analogue), checks the protocol the device relies on and supports the
// host-spill trio.  At the end every request that finished on its own must hold exactly the tokens a replay of it alone produces.
// Built and run by tests/test_scheduler_sanitizers.py; exit code 0 = clean.
#include "../../jlama_b200/csrc/jl_sched.cu"

#include <atomic>
#include <map>
#include <random>
#include <thread>

// the jl_model backend of jl_sched_create is not under test here: satisfy the linker
extern "C" {
int jl_model_reset_session(jl_model *, int) { return JL_ERR_INVALID; }
int jl_model_batch_forward(jl_model *, int, const int32_t *, int, int) { return JL_ERR_INVALID; }
int jl_model_sample(jl_model *, int, float, float, int32_t *, float *) { return JL_ERR_INVALID; }
int jl_model_decode_sample(jl_model *, int, const int32_t *, const int32_t *, const int32_t *, const float *, const float *, int32_t *, float *) {
    return JL_ERR_INVALID;
}
int64_t jl_model_kv_offload(jl_model *, int) { return JL_ERR_INVALID; }
int jl_model_kv_restore(jl_model *, int, int64_t) { return JL_ERR_INVALID; }
int jl_model_kv_discard(jl_model *, int64_t) { return JL_ERR_INVALID; }
}
void jl_model_limits(jl_model *, int out[4]) { out[0] = out[1] = out[2] = out[3] = 0; }

static const int VOCAB = 97, CONTEXT = 80;
#define REQUIRE(cond)                                                              \
    do {                                                                           \
        if (!(cond)) {                                                             \
            fprintf(stderr, "REQUIRE failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            abort();                                                               \
        }                                                                          \
    } while (0)

static int32_t toy_next(const std::vector<int32_t> &h, float T, float u) {
    uint64_t x = 1469598103934665603ull;
    for (int32_t t : h) x = (x ^ (uint64_t)(t + 1)) * 1099511628211ull;
    if (T != 0.0f) x += (uint64_t)(u * 16777216.0f) * 2654435761ull + (uint64_t)(T * 1000.0f);
    return (int32_t)(x % VOCAB);
}

struct Toy {
    std::vector<std::vector<int32_t>> hist;
    std::vector<char> live;
    std::map<int64_t, std::vector<int32_t>> store;
    int64_t next_handle = 1;
    int max_rows;
    static int reset(void *u, int s) {
        Toy *t = (Toy *)u;
        t->hist[(size_t)s].clear(), t->live[(size_t)s] = 1;
        return JL_OK;
    }
    static int forward(void *u, int s, const int32_t *tok, int n, int start) {
        Toy *t = (Toy *)u;
        REQUIRE(t->live[(size_t)s] && start == (int)t->hist[(size_t)s].size());
        t->hist[(size_t)s].insert(t->hist[(size_t)s].end(), tok, tok + n);
        return JL_OK;
    }
    static int sample(void *u, int s, float T, float uni, int32_t *out) {
        Toy *t = (Toy *)u;
        *out = toy_next(t->hist[(size_t)s], T, uni);
        return JL_OK;
    }
    static int decode(void *u, int n, const int32_t *sess, const int32_t *tok, const int32_t *pos, const float *T, const float *uni, int32_t *next) {
        Toy *t = (Toy *)u;
        REQUIRE(n >= 1 && n <= t->max_rows);
        for (int i = 0; i < n; i++) {
            for (int j = 0; j < i; j++) REQUIRE(sess[i] != sess[j]);
            auto &h = t->hist[(size_t)sess[i]];
            REQUIRE(t->live[(size_t)sess[i]] && pos[i] == (int)h.size() && pos[i] < CONTEXT);
            h.push_back(tok[i]);
            next[i] = toy_next(h, T[i], uni[i]);
        }
        return JL_OK;
    }
    static int offload(void *u, int s, int64_t *handle) {
        Toy *t = (Toy *)u;
        REQUIRE(t->live[(size_t)s]);
        *handle = t->next_handle++;
        t->store[*handle] = std::move(t->hist[(size_t)s]);
        t->hist[(size_t)s].clear(), t->live[(size_t)s] = 0;
        return JL_OK;
    }
    static int restore(void *u, int s, int64_t handle) {
        Toy *t = (Toy *)u;
        auto it = t->store.find(handle);
        REQUIRE(it != t->store.end());
        t->hist[(size_t)s] = std::move(it->second), t->live[(size_t)s] = 1;
        t->store.erase(it);
        return JL_OK;
    }
    static int discard(void *u, int64_t handle) {
        Toy *t = (Toy *)u;
        REQUIRE(t->store.erase(handle) == 1);
        return JL_OK;
    }
};

struct Meta {
    std::vector<int32_t> prompt;
    int max_new;
    int32_t stop;
    float T;
    uint64_t seed;
    int64_t parent;
    bool keep;
};

int main() {
    const int SLOTS = 3;
    Toy toy;
    toy.hist.resize(SLOTS), toy.live.assign(SLOTS, 0), toy.max_rows = 2;
    jl_sched_backend be = {Toy::reset, Toy::forward, Toy::sample, Toy::decode, Toy::offload, Toy::restore, Toy::discard};
    jl_sched *s = nullptr;
    REQUIRE(jl_sched_create_backend(&be, &toy, SLOTS, 2, CONTEXT, 5, &s) == JL_OK);

    std::mutex meta_mu;
    std::map<int64_t, Meta> meta;
    std::vector<int64_t> kept; // finished kept requests open for a follow-up (guarded by meta_mu)
    std::atomic<bool> producing{true};
    std::atomic<int> submitted{0};

    auto producer = [&](unsigned seed) {
        std::mt19937 rng(seed);
        for (int it = 0; it < 120; it++) {
            Meta m;
            m.parent = -1;
            {
                std::lock_guard<std::mutex> lk(meta_mu);
                if (!kept.empty() && rng() % 2) {
                    const size_t k = rng() % kept.size();
                    m.parent = kept[k];
                    kept.erase(kept.begin() + (long)k);
                }
            }
            if (it % 11 == 0) jl_sched_submit(s, nullptr, 0, 0, nullptr, 0, 0, -1, 0.0f, 0); // a rejected submit writes the error text
            const int n = 1 + (int)(rng() % 7);
            for (int i = 0; i < n; i++) m.prompt.push_back((int32_t)(rng() % VOCAB));
            m.max_new = 1 + (int)(rng() % 6);
            m.stop = rng() % 3 ? -1 : (int32_t)(rng() % VOCAB);
            m.T = rng() % 3 ? 0.0f : 0.7f;
            m.seed = rng();
            m.keep = rng() % 2;
            const int64_t id = jl_sched_submit(s, m.prompt.data(), n, m.max_new, m.stop >= 0 ? &m.stop : nullptr, m.stop >= 0 ? 1 : 0,
                                               m.keep ? JL_SCHED_KEEP_SESSION : 0, m.parent, m.T, m.seed);
            if (id > 0) {
                std::lock_guard<std::mutex> lk(meta_mu);
                meta[id] = m;
                submitted++;
            }
            if (it % 8 == 0) std::this_thread::yield();
        }
    };
    auto canceller = [&] {
        std::mt19937 rng(99);
        while (producing) {
            int64_t id = -1;
            {
                std::lock_guard<std::mutex> lk(meta_mu);
                if (!meta.empty() && rng() % 4 == 0) {
                    auto it = meta.begin();
                    std::advance(it, (long)(rng() % meta.size()));
                    id = it->first;
                }
            }
            if (id > 0) jl_sched_cancel(s, id);
            std::this_thread::yield();
        }
    };
    auto poller = [&] { // streams results like a serving thread; collects finished kept requests for follow-ups
        std::vector<int32_t> buf(64);
        std::map<int64_t, int> seen;
        while (producing) {
            std::vector<int64_t> ids;
            {
                std::lock_guard<std::mutex> lk(meta_mu);
                for (auto &kv : meta) ids.push_back(kv.first);
            }
            for (int64_t id : ids) {
                int n = 0, state = 0, reason = 0;
                if (jl_sched_result(s, id, buf.data(), (int)buf.size(), &n, &state, &reason) != JL_OK) continue;
                jl_sched_request_info_t info;
                jl_sched_request_info(s, id, &info);
                if (jl_sched_last_error(s)[0] == 1) abort(); // reads the error text while other threads may be failing (bad submits)
                if (state == JL_SCHED_FINISHED && reason != JL_FINISH_CANCELLED && !seen[id]) {
                    seen[id] = 1;
                    std::lock_guard<std::mutex> lk(meta_mu);
                    if (meta[id].keep) kept.push_back(id);
                }
            }
            std::this_thread::yield();
        }
    };

    std::atomic<int> producers_done{0};
    std::vector<std::thread> threads;
    for (unsigned p = 0; p < 3; p++)
        threads.emplace_back([&, p] {
            producer(1000u + p);
            producers_done++;
        });
    std::thread tc(canceller), tp(poller);
    // the serving loop: step while the producers run, then drain
    int spilled = 0;
    while (producers_done.load() < 3) {
        jl_sched_stats st;
        jl_sched_step(s, &st);
        REQUIRE(st.active <= SLOTS);
        spilled += st.spilled;
    }
    for (auto &t : threads) t.join();
    jl_sched_stats totals;
    jl_sched_run(s, 0, &totals);
    spilled += totals.spilled;
    producing = false;
    tc.join(), tp.join();
    jl_sched_run(s, 0, &totals); // a cancel or a follow-up may have landed after the drain
    spilled += totals.spilled;

    // ---- every request that finished on its own: tokens == a replay of the request alone on its history chain ---------------------
    std::map<int64_t, std::vector<int32_t>> toks;
    std::map<int64_t, std::pair<int, int>> fin;
    for (auto &kv : meta) {
        std::vector<int32_t> buf(64);
        int n = 0, state = 0, reason = 0;
        REQUIRE(jl_sched_result(s, kv.first, buf.data(), 64, &n, &state, &reason) == JL_OK);
        REQUIRE(state == JL_SCHED_FINISHED || state == JL_SCHED_FAILED);
        buf.resize((size_t)n);
        toks[kv.first] = buf, fin[kv.first] = {state, reason};
    }
    int checked = 0;
    for (auto &kv : meta) {
        const Meta &m = kv.second;
        if (fin[kv.first].first != JL_SCHED_FINISHED || fin[kv.first].second == JL_FINISH_CANCELLED) continue;
        // history: ancestors' prompts and forwarded tokens, oldest first
        std::vector<int64_t> chain;
        for (int64_t q = m.parent; q >= 0; q = meta[q].parent) chain.push_back(q);
        std::vector<int32_t> hist;
        for (size_t i = chain.size(); i-- > 0;) {
            const int64_t q = chain[i];
            hist.insert(hist.end(), meta[q].prompt.begin(), meta[q].prompt.end());
            hist.insert(hist.end(), toks[q].begin(), toks[q].end() - 1);
        }
        hist.insert(hist.end(), m.prompt.begin(), m.prompt.end());
        uint64_t x = m.seed;
        std::vector<int32_t> out;
        for (;;) {
            float u = 0.0f;
            if (m.T != 0.0f) {
                uint64_t z = (x += 0x9E3779B97F4A7C15ull);
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                z ^= z >> 31;
                u = (float)(z >> 40) * (1.0f / 16777216.0f);
            }
            out.push_back(toy_next(hist, m.T, u));
            if ((out.size() > 1 && out.back() == m.stop) || (int)out.size() >= m.max_new || (int)hist.size() >= CONTEXT) break;
            hist.push_back(out.back());
        }
        REQUIRE(out == toks[kv.first]);
        checked++;
    }
    REQUIRE(checked >= 20);
    // ---- release everything (children before parents would be refused only while queued: nothing is queued now) -------------------
    for (int pass = 0; pass < 2; pass++)
        for (auto &kv : meta) jl_sched_release(s, kv.first);
    jl_sched_stats st;
    jl_sched_step(s, &st);
    int q = 0, a = 0, f = 0;
    jl_sched_counts(s, &q, &a, &f);
    REQUIRE(q == 0 && a == 0 && f == SLOTS && toy.store.empty());
    REQUIRE(jl_sched_free(s) == JL_OK);
    REQUIRE(spilled > 0);
    printf("ok: %d requests, %d replayed, %d spilled\n", (int)meta.size(), checked, spilled);
    return 0;
}
