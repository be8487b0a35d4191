// Concurrent-session scheduler: iteration-level ("continuous") batching of generation requests into the batched decode step.
//
// The reference serves concurrent requests with one thread per request, each running AbstractModel.generate() on its own KvBuffer
// (core/tensor/KvBufferCache.java:58-60; jlama-net/.../openai/OpenAIChatService.java:64-74,107-160): the weights are streamed once
// per thread per token.  On the GPU the unit that shares one weight stream is the *decode step over N sessions* (jl_model_decode,
// csrc/jl_gemm8.cu), so concurrency becomes a queue in front of that step:
//
//   every jl_sched_step():   admit queued requests into free session slots (FIFO)            KvBufferCache.getKvBuffer
//                            forward prompt chunks of the admitted requests (token budget)    AbstractModel.batchForward :295-312
//                            sample their first token when a prompt is complete               AbstractModel.generate :576
//                            ONE decode step for every generating request, <= max_rows rows   AbstractModel.generate :590-600
//                              per call, requests of different lengths side by side
//                            retire requests that hit a stop token / their token limit        :604-608, FinishReason
//                            -> their slots are free for the next step's admissions
//
// The scheduler is pure host logic above four backend calls (reset_session / batch_forward / sample / decode).  jl_sched_create
// binds them to a jl_model; jl_sched_create_backend takes them as function pointers, which is how the CPU tests drive the very same
// policy code with the oracle as the "device" (tests/test_scheduler.py) and how a host that owns its own model object would plug in.
// Kept sessions and host spill: a finished JL_SCHED_KEEP_SESSION request holds its slot and KV for a follow-up turn.  When a request
// finds no free slot and the backend can offload (jl_model_kv_offload), the least recently finished kept session is moved to host
// memory and its slot is reused; its continuation later restores the pages into whatever slot is free (jl_model_kv_restore).
// Locking: `step_mu` serialises steps; `mu` guards the request table and is NOT held across backend calls, so submit / result /
// cancel from other threads never wait for GPU work.
#include <chrono>
#include <deque>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/jlama_b200.h"

void jl_model_limits(jl_model *m, int out[4]); // jl_model.cu: {max_sessions, max_context, max_batch, max rows per decode call}

namespace {

struct Request {
    int64_t id = 0;
    std::vector<int32_t> prompt;
    std::vector<int32_t> stop;
    int max_new = 0;
    int flags = 0;
    float temperature = 0.0f;
    uint64_t rng = 0; // splitmix64 state: the k-th sampled token of a request uses its k-th draw, however the steps were batched
    int state = JL_SCHED_QUEUED;
    int reason = JL_FINISH_NONE;
    int session = -1;
    int start_pos = 0;  // context position of prompt[0] (> 0 when the request continues a kept session)
    int prefilled = 0;  // prompt tokens already forwarded
    int forwarded = 0;  // generated tokens already fed back through a decode step
    bool cancel = false;
    int64_t parent = -1;     // finished request whose kept session (and KV) this request continues
    int64_t child = -1;      // pending / admitted continuation of this request
    bool spilling = false;   // chosen as the victim of this step's offload; `spill` is set when the backend call returns
    int64_t spill = -1;      // backend handle of this finished request's KV after it was offloaded to host memory (session == -1 then)
    std::vector<int32_t> out;
    uint64_t submit_step = 0, first_token_step = 0, finish_step = 0;
    // host clock, like Generator.Response's promptTimeMs / generateTimeMs (AbstractModel.java:561-568,589,623): queued -> admitted ->
    // first token -> finished
    std::chrono::steady_clock::time_point t_submit, t_admit, t_first, t_finish;
    bool admitted = false;
    int next_pos() const { return start_pos + (int)prompt.size() + forwarded; } // position the next decode step writes
};

} // namespace

struct jl_sched {
    jl_sched_backend be;
    void *user = nullptr;
    int n_sessions = 0, max_rows = 0, max_context = 0, prefill_budget = 0;
    std::mutex mu, step_mu;
    std::unordered_map<int64_t, Request> reqs;
    std::deque<int64_t> queue;        // admission order
    std::vector<int64_t> active;      // admitted, unfinished requests in admission order
    std::vector<int64_t> slot_owner;  // session slot -> request id, -1 = free; a kept session stays owned by its finished request
    int64_t next_id = 1;
    uint64_t step_no = 0;
    std::string last_error;
    jl_sched_stats totals = {};
    std::vector<int64_t> discards; // spill handles of released requests, dropped by the next step (backend calls never run under `mu`)
};

static int sched_error(jl_sched *s, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    s->last_error = buf;
    return code;
}

// ThreadLocalRandom.current().nextFloat() of AbstractModel.generate (:576, :594): 24 random bits / 2^24 in [0, 1), from a per-request
// splitmix64 stream so that a (seed, prompt) pair reproduces its tokens.
static float next_uniform(Request &r) {
    uint64_t z = (r.rng += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// ---- jl_model backend ---------------------------------------------------------------------------------------------------------
static int mb_reset(void *u, int session) { return jl_model_reset_session((jl_model *)u, session); }
static int mb_forward(void *u, int session, const int32_t *tokens, int n, int start_pos) {
    return jl_model_batch_forward((jl_model *)u, session, tokens, n, start_pos);
}
static int mb_sample(void *u, int session, float temperature, float uniform, int32_t *token) {
    return jl_model_sample((jl_model *)u, session, temperature, uniform, token, nullptr);
}
static int mb_decode(void *u, int n, const int32_t *sessions, const int32_t *tokens, const int32_t *positions, const float *temperatures,
                     const float *uniforms, int32_t *next) {
    return jl_model_decode_sample((jl_model *)u, n, sessions, tokens, positions, temperatures, uniforms, next, nullptr);
}

static int mb_offload(void *u, int session, int64_t *handle) {
    const int64_t h = jl_model_kv_offload((jl_model *)u, session);
    if (h < 0) return (int)h;
    *handle = h;
    return JL_OK;
}
static int mb_restore(void *u, int session, int64_t handle) { return jl_model_kv_restore((jl_model *)u, session, handle); }
static int mb_discard(void *u, int64_t handle) { return jl_model_kv_discard((jl_model *)u, handle); }

extern "C" int jl_sched_create_backend(const jl_sched_backend *be, void *user, int n_sessions, int max_rows, int max_context,
                                       int prefill_tokens_per_step, jl_sched **out) {
    if (!be || !out || !be->reset_session || !be->batch_forward || !be->sample || !be->decode || n_sessions <= 0 || max_rows <= 0 ||
        max_context <= 1)
        return JL_ERR_INVALID;
    if ((be->offload != nullptr) != (be->restore != nullptr) || (be->offload != nullptr) != (be->discard != nullptr))
        return JL_ERR_INVALID; // host spill is all three calls or none
    jl_sched *s = new jl_sched();
    s->be = *be;
    s->user = user;
    s->n_sessions = n_sessions;
    s->max_rows = max_rows;
    s->max_context = max_context;
    s->prefill_budget = prefill_tokens_per_step > 0 ? prefill_tokens_per_step : 0;
    s->slot_owner.assign((size_t)n_sessions, -1);
    *out = s;
    return JL_OK;
}

extern "C" int jl_sched_create(jl_model *m, int max_active, int prefill_tokens_per_step, jl_sched **out) {
    if (!m || !out) return JL_ERR_INVALID;
    int lim[4] = {0, 0, 0, 0};
    jl_model_limits(m, lim);
    if (lim[0] <= 0) return JL_ERR_INVALID; // not finalized
    const int n = max_active > 0 && max_active < lim[0] ? max_active : lim[0];
    static const jl_sched_backend be = {mb_reset, mb_forward, mb_sample, mb_decode, mb_offload, mb_restore, mb_discard};
    return jl_sched_create_backend(&be, m, n, lim[3], lim[1], prefill_tokens_per_step, out);
}

extern "C" int jl_sched_free(jl_sched *s) {
    if (!s) return JL_ERR_INVALID;
    {
        std::lock_guard<std::mutex> a(s->step_mu); // a step in flight finishes first
    }
    // host copies of spilled sessions nobody will restore any more
    if (s->be.discard) {
        for (int64_t h : s->discards) s->be.discard(s->user, h);
        for (auto &kv : s->reqs)
            if (kv.second.spill >= 0) s->be.discard(s->user, kv.second.spill);
    }
    delete s;
    return JL_OK;
}

// a copy the calling thread owns (the shared string is rewritten under `mu` by whichever thread fails next)
extern "C" const char *jl_sched_last_error(jl_sched *s) {
    static thread_local std::string mine;
    if (!s) return "null scheduler";
    std::lock_guard<std::mutex> lk(s->mu);
    mine = s->last_error;
    return mine.c_str();
}

extern "C" int64_t jl_sched_submit(jl_sched *s, const int32_t *prompt, int n_prompt, int max_new, const int32_t *stop_tokens,
                                   int n_stop, int flags, int64_t continue_request, float temperature, uint64_t seed) {
    if (!s) return -1;
    std::lock_guard<std::mutex> lk(s->mu);
    if (!prompt || n_prompt <= 0 || max_new <= 0 || n_stop < 0 || (n_stop > 0 && !stop_tokens) || !(temperature >= 0.0f))
        return sched_error(s, -1, "submit: bad arguments (n_prompt=%d, max_new=%d, temperature=%g)", n_prompt, max_new, (double)temperature);
    Request r;
    if (continue_request >= 0) {
        // AbstractModel.generate :533 startPos = kvmem.getCurrentContextPosition(): a follow-up on the same session appends to its KV
        auto it = s->reqs.find(continue_request);
        if (it == s->reqs.end() || it->second.state != JL_SCHED_FINISHED || !(it->second.flags & JL_SCHED_KEEP_SESSION) ||
            (it->second.session < 0 && it->second.spill < 0 && !it->second.spilling))
            return sched_error(s, -1, "submit: request %lld is not a finished request that kept its session", (long long)continue_request);
        if (it->second.child >= 0)
            return sched_error(s, -1, "submit: request %lld already has a continuation", (long long)continue_request);
        r.start_pos = it->second.next_pos();
        r.parent = continue_request;
    }
    // Preconditions :530: the prompt must leave room for at least the token sampled from it
    if (r.start_pos + n_prompt >= s->max_context)
        return sched_error(s, -1, "submit: prompt of %d tokens at position %d exceeds the context of %d", n_prompt, r.start_pos,
                           s->max_context);
    r.id = s->next_id++;
    r.prompt.assign(prompt, prompt + n_prompt);
    if (n_stop > 0) r.stop.assign(stop_tokens, stop_tokens + n_stop);
    r.max_new = max_new;
    r.flags = flags;
    r.temperature = temperature;
    r.rng = seed;
    r.submit_step = s->step_no;
    r.t_submit = r.t_admit = r.t_first = r.t_finish = std::chrono::steady_clock::now();
    r.out.reserve((size_t)(max_new < 4096 ? max_new : 4096));
    const int64_t id = r.id;
    if (r.parent >= 0) s->reqs[r.parent].child = id;
    s->reqs.emplace(id, std::move(r));
    s->queue.push_back(id);
    return id;
}

extern "C" int jl_sched_cancel(jl_sched *s, int64_t request) {
    if (!s) return JL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(s->mu);
    auto it = s->reqs.find(request);
    if (it == s->reqs.end()) return sched_error(s, JL_ERR_INVALID, "cancel: unknown request %lld", (long long)request);
    if (it->second.state == JL_SCHED_FINISHED || it->second.state == JL_SCHED_FAILED) return JL_OK;
    it->second.cancel = true; // takes effect at the next step boundary
    return JL_OK;
}

extern "C" int jl_sched_result(jl_sched *s, int64_t request, int32_t *tokens, int cap, int *n_tokens, int *state, int *finish_reason) {
    if (!s || cap < 0 || (cap > 0 && !tokens)) return JL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(s->mu);
    auto it = s->reqs.find(request);
    if (it == s->reqs.end()) return sched_error(s, JL_ERR_INVALID, "result: unknown request %lld", (long long)request);
    const Request &r = it->second;
    const int n = (int)r.out.size();
    if (tokens && cap > 0) memcpy(tokens, r.out.data(), sizeof(int32_t) * (size_t)(n < cap ? n : cap));
    if (n_tokens) *n_tokens = n;
    if (state) *state = r.state;
    if (finish_reason) *finish_reason = r.reason;
    return JL_OK;
}

extern "C" int jl_sched_request_info(jl_sched *s, int64_t request, jl_sched_request_info_t *info) {
    if (!s || !info) return JL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(s->mu);
    auto it = s->reqs.find(request);
    if (it == s->reqs.end()) return sched_error(s, JL_ERR_INVALID, "info: unknown request %lld", (long long)request);
    const Request &r = it->second;
    info->state = r.state, info->finish_reason = r.reason, info->session = r.session, info->start_pos = r.start_pos;
    info->n_prompt = (int)r.prompt.size(), info->n_prefilled = r.prefilled, info->n_generated = (int)r.out.size();
    info->next_position = r.next_pos();
    info->spilled = r.spill >= 0 ? 1 : 0;
    info->submit_step = (int64_t)r.submit_step, info->first_token_step = (int64_t)r.first_token_step;
    info->finish_step = (int64_t)r.finish_step;
    // phases that have not ended yet are measured up to now
    const auto now = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const bool admitted = r.admitted;
    const bool has_first = !r.out.empty();
    const bool done = r.state == JL_SCHED_FINISHED || r.state == JL_SCHED_FAILED;
    info->queue_ms = ms(r.t_submit, admitted ? r.t_admit : (done ? r.t_finish : now));
    info->prompt_ms = admitted ? ms(r.t_admit, has_first ? r.t_first : (done ? r.t_finish : now)) : 0.0;
    info->generate_ms = has_first ? ms(r.t_first, done ? r.t_finish : now) : 0.0;
    return JL_OK;
}

// free the slot a finished request kept (or drop its bookkeeping)
extern "C" int jl_sched_release(jl_sched *s, int64_t request) {
    if (!s) return JL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(s->mu);
    auto it = s->reqs.find(request);
    if (it == s->reqs.end()) return sched_error(s, JL_ERR_INVALID, "release: unknown request %lld", (long long)request);
    Request &r = it->second;
    if (r.state != JL_SCHED_FINISHED && r.state != JL_SCHED_FAILED)
        return sched_error(s, JL_ERR_INVALID, "release: request %lld is still running (cancel it first)", (long long)request);
    if (r.child >= 0) {
        auto ch = s->reqs.find(r.child);
        if (ch != s->reqs.end() && ch->second.state == JL_SCHED_QUEUED)
            return sched_error(s, JL_ERR_INVALID, "release: request %lld has a queued continuation", (long long)request);
        if (ch != s->reqs.end()) ch->second.parent = -1;
    }
    if (r.parent >= 0) {
        auto pa = s->reqs.find(r.parent);
        if (pa != s->reqs.end()) pa->second.child = -1;
    }
    if (r.session >= 0 && s->slot_owner[(size_t)r.session] == request) s->slot_owner[(size_t)r.session] = -1;
    if (r.spill >= 0) s->discards.push_back(r.spill);
    s->reqs.erase(it);
    return JL_OK;
}

extern "C" int jl_sched_counts(jl_sched *s, int *queued, int *active, int *free_slots) {
    if (!s) return JL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(s->mu);
    if (queued) *queued = (int)s->queue.size();
    if (active) *active = (int)s->active.size();
    if (free_slots) {
        int f = 0;
        for (int64_t o : s->slot_owner) f += o < 0;
        *free_slots = f;
    }
    return JL_OK;
}

// must hold s->mu
static void finish(jl_sched *s, Request &r, int state, int reason) {
    r.state = state;
    r.reason = reason;
    r.finish_step = s->step_no;
    r.t_finish = std::chrono::steady_clock::now();
    for (size_t i = 0; i < s->active.size(); i++)
        if (s->active[i] == r.id) {
            s->active.erase(s->active.begin() + (long)i);
            break;
        }
    if (r.session >= 0) {
        const bool keep = state == JL_SCHED_FINISHED && (r.flags & JL_SCHED_KEEP_SESSION) && reason != JL_FINISH_CANCELLED;
        if (keep) {
            s->slot_owner[(size_t)r.session] = r.id; // the finished request holds the KV until released or continued
        } else if (s->slot_owner[(size_t)r.session] == r.id) {
            s->slot_owner[(size_t)r.session] = -1;
        }
    }
}

// must hold s->mu.  After a token was appended: AbstractModel.generate's exits (:590 position limit, :604-608 stop token).
static void check_done(jl_sched *s, Request &r, bool stop_check) {
    const int32_t t = r.out.back();
    if (stop_check)
        for (int32_t e : r.stop)
            if (e == t) return finish(s, r, JL_SCHED_FINISHED, JL_FINISH_STOP_TOKEN);
    if ((int)r.out.size() >= r.max_new || r.next_pos() >= s->max_context) finish(s, r, JL_SCHED_FINISHED, JL_FINISH_MAX_TOKENS);
}

namespace {
struct SlotJob { // what has to happen to a session slot before its new owner's first prompt chunk
    int64_t victim = -1; // kept request whose KV is offloaded out of the slot first
    int64_t restore_from = -1; // request whose spilled KV is restored into the slot (continuation of a spilled session)
    bool reset = false;        // fresh request: zero the slot
};
struct PrefillJob {
    int64_t id;
    int session, start, n, pos;
    bool last;
    float temperature = 0.0f, uniform = 0.0f; // for the token sampled from the prompt's last row
    SlotJob slot;
    std::vector<int32_t> tokens;
};
struct Row {
    int64_t id;
    int32_t session, token, position;
    float temperature, uniform;
};
} // namespace

// must hold s->mu.  A slot for a request that needs one: a free slot, else (backend permitting) the slot of the kept session that
// finished longest ago -- sessions whose follow-up is already queued go last.  *victim = the request to offload, or -1.
static int take_slot(jl_sched *s, int64_t *victim) {
    *victim = -1;
    for (int k = 0; k < s->n_sessions; k++)
        if (s->slot_owner[(size_t)k] < 0) return k;
    if (!s->be.offload) return -1;
    int best = -1;
    for (int k = 0; k < s->n_sessions; k++) {
        auto it = s->reqs.find(s->slot_owner[(size_t)k]);
        if (it == s->reqs.end() || it->second.state != JL_SCHED_FINISHED || it->second.session != k) continue; // running, not kept
        if (best < 0) {
            best = k;
            continue;
        }
        const Request &a = it->second, &b = s->reqs[s->slot_owner[(size_t)best]];
        const bool a_waits = a.child >= 0, b_waits = b.child >= 0;
        if (a_waits != b_waits ? !a_waits : a.finish_step < b.finish_step) best = k;
    }
    if (best >= 0) *victim = s->slot_owner[(size_t)best];
    return best;
}

extern "C" int jl_sched_step(jl_sched *s, jl_sched_stats *stats) {
    if (!s) return JL_ERR_INVALID;
    std::lock_guard<std::mutex> step_lock(s->step_mu);
    jl_sched_stats st = {};
    int first_error = JL_OK;
    std::vector<PrefillJob> jobs;
    std::unordered_map<int64_t, SlotJob> slot_jobs; // admitted this step: request id -> slot preparation
    std::vector<int64_t> discards;

    // ---- plan: cancellations, admissions, prompt chunks ----------------------------------------------------------------------
    {
        std::lock_guard<std::mutex> lk(s->mu);
        s->step_no++;
        discards.swap(s->discards);
        for (size_t i = 0; i < s->active.size();) {
            Request &r = s->reqs[s->active[i]];
            if (r.cancel) finish(s, r, JL_SCHED_FINISHED, JL_FINISH_CANCELLED), st.finished++;
            else i++;
        }
        // FIFO admission.  A continuation whose parent still sits in its slot returns to that slot (always possible); every other
        // request needs a slot from take_slot, and the first one that finds none blocks the slot-seeking requests behind it.
        bool blocked = false;
        for (size_t qi = 0; qi < s->queue.size();) {
            Request &r = s->reqs[s->queue[qi]];
            auto pa = r.parent >= 0 ? s->reqs.find(r.parent) : s->reqs.end();
            if (r.cancel) {
                if (pa != s->reqs.end()) pa->second.child = -1; // the parent still holds the KV and may be continued again
                finish(s, r, JL_SCHED_FINISHED, JL_FINISH_CANCELLED), st.finished++;
                s->queue.erase(s->queue.begin() + (long)qi);
                continue;
            }
            SlotJob sj;
            if (pa != s->reqs.end() && pa->second.session >= 0) {
                r.session = pa->second.session; // the parent's slot with its KV; ownership moves to the continuation
                pa->second.session = -1;
            } else {
                int64_t victim = -1;
                const int slot = blocked ? -1 : take_slot(s, &victim);
                if (slot < 0) {
                    blocked = true;
                    qi++;
                    continue;
                }
                if (victim >= 0) {
                    Request &v = s->reqs[victim];
                    v.session = -1; // its KV leaves the device in this step (spill handle set when the offload returns)
                    v.spilling = true;
                    sj.victim = victim;
                    st.spilled++;
                }
                r.session = slot;
                if (pa != s->reqs.end()) sj.restore_from = r.parent; // continuation of a spilled session
                else sj.reset = true;
            }
            s->slot_owner[(size_t)r.session] = r.id;
            r.state = JL_SCHED_PREFILL;
            r.t_admit = std::chrono::steady_clock::now();
            r.admitted = true;
            s->active.push_back(r.id);
            s->queue.erase(s->queue.begin() + (long)qi);
            slot_jobs[r.id] = sj;
            st.admitted++;
        }
        int budget = s->prefill_budget > 0 ? s->prefill_budget : 0x7fffffff;
        for (int64_t id : s->active) {
            Request &r = s->reqs[id];
            if (r.state != JL_SCHED_PREFILL) continue;
            auto sj = slot_jobs.find(id);
            if (budget <= 0 && sj == slot_jobs.end()) continue;
            const int left = (int)r.prompt.size() - r.prefilled;
            const int n = left < budget ? left : (budget > 0 ? budget : 0);
            PrefillJob j;
            j.id = id, j.session = r.session, j.start = r.prefilled, j.n = n, j.pos = r.start_pos + r.prefilled;
            if (sj != slot_jobs.end()) j.slot = sj->second; // the slot is prepared in the admission step even if the budget is spent
            j.last = n > 0 && r.prefilled + n == (int)r.prompt.size();
            if (j.last && r.temperature != 0.0f) j.temperature = r.temperature, j.uniform = next_uniform(r);
            j.tokens.assign(r.prompt.begin() + r.prefilled, r.prompt.begin() + r.prefilled + n);
            jobs.push_back(std::move(j));
            budget -= n;
        }
    }

    // ---- backend calls, table unlocked: dropped spills, slot preparation, prompt chunks ---------------------------------------------
    for (int64_t h : discards) s->be.discard(s->user, h);
    for (PrefillJob &j : jobs) {
        int rc = JL_OK;
        int32_t tok = 0;
        if (j.slot.victim >= 0) {
            int64_t handle = -1;
            const int orc = s->be.offload(s->user, j.session, &handle);
            std::lock_guard<std::mutex> lk(s->mu);
            auto v = s->reqs.find(j.slot.victim);
            if (v != s->reqs.end()) v->second.spilling = false;
            if (orc == JL_OK && v != s->reqs.end()) v->second.spill = handle;
            else if (orc == JL_OK) s->discards.push_back(handle); // released meanwhile
            else {
                // the KV of the kept session could not be saved: it can no longer be continued; the slot is reused regardless
                if (v != s->reqs.end()) v->second.flags &= ~JL_SCHED_KEEP_SESSION;
                if (first_error == JL_OK) first_error = sched_error(s, orc, "offload of kept request %lld failed (%d): its session is dropped", (long long)j.slot.victim, orc);
            }
        }
        if (j.slot.restore_from >= 0) {
            int64_t handle = -1;
            {
                std::lock_guard<std::mutex> lk(s->mu);
                auto pa = s->reqs.find(j.slot.restore_from);
                if (pa != s->reqs.end()) handle = pa->second.spill, pa->second.spill = -1;
            }
            rc = handle >= 0 ? s->be.restore(s->user, j.session, handle) : JL_ERR_INVALID;
        } else if (j.slot.reset) {
            rc = s->be.reset_session(s->user, j.session);
        }
        if (rc == JL_OK && j.n > 0) rc = s->be.batch_forward(s->user, j.session, j.tokens.data(), j.n, j.pos);
        if (rc == JL_OK && j.last) rc = s->be.sample(s->user, j.session, j.temperature, j.uniform, &tok);
        std::lock_guard<std::mutex> lk(s->mu);
        Request &r = s->reqs[j.id];
        if (rc != JL_OK) {
            if (first_error == JL_OK) first_error = sched_error(s, rc, "request %lld: prompt forward failed (%d)", (long long)j.id, rc);
            finish(s, r, JL_SCHED_FAILED, JL_FINISH_ERROR), st.finished++;
            continue;
        }
        r.prefilled += j.n;
        st.prefill_tokens += j.n;
        if (j.last) {
            r.out.push_back(tok);
            r.first_token_step = s->step_no;
            r.t_first = std::chrono::steady_clock::now();
            r.state = JL_SCHED_DECODING;
            // the token sampled from the prompt is not stop-checked by the reference (AbstractModel.java:576-589)
            check_done(s, r, false);
            if (r.state != JL_SCHED_DECODING) st.finished++;
        }
    }

    // ---- one decode step for every generating request, max_rows rows per backend call ------------------------------------------------
    std::vector<Row> rows;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        for (int64_t id : s->active) {
            Request &r = s->reqs[id];
            if (r.state != JL_SCHED_DECODING) continue;
            const float u = r.temperature != 0.0f ? next_uniform(r) : 0.0f;
            rows.push_back(Row{id, r.session, r.out.back(), r.next_pos(), r.temperature, u});
        }
    }
    std::vector<int32_t> sess((size_t)s->max_rows), toks((size_t)s->max_rows), pos((size_t)s->max_rows), next((size_t)s->max_rows);
    std::vector<float> temps((size_t)s->max_rows), unis((size_t)s->max_rows);
    for (size_t g = 0; g < rows.size(); g += (size_t)s->max_rows) {
        const int n = (int)(rows.size() - g < (size_t)s->max_rows ? rows.size() - g : (size_t)s->max_rows);
        for (int i = 0; i < n; i++) {
            const Row &w = rows[g + (size_t)i];
            sess[(size_t)i] = w.session, toks[(size_t)i] = w.token, pos[(size_t)i] = w.position, temps[(size_t)i] = w.temperature, unis[(size_t)i] = w.uniform;
        }
        const int rc = s->be.decode(s->user, n, sess.data(), toks.data(), pos.data(), temps.data(), unis.data(), next.data());
        st.decode_calls++;
        std::lock_guard<std::mutex> lk(s->mu);
        for (int i = 0; i < n; i++) {
            Request &r = s->reqs[rows[g + i].id];
            if (rc != JL_OK) {
                if (first_error == JL_OK) first_error = sched_error(s, rc, "decode step of %d sessions failed (%d)", n, rc);
                finish(s, r, JL_SCHED_FAILED, JL_FINISH_ERROR), st.finished++;
                continue;
            }
            r.forwarded++;
            r.out.push_back(next[(size_t)i]);
            st.decode_rows++;
            check_done(s, r, true);
            if (r.state != JL_SCHED_DECODING) st.finished++;
        }
    }

    {
        std::lock_guard<std::mutex> lk(s->mu);
        st.active = (int)s->active.size();
        st.queued = (int)s->queue.size();
        s->totals.admitted += st.admitted, s->totals.prefill_tokens += st.prefill_tokens, s->totals.decode_rows += st.decode_rows;
        s->totals.decode_calls += st.decode_calls, s->totals.finished += st.finished, s->totals.spilled += st.spilled;
        s->totals.active = st.active, s->totals.queued = st.queued;
    }
    if (stats) *stats = st;
    return first_error;
}

extern "C" int jl_sched_run(jl_sched *s, int max_steps, jl_sched_stats *totals) {
    if (!s) return JL_ERR_INVALID;
    int rc_all = JL_OK;
    jl_sched_stats sum = {};
    for (int i = 0; max_steps <= 0 || i < max_steps; i++) {
        jl_sched_stats st;
        const int rc = jl_sched_step(s, &st);
        if (rc != JL_OK && rc_all == JL_OK) rc_all = rc;
        sum.admitted += st.admitted, sum.prefill_tokens += st.prefill_tokens, sum.decode_rows += st.decode_rows;
        sum.decode_calls += st.decode_calls, sum.finished += st.finished, sum.spilled += st.spilled, sum.active = st.active, sum.queued = st.queued;
        if (st.active == 0 && st.queued == 0) break;
        // nothing admitted, forwarded or decoded although work is queued: every slot is held by a kept session
        if (st.admitted == 0 && st.prefill_tokens == 0 && st.decode_rows == 0 && st.finished == 0) {
            if (rc_all == JL_OK) rc_all = sched_error(s, JL_ERR_INVALID, "run: %d queued requests but no session slot can be freed", st.queued);
            break;
        }
    }
    if (totals) *totals = sum;
    return rc_all;
}
