"""The concurrent-session scheduler (csrc/jl_sched.cu) on CPU: the native policy code runs over Python backends.

1. a toy deterministic "model" that checks the protocol (KV positions appended in order, distinct sessions per step, rows per call,
   reset before reuse) while the scheduler batches requests of different lengths;
2. the CPU oracle (one OracleLlama per session slot) as the device: continuous batching must produce exactly the tokens of
   AbstractModel.generate() run request by request (core/model/AbstractModel.java:516-646).
"""
import threading

import numpy as np
import pytest

from jlama_b200 import native, synth
from jlama_b200.scheduler import SessionScheduler

VOCAB = 97


def _toy_next(history, temperature=0.0, uniform=0.0):
    h = 1469598103934665603
    for t in history:
        h = ((h ^ (int(t) + 1)) * 1099511628211) % (1 << 64)
    if temperature != 0.0:  # a "sampled" token depends on the draw it was given
        h = (h + int(round(float(uniform) * (1 << 24))) * 2654435761 + int(round(float(temperature) * 1000))) % (1 << 64)
    return int(h % VOCAB)


def _toy_generate(prompt, n_new, stop=(), context=1 << 30, start_hist=()):
    """What one request must produce: the reference's loop (first token not stop-checked, :576-608)."""
    hist = list(start_hist) + [int(t) for t in prompt]
    out = [_toy_next(hist)]
    while len(out) < n_new and len(hist) < context:
        hist.append(out[-1])
        out.append(_toy_next(hist))
        if out[-1] in stop:
            break
    return out, hist


class ToyBackend:
    """Sessions are token histories (the KV analogue); every call asserts the protocol the GPU model relies on."""

    def __init__(self, n_sessions, max_rows, fail_session=None):
        self.hist = [None] * n_sessions
        self.max_rows = max_rows
        self.calls = []
        self.fail_session = fail_session
        self.store, self.next_handle, self.fail_offload = {}, 100, False
        self.draws = []

    def reset_session(self, s):
        self.hist[s] = []
        self.calls.append(("reset", s))

    def batch_forward(self, s, tokens, start_pos):
        assert self.hist[s] is not None, "forward on a session that was never reset"
        assert start_pos == len(self.hist[s]), (start_pos, len(self.hist[s]))
        if s == self.fail_session:
            raise RuntimeError("injected failure")
        self.hist[s].extend(int(t) for t in tokens)
        self.calls.append(("forward", s, len(tokens), start_pos))

    def sample(self, s, temperature, uniform):
        self.calls.append(("sample", s))
        self.draws.append((s, float(temperature), float(uniform)))
        return _toy_next(self.hist[s], temperature, uniform)

    def decode(self, sessions, tokens, positions, temperatures, uniforms):
        assert 1 <= len(sessions) <= self.max_rows
        assert len(set(sessions.tolist())) == len(sessions), "a session twice in one step"
        out = []
        for s, t, p, T, u in zip(sessions.tolist(), tokens.tolist(), positions.tolist(), temperatures.tolist(), uniforms.tolist()):
            assert p == len(self.hist[s]), (s, p, len(self.hist[s]))
            self.hist[s].append(t)
            self.draws.append((s, T, u))
            out.append(_toy_next(self.hist[s], T, u))
        self.calls.append(("decode", tuple(sessions.tolist())))
        return out


    # host spill: the history (the KV analogue) leaves the slot and comes back through a handle
    def offload(self, s):
        if self.fail_offload:
            raise RuntimeError("injected offload failure")
        assert self.hist[s] is not None
        self.next_handle += 1
        self.store[self.next_handle] = self.hist[s]
        self.hist[s] = None
        self.calls.append(("offload", s, self.next_handle))
        return self.next_handle

    def restore(self, s, handle):
        self.hist[s] = self.store.pop(handle)
        self.calls.append(("restore", s, handle))

    def discard(self, handle):
        del self.store[handle]
        self.calls.append(("discard", handle))


def _toy_sched(n_sessions=4, max_rows=3, max_context=4096, budget=0, spill=False, **kw):
    be = ToyBackend(n_sessions, max_rows, **kw)
    extra = dict(offload=be.offload, restore=be.restore, discard=be.discard) if spill else {}
    s = SessionScheduler.over_backend(be.reset_session, be.batch_forward, be.sample, be.decode, n_sessions, max_rows, max_context, budget, **extra)
    return be, s


def test_continuous_batching_reproduces_sequential_generation_and_reuses_slots():
    rng = np.random.default_rng(7)
    be, s = _toy_sched(n_sessions=4, max_rows=3)
    reqs = []
    for i in range(23):
        prompt = rng.integers(0, VOCAB, size=int(rng.integers(1, 40)))
        n_new = int(rng.integers(1, 30))
        reqs.append((s.submit(prompt, n_new), prompt, n_new))
    max_active = 0
    steps = 0
    while True:
        st = s.step()
        steps += 1
        max_active = max(max_active, st.active)
        assert st.active <= 4 and st.decode_rows <= 4 and st.decode_calls <= 2  # 4 slots, 3 rows per call
        if st.active == 0 and st.queued == 0:
            break
        assert steps < 1000
    assert max_active == 4
    sessions_used = {}
    first_steps = []
    for rid, prompt, n_new in reqs:
        toks, state, reason = s.result(rid)
        want, _ = _toy_generate(prompt, n_new)
        assert state == native.SCHED_FINISHED and reason == native.FINISH_MAX_TOKENS
        assert toks.tolist() == want
        inf = s.info(rid)
        sessions_used.setdefault(inf.session, []).append(rid)
        first_steps.append(inf.first_token_step)
        assert inf.n_prefilled == len(prompt) and inf.n_generated == n_new
    assert first_steps == sorted(first_steps)  # FIFO admission
    assert max(len(v) for v in sessions_used.values()) >= 4  # 23 requests over 4 slots: slots are reused...
    resets = [c for c in be.calls if c[0] == "reset"]
    assert len(resets) == 23  # ...and zeroed before every reuse
    # steps with several generating requests share one decode call
    assert any(c[0] == "decode" and len(c[1]) == 3 for c in be.calls)
    q, a, f = s.counts()
    assert (q, a, f) == (0, 0, 4)
    s.close()


def test_stop_tokens_follow_the_reference_loop():
    be, s = _toy_sched()
    prompt = [3, 1, 4, 1, 5]
    free, _ = _toy_generate(prompt, 40)
    stop = free[7]
    first_hit = next(i for i in range(1, 40) if free[i] == stop)  # the token sampled from the prompt is not stop-checked
    rid = s.submit(prompt, 40, stop=[stop, 1000])
    # a stop token equal to the very first sampled token must not stop the request (AbstractModel.java:576-589)
    rid2 = s.submit(prompt, 5, stop=[free[0]])
    s.run()
    toks, state, reason = s.result(rid)
    assert toks.tolist() == free[:first_hit + 1] and reason == native.FINISH_STOP_TOKEN
    toks2, _, reason2 = s.result(rid2)
    want2, _ = _toy_generate(prompt, 5, stop=[free[0]])
    assert toks2.tolist() == want2 and toks2[0] == free[0]
    assert len(toks2) == 5 or reason2 == native.FINISH_STOP_TOKEN
    s.close()


def test_context_limit_and_argument_checks():
    be, s = _toy_sched(max_context=16)
    with pytest.raises(native.JlamaNativeError):
        s.submit(list(range(16)), 4)  # Preconditions :530 -- no room for a generated token
    with pytest.raises(native.JlamaNativeError):
        s.submit([], 4)
    with pytest.raises(native.JlamaNativeError):
        s.submit([1, 2], 0)
    rid = s.submit(list(range(10)), 100)
    s.run()
    toks, state, reason = s.result(rid)
    want, hist = _toy_generate(list(range(10)), 100, context=16)
    assert toks.tolist() == want and reason == native.FINISH_MAX_TOKENS and len(hist) == 16  # positions 0..15 written, none beyond
    assert s.info(rid).next_position == 16
    with pytest.raises(native.JlamaNativeError):
        s.result(12345)
    s.close()


def test_chunked_prefill_budget_interleaves_decode_steps():
    be, s = _toy_sched(n_sessions=2, max_rows=2, budget=8)
    short = s.submit([5, 6, 7], 12)
    s.step()  # the short request is generating before the long prompt arrives
    long_prompt = list(range(1, 31))
    long = s.submit(long_prompt, 3)
    per_step = []
    while True:
        st = s.step()
        per_step.append((st.prefill_tokens, st.decode_rows))
        if st.active == 0 and st.queued == 0:
            break
    assert all(p <= 8 for p, _ in per_step)
    assert [p for p, _ in per_step[:4]] == [8, 8, 8, 6]
    assert all(d >= 1 for _, d in per_step[:4])  # the running request keeps decoding while the long prompt is forwarded in chunks
    fw = [c for c in be.calls if c[0] == "forward" and c[2] != 3]
    assert [(c[2], c[3]) for c in fw] == [(8, 0), (8, 8), (8, 16), (6, 24)]
    assert s.result(long)[0].tolist() == _toy_generate(long_prompt, 3)[0]
    assert s.result(short)[0].tolist() == _toy_generate([5, 6, 7], 12)[0]
    s.close()


def test_cancel_queued_and_running_requests():
    be, s = _toy_sched(n_sessions=1, max_rows=1)
    a = s.submit([1, 2, 3], 50)
    b = s.submit([4, 5], 50)
    c = s.submit([6], 4)
    s.step()
    s.step()
    s.cancel(b)  # still queued
    s.cancel(a)  # running
    s.run()
    ta, sa, ra = s.result(a)
    assert ra == native.FINISH_CANCELLED and 1 <= len(ta) < 50
    assert ta.tolist() == _toy_generate([1, 2, 3], 50)[0][:len(ta)]
    tb, sb, rb = s.result(b)
    assert rb == native.FINISH_CANCELLED and len(tb) == 0 and s.info(b).session == -1
    assert s.result(c)[0].tolist() == _toy_generate([6], 4)[0]
    with pytest.raises(native.JlamaNativeError):
        s.cancel(999)
    s.close()


def test_kept_sessions_continue_at_their_context_position():
    be, s = _toy_sched(n_sessions=2, max_rows=2)
    a = s.submit([9, 8, 7], 5, keep_session=True)
    s.run()
    ta, _, _ = s.result(a)
    want_a, hist_a = _toy_generate([9, 8, 7], 5)
    assert ta.tolist() == want_a
    assert s.counts() == (0, 0, 1)  # the finished request still holds its slot
    ia = s.info(a)
    assert ia.next_position == len(hist_a) == 3 + 4
    # follow-up turn: appended to the same KV at the session's position (AbstractModel.java:533); the last sampled token of the first
    # turn was never forwarded, exactly like the reference
    b = s.submit([11, 12], 6, continue_request=a)
    with pytest.raises(native.JlamaNativeError):
        s.submit([1], 2, continue_request=a)  # one continuation per kept session
    with pytest.raises(native.JlamaNativeError):
        s.release(a)  # its continuation is queued
    other = s.submit([1, 2, 3, 4], 3)
    s.run()
    want_b, _ = _toy_generate([11, 12], 6, start_hist=hist_a)
    assert s.result(b)[0].tolist() == want_b
    ib = s.info(b)
    assert ib.session == ia.session and ib.start_pos == ia.next_position
    assert [c for c in be.calls if c[0] == "reset"].count(("reset", ia.session)) == 1  # the continuation did not zero the session
    assert s.result(other)[0].tolist() == _toy_generate([1, 2, 3, 4], 3)[0]
    assert s.counts() == (0, 0, 2)  # b did not ask to keep the session
    with pytest.raises(native.JlamaNativeError):
        s.submit([1], 2, continue_request=b)
    s.release(a)
    s.release(b)
    with pytest.raises(native.JlamaNativeError):
        s.result(a)
    s.close()


def test_run_reports_a_queue_that_can_never_be_admitted():
    be, s = _toy_sched(n_sessions=1, max_rows=1)
    a = s.submit([1, 2], 2, keep_session=True)
    s.run()
    b = s.submit([3], 2)
    with pytest.raises(native.JlamaNativeError):
        s.run()  # the only slot is held by a's kept session
    assert s.result(b)[1] == native.SCHED_QUEUED
    s.release(a)
    s.run()
    assert s.result(b)[0].tolist() == _toy_generate([3], 2)[0]
    s.close()


def test_kept_sessions_spill_to_host_when_slots_run_out_and_come_back_in_any_slot():
    be, s = _toy_sched(n_sessions=2, max_rows=2, spill=True)
    a = s.submit([1, 2, 3], 4, keep_session=True)
    s.run()
    b = s.submit([4, 5], 6, keep_session=True)
    s.run()
    assert s.counts() == (0, 0, 0) and s.info(a).session == 0 and s.info(b).session == 1
    want_a, hist_a = _toy_generate([1, 2, 3], 4)
    want_b, hist_b = _toy_generate([4, 5], 6)
    # both slots are held by idle kept sessions; a fresh request takes the slot of the one that finished first
    c = s.submit([7, 8, 9, 10], 5)
    st = s.run()
    assert st.spilled == 1 and ("offload", 0, 101) in be.calls
    ia = s.info(a)
    assert ia.spilled == 1 and ia.session == -1 and s.info(b).spilled == 0 and s.info(c).session == 0
    assert s.result(c)[0].tolist() == _toy_generate([7, 8, 9, 10], 5)[0]
    assert be.calls.index(("offload", 0, 101)) < be.calls.index(("reset", 0), 2)  # saved before the slot is zeroed for its new user
    # a's follow-up turn: restored into the slot c left, at a's context position, history intact
    a2 = s.submit([20, 21], 3, continue_request=a)
    s.run()
    assert ("restore", 0, 101) in be.calls and not be.store
    assert s.result(a2)[0].tolist() == _toy_generate([20, 21], 3, start_hist=hist_a)[0]
    assert s.info(a2).start_pos == len(hist_a) and s.info(a).spilled == 0
    # b's follow-up goes back to b's own slot without any copy
    n_calls = len(be.calls)
    b2 = s.submit([30], 2, continue_request=b)
    s.run()
    assert s.result(b2)[0].tolist() == _toy_generate([30], 2, start_hist=hist_b)[0] and s.info(b2).session == 1
    assert not any(c2[0] in ("offload", "restore", "reset") for c2 in be.calls[n_calls:])
    assert s.result(a)[0].tolist() == want_a and s.result(b)[0].tolist() == want_b
    s.close()


def test_spill_prefers_sessions_without_a_queued_follow_up_and_release_discards_the_host_copy():
    be, s = _toy_sched(n_sessions=2, max_rows=2, spill=True)
    a = s.submit([1], 2, keep_session=True)
    s.run()
    b = s.submit([2], 2, keep_session=True)
    s.run()
    # a fresh request, then a's follow-up: a finished first, but its continuation is queued, so b is the one to go
    fresh = [s.submit([5], 2)]
    a2 = s.submit([9], 2, continue_request=a)
    s.step()
    assert s.info(b).spilled == 1 and s.info(a).spilled == 0 and s.info(a2).session == 0 and s.info(fresh[0]).session == 1
    fresh += [s.submit([6], 2), s.submit([7], 2)]
    s.run()
    for i, f in enumerate(fresh):
        assert s.result(f)[0].tolist() == _toy_generate([5 + i], 2)[0]
    assert s.result(a2)[0].tolist() == _toy_generate([9], 2, start_hist=_toy_generate([1], 2)[1])[0]
    # b is never continued: releasing it drops the host copy at the next step
    assert len(be.store) == 1
    s.release(b)
    s.step()
    assert not be.store and be.calls[-1][0] == "discard"
    s.close()


def test_a_failed_offload_drops_the_kept_session_but_not_the_new_request():
    be, s = _toy_sched(n_sessions=1, max_rows=1, spill=True)
    a = s.submit([1, 2], 3, keep_session=True)
    s.run()
    be.fail_offload = True
    b = s.submit([3, 4], 3)
    with pytest.raises(native.JlamaNativeError, match="offload"):
        s.run()
    s.run()
    assert s.result(b)[0].tolist() == _toy_generate([3, 4], 3)[0]  # the slot was reused regardless
    with pytest.raises(native.JlamaNativeError):
        s.submit([5], 2, continue_request=a)  # a's KV is gone, it cannot be continued
    assert s.info(a).spilled == 0 and s.info(a).session == -1
    s.close()


def test_request_timings_split_queueing_prompt_and_generation():
    import time
    be, s = _toy_sched(n_sessions=1, max_rows=1)
    a = s.submit([1, 2, 3], 4)
    b = s.submit([4], 2)
    c = s.submit([5], 2)
    s.cancel(c)
    time.sleep(0.02)
    assert s.info(a).queue_ms >= 15 and s.info(a).prompt_ms == 0 and s.info(a).generate_ms == 0  # still queued: measured up to now
    s.step()
    time.sleep(0.02)
    ia = s.info(a)
    assert ia.queue_ms >= 15 and ia.prompt_ms >= 0 and ia.generate_ms >= 15  # first token is out, generation is running
    s.run()
    ia, ib, ic = s.info(a), s.info(b), s.info(c)
    frozen = (ia.queue_ms, ia.prompt_ms, ia.generate_ms)
    time.sleep(0.01)
    ia = s.info(a)
    assert (ia.queue_ms, ia.prompt_ms, ia.generate_ms) == frozen  # finished: the clocks stand
    assert ib.queue_ms >= ia.queue_ms + ia.prompt_ms  # b waited for a's slot
    assert ic.prompt_ms == 0 and ic.generate_ms == 0 and ic.queue_ms > 0  # cancelled in the queue
    s.close()


def test_closing_the_scheduler_drops_the_host_copies_it_still_holds():
    be, s = _toy_sched(n_sessions=1, max_rows=1, spill=True)
    a = s.submit([1, 2], 2, keep_session=True)
    s.run()
    s.submit([3], 2, keep_session=True)
    s.run()
    assert len(be.store) == 1 and s.info(a).spilled == 1
    s.close()
    assert not be.store and be.calls[-1][0] == "discard"


def test_a_failing_backend_call_fails_only_its_request():
    be, s = _toy_sched(n_sessions=3, max_rows=3, fail_session=1)
    ids = [s.submit([i + 1, i + 2], 4) for i in range(5)]
    st = s.run(check=False)
    states = [s.result(r)[1] for r in ids]
    reasons = [s.result(r)[2] for r in ids]
    failed = [i for i, x in enumerate(states) if x == native.SCHED_FAILED]
    assert failed and all(s.info(ids[i]).session == 1 for i in failed) and all(reasons[i] == native.FINISH_ERROR for i in failed)
    for i, r in enumerate(ids):
        if i not in failed:
            assert s.result(r)[0].tolist() == _toy_generate([i + 1, i + 2], 4)[0]
    assert len(failed) + sum(x == native.SCHED_FINISHED for x in states) == 5
    assert s.backend_errors and "injected" in str(s.backend_errors[0])
    assert s.counts() == (0, 0, 3)  # failed requests do not leak their slot
    s.close()


def test_submit_from_other_threads_while_stepping():
    be, s = _toy_sched(n_sessions=4, max_rows=4)
    rng = np.random.default_rng(3)
    work = [(rng.integers(0, VOCAB, size=int(rng.integers(1, 20))), int(rng.integers(1, 20))) for _ in range(40)]
    ids = [None] * len(work)

    def producer(lo, hi):
        for i in range(lo, hi):
            ids[i] = s.submit(work[i][0], work[i][1])

    threads = [threading.Thread(target=producer, args=(i * 10, i * 10 + 10)) for i in range(4)]
    for t in threads:
        t.start()
    done = False
    while not done:
        st = s.step()
        alive = any(t.is_alive() for t in threads)
        done = not alive and st.active == 0 and st.queued == 0
    for t in threads:
        t.join()
    s.run()
    for rid, (prompt, n_new) in zip(ids, work):
        toks, state, _ = s.result(rid)
        assert state == native.SCHED_FINISHED and toks.tolist() == _toy_generate(prompt, n_new)[0]
    s.close()


def test_oracle_as_the_device_matches_generate_request_by_request(oracle):
    """The same native scheduler, with the CPU restatement of the model in place of the GPU: one OracleLlama per session slot.
    Every request must come out token for token as AbstractModel.generate() produces it on its own."""
    cfg = synth.get_config("tiny")
    w = synth.make_weights(cfg)
    n_slots = 3
    slots = [oracle.OracleLlama(cfg, w, act_q8=True) for _ in range(n_slots)]
    hidden = [None] * n_slots

    def reset(s):
        slots[s].reset()

    def forward(s, tokens, start_pos):
        hidden[s] = slots[s].batch_forward(tokens, start_pos)

    def pick(s, hid, T, u):
        tok, logits = slots[s].sample(hid)
        return tok if T == 0.0 else oracle.sample_temperature(logits, T, u)

    def sample(s, T, u):
        return pick(s, hidden[s], T, u)

    def decode(sessions, tokens, positions, temperatures, uniforms):
        return [pick(s, slots[s].batch_forward([t], p), T, u)
                for s, t, p, T, u in zip(sessions.tolist(), tokens.tolist(), positions.tolist(), temperatures.tolist(), uniforms.tolist())]

    sched = SessionScheduler.over_backend(reset, forward, sample, decode, n_slots, 2, cfg["ctx"], prefill_tokens_per_step=16)
    reqs = []
    for i in range(7):
        prompt = synth.random_prompt(cfg, 4 + 5 * i, seed=300 + i)
        n_new = 3 + (i * 2) % 7
        reqs.append((sched.submit(prompt, n_new), prompt, n_new))
    sched.run()
    ref = oracle.OracleLlama(cfg, w, act_q8=True)
    for rid, prompt, n_new in reqs:
        toks, state, reason = sched.result(rid)
        want, _ = ref.generate(prompt, n_new)
        assert state == native.SCHED_FINISHED and toks.tolist() == list(want), (rid, toks.tolist(), list(want))
    sched.close()
    ref.close()
    for o in slots:
        o.close()


def test_sampled_requests_draw_a_reproducible_uniform_stream_independent_of_batching():
    """temperature != 0: the k-th token of a request is drawn with the k-th value of its own seeded stream (the reference draws
    ThreadLocalRandom.nextFloat() per sample() call, AbstractModel.java:576,594), so the tokens do not depend on how many other
    requests shared its decode steps; greedy requests in the same steps get temperature 0 and are unaffected."""
    def run(n_sessions, max_rows, budget):
        be, s = _toy_sched(n_sessions=n_sessions, max_rows=max_rows, budget=budget)
        ids = [s.submit([3, 4, 5], 9, temperature=0.8, seed=42),
               s.submit([6, 7], 7),
               s.submit([3, 4, 5], 9, temperature=0.8, seed=43),
               s.submit([8, 9, 10, 11, 12], 6, temperature=1.5, seed=42),
               s.submit([3, 4, 5], 9, temperature=0.8, seed=42)]
        s.run()
        out = [s.result(r)[0].tolist() for r in ids]
        draws = be.draws
        s.close()
        return out, draws
    a, draws = run(5, 5, 0)
    b, _ = run(2, 1, 3)   # different slot count, one row per decode call, chunked prefill
    assert a == b
    assert a[0] == a[4] and a[0] != a[2]            # same seed and prompt reproduce; another seed differs
    assert a[1] == _toy_generate([6, 7], 7)[0]     # the greedy request is what it would be alone
    us = [u for (_, T, u) in draws if T != 0.0]
    assert len(us) == 9 + 9 + 6 + 9 and all(0.0 <= u < 1.0 for u in us) and len(set(us)) == 18  # two distinct seeds, at most 9 draws each
    assert all(u == 0.0 for (_, T, u) in draws if T == 0.0)
    # the stream itself: splitmix64 from the seed, top 24 bits / 2^24 (tests/test_gpu_scheduler.py replays it for generate_sample)
    x, M, want = 42, (1 << 64) - 1, []
    for _ in range(9):
        x = (x + 0x9E3779B97F4A7C15) & M
        z = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        want.append(float(np.float32((z ^ (z >> 31)) >> 40) * np.float32(1.0 / 16777216.0)))
    first = [u for (s_, T, u) in draws if T != 0.0 and s_ == 0][:9]  # request 0 sits in slot 0 when every request has its own slot
    assert first == want
    be, s = _toy_sched()
    with pytest.raises(native.JlamaNativeError):
        s.submit([1], 2, temperature=-0.5)
    with pytest.raises(native.JlamaNativeError):
        s.submit([1], 2, temperature=float("nan"))
    s.close()


def test_oracle_temperature_rule_matches_a_numpy_restatement(oracle):
    """AbstractModel.sample :475-489 restated twice (C in the oracle, numpy here): float exponentials of (logit - max) / T computed in
    double, float running sums in index order, first index whose cumulative probability reaches the uniform sample."""
    rng = np.random.default_rng(5)
    for V, T in ((17, 0.7), (512, 1.0), (4096, 0.3)):
        logits = (rng.standard_normal(V) * 3).astype(np.float32)
        e = np.exp((logits.astype(np.float64) - float(logits.max())) / T).astype(np.float32)
        total = np.float32(0)
        for v in e:
            total = np.float32(total + v)
        p = (e / total).astype(np.float32)
        cdf = np.zeros(V, dtype=np.float32)
        acc = np.float32(0)
        for i, v in enumerate(p):
            acc = np.float32(acc + v)
            cdf[i] = acc
        for u in (0.0, 1e-6, 0.2, 0.5, 0.77, 0.999, 1.0):
            hit = np.nonzero(cdf >= np.float32(u))[0]
            want = int(hit[0]) if len(hit) else V - 1
            assert oracle.sample_temperature(logits, T, u) == want, (V, T, u)
    # u above the final cumulative sum (float rounding can leave it below 1): the last index
    assert oracle.sample_temperature(np.zeros(8, np.float32), 1.0, 1.5) == 7


@pytest.mark.parametrize("seed", range(6))
def test_randomised_serving_scenarios_keep_every_request_on_its_own_trajectory(seed):
    """Model-based stress test: random submissions (fresh, kept, follow-up turns, sampled), cancellations and releases interleaved with
    steps, over few slots with host spill.  The toy backend asserts the device protocol on every call; here every request that finished on
    its own must hold exactly the tokens its history chain dictates, whatever was batched, spilled or restored around it."""
    rng = np.random.default_rng(1000 + seed)
    n_slots = int(rng.integers(1, 4))
    be, s = _toy_sched(n_sessions=n_slots, max_rows=int(rng.integers(1, 4)), budget=int(rng.integers(0, 12)), spill=True, max_context=96)
    meta = {}      # id -> dict(prompt, max_new, stop, T, seed, parent, keep)
    open_kept = []  # finished kept requests that may still be continued
    live = []
    for _ in range(int(rng.integers(40, 90))):
        op = rng.random()
        if op < 0.45:
            parent = -1
            if open_kept and rng.random() < 0.5:
                parent = open_kept.pop(int(rng.integers(0, len(open_kept))))
            prompt = rng.integers(0, VOCAB, size=int(rng.integers(1, 9))).tolist()
            m = dict(prompt=prompt, max_new=int(rng.integers(1, 8)), stop=[int(rng.integers(0, VOCAB))] if rng.random() < 0.3 else [],
                     T=float(rng.choice([0.0, 0.0, 0.7])), seed=int(rng.integers(0, 1 << 30)), parent=parent, keep=bool(rng.random() < 0.5))
            try:
                rid = s.submit(prompt, m["max_new"], stop=m["stop"], keep_session=m["keep"], continue_request=parent, temperature=m["T"], seed=m["seed"])
            except native.JlamaNativeError:
                continue  # context exhausted for this chain, or the parent's KV was dropped
            meta[rid] = m
            live.append(rid)
        elif op < 0.55 and live:
            s.cancel(live[int(rng.integers(0, len(live)))])
        elif op < 0.62 and open_kept:
            s.release(open_kept.pop(int(rng.integers(0, len(open_kept)))))
        else:
            st = s.step(check=False)
            assert st.active <= n_slots
        for rid in list(live):
            _, state, reason = s.result(rid)
            if state in (native.SCHED_FINISHED, native.SCHED_FAILED):
                live.remove(rid)
                if state == native.SCHED_FINISHED and reason != native.FINISH_CANCELLED and meta[rid]["keep"]:
                    open_kept.append(rid)
    s.run(check=False)
    assert not be.__dict__.get("errors") and not s.backend_errors, s.backend_errors

    def history_before(rid):
        """tokens in the KV when request rid starts: its parent's chain, prompt and forwarded tokens"""
        m = meta[rid]
        if m["parent"] < 0:
            return []
        p = m["parent"]
        toks = s_tokens[p]
        return history_before(p) + meta[p]["prompt"] + toks[:len(toks) - 1]

    s_tokens, s_state = {}, {}
    for rid in meta:
        try:
            t, state, reason = s.result(rid)
        except native.JlamaNativeError:
            continue  # released
        s_tokens[rid], s_state[rid] = t.tolist(), (state, reason)
    checked = 0
    for rid, m in meta.items():
        if rid not in s_tokens or s_state[rid][0] != native.SCHED_FINISHED or s_state[rid][1] == native.FINISH_CANCELLED:
            continue
        chain_ok, q = True, m["parent"]
        while q >= 0:
            chain_ok = chain_ok and q in s_tokens
            q = meta[q]["parent"]
        if not chain_ok:
            continue  # an ancestor was released: its tokens are no longer readable (the KV chain itself was still intact)
        hist = history_before(rid) + m["prompt"]
        # replay the request alone: k-th token with the k-th draw of its stream
        x, M, out = m["seed"], (1 << 64) - 1, []
        while True:
            u = 0.0
            if m["T"] != 0.0:
                x = (x + 0x9E3779B97F4A7C15) & M
                z = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
                z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
                u = float(np.float32((z ^ (z >> 31)) >> 40) * np.float32(1.0 / 16777216.0))
            out.append(_toy_next(hist, np.float32(m["T"]), u))
            if (len(out) > 1 and out[-1] in m["stop"]) or len(out) >= m["max_new"] or len(hist) >= 96:
                break
            hist = hist + [out[-1]]
        assert s_tokens[rid] == out, (rid, m, s_tokens[rid], out)
        checked += 1
    assert checked >= 5
    # everything released: no slot, no host copy may be left behind
    for rid in list(meta):
        try:
            s.release(rid)
        except native.JlamaNativeError:
            pass
    for rid in list(meta):  # parents whose continuation was still pending at the first pass
        try:
            s.release(rid)
        except native.JlamaNativeError:
            pass
    s.step()
    assert s.counts() == (0, 0, n_slots) and not be.store
    s.close()
