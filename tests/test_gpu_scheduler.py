"""Concurrent sessions on the GPU: the request scheduler (csrc/jl_sched.cu) over the batched decode step, and generate() called from
several threads (the reference's concurrency model: one thread per request, KvBufferCache.java:58-60, OpenAIChatService.java:107-160).
Every request must come out as AbstractModel.generate() produces it on its own -- checked against the CPU oracle, request by request.
tests/test_scheduler.py runs the same policy code on CPU."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _check_against_oracle(om, results, near_tie=1e-2):
    """token for token; a different token is only acceptable where the oracle's own top two logits are a near tie"""
    exact = 0
    for prompt, n_new, toks in results:
        ot, ol = om.generate(prompt, n_new)
        assert len(toks) == len(ot)
        for i in range(len(ot)):
            if toks[i] != ot[i]:
                top = np.sort(ol[i])[-2:]
                assert top[1] - top[0] <= near_tie * np.abs(ol[i]).max(), (i, int(toks[i]), int(ot[i]), float(top[1] - top[0]))
                break
        else:
            exact += 1
    assert exact >= len(results) - 1  # near ties are rare
    return exact


@pytest.mark.parametrize("name,act_q8,slots,budget", [("small", True, 4, 24), ("tiny", True, 3, 0), ("small", True, 1, 0)])
def test_scheduler_batches_requests_like_sequential_generate(cuda_ctx, oracle, name, act_q8, slots, budget):
    from jlama_b200 import native, synth
    from jlama_b200.model import LlamaModel
    from jlama_b200.scheduler import SessionScheduler
    cfg = synth.get_config(name)
    w = synth.make_weights(cfg)
    gm = LlamaModel(cuda_ctx, cfg, w, working_qtype=native.I8 if act_q8 else native.F32, max_sessions=slots)
    om = oracle.OracleLlama(cfg, w, act_q8=act_q8)
    n_req = 9 if slots > 1 else 3
    work = [(synth.random_prompt(cfg, 5 + 4 * i, seed=500 + i), 3 + (5 * i) % 12) for i in range(n_req)]
    with SessionScheduler(gm, prefill_tokens_per_step=budget) as sched:
        ids = [sched.submit(p, n) for p, n in work]
        max_active, rows_seen = 0, set()
        while True:
            st = sched.step()
            max_active = max(max_active, st.active)
            rows_seen.add(st.decode_rows)
            assert st.active <= slots and st.decode_rows <= slots
            if budget:
                assert st.prefill_tokens <= budget
            if st.active == 0 and st.queued == 0:
                break
        assert max_active == slots
        if slots > 1:
            assert max(rows_seen) > 1  # requests really shared decode steps
        results = []
        sessions = set()
        for rid, (p, n) in zip(ids, work):
            toks, state, reason = sched.result(rid)
            assert state == native.SCHED_FINISHED and reason == native.FINISH_MAX_TOKENS and len(toks) == n
            sessions.add(sched.info(rid).session)
            results.append((p, n, toks))
        assert len(sessions) == slots  # all slots used, and reused: n_req > slots
        assert sched.counts() == (0, 0, slots)
    _check_against_oracle(om, results)
    gm.close()
    om.close()


def test_scheduler_stop_tokens_and_kept_sessions(cuda_ctx, oracle):
    from jlama_b200 import native, synth
    from jlama_b200.model import LlamaModel
    from jlama_b200.scheduler import SessionScheduler
    cfg = synth.get_config("small")
    w = synth.make_weights(cfg)
    gm = LlamaModel(cuda_ctx, cfg, w, max_sessions=2)
    om = oracle.OracleLlama(cfg, w, act_q8=True)
    p1 = synth.random_prompt(cfg, 11, seed=41)
    with SessionScheduler(gm) as sched:
        b = sched.submit(p1, 10, keep_session=True)
        sched.run()
        tb, _, rb = sched.result(b)
        assert len(tb) == 10 and rb == native.FINISH_MAX_TOKENS
        # the same request again, alone like the first (same kernels, same inputs: bit-identical), now with a stop token
        stop = int(tb[4])
        hit = next(i for i in range(1, 10) if tb[i] == stop)  # the token sampled from the prompt is not stop-checked (:576-589)
        a = sched.submit(p1, 10, stop=[stop])
        sched.run()
        ta, _, ra = sched.result(a)
        assert ta.tolist() == tb[:hit + 1].tolist() and ra == native.FINISH_STOP_TOKEN
        # a second turn on the kept session appends to its KV (AbstractModel.java:533): the same tokens as one generate() over
        # prompt + the first turn's forwarded tokens + the new prompt
        p2 = synth.random_prompt(cfg, 5, seed=42)
        c = sched.submit(p2, 5, continue_request=b)
        sched.run()
        tc, _, _ = sched.result(c)
        assert sched.info(c).start_pos == len(p1) + 9 and sched.info(c).session == 0
        joined = np.concatenate([p1, tb[:9], p2]).astype(np.int32)
    # the oracle forwards the joined prompt in one batch; the GPU did it in two turns with decode steps in between: same KV, same tokens
    _check_against_oracle(om, [(p1, 10, tb), (joined, 5, tc)])
    gm.close()
    om.close()


def test_generate_from_two_threads_on_two_sessions(cuda_ctx, oracle):
    """The reference's concurrency model: request threads call generate() on one model, each with its own session; the entry points
    of a jl_model serialise on the model's lock (they share the stream and the staging buffers)."""
    from jlama_b200 import synth
    from jlama_b200.model import LlamaModel
    cfg = synth.get_config("small")
    w = synth.make_weights(cfg)
    gm = LlamaModel(cuda_ctx, cfg, w, max_sessions=2)
    om = oracle.OracleLlama(cfg, w, act_q8=True)
    prompts = [synth.random_prompt(cfg, 13, seed=61), synth.random_prompt(cfg, 21, seed=62)]
    outs, errs = [None, None], []

    def worker(i):
        try:
            for _ in range(3):
                outs[i] = gm.generate(prompts[i], 12, session=i)[0]
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    _check_against_oracle(om, [(prompts[i], 12, outs[i]) for i in range(2)])
    gm.close()
    om.close()


def test_decode_rejects_a_session_twice_in_one_step(cuda_ctx):
    from jlama_b200 import native, synth
    from jlama_b200.model import LlamaModel
    cfg = synth.get_config("tiny")
    gm = LlamaModel(cuda_ctx, cfg, synth.make_weights(cfg), max_sessions=2)
    gm.batch_forward(synth.random_prompt(cfg, 4), 0, session=0)
    with pytest.raises(native.JlamaNativeError):
        gm.decode([1, 2], [4, 4], sessions=[0, 0])
    gm.close()


def test_kv_offload_frees_the_slot_and_restore_resumes_in_another_one(cuda_ctx, oracle):
    """Host spill: the pages of an idle session leave HBM, the slot serves another request, and the session resumes token
    for token after its pages come back into a different slot."""
    from jlama_b200 import native, synth
    from jlama_b200.model import LlamaModel
    cfg = synth.get_config("small")
    w = synth.make_weights(cfg)
    gm = LlamaModel(cuda_ctx, cfg, w, max_sessions=2)
    om = oracle.OracleLlama(cfg, w, act_q8=True)
    prompt = synth.random_prompt(cfg, 23)
    gm.reset_session(0)
    gm.batch_forward(prompt, 0, session=0)
    toks = [gm.sample(session=0, want_logits=False)[0]]
    for i in range(4):
        toks.append(int(gm.decode([toks[-1]], [len(prompt) + i], sessions=[0])[0][0]))
    pages = gm.kv_pages(0)
    assert pages >= 1
    k_before = gm.read_kv(1, 5, 0, session=0)
    h = gm.kv_offload(0)
    assert h > 0 and gm.kv_pages(0) == 0
    with pytest.raises(native.JlamaNativeError):
        gm.read_kv(1, 5, 0, session=0)  # the page is gone from the device
    # the emptied slot serves somebody else in the meantime
    other = synth.random_prompt(cfg, 9, seed=77)
    got_other = gm.generate(other, 6, session=0)[0]
    gm.kv_restore(h, session=1)
    assert gm.kv_pages(1) == pages and np.array_equal(gm.read_kv(1, 5, 0, session=1), k_before)
    with pytest.raises(native.JlamaNativeError):
        gm.kv_restore(h, session=1)  # a handle restores once
    for i in range(4, 9):
        toks.append(int(gm.decode([toks[-1]], [len(prompt) + i], sessions=[1])[0][0]))
    _check_against_oracle(om, [(prompt, 10, np.array(toks)), (other, 6, got_other)])
    h2 = gm.kv_offload(1)
    gm.kv_discard(h2)
    with pytest.raises(native.JlamaNativeError):
        gm.kv_discard(h2)
    gm.close()
    om.close()


def test_scheduler_spills_an_idle_kept_session_and_restores_it_for_its_follow_up(cuda_ctx, oracle):
    from jlama_b200 import synth
    from jlama_b200.model import LlamaModel
    from jlama_b200.scheduler import SessionScheduler
    cfg = synth.get_config("small")
    w = synth.make_weights(cfg)
    gm = LlamaModel(cuda_ctx, cfg, w, max_sessions=1)
    om = oracle.OracleLlama(cfg, w, act_q8=True)
    p1, p2, p3 = (synth.random_prompt(cfg, n, seed=s) for n, s in ((12, 81), (7, 82), (4, 83)))
    with SessionScheduler(gm) as sched:
        a = sched.submit(p1, 6, keep_session=True)
        sched.run()
        b = sched.submit(p2, 5)  # the only slot is held by a's idle session: a goes to host memory
        st = sched.run()
        assert st.spilled == 1 and sched.info(a).spilled == 1
        a2 = sched.submit(p3, 5, continue_request=a)
        sched.run()
        ta, tb, ta2 = sched.result(a)[0], sched.result(b)[0], sched.result(a2)[0]
        assert sched.info(a2).start_pos == len(p1) + 5 and sched.info(a).spilled == 0
    joined = np.concatenate([p1, ta[:5], p3]).astype(np.int32)
    _check_against_oracle(om, [(p1, 6, ta), (p2, 5, tb), (joined, 5, ta2)])
    gm.close()
    om.close()


# ---- temperature sampling in decode steps ----
def _splitmix_uniforms(seed, n):
    """the scheduler's per-request stream (csrc/jl_sched.cu next_uniform): splitmix64, top 24 bits / 2^24"""
    out, x, M = [], seed, (1 << 64) - 1
    for _ in range(n):
        x = (x + 0x9E3779B97F4A7C15) & M
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        z ^= z >> 31
        out.append(np.float32(z >> 40) * np.float32(1.0 / 16777216.0))
    return np.array(out, dtype=np.float32)


def test_decode_sample_draws_rows_with_the_reference_rule(cuda_ctx, oracle):
    """jl_model_decode_sample: rows with a temperature are drawn by AbstractModel.sample's rule (:475-489) from their own logits row,
    rows at temperature 0 keep the arg-max; the logits handed back are the raw ones."""
    from jlama_b200 import synth
    from jlama_b200.model import LlamaModel
    cfg = synth.get_config("small")
    gm = LlamaModel(cuda_ctx, cfg, synth.make_weights(cfg), max_sessions=3)
    firsts = []
    prompts = [synth.random_prompt(cfg, 6 + 2 * s, seed=900 + s) for s in range(3)]
    for s, p in enumerate(prompts):
        gm.reset_session(s)
        gm.batch_forward(p, 0, session=s)
        firsts.append(gm.sample(session=s, want_logits=False)[0])
    pos = [len(p) for p in prompts]
    greedy, lg = gm.decode(firsts, pos, want_logits=True)  # re-decoding a position rewrites the same KV row: idempotent
    T = np.array([0.7, 0.0, 1.3], dtype=np.float32)
    for us in ((0.0, 0.5, 0.25), (0.9, 0.1, 0.6), (0.4, 0.4, 0.999)):
        toks, lg2 = gm.decode(firsts, pos, want_logits=True, temperatures=T, uniforms=np.array(us, dtype=np.float32))
        assert np.abs(lg2 - lg).max() <= 1e-5 * np.abs(lg).max()  # the raw logits, copied out before the rows are exponentiated in place
        assert toks[1] == greedy[1]
        for i in (0, 2):
            expect = oracle.sample_temperature(lg[i], float(T[i]), us[i])
            e = np.exp((lg[i].astype(np.float64) - lg[i].max()) / float(T[i])).astype(np.float32)
            cdf = np.cumsum((e / e.sum(dtype=np.float32)).astype(np.float32), dtype=np.float32)
            # the device sums the exponentials in another order than the reference's sequential loop: a draw that lands within
            # rounding of a cdf step may pick the neighbour
            assert toks[i] == expect or abs(int(toks[i]) - expect) <= 1 or abs(cdf[toks[i]] - us[i]) < 1e-4, (i, us, int(toks[i]), expect)
    # a single sampled row goes through the persistent kernel's logits
    t1, _ = gm.decode(firsts[:1], pos[:1], sessions=[0], temperatures=[0.7], uniforms=[0.5])
    e1 = oracle.sample_temperature(lg[0], 0.7, 0.5)
    assert abs(int(t1[0]) - e1) <= 1
    gm.close()


def test_scheduler_sampled_request_equals_generate_sample_with_the_same_stream(cuda_ctx):
    from jlama_b200 import synth
    from jlama_b200.model import LlamaModel
    from jlama_b200.scheduler import SessionScheduler
    cfg = synth.get_config("small")
    gm = LlamaModel(cuda_ctx, cfg, synth.make_weights(cfg), max_sessions=1)
    prompt = synth.random_prompt(cfg, 10, seed=950)
    with SessionScheduler(gm) as sched:
        a = sched.submit(prompt, 12, temperature=0.9, seed=1234)
        sched.run()
        b = sched.submit(prompt, 12, temperature=0.9, seed=1234)
        sched.run()
        c = sched.submit(prompt, 12)
        sched.run()
        ta, tb, tc = sched.result(a)[0], sched.result(b)[0], sched.result(c)[0]
    assert ta.tolist() == tb.tolist()  # (seed, prompt) reproduces
    # one slot: every step is a one-row decode, the same kernels generate_sample runs -> bit-identical logits, identical draws
    assert gm.generate_sample(prompt, 12, 0.9, _splitmix_uniforms(1234, 12)).tolist() == ta.tolist()
    assert gm.generate(prompt, 12)[0].tolist() == tc.tolist()
    assert ta.tolist() != tc.tolist()  # twelve draws at T = 0.9 do not all land on the arg-max
    gm.close()
