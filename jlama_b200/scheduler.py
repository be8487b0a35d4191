"""Host-side mirror of the concurrent-session scheduler (csrc/jl_sched.cu, C ABI jl_sched_*).

The reference serves concurrent requests with one thread per request, each running AbstractModel.generate() on its own KvBuffer
(core/tensor/KvBufferCache.java:58-60, jlama-net/.../openai/OpenAIChatService.java:107-160).  Here requests queue in front of the
batched decode step and are batched per iteration:

    sched = SessionScheduler(model)                       # all of the model's session slots
    a = sched.submit(prompt_a, max_new=64, stop=[eos])
    b = sched.submit(prompt_b, max_new=16)
    sched.run()                                           # or step() from a serving loop while other threads submit
    tokens, state, reason = sched.result(a)

`SessionScheduler.over_backend` runs the same native policy code over Python callables (reset_session / batch_forward / sample /
decode); tests/test_scheduler.py uses it with the CPU oracle standing in for the device.
"""
import ctypes as C

import numpy as np

from . import native


class SessionScheduler:
    def __init__(self, model, max_active=0, prefill_tokens_per_step=0):
        self.lib = model.lib
        self._model = model  # the scheduler borrows the model: keep it alive
        self._keep = None
        h = C.c_void_p()
        rc = self.lib.jl_sched_create(model.h, max_active, prefill_tokens_per_step, C.byref(h))
        if rc != native.JL_OK:
            raise native.JlamaNativeError(rc, "jl_sched_create failed (is the model finalized?)")
        self.h = h

    @classmethod
    def over_backend(cls, reset_session, batch_forward, sample, decode, n_sessions, max_rows, max_context, prefill_tokens_per_step=0,
                     offload=None, restore=None, discard=None):
        """reset_session(session); batch_forward(session, tokens: np.int32[n], start_pos); sample(session, temperature, uniform) -> token;
        decode(sessions, tokens, positions: np.int32[n], temperatures, uniforms: np.float32[n]) -> next tokens; optionally the host-spill trio offload(session) -> handle,
        restore(session, handle), discard(handle).  A raised exception fails the affected requests (JL_ERR_INVALID), exactly like a
        backend error code."""
        self = cls.__new__(cls)
        self.lib = native.load()
        self._model = None
        self.backend_errors = []

        def guard(fn):
            def wrapped(*a):
                try:
                    return fn(*a)
                except Exception as e:  # noqa: BLE001 -- a Python exception must not unwind through the C caller
                    self.backend_errors.append(e)
                    return native.JL_ERR_INVALID
            return wrapped

        @guard
        def _reset(_u, session):
            reset_session(session)
            return native.JL_OK

        @guard
        def _forward(_u, session, tokens, n, start_pos):
            batch_forward(session, np.ctypeslib.as_array(tokens, shape=(n,)).copy(), start_pos)
            return native.JL_OK

        @guard
        def _sample(_u, session, temperature, uniform, out):
            out[0] = int(sample(session, temperature, uniform))
            return native.JL_OK

        @guard
        def _decode(_u, n, sessions, tokens, positions, temperatures, uniforms, nxt):
            arr = lambda p: np.ctypeslib.as_array(p, shape=(n,)).copy()  # noqa: E731
            res = decode(arr(sessions), arr(tokens), arr(positions), arr(temperatures), arr(uniforms))
            for i in range(n):
                nxt[i] = int(res[i])
            return native.JL_OK

        @guard
        def _offload(_u, session, handle_out):
            handle_out[0] = int(offload(session))
            return native.JL_OK

        @guard
        def _restore(_u, session, handle):
            restore(session, handle)
            return native.JL_OK

        @guard
        def _discard(_u, handle):
            discard(handle)
            return native.JL_OK

        be = native.SchedBackend(native.SCHED_RESET_FN(_reset), native.SCHED_FORWARD_FN(_forward), native.SCHED_SAMPLE_FN(_sample),
                                 native.SCHED_DECODE_FN(_decode))
        if offload is not None:
            be.offload, be.restore, be.discard = native.SCHED_OFFLOAD_FN(_offload), native.SCHED_RESTORE_FN(_restore), native.SCHED_DISCARD_FN(_discard)
        self._keep = be  # the C side copies the struct, the CFUNCTYPE objects must outlive it
        h = C.c_void_p()
        rc = self.lib.jl_sched_create_backend(C.byref(be), None, n_sessions, max_rows, max_context, prefill_tokens_per_step, C.byref(h))
        if rc != native.JL_OK:
            raise native.JlamaNativeError(rc, "jl_sched_create_backend: bad arguments")
        self.h = h
        return self

    def _check(self, rc):
        if rc != native.JL_OK:
            raise native.JlamaNativeError(rc, self.lib.jl_sched_last_error(self.h).decode())

    def submit(self, prompt, max_new, stop=(), keep_session=False, continue_request=-1, temperature=0.0, seed=0):
        """temperature 0 = arg-max; otherwise tokens are drawn with the reference's rule from a uniform stream seeded with `seed`."""
        prompt = np.ascontiguousarray(prompt, dtype=np.int32)
        stop = np.ascontiguousarray(list(stop), dtype=np.int32)
        rid = self.lib.jl_sched_submit(self.h, native.ptr(prompt), len(prompt), max_new, native.ptr(stop) if len(stop) else None,
                                       len(stop), native.SCHED_KEEP_SESSION if keep_session else 0, continue_request,
                                       C.c_float(temperature), C.c_uint64(seed))
        if rid < 0:
            raise native.JlamaNativeError(native.JL_ERR_INVALID, self.lib.jl_sched_last_error(self.h).decode())
        return rid

    def cancel(self, request):
        self._check(self.lib.jl_sched_cancel(self.h, request))

    def step(self, check=True):
        st = native.SchedStats()
        rc = self.lib.jl_sched_step(self.h, C.byref(st))
        if check:
            self._check(rc)
        return st

    def run(self, max_steps=0, check=True):
        st = native.SchedStats()
        rc = self.lib.jl_sched_run(self.h, max_steps, C.byref(st))
        if check:
            self._check(rc)
        return st

    def result(self, request):
        """(tokens generated so far, state, finish reason)"""
        n, state, reason = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.jl_sched_result(self.h, request, None, 0, C.byref(n), C.byref(state), C.byref(reason)))
        out = np.empty(n.value, dtype=np.int32)
        if n.value:
            self._check(self.lib.jl_sched_result(self.h, request, native.ptr(out), n.value, C.byref(n), C.byref(state), C.byref(reason)))
        return out[:min(n.value, len(out))], state.value, reason.value

    def info(self, request):
        inf = native.SchedRequestInfo()
        self._check(self.lib.jl_sched_request_info(self.h, request, C.byref(inf)))
        return inf

    def release(self, request):
        self._check(self.lib.jl_sched_release(self.h, request))

    def counts(self):
        q, a, f = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.jl_sched_counts(self.h, C.byref(q), C.byref(a), C.byref(f)))
        return q.value, a.value, f.value

    def close(self):
        if self.h:
            self.lib.jl_sched_free(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
