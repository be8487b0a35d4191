/*
 * Reference-side binding for the concurrent-session scheduler of libjlama_b200.so (include/jlama_b200.h, jl_sched_*).
 *
 * Jlama serves concurrent requests with one thread per request, each inside AbstractModel.generate(UUID session, ...) on its own KvBuffer
 * (core/model/AbstractModel.java:516-646, core/tensor/KvBufferCache.java:58-60, jlama-net/.../openai/OpenAIChatService.java:64-74,107-160).
 * With a device-resident model those threads would serialise on the GPU and stream the weights once per request per token.  This class
 * keeps the call shape -- generate(session, promptTokens, maxNew, eosTokens, onToken) blocks its calling thread and streams tokens to the
 * callback -- but every call only queues a request; ONE stepping thread runs jl_sched_step, which batches all generating requests into
 * one decode step (the weights are read once for all of them), forwards prompt chunks under a token budget and reuses session slots as
 * requests finish.  A session UUID that comes back (the X-Jlama-Session header) continues on its kept KV at
 * kvmem.getCurrentContextPosition() exactly like AbstractModel.java:533.
 *
 * SOURCE ONLY (no JDK in this image): jlama_b200/scheduler.py binds the same entry points and is what tests/test_scheduler.py (CPU,
 * the native policy over the oracle) and tests/test_gpu_scheduler.py exercise.
 */
package com.github.tjake.jlama.model;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;
import java.util.Map;
import java.util.UUID;
import java.util.concurrent.ConcurrentHashMap;
import java.util.function.IntConsumer;

import static java.lang.foreign.ValueLayout.*;

public final class CudaSessionScheduler implements AutoCloseable {
    public static final int QUEUED = 0, PREFILL = 1, DECODING = 2, FINISHED = 3, FAILED = 4;          // JL_SCHED_*
    public static final int MAX_TOKENS = 1, STOP_TOKEN = 2, CANCELLED = 3, ERROR = 4;                    // JL_FINISH_* (Generator.FinishReason)
    private static final int KEEP_SESSION = 1;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB = SymbolLookup.loaderLookup().or(LINKER.defaultLookup());
    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
    }
    private static final MethodHandle jl_sched_create = h("jl_sched_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
    private static final MethodHandle jl_sched_free = h("jl_sched_free", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    private static final MethodHandle jl_sched_last_error = h("jl_sched_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle jl_sched_submit =
        h("jl_sched_submit", FunctionDescriptor.of(JAVA_LONG, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_LONG, JAVA_FLOAT, JAVA_LONG));
    private static final MethodHandle jl_sched_cancel = h("jl_sched_cancel", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG));
    private static final MethodHandle jl_sched_step = h("jl_sched_step", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    private static final MethodHandle jl_sched_result =
        h("jl_sched_result", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    private static final MethodHandle jl_sched_release = h("jl_sched_release", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG));

    private final Arena arena = Arena.ofShared();
    private final MemorySegment sched;
    private final Map<UUID, Long> kept = new ConcurrentHashMap<>(); // session UUID -> its last finished request (holds the KV slot)
    private final Thread stepper;
    private volatile boolean running = true;
    private final Object tick = new Object();

    /** prefillTokensPerStep bounds how long a new prompt may delay the decode step of the running requests (0 = whole prompts). */
    public CudaSessionScheduler(CudaLlamaModel model, int maxActive, int prefillTokensPerStep) {
        try {
            MemorySegment out = arena.allocate(ADDRESS);
            int rc = (int) jl_sched_create.invokeExact(model.handle(), maxActive, prefillTokensPerStep, out);
            if (rc != 0) throw new IllegalStateException("jl_sched_create: " + rc);
            sched = out.get(ADDRESS, 0);
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
        stepper = new Thread(this::loop, "jlama-cuda-scheduler");
        stepper.setDaemon(true);
        stepper.start();
    }

    /** The serving loop: step while anything is queued or running, sleep otherwise.  jl_sched_stats = 8 ints ([5] active, [6] queued). */
    private void loop() {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment st = a.allocate(JAVA_INT, 8);
            while (running) {
                int rc = (int) jl_sched_step.invokeExact(sched, st);   // a failing request is marked FAILED; the others go on
                boolean idle = st.getAtIndex(JAVA_INT, 5) == 0 && st.getAtIndex(JAVA_INT, 6) == 0;
                synchronized (tick) {
                    tick.notifyAll();                                  // wake the generate() callers: new tokens may be there
                    if (idle) tick.wait(5);
                }
            }
        } catch (Throwable t) { running = false; synchronized (tick) { tick.notifyAll(); } }
    }

    /**
     * AbstractModel.generate over token ids: blocks until the request finishes, calls onToken for every generated token in order, returns
     * the FinishReason code.  keepSession = the caller sent a session id and may come back with a follow-up prompt.  temperature 0 = arg-max;
     * otherwise every token is drawn with AbstractModel.sample's rule (:475-489) from a stream seeded here (the reference draws
     * ThreadLocalRandom.current().nextFloat() per token, :576,:594).
     */
    public int generate(UUID session, int[] promptTokens, float temperature, int maxNew, int[] eosTokens, boolean keepSession, IntConsumer onToken) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment pr = a.allocateFrom(JAVA_INT, promptTokens);
            MemorySegment eos = eosTokens.length == 0 ? MemorySegment.NULL : a.allocateFrom(JAVA_INT, eosTokens);
            Long parent = kept.remove(session);
            long id = (long) jl_sched_submit.invokeExact(sched, pr, promptTokens.length, maxNew, eos, eosTokens.length,
                                                          keepSession ? KEEP_SESSION : 0, parent == null ? -1L : (long) parent,
                                                          temperature, java.util.concurrent.ThreadLocalRandom.current().nextLong());
            if (id < 0) throw new IllegalArgumentException(((MemorySegment) jl_sched_last_error.invokeExact(sched)).reinterpret(512).getString(0));
            synchronized (tick) { tick.notifyAll(); }
            MemorySegment buf = a.allocate(JAVA_INT, maxNew), n = a.allocate(JAVA_INT), state = a.allocate(JAVA_INT), reason = a.allocate(JAVA_INT);
            int seen = 0;
            for (;;) {
                int rc = (int) jl_sched_result.invokeExact(sched, id, buf, maxNew, n, state, reason);
                if (rc != 0) throw new IllegalStateException("jl_sched_result: " + rc);
                for (int have = Math.min(n.get(JAVA_INT, 0), maxNew); seen < have; seen++) onToken.accept(buf.getAtIndex(JAVA_INT, seen));
                int s = state.get(JAVA_INT, 0);
                if (s == FINISHED || s == FAILED) break;
                if (!running) throw new IllegalStateException("scheduler stopped");
                synchronized (tick) { tick.wait(2); }
            }
            int why = reason.get(JAVA_INT, 0);
            if (parent != null) { int ignored = (int) jl_sched_release.invokeExact(sched, (long) parent); } // its slot moved to this request
            if (keepSession && state.get(JAVA_INT, 0) == FINISHED && why != CANCELLED) kept.put(session, id);
            else { int ignored = (int) jl_sched_release.invokeExact(sched, id); }
            if (state.get(JAVA_INT, 0) == FAILED)
                throw new RuntimeException(((MemorySegment) jl_sched_last_error.invokeExact(sched)).reinterpret(512).getString(0));
            return why;
        } catch (RuntimeException | Error e) { throw e; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** KvBufferCache.KvBuffer.close for a kept session: its slot returns to the pool. */
    public void closeSession(UUID session) {
        Long id = kept.remove(session);
        if (id != null) try { int rc = (int) jl_sched_release.invokeExact(sched, (long) id); } catch (Throwable ignored) { }
    }

    @Override public void close() {
        running = false;
        synchronized (tick) { tick.notifyAll(); }
        try { stepper.join(1000); int rc = (int) jl_sched_free.invokeExact(sched); } catch (Throwable ignored) { }
        arena.close();
    }
}
