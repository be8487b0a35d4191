// Persistent decode kernel: ONE cooperative launch = one decoded token (M = 1), one 512-thread CTA per SM.
//
// It executes AbstractModel.forward + sample (core/model/AbstractModel.java:314-329,443-473) for one session:
// embedding row, then per layer TransformerBlock.forward (TransformerBlock.java:158-215) = RMSNorm -> Q8 -> QKV
// (CausalSelfAttention.java:161-171) -> RoPE / KV append / attention over the pages (:199-356) -> Q8 -> o_proj + residual
// (:363-378) -> RMSNorm -> Q8 -> gate/up + SiLU*up (MLPBlock.java:117-141) -> Q8 -> down_proj + residual (:144-160), then
// the final norm, the F32 x Q4 lm_head GEMV and the arg-max (strict '>', lowest index).
//
// Why one kernel (DESIGN.md sections 5 and 9.1): in the kernel-per-op graph every GEMV launch paid ~0.85 us of launch
// gap plus ~3 us until its first weights arrived; 128 launches per token = 0.5 ms of a 1.74 ms step.  Here the five ops
// of a layer are PHASES of one resident grid:
//   * the GEMV body is the register-ring body of jl_gemv.cu (128-bit evict-first loads straight into a 3-deep ring of
//     register chunks, dp4a against Q8 activations staged in shared memory, warp-shuffle reduction, fused epilogues);
//   * a phase ends with `bar.sync; red.release.gpu` on a cumulative counter; the next phase's weights do not depend on
//     activations, so every warp puts its first ring chunks in flight BEFORE it polls the counter (ld.acquire.gpu): the
//     weight stream keeps running through the barrier and the ~3 us first-data latency of a fresh kernel is hidden;
//   * activations written by other CTAs are only ever read through L2 (ld.global.cg); residual rows are re-read by the
//     CTA that wrote them (same row partition in o_proj and down_proj);
//   * attention runs as (kv head, context split) tasks on the first CTAs while all others sit at the barrier with
//     their o_proj ring primed;
//   * tensor parallelism (jlama-net model shards, DistributedContext.java:79-98): the o_proj / down_proj partial sums
//     are exchanged INSIDE the kernel over NVLink peer memory with 8-byte self-validating {value, tag} stores (the LL
//     protocol: no fence round trip), reduced in rank order by the CTA that owns the rows, residual added in the same
//     pass -- this replaces JlamaService.combine (jlama-net .../JlamaService.java:300-359) and the round-1
//     ncclAllReduce + copy + accumulate; the lm_head is sharded by vocabulary rows with an exchange of per-rank
//     (max, argmax) candidates.
// Every spin is bounded: a protocol error sets sync[1] and the launch drains instead of hanging the GPU.
#include "jl_pdecode.cuh"
#include "jl_gemv_body.cuh"
#include "jl_attn_task.cuh"

#define PD_NT PD_THREADS
#define PD_NWARP (PD_NT / 32)
#define PD_CH(WDT) ((WDT) == JL_Q4 ? 4 : 2) // 32-block groups per lane per chunk: an int8 block is two 128-bit loads
#define PD_NBUF 3
#define PD_SPIN_LOCAL (1u << 22)
#define PD_SPIN_PEER (1u << 26)
#define EPI_LL 3 // partial sums to every rank's LL receive buffer (tensor parallel o_proj / down_proj)

// ---- cross-CTA ordering ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pd_ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long pd_ld_volatile(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// all threads of the CTA call; publishes everything the CTA wrote in this phase
__device__ __forceinline__ void pd_arrive(const PdParams &P, int which) {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(P.sync + which) : "memory");
}
// all threads of the CTA call; sync = P.sync
__device__ __forceinline__ void pd_wait_raw(unsigned long long *sync, int which, unsigned long long target) {
    if (threadIdx.x == 0) {
        const unsigned long long *c = sync + which;
        unsigned n = 0;
        while (pd_ld_acquire(c) < target) {
            if ((++n & 0x3ff) == 0) {
                if (pd_ld_volatile(sync + 1) != 0) break; // another CTA gave up: drain
                if (n > PD_SPIN_LOCAL) {
                    sync[1] = (unsigned long long)which;
                    break;
                }
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void pd_wait(const PdParams &P, int which, unsigned long long target) { pd_wait_raw(P.sync, which, target); }
// dependency of a phase: counter index + cumulative target (+ optional timeline stamp written once the wait is over)
struct PdDep {
    unsigned long long *sync;
    int which;
    unsigned long long target;
    unsigned long long *stamp; // nullptr or the slot of CTA 0
};
__device__ __forceinline__ void pd_dep_wait(const PdDep &d) {
    pd_wait_raw(d.sync, d.which, d.target);
    if (d.stamp && blockIdx.x == 0 && threadIdx.x == 0) *d.stamp = globaltimer_ns();
}

// ---- LL exchange ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint4 *const *bufs, int world, int src_rank, int E, int row, float v, uint32_t tag) {
    // half-line (8 bytes) of row `row` in the [src_rank] block of every rank's receive buffer
#pragma unroll 1
    for (int d = 0; d < world; d++) {
        uint32_t *dst = (uint32_t *)bufs[d] + ((size_t)src_rank * E + row) * 2;
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(__float_as_uint(v)), "r"(tag) : "memory");
    }
}
__device__ __forceinline__ float ll_load(unsigned long long *sync, const uint4 *mine, int src_rank, int E, int row, uint32_t tag) {
    const uint32_t *src = (const uint32_t *)mine + ((size_t)src_rank * E + row) * 2;
    uint32_t v, t;
    unsigned n = 0;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(t) : "l"(src) : "memory");
        if (t == tag) break;
        if ((++n & 0x3ff) == 0) {
            if (pd_ld_volatile(sync + 1) != 0) break;
            if (n > PD_SPIN_PEER) {
                sync[1] = 100 + (unsigned long long)src_rank;
                break;
            }
        }
    }
    return __uint_as_float(v);
}

__device__ __forceinline__ int item_owner_pd(int i, int items, int nwarp) { return (int)((((long long)(i + 1)) * nwarp - 1) / items); }

struct PdEpi {
    // EPI_LL
    uint4 *const *ll;
    int world, rank, E;
    uint32_t tag;
};
// weight / output segments of a phase as scalars (a runtime-indexed array would force the whole descriptor into local memory)
struct PdSegs {
    const uint8_t *w0, *w1, *w2;
    const float *s0, *s1, *s2;
    float *o0, *o1, *o2;
    int r0, r1, nseg;
    const float *residual;
};
__device__ __forceinline__ void pd_seg_lookup(const PdSegs &S, int row, int &seg, int &local) {
    seg = 0, local = row;
    if (S.nseg > 1 && local >= S.r0) {
        local -= S.r0, seg = 1;
        if (S.nseg > 2 && local >= S.r1) local -= S.r1, seg = 2;
    }
}

__device__ __forceinline__ unsigned long long pd_pack_arg(float v, int idx) {
    if (!(v == v)) return 0ull; // NaN never wins (AbstractModel.java:465: 'v > maxv' is false)
    uint32_t b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}

__device__ __forceinline__ void pd_stamp(const PdParams &P, int idx) {
    if (P.trace && blockIdx.x == 0 && threadIdx.x == 0) P.trace[idx] = globaltimer_ns();
}

// ---- one quantised GEMV phase (M = 1, Q8 activations) -------------------------------------------------------------------
// Work split as in gemv_decode_kernel (jl_gemv.cu): output rows are dealt to the CTAs, inside a CTA the (row, weight-row,
// chunk) items to the warps; LONG rows (K > 4096) are split at chunk granularity with per-warp partial sums combined in
// chunk order.  `dep()` is called by all threads after the ring has been primed and before the activations are touched.
// __noinline__: every phase gets its own register allocation (inlined into one kernel the ring buffers spilled).
template <int WDT, int EPI, int PRO, bool LONG>
__device__ __noinline__ void pd_gemv(const GemvParams &p, const PdSegs &S, const PdEpi &X, unsigned char *smem, const PdDep dep) {
    constexpr int NT = PD_NT, NWARP = PD_NWARP, CH = PD_CH(WDT), NBUF = PD_NBUF;
    constexpr bool NORM = PRO == PRO_RMSNORM_QUANT;
    constexpr int NW = (EPI == EPI_SILU_MUL) ? 2 : 1; // weight rows per output row
    constexpr int WB = (WDT == JL_Q4) ? 16 : 32;      // weight bytes per 32-element block
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = p.K / 32;
    const int nchunks = LONG ? (nblk + 32 * CH - 1) / (32 * CH) : 1;
    const int R0 = (int)(((long long)p.total_rows * blockIdx.x) / gridDim.x);
    const int R1 = (int)(((long long)p.total_rows * (blockIdx.x + 1)) / gridDim.x);
    const int nrows = R1 - R0;
    const int per_row = NW * nchunks;
    const int items = nrows * per_row;
    int i0, i1;
    if (LONG) {
        i0 = (int)(((long long)items * warp) / NWARP);
        i1 = (int)(((long long)items * (warp + 1)) / NWARP);
    } else {
        i0 = (int)(((long long)nrows * warp) / NWARP) * per_row;
        i1 = (int)(((long long)nrows * (warp + 1)) / NWARP) * per_row;
    }
    struct It {
        int r, wr, c;
    };
    const size_t row_blocks = (size_t)nblk;
    auto row_ptrs = [&](const It &it, const uint8_t *&wrow, const float *&srow) {
        int seg, local;
        if (EPI == EPI_SILU_MUL) {
            seg = it.wr;
            local = R0 + it.r;
        } else {
            pd_seg_lookup(S, R0 + it.r, seg, local);
        }
        const size_t blk = (size_t)local * row_blocks;
        wrow = (seg == 0 ? S.w0 : (seg == 1 ? S.w1 : S.w2)) + blk * WB;
        srow = (seg == 0 ? S.s0 : (seg == 1 ? S.s1 : S.s2)) + blk;
    };
    auto advance = [&](It &it) {
        if (!LONG || ++it.c == nchunks) {
            it.c = 0;
            if (++it.wr == NW) {
                it.wr = 0;
                ++it.r;
            }
        }
    };
    It cur, ld;
    {
        cur.r = i0 / per_row;
        const int rem = i0 - cur.r * per_row;
        cur.wr = rem / nchunks;
        cur.c = rem - cur.wr * nchunks;
        ld = cur;
    }
    WBuf<WDT, CH> buf[NBUF];
    const uint8_t *wrow;
    const float *srow;
    const unsigned long long pol = l2_evict_first_policy();
    int ci = i0, li = i0;
    // the ring is primed before the dependency is polled: the weight stream does not depend on the previous phase
#pragma unroll
    for (int b = 0; b < NBUF - 1; b++) {
        if (li < i1) {
            row_ptrs(ld, wrow, srow);
            load_chunk<WDT, CH>(buf[b], wrow, srow, ld.c * 32 * CH, nblk, lane, pol);
            advance(ld);
            li++;
        }
    }
    pd_dep_wait(dep);
    {
        StageRegs<NORM, LONG && !NORM> sr;
        stage_q8_issue<NORM, LONG && !NORM, NT>(p, sr);
        stage_q8_finish<NORM, LONG && !NORM, NT>(p, sr, smem, nblk);
    }
    float *parts = (float *)(smem + (((size_t)nblk * 40 + 15) & ~(size_t)15));
    const int maxsplit = nchunks < NWARP ? nchunks : NWARP;
    float acc[1] = {0.0f};
    float park0 = 0.0f, park1 = 0.0f; // value (or gate), up
    int nparked = 0, park_row0 = R0 + cur.r;
    auto store_row = [&](int row, float v0, float v1) { // row = index in the concatenated row space of the launch
        int seg = 0, local = row;
        if (EPI != EPI_SILU_MUL) pd_seg_lookup(S, row, seg, local);
        float v = NW == 2 ? v1 : v0;
        if (EPI == EPI_ADD_RESIDUAL) v = __fadd_rn(v, __ldcg(S.residual + local));
        if (EPI == EPI_SILU_MUL) v = __fmul_rn(silu_ref(v0), v);
        if (EPI == EPI_LL) ll_store(X.ll, X.world, X.rank, X.E, local, v, X.tag);
        else (seg == 0 ? S.o0 : (seg == 1 ? S.o1 : S.o2))[local] = v;
    };
    auto flush = [&]() {
        if (lane < nparked) store_row(park_row0 + lane, park0, park1);
        park_row0 += nparked;
        nparked = 0;
    };
    while (ci < i1) {
#pragma unroll
        for (int b = 0; b < NBUF; b++) {
            if (ci < i1) {
                if (li < i1) {
                    row_ptrs(ld, wrow, srow);
                    load_chunk<WDT, CH>(buf[(b + NBUF - 1) % NBUF], wrow, srow, ld.c * 32 * CH, nblk, lane, pol);
                    advance(ld);
                    li++;
                }
                compute_chunk<WDT, true, 1, CH>(buf[b], acc, smem, cur.c * 32 * CH, nblk, lane);
                if (!LONG) {
                    const float v = warp_sum(acc[0]);
                    acc[0] = 0.0f;
                    if (NW == 2 && cur.wr == 0) {
                        if (lane == nparked) park0 = v;
                    } else {
                        if (lane == nparked) (NW == 2 ? park1 : park0) = v;
                        ++nparked;
                    }
                } else if (cur.c == nchunks - 1 || ci == i1 - 1) {
                    const float v = warp_sum(acc[0]);
                    const int rw = cur.r * NW + cur.wr;
                    if (lane == 0) parts[rw * maxsplit + (warp - item_owner_pd(rw * nchunks, items, NWARP))] = v;
                    acc[0] = 0.0f;
                }
                advance(cur);
                ci++;
            }
        }
        if (!LONG && nparked > 32 - NBUF) flush();
    }
    if (!LONG) flush();
    if (LONG) {
        __syncthreads();
        for (int o = tid; o < nrows; o += NT) {
            float sums[NW];
#pragma unroll
            for (int wr = 0; wr < NW; wr++) {
                const int rw = o * NW + wr;
                const int wa = item_owner_pd(rw * nchunks, items, NWARP), wb = item_owner_pd(rw * nchunks + nchunks - 1, items, NWARP);
                float t = 0.0f;
                for (int w = wa; w <= wb; w++) t = __fadd_rn(t, parts[rw * maxsplit + (w - wa)]);
                sums[wr] = t;
            }
            store_row(R0 + o, sums[0], sums[NW - 1]);
        }
    }
}

// ---- lm_head: F32 activations (RMSNorm, not re-quantised: AbstractModel.java:444-449) x quantised rows + running arg-max --
template <int WDT>
__device__ __noinline__ void pd_lm_head(const PdParams &P, unsigned char *smem, const PdDep dep, unsigned long long &cta_best) {
    constexpr int NT = PD_NT, NWARP = PD_NWARP, CH = PD_CH(WDT), NBUF = PD_NBUF;
    constexpr int WB = (WDT == JL_Q4) ? 16 : 32;
    __shared__ double red[NWARP];
    __shared__ unsigned long long wbest[NWARP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = P.E, nblk = K / 32;
    const int R0 = (int)(((long long)P.vocab_rows * blockIdx.x) / gridDim.x);
    const int R1 = (int)(((long long)P.vocab_rows * (blockIdx.x + 1)) / gridDim.x);
    const int nrows = R1 - R0;
    const int r0 = R0 + (int)(((long long)nrows * warp) / NWARP), r1 = R0 + (int)(((long long)nrows * (warp + 1)) / NWARP);
    const int nchunks = (nblk + 32 * CH - 1) / (32 * CH);
    WBuf<WDT, CH> buf[NBUF];
    const unsigned long long pol = l2_evict_first_policy();
    int lr = r0, lc = 0, cr = r0, cc = 0;
    auto issue = [&](WBuf<WDT, CH> &b) {
        const size_t blk = (size_t)lr * nblk;
        load_chunk<WDT, CH>(b, P.lm_w + blk * WB, P.lm_s + blk, lc * 32 * CH, nblk, lane, pol);
        if (++lc == nchunks) lc = 0, ++lr;
    };
#pragma unroll
    for (int b = 0; b < NBUF - 1; b++)
        if (lr < r1) issue(buf[b]);
    pd_dep_wait(dep);
    // final RMSNorm into the [c4(8)][blk][4] float layout of compute_chunk<.., ACTQ8 = false>
    {
        float4 *af4 = (float4 *)smem;
        double ss = 0.0;
        for (int i4 = tid; i4 < K / 4; i4 += NT) {
            const float4 v = __ldcg((const float4 *)(P.x + i4 * 4));
            ss += (double)__fmul_rn(v.x, v.x);
            ss += (double)__fmul_rn(v.y, v.y);
            ss += (double)__fmul_rn(v.z, v.z);
            ss += (double)__fmul_rn(v.w, v.w);
        }
        ss = warp_sum_d(ss);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NWARP; w++) t += red[w];
        const float rsf = rms_scale(t, 1.0 / (double)P.E, P.eps);
        for (int i4 = tid; i4 < K / 4; i4 += NT) {
            const float4 v = __ldcg((const float4 *)(P.x + i4 * 4));
            float w[4];
            if (P.out_norm_dt == JL_BF16) {
                const uint2 u = __ldg((const uint2 *)((const uint16_t *)P.out_norm + i4 * 4));
                w[0] = __uint_as_float(u.x << 16), w[1] = __uint_as_float(u.x & 0xffff0000u);
                w[2] = __uint_as_float(u.y << 16), w[3] = __uint_as_float(u.y & 0xffff0000u);
            } else {
                const float4 f = __ldg((const float4 *)((const float *)P.out_norm + i4 * 4));
                w[0] = f.x, w[1] = f.y, w[2] = f.z, w[3] = f.w;
            }
            const int b = i4 >> 3, c4 = i4 & 7;
            af4[(size_t)c4 * nblk + b] = make_float4(__fmul_rn(__fadd_rn(0.0f, w[0]), __fmul_rn(rsf, v.x)), __fmul_rn(__fadd_rn(0.0f, w[1]), __fmul_rn(rsf, v.y)),
                                                     __fmul_rn(__fadd_rn(0.0f, w[2]), __fmul_rn(rsf, v.z)), __fmul_rn(__fadd_rn(0.0f, w[3]), __fmul_rn(rsf, v.w)));
        }
        __syncthreads();
    }
    float acc[1] = {0.0f};
    float best_v = -INFINITY;
    int best_i = 0x7fffffff;
    while (cr < r1) {
#pragma unroll
        for (int b = 0; b < NBUF; b++) {
            if (cr < r1) {
                if (lr < r1) issue(buf[(b + NBUF - 1) % NBUF]);
                compute_chunk<WDT, false, 1, CH>(buf[b], acc, smem, cc * 32 * CH, nblk, lane);
                if (++cc == nchunks) {
                    const float v = warp_sum(acc[0]);
                    acc[0] = 0.0f;
                    const int grow = P.vocab0 + cr;
                    if (lane == 0) {
                        P.logits[grow] = v;
                        if (P.want_logits)
                            for (int d = 0; d < P.world; d++)
                                if (d != P.rank) P.logits_peer[d][grow] = v;
                    }
                    if (v > best_v) best_v = v, best_i = grow; // rows ascend: strict '>' keeps the lowest index
                    cc = 0, ++cr;
                }
            }
        }
    }
    if (lane == 0) wbest[warp] = best_i == 0x7fffffff ? 0ull : pd_pack_arg(best_v, best_i);
    __syncthreads();
    if (tid == 0) {
        unsigned long long b = 0ull;
        for (int w = 0; w < NWARP; w++) b = wbest[w] > b ? wbest[w] : b;
        cta_best = b;
    }
}

__device__ __forceinline__ float pd_embed_value(const PdParams &P, size_t tok, int c) {
    // LlamaModel.java:68-100 (Q4 / I8 rows are read through get(): Q4ByteBufferTensor.java:179-197)
    if (P.embed_dt == JL_F32) return ((const float *)P.embed_w)[tok * P.E + c];
    if (P.embed_dt == JL_BF16) return bf16_bits_to_f32(((const uint16_t *)P.embed_w)[tok * P.E + c]);
    if (P.embed_dt == JL_Q4) {
        const int blk = c / 32, in = c % 32;
        const uint8_t byte = ((const uint8_t *)P.embed_w)[(tok * P.E + blk * 32) / 2 + (in & 15)];
        const int nib = in < 16 ? (byte & 0x0F) : (byte >> 4);
        return __fmul_rn((float)(nib - 8), P.embed_s[tok * (P.E / 32) + blk]);
    }
    return __fmul_rn((float)((const int8_t *)P.embed_w)[tok * P.E + c], P.embed_s[tok * (P.E / 32) + c / 32]);
}

// reduce the LL partials of this CTA's rows in rank order, add the residual, store the new hidden rows (local global)
__device__ __noinline__ void pd_ll_reduce(const PdParams &P, const uint4 *mine, uint32_t tag, const float *residual, float *out,
                                             float *scratch /* smem [rows][world] */) {
    const int tid = threadIdx.x;
    const int R0 = (int)(((long long)P.E * blockIdx.x) / gridDim.x), R1 = (int)(((long long)P.E * (blockIdx.x + 1)) / gridDim.x);
    const int nrows = R1 - R0, W = P.world;
    for (int t = tid; t < nrows * W; t += PD_NT) {
        const int r = t / W, src = t - r * W;
        scratch[t] = ll_load(P.sync, mine, src, P.E, R0 + r, tag);
    }
    __syncthreads();
    for (int r = tid; r < nrows; r += PD_NT) {
        float s = scratch[r * W];
        for (int src = 1; src < W; src++) s = __fadd_rn(s, scratch[r * W + src]); // rank order: identical on every rank
        out[R0 + r] = __fadd_rn(s, __ldcg(residual + R0 + r));
    }
}

template <int HS>
__device__ __noinline__ void pd_attention(const AttnTask &at, int layer, int kvh, int split, unsigned char *smem) {
    attention_task<HS, PD_NT, -1>(at, layer, 0, kvh, split, smem);
}

template <int WDT, int HS>
__global__ void __launch_bounds__(PD_NT, 1) pdecode_kernel(const PdParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int G = gridDim.x, cta = blockIdx.x;
    const unsigned long long epoch = pd_ld_volatile(P.sync); // tokens decoded by this model so far
    const unsigned long long uG = (unsigned long long)G;
    const bool tp = P.world > 1;
    const bool lng_E = P.E > 32 * 32 * PD_CH(WDT), lng_A = P.attn_seg > 32 * 32 * PD_CH(WDT); // rows longer than one register chunk
    auto stamp_slot = [&](int idx) -> unsigned long long * { return P.trace ? P.trace + idx : nullptr; };
    pd_stamp(P, 0);

    // ---- embedding row (columns split over the grid) ----
    {
        const int c_a = (int)(((long long)P.E * cta) / G), c_b = (int)(((long long)P.E * (cta + 1)) / G);
        const size_t tok = (size_t)P.tokens[0];
        for (int c = c_a + tid; c < c_b; c += PD_NT) P.x[c] = pd_embed_value(P, tok, c);
        pd_arrive(P, PC_EMBED);
    }
    AttnTask at;
    at.heads = P.heads, at.kv_heads = P.kv_heads, at.head_size = P.head_size, at.attn_seg = P.attn_seg, at.kv_seg = P.kv_seg;
    at.kv_head0_global = P.kv_head0_global, at.splits = P.splits, at.attn_scale = P.attn_scale;
    at.q = P.q, at.k = P.k, at.v = P.v, at.att = P.att, at.attn_ws = P.attn_ws, at.rope = P.rope, at.kv = P.kv;
    at.sessions = P.sessions, at.positions = P.positions;
    const int ntasks = P.kv_heads * P.splits;
    PdEpi X;
    X.world = P.world, X.rank = P.rank, X.E = P.E, X.ll = nullptr, X.tag = 0;

    for (int L = 0; L < P.layers; L++) {
        const PdLayer &lw = P.lw[L];
        const unsigned long long use = epoch * (unsigned long long)P.layers + (unsigned long long)L + 1; // 1-based use count
        const uint32_t tag_o = (uint32_t)(use * 2), tag_d = (uint32_t)(use * 2 + 1);
        // ---- QKV: RMSNorm(x) -> Q8 -> q | k | v ----
        {
            GemvParams p = {};
            PdSegs S = {lw.w[PW_Q], lw.w[PW_K], lw.w[PW_V], lw.s[PW_Q], lw.s[PW_K], lw.s[PW_V], P.q, P.k, P.v, P.attn_seg, P.kv_seg, 3, nullptr};
            p.K = P.E, p.a = P.x, p.lda = P.E;
            p.norm_w = lw.attn_norm, p.norm_w_dtype = lw.attn_norm_dt, p.norm_eps = P.eps, p.norm_E = P.E, p.norm_inv_E = 1.0 / (double)P.E;
            p.total_rows = P.attn_seg + 2 * P.kv_seg;
            const PdDep dep = {P.sync, L == 0 ? PC_EMBED : PC_DOWN, L == 0 ? (epoch + 1) * uG : (use - 1) * uG, stamp_slot(1 + L * 8 + 0)};
            if (lng_E) pd_gemv<WDT, EPI_STORE, PRO_RMSNORM_QUANT, true>(p, S, X, smem, dep);
            else pd_gemv<WDT, EPI_STORE, PRO_RMSNORM_QUANT, false>(p, S, X, smem, dep);
            pd_arrive(P, PC_QKV);
            pd_stamp(P, 1 + L * 8 + 1);
        }
        // ---- attention tasks on the first CTAs (RoPE, KV append, scores, softmax, P.V) ----
        if (cta < ntasks) {
            pd_wait(P, PC_QKV, use * uG);
            const int split = cta % P.splits, kvh = cta / P.splits;
            pd_attention<HS>(at, L, kvh, split, smem);
            bool signal = true;
            if (P.splits > 1) {
                __syncthreads();
                if (tid == 0) {
                    __threadfence();
                    unsigned *c = &P.att_done[kvh];
                    const unsigned old = atomicAdd(c, 1u);
                    s_last = (old == (unsigned)P.splits - 1);
                    if (s_last) *c = 0;
                    __threadfence();
                }
                __syncthreads();
                signal = s_last != 0;
                if (signal) attention_merge<HS, PD_NT>(at, 0, kvh, smem);
            }
            __syncthreads();
            if (signal && tid == 0) asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(P.sync + PC_ATT) : "memory");
            pd_stamp(P, 1 + L * 8 + 2);
        }
        // ---- o_proj: Q8(att) -> x + o (single rank) or LL partials -> rank-ordered reduce + residual (tensor parallel) ----
        {
            GemvParams p = {};
            PdSegs S = {lw.w[PW_O], nullptr, nullptr, lw.s[PW_O], nullptr, nullptr, P.xb, nullptr, nullptr, P.E, 0, 1, P.x};
            p.K = P.attn_seg, p.a = P.att, p.lda = P.attn_seg, p.total_rows = P.E;
            const PdDep dep = {P.sync, PC_ATT, use * (unsigned long long)P.kv_heads, stamp_slot(1 + L * 8 + 3)};
            if (tp) {
                X.ll = P.ll_o, X.tag = tag_o;
                if (lng_A) pd_gemv<WDT, EPI_LL, PRO_F32_QUANT, true>(p, S, X, smem, dep);
                else pd_gemv<WDT, EPI_LL, PRO_F32_QUANT, false>(p, S, X, smem, dep);
                __syncthreads();
                pd_ll_reduce(P, P.ll_o[P.rank], tag_o, P.x, P.xb, (float *)smem);
            } else {
                if (lng_A) pd_gemv<WDT, EPI_ADD_RESIDUAL, PRO_F32_QUANT, true>(p, S, X, smem, dep);
                else pd_gemv<WDT, EPI_ADD_RESIDUAL, PRO_F32_QUANT, false>(p, S, X, smem, dep);
            }
            pd_arrive(P, PC_O);
            pd_stamp(P, 1 + L * 8 + 4);
        }
        // ---- gate / up: RMSNorm(xb) -> Q8 -> silu(gate) * up ----
        {
            GemvParams p = {};
            PdSegs S = {lw.w[PW_GATE], lw.w[PW_UP], nullptr, lw.s[PW_GATE], lw.s[PW_UP], nullptr, P.h, P.h, nullptr, P.H, P.H, 2, nullptr};
            p.K = P.E, p.a = P.xb, p.lda = P.E;
            p.norm_w = lw.ffn_norm, p.norm_w_dtype = lw.ffn_norm_dt, p.norm_eps = P.eps, p.norm_E = P.E, p.norm_inv_E = 1.0 / (double)P.E;
            p.total_rows = P.H;
            const PdDep dep = {P.sync, PC_O, use * uG, stamp_slot(1 + L * 8 + 5)};
            if (lng_E) pd_gemv<WDT, EPI_SILU_MUL, PRO_RMSNORM_QUANT, true>(p, S, X, smem, dep);
            else pd_gemv<WDT, EPI_SILU_MUL, PRO_RMSNORM_QUANT, false>(p, S, X, smem, dep);
            pd_arrive(P, PC_GU);
            pd_stamp(P, 1 + L * 8 + 6);
        }
        // ---- down_proj: Q8(h) -> xb + down ----
        {
            GemvParams p = {};
            PdSegs S = {lw.w[PW_DOWN], nullptr, nullptr, lw.s[PW_DOWN], nullptr, nullptr, P.x, nullptr, nullptr, P.E, 0, 1, P.xb};
            p.K = P.H, p.a = P.h, p.lda = P.H, p.total_rows = P.E;
            const PdDep dep = {P.sync, PC_GU, use * uG, stamp_slot(1 + L * 8 + 7)};
            const bool lng = P.H > 32 * 32 * PD_CH(WDT);
            if (tp) {
                X.ll = P.ll_d, X.tag = tag_d;
                if (lng) pd_gemv<WDT, EPI_LL, PRO_F32_QUANT, true>(p, S, X, smem, dep);
                else pd_gemv<WDT, EPI_LL, PRO_F32_QUANT, false>(p, S, X, smem, dep);
                __syncthreads();
                pd_ll_reduce(P, P.ll_d[P.rank], tag_d, P.xb, P.x, (float *)smem);
            } else {
                if (lng) pd_gemv<WDT, EPI_ADD_RESIDUAL, PRO_F32_QUANT, true>(p, S, X, smem, dep);
                else pd_gemv<WDT, EPI_ADD_RESIDUAL, PRO_F32_QUANT, false>(p, S, X, smem, dep);
            }
            pd_arrive(P, PC_DOWN);
        }
    }
    // ---- final norm + lm_head + arg-max ----
    unsigned long long cta_best = 0ull;
    {
        const PdDep dep = {P.sync, PC_DOWN, (epoch + 1) * (unsigned long long)P.layers * uG, stamp_slot(1 + P.layers * 8 + 0)};
        pd_lm_head<WDT>(P, smem, dep, cta_best);
    }
    if (tid == 0) P.argmax_slots[cta] = cta_best;
    pd_arrive(P, PC_LM);
    if (cta == 0) {
        pd_wait(P, PC_LM, (epoch + 1) * uG);
        if (tid < 32) {
            unsigned long long b = 0ull;
            for (int i = tid; i < G; i += 32) {
                const unsigned long long c = __ldcg(&P.argmax_slots[i]);
                b = c > b ? c : b;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long c = __shfl_xor_sync(0xffffffffu, b, o);
                b = c > b ? c : b;
            }
            if (tp) {
                // exchange the per-rank candidates (lm_head is sharded by vocabulary rows): LL line {lo, tag, hi, tag}
                const uint32_t tag = (uint32_t)(epoch + 1);
                if (tid < P.world) {
                    uint4 line = make_uint4((uint32_t)b, tag, (uint32_t)(b >> 32), tag);
                    uint4 *dst = P.ll_a[tid] + P.rank;
                    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(line.x), "r"(line.y), "r"(line.z), "r"(line.w)
                                 : "memory");
                }
                unsigned long long c = 0ull;
                if (tid < P.world) {
                    const uint4 *src = P.ll_a[P.rank] + tid;
                    uint4 v;
                    unsigned n = 0;
                    for (;;) {
                        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(src) : "memory");
                        if (v.y == tag && v.w == tag) break;
                        if ((++n & 0x3ff) == 0) {
                            if (pd_ld_volatile(P.sync + 1) != 0) break;
                            if (n > PD_SPIN_PEER) {
                                P.sync[1] = 200 + (unsigned long long)tid;
                                break;
                            }
                        }
                    }
                    c = ((unsigned long long)v.z << 32) | v.x;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long d = __shfl_xor_sync(0xffffffffu, c, o);
                    c = d > c ? d : c;
                }
                b = c;
            }
            if (tid == 0) {
                const int tok = b == 0ull ? 0 : (int)(0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFull));
                P.next[0] = tok;
                if (P.resident) {
                    const int cntv = *P.counter;
                    P.tokens[0] = tok;
                    P.positions[0] += 1;
                    if (cntv < P.hist_cap) P.hist[cntv] = tok;
                    *P.counter = cntv + 1;
                }
                P.sync[0] = epoch + 1;
                pd_stamp(P, 1 + P.layers * 8 + 1);
            }
        }
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
static size_t pd_smem_bytes(const PdParams &p) {
    int Kmax = p.E;
    if (p.H > Kmax) Kmax = p.H;
    if (p.attn_seg > Kmax) Kmax = p.attn_seg;
    const int G = 148;
    int Rmax = p.E > p.H ? p.E : p.H; // rows of the largest phase that may run in the split-row (LONG) form; gate+up has two weight rows
    if (p.attn_seg + 2 * p.kv_seg > Rmax) Rmax = p.attn_seg + 2 * p.kv_seg;
    size_t acts = (((size_t)(Kmax / 32) * 40 + 15) & ~(size_t)15) + (size_t)((Rmax + G - 1) / G + 2) * 2 * PD_NWARP * 4 + 64;
    const size_t lm = (size_t)p.E * 4;
    if (lm > acts) acts = lm;
    const int hs = p.head_size;
    size_t att = hs == 32 ? attention_task_smem<32, PD_NT>() : (hs == 64 ? attention_task_smem<64, PD_NT>() : attention_task_smem<128, PD_NT>());
    att += 1024; // merge factors
    size_t m = acts > att ? acts : att;
    const size_t red = (size_t)((p.E + G - 1) / G + 2) * PD_MAX_TP * 4;
    if (red > m) m = red;
    return (m + 1023) & ~(size_t)1023;
}

bool jl_pdecode_supported(const PdParams &p, int w_dtype, int grid) {
    if (w_dtype != JL_Q4 && w_dtype != JL_I8) return false;
    if (p.head_size != 32 && p.head_size != 64 && p.head_size != 128) return false;
    if (p.heads % p.kv_heads || p.heads / p.kv_heads > MG_MAX_GROUP) return false;
    if ((p.E % 32) || (p.H % 32) || (p.attn_seg % 32) || (p.E % 8)) return false;
    if (p.E > 16 * PD_NT) return false;        // RMSNorm prologue keeps the row in registers (16 floats per thread)
    if (p.kv_heads * 1 > grid) return false;
    if (p.world > PD_MAX_TP) return false;
    if (PD_NT % p.head_size) return false;
    return pd_smem_bytes(p) <= 200 * 1024;
}

template <int WDT, int HS>
static int launch_pd(jl_ctx *ctx, cudaStream_t stream, const PdParams &p) {
    auto kern = pdecode_kernel<WDT, HS>;
    const size_t smem = pd_smem_bytes(p);
    static size_t configured[JL_MAX_DEVICES] = {};
    JL_CUDA_CHECK(ctx, jl_ensure_dyn_smem(kern, ctx->device, smem, configured));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctx->sm_count);
    cfg.blockDim = dim3(PD_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; // all CTAs co-resident: the (bounded) spin waits cannot starve each other
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    JL_CUDA_CHECK(ctx, cudaLaunchKernelEx(&cfg, kern, p));
    ctx->launches++;
    return JL_OK;
}

template <int WDT>
static int launch_pd_hs(jl_ctx *ctx, cudaStream_t stream, const PdParams &p) {
    switch (p.head_size) {
        case 32: return launch_pd<WDT, 32>(ctx, stream, p);
        case 64: return launch_pd<WDT, 64>(ctx, stream, p);
        default: return launch_pd<WDT, 128>(ctx, stream, p);
    }
}

int jl_launch_pdecode(jl_ctx *ctx, cudaStream_t stream, const PdParams &p, int w_dtype) {
    if (!jl_pdecode_supported(p, w_dtype, ctx->sm_count)) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "persistent decode: unsupported shape");
    return w_dtype == JL_Q4 ? launch_pd_hs<JL_Q4>(ctx, stream, p) : launch_pd_hs<JL_I8>(ctx, stream, p);
}
