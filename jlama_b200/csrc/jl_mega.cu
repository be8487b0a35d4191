// Persistent cooperative decode megakernel: one launch = one decoded token for M <= 4 sessions.
//
// It executes the whole of AbstractModel.forward + sample (core/model/AbstractModel.java:314-329,443-473;
// TransformerBlock.java:158-215; CausalSelfAttention.java:145-385; MLPBlock.java:106-166) in a single grid of
// one CTA per SM.  Why: at batch 1 every Llama-3-8B GEMV is 2-11 us of HBM streaming, so per-kernel launch,
// prologue and drain latencies (measured ~10 us per launch) dominate a kernel-per-op design.  Here
//
//   * one PRODUCER warp per CTA walks the CTA's static share of every weight matrix of every layer and streams
//     it with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) into a shared-memory ring; it never
//     waits for activations, only for free ring slots, and when the ring is full it keeps HBM busy by issuing
//     L2 prefetches (cp.async.bulk.prefetch.L2) for the next stages of its schedule;
//   * 16 CONSUMER warps wait on the ring's mbarriers, run the dp4a block dot products against the Q8
//     activations staged in shared memory, and apply the fused epilogues (residual add, SiLU*up, arg-max);
//   * ops are ordered by per-op completion counters in global memory (release/acquire), not kernel boundaries;
//     RMSNorm + Q8 quantisation are recomputed per CTA in the op prologue (one L2 round trip);
//   * attention (RoPE, KV append, scores, softmax, P.V over the paged KV cache) runs as (row, kv-head, split)
//     tasks on the first CTAs while every producer keeps prefetching the following matrices.
//
// Arithmetic is the same as the kernel-per-op path (jl_gemv.cu / jl_attention.cu): identical per-lane block
// order, so GEMV results are bit-identical; the launch is cooperative so the spin waits cannot deadlock.
#include "jl_mega.cuh"

#define MG_CWARPS 16
#define MG_CONSUMERS (MG_CWARPS * 32)
#define MG_THREADS (MG_CONSUMERS + 32)
#define MG_ROWS 16
#define MG_SLICE 4096
#define MG_PSLICE 2048
#define MG_STAGE_NIB 32768
#define MG_STAGE_SC 8192
#define MG_STAGE_BYTES (MG_STAGE_NIB + MG_STAGE_SC)
#define MG_L2_AHEAD 8
#define MG_ATT_TILE 32
#define MG_MAX_GROUP 8

// ---- PTX helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(MG_CONSUMERS) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// ---- static schedule ------------------------------------------------------------------------------------------------
// ops: for each layer QKV, O, GATEUP, DOWN; then LMHEAD.
enum { OP_QKV = 0, OP_O, OP_GATEUP, OP_DOWN, OP_LMHEAD };
struct OpInfo {
    int type, layer, nseg, K, pair;
    int seg_rows[3];
    const uint8_t *w[3];
    const float *s[3];
};
__device__ __forceinline__ OpInfo op_info(const MegaParams &P, int op) {
    OpInfo o;
    if (op >= P.layers * 4) {
        o.type = OP_LMHEAD, o.layer = P.layers, o.nseg = 1, o.K = P.E, o.pair = 0;
        o.seg_rows[0] = P.vocab, o.w[0] = P.lm_w, o.s[0] = P.lm_s;
        return o;
    }
    o.layer = op >> 2, o.type = op & 3, o.pair = 0;
    const MegaLayer &L = P.lw[o.layer];
    switch (o.type) {
        case OP_QKV:
            o.nseg = 3, o.K = P.E;
            o.seg_rows[0] = P.attn_seg, o.seg_rows[1] = P.kv_seg, o.seg_rows[2] = P.kv_seg;
            o.w[0] = L.w[MW_Q], o.w[1] = L.w[MW_K], o.w[2] = L.w[MW_V];
            o.s[0] = L.s[MW_Q], o.s[1] = L.s[MW_K], o.s[2] = L.s[MW_V];
            break;
        case OP_O:
            o.nseg = 1, o.K = P.attn_seg, o.seg_rows[0] = P.E, o.w[0] = L.w[MW_O], o.s[0] = L.s[MW_O];
            break;
        case OP_GATEUP: // seg0 = gate rows, w[1]/s[1] = up rows of the same pair
            o.nseg = 1, o.K = P.E, o.pair = 1, o.seg_rows[0] = P.H;
            o.w[0] = L.w[MW_GATE], o.s[0] = L.s[MW_GATE], o.w[1] = L.w[MW_UP], o.s[1] = L.s[MW_UP];
            break;
        default:
            o.nseg = 1, o.K = P.H, o.seg_rows[0] = P.E, o.w[0] = L.w[MW_DOWN], o.s[0] = L.s[MW_DOWN];
            break;
    }
    return o;
}

struct Sched {
    int op, seg, rg, c0;
};
struct Stage {
    int op, seg, row0, nrows, c0, ncols, K, pair, last_slice, type;
    const uint8_t *w0, *w1;
    const float *s0, *s1;
    uint32_t bytes;
};
// Normalise the cursor to the next non-empty stage of this CTA; false at the end of the schedule.
__device__ bool sched_get(const MegaParams &P, Sched &s, Stage &d, int limit_op = 1 << 30) {
    const int G = gridDim.x, cta = blockIdx.x;
    const int n_ops = P.layers * 4 + 1;
    while (s.op < n_ops && s.op < limit_op) {
        const OpInfo oi = op_info(P, s.op);
        long long T = 0;
        for (int i = 0; i < oi.nseg; i++) T += oi.seg_rows[i];
        const int a = (int)((T * cta) / G), b = (int)((T * (cta + 1)) / G);
        int seg_start = 0;
        for (int i = 0; i < s.seg; i++) seg_start += oi.seg_rows[i];
        while (s.seg < oi.nseg) {
            const int seg_end = seg_start + oi.seg_rows[s.seg];
            const int p0 = max(a, seg_start) - seg_start, p1 = min(b, seg_end) - seg_start;
            if (s.rg < p0) s.rg = p0;
            if (s.rg < p1) {
                const int slice = oi.pair ? MG_PSLICE : MG_SLICE;
                d.op = s.op, d.seg = s.seg, d.row0 = s.rg, d.nrows = min(MG_ROWS, p1 - s.rg);
                d.c0 = s.c0, d.ncols = min(slice, oi.K - s.c0), d.K = oi.K, d.pair = oi.pair, d.type = oi.type;
                d.last_slice = (s.c0 + slice >= oi.K);
                d.w0 = oi.w[s.seg], d.s0 = oi.s[s.seg];
                d.w1 = oi.pair ? oi.w[1] : nullptr, d.s1 = oi.pair ? oi.s[1] : nullptr;
                d.bytes = (uint32_t)d.nrows * (uint32_t)(d.ncols / 2 + d.ncols / 8) * (oi.pair ? 2u : 1u);
                return true;
            }
            seg_start = seg_end;
            s.seg++, s.rg = 0, s.c0 = 0;
        }
        s.op++, s.seg = 0, s.rg = 0, s.c0 = 0;
    }
    return false;
}
__device__ __forceinline__ void sched_advance(Sched &s, const Stage &d) {
    s.c0 += d.pair ? MG_PSLICE : MG_SLICE;
    if (s.c0 >= d.K) s.c0 = 0, s.rg += MG_ROWS;
}

// ---- producer ----------------------------------------------------------------------------------------------------------
// slot layout: normal stage: row r nibbles at r*2048, scales at r*512 bytes;
//              pair stage:   gate r at r*2048 / r*512, up r at r*2048+1024 / r*512+256.
template <bool PREFETCH>
__device__ __forceinline__ void issue_stage(const Stage &d, unsigned char *slot, uint64_t *full, int lane) {
    const int ncopies = d.nrows * (d.pair ? 2 : 1);
    const uint32_t nb = d.ncols / 2, sb = d.ncols / 8;
    for (int i = lane; i < ncopies; i += 32) {
        const int r = d.pair ? (i >> 1) : i, up = d.pair ? (i & 1) : 0;
        const uint8_t *w = up ? d.w1 : d.w0;
        const float *s = up ? d.s1 : d.s0;
        const size_t row = (size_t)(d.row0 + r);
        const uint8_t *src_n = w + row * (size_t)(d.K / 2) + d.c0 / 2;
        const float *src_s = s + row * (size_t)(d.K / 32) + d.c0 / 32;
        if (PREFETCH) {
            l2_prefetch(src_n, nb);
            l2_prefetch(src_s, sb);
        } else {
            tma_load_1d(slot + r * 2048 + up * 1024, src_n, nb, full);
            tma_load_1d(slot + MG_STAGE_NIB + r * 512 + up * 256, src_s, sb, full);
        }
    }
}

template <int NSTAGE>
__device__ void producer_loop(const MegaParams &P, unsigned char *ring, uint64_t *full, uint64_t *empty, int lane) {
    Sched rs = {0, 0, 0, 0}, ls = {0, 0, 0, 0};
    Stage rd, ld;
    bool rvalid = sched_get(P, rs, rd), lvalid = sched_get(P, ls, ld);
    unsigned it = 0;   // stages put into the ring
    unsigned lit = 0;  // stages prefetched into L2 (>= it)
    while (rvalid) {
        const unsigned slot = it % NSTAGE, use = it / NSTAGE;
        const bool free_slot = __shfl_sync(0xffffffffu, lane == 0 ? (int)mbar_try_wait(&empty[slot], (use & 1) ^ 1) : 0, 0) != 0;
        if (free_slot) {
            if (lane == 0) mbar_expect_tx(&full[slot], rd.bytes);
            __syncwarp();
            issue_stage<false>(rd, ring + (size_t)slot * MG_STAGE_BYTES, &full[slot], lane);
            sched_advance(rs, rd);
            rvalid = sched_get(P, rs, rd);
            it++;
            if (lit < it) { // keep the L2 cursor at or ahead of the ring cursor
                sched_advance(ls, ld);
                lvalid = sched_get(P, ls, ld);
                lit = it;
            }
        } else if (lvalid && lit < it + MG_L2_AHEAD) {
            issue_stage<true>(ld, nullptr, nullptr, lane);
            sched_advance(ls, ld);
            lvalid = sched_get(P, ls, ld);
            lit++;
        } else {
            __nanosleep(100);
        }
    }
}

// ---- cross-CTA ordering ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void op_signal(unsigned *cnt) { // all consumer threads call
    consumer_bar();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(cnt, 1u);
    }
}
__device__ __forceinline__ void op_wait(const unsigned *cnt, unsigned expected) { // all consumer threads call
    if (threadIdx.x == 0) {
        while (ld_acquire(cnt) < expected) {
        }
        __threadfence();
    }
    consumer_bar();
}

// ---- activation prologues (512 consumer threads) ------------------------------------------------------------------------
struct ActView {
    int8_t *aq;   // [MM][2][nblk][16]
    float *asc;   // [MM][nblk]
    int *asum;    // [MM][nblk]
    float4 *af4;  // f32 layout [MM][8][nblk]
};
template <int MM>
__device__ __forceinline__ ActView act_view(unsigned char *acts, int nblk) {
    ActView v;
    v.aq = (int8_t *)acts;
    v.asc = (float *)(acts + (size_t)MM * nblk * 32);
    v.asum = (int *)(acts + (size_t)MM * nblk * 32 + (size_t)MM * nblk * 4);
    v.af4 = (float4 *)acts;
    return v;
}

__device__ __forceinline__ float4 ldcg4(const float *p) { return __ldcg((const float4 *)p); }

// src: [M, ld] f32 in global (written by other CTAs -> read through L2).  norm_w == nullptr: no RMSNorm.
template <int MM, bool ACTQ8>
__device__ void stage_acts(const MegaParams &P, unsigned char *acts, const float *src, int ld, int K, const void *norm_w,
                           int norm_dt, double *red /*[MM][16]*/, float *rs /*[MM]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = K / 32;
    ActView av = act_view<MM>(acts, nblk);
    if (norm_w) { // RMSNorm.java:41-52
        double ss[MM];
#pragma unroll
        for (int m = 0; m < MM; m++) ss[m] = 0.0;
        for (int i4 = tid; i4 < K / 4; i4 += MG_CONSUMERS) {
            float4 v[MM];
#pragma unroll
            for (int m = 0; m < MM; m++) v[m] = m < P.M ? ldcg4(src + (size_t)m * ld + i4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int m = 0; m < MM; m++) {
                ss[m] += (double)__fmul_rn(v[m].x, v[m].x);
                ss[m] += (double)__fmul_rn(v[m].y, v[m].y);
                ss[m] += (double)__fmul_rn(v[m].z, v[m].z);
                ss[m] += (double)__fmul_rn(v[m].w, v[m].w);
            }
        }
#pragma unroll
        for (int m = 0; m < MM; m++) {
            ss[m] = warp_sum_d(ss[m]);
            if (lane == 0) red[m * MG_CWARPS + warp] = ss[m];
        }
        consumer_bar();
        if (tid < MM) {
            double t = 0;
            for (int w = 0; w < MG_CWARPS; w++) t += red[tid * MG_CWARPS + w];
            t /= (double)P.E;
            t += (double)P.eps;
            rs[tid] = (float)(1.0 / sqrt(t));
        }
        consumer_bar();
    }
    for (int idx = tid; idx < MM * nblk; idx += MG_CONSUMERS) {
        const int m = idx / nblk, b = idx - m * nblk;
        float v[32];
        if (m < P.M) {
            const float *sp = src + (size_t)m * ld + b * 32;
            float4 x4[8];
#pragma unroll
            for (int i = 0; i < 8; i++) x4[i] = ldcg4(sp + i * 4);
#pragma unroll
            for (int i = 0; i < 8; i++) v[i * 4] = x4[i].x, v[i * 4 + 1] = x4[i].y, v[i * 4 + 2] = x4[i].z, v[i * 4 + 3] = x4[i].w;
            if (norm_w) {
                const float rsf = rs[m];
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const float w = norm_dt == JL_BF16 ? bf16_bits_to_f32(((const uint16_t *)norm_w)[b * 32 + i])
                                                       : ((const float *)norm_w)[b * 32 + i];
                    v[i] = __fmul_rn(__fadd_rn(0.0f, w), __fmul_rn(rsf, v[i]));
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = 0.0f;
        }
        if (ACTQ8) { // PanamaTensorOperations.java:1696-1710
            float mx = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; i++) mx = fmaxf(mx, fabsf(v[i]));
            const float d = __fdiv_rn(mx, 127.0f);
            const float id = mx != 0.0f ? __fdiv_rn(127.0f, mx) : 0.0f;
            uint32_t w[8];
            int sum = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const int q0 = (int)__fadd_rn(__fmul_rn(v[i * 4], id), 0.5f);
                const int q1 = (int)__fadd_rn(__fmul_rn(v[i * 4 + 1], id), 0.5f);
                const int q2 = (int)__fadd_rn(__fmul_rn(v[i * 4 + 2], id), 0.5f);
                const int q3 = (int)__fadd_rn(__fmul_rn(v[i * 4 + 3], id), 0.5f);
                sum += q0 + q1 + q2 + q3;
                w[i] = (uint32_t)(q0 & 0xFF) | ((uint32_t)(q1 & 0xFF) << 8) | ((uint32_t)(q2 & 0xFF) << 16) |
                       ((uint32_t)(q3 & 0xFF) << 24);
            }
            *(uint4 *)(av.aq + (((size_t)m * 2 + 0) * nblk + b) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            *(uint4 *)(av.aq + (((size_t)m * 2 + 1) * nblk + b) * 16) = make_uint4(w[4], w[5], w[6], w[7]);
            av.asc[m * nblk + b] = d;
            av.asum[m * nblk + b] = sum;
        } else {
#pragma unroll
            for (int c4 = 0; c4 < 8; c4++)
                av.af4[((size_t)m * 8 + c4) * nblk + b] = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
        }
    }
    consumer_bar();
}

// ---- consumer math on one ring stage -----------------------------------------------------------------------------------
// Q8 activations x Q4 weights; acc[0] = row (or gate), acc[1] = up (pair stages)
template <int MM>
__device__ __forceinline__ void consume_q8(const Stage &d, const unsigned char *slot, const ActView &av, int nblk_total,
                                           float (&acc)[2][MM], int warp, int lane) {
    if (warp >= d.nrows) return;
    const int nb = d.ncols / 32, blk0 = d.c0 / 32;
    const int nsub = d.pair ? 2 : 1;
#pragma unroll
    for (int u = 0; u < 2; u++) {
        if (u >= nsub) break;
        const unsigned char *nib = slot + warp * 2048 + u * 1024;
        const float *sc = (const float *)(slot + MG_STAGE_NIB + warp * 512 + u * 256);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int b = lane + 32 * j;
            if (b >= nb) break;
            const uint4 q = *(const uint4 *)(nib + b * 16);
            const float sb = sc[b];
            const int gb = blk0 + b;
#pragma unroll
            for (int m = 0; m < MM; m++) {
                const uint4 alo = *(const uint4 *)(av.aq + (((size_t)m * 2 + 0) * nblk_total + gb) * 16);
                const uint4 ahi = *(const uint4 *)(av.aq + (((size_t)m * 2 + 1) * nblk_total + gb) * 16);
                int s = 0;
                s = __dp4a((int)(q.x & 0x0F0F0F0Fu), (int)alo.x, s);
                s = __dp4a((int)((q.x >> 4) & 0x0F0F0F0Fu), (int)ahi.x, s);
                s = __dp4a((int)(q.y & 0x0F0F0F0Fu), (int)alo.y, s);
                s = __dp4a((int)((q.y >> 4) & 0x0F0F0F0Fu), (int)ahi.y, s);
                s = __dp4a((int)(q.z & 0x0F0F0F0Fu), (int)alo.z, s);
                s = __dp4a((int)((q.z >> 4) & 0x0F0F0F0Fu), (int)ahi.z, s);
                s = __dp4a((int)(q.w & 0x0F0F0F0Fu), (int)alo.w, s);
                s = __dp4a((int)((q.w >> 4) & 0x0F0F0F0Fu), (int)ahi.w, s);
                s -= 8 * av.asum[m * nblk_total + gb];
                acc[u][m] = fmaf(__fmul_rn(av.asc[m * nblk_total + gb], sb), (float)s, acc[u][m]);
            }
        }
    }
}

// F32 activations x Q4 weights (lm_head; AbstractModel.java:444-449)
template <int MM>
__device__ __forceinline__ void consume_f32(const Stage &d, const unsigned char *slot, const ActView &av, int nblk_total,
                                            float (&acc)[2][MM], int warp, int lane) {
    if (warp >= d.nrows) return;
    const int nb = d.ncols / 32, blk0 = d.c0 / 32;
    const unsigned char *nib = slot + warp * 2048;
    const float *sc = (const float *)(slot + MG_STAGE_NIB + warp * 512);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int b = lane + 32 * j;
        if (b >= nb) break;
        const uint4 q = *(const uint4 *)(nib + b * 16);
        const float sb = sc[b];
        const int gb = blk0 + b;
        float wf[32];
        const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t lo = qw[i] & 0x0F0F0F0Fu, hi = (qw[i] >> 4) & 0x0F0F0F0Fu;
#pragma unroll
            for (int t = 0; t < 4; t++) {
                wf[i * 4 + t] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7540 | t)) - 8388616.0f;
                wf[16 + i * 4 + t] = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7540 | t)) - 8388616.0f;
            }
        }
#pragma unroll
        for (int m = 0; m < MM; m++) {
            float part = 0.0f;
#pragma unroll
            for (int c4 = 0; c4 < 8; c4++) {
                const float4 a4 = av.af4[((size_t)m * 8 + c4) * nblk_total + gb];
                part = fmaf(a4.x, wf[c4 * 4 + 0], part);
                part = fmaf(a4.y, wf[c4 * 4 + 1], part);
                part = fmaf(a4.z, wf[c4 * 4 + 2], part);
                part = fmaf(a4.w, wf[c4 * 4 + 3], part);
            }
            acc[0][m] = fmaf(sb, part, acc[0][m]);
        }
    }
}

// ---- attention task (all 512 consumer threads of one CTA) ---------------------------------------------------------------
__device__ __forceinline__ const char *mg_kv_row(const KvLayout &kv, int session, int layer, int pos, int which) {
    const int lp = layer / kv.layers_per_page, rl = layer % kv.layers_per_page;
    const int cp = pos / kv.ctx_per_page, rc = pos % kv.ctx_per_page;
    const char *base = (const char *)kv.page_table[((size_t)session * kv.n_layer_pages + lp) * kv.n_ctx_pages + cp];
    const size_t elem = (((size_t)rl * 2 + which) * kv.ctx_per_page + rc) * kv.kv_len;
    return base + elem * (kv.kv_dtype == JL_F32 ? 4 : 2);
}
__device__ __forceinline__ uint16_t mg_bf16(float n) {
    const uint32_t nbits = __float_as_uint(n);
    const uint32_t s = (nbits >> 16) & 0x8000u, e = (nbits >> 16) & 0x7f80u, m = nbits & 0x7fffffu;
    if (e != 0x7f80u) {
        const int mshift = (int)(m >> 16), masked = (int)(m & 0xffff), cmp = masked - 0x8000;
        const int m1 = cmp > 0 ? mshift + 1 : (cmp < 0 ? mshift : ((mshift & 1) ? mshift + 1 : mshift));
        return (uint16_t)(s | (e + (uint32_t)m1));
    }
    return m != 0 ? (uint16_t)0x7fc0 : (uint16_t)(nbits >> 16);
}

// One (row m, kv head, split) task.  smem `u` (sized by uarea_bytes) aliases the
// activation staging area, which is dead between the QKV stages and the o_proj prologue.
template <int HS>
__device__ void attention_task(const MegaParams &P, int layer, int m, int kvh, int split, unsigned char *u) {
    constexpr int C4 = HS / 4;
    constexpr int PARTS = MG_CONSUMERS / HS; // P.V position groups
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = P.heads / P.kv_heads;
    float4 *Ks = (float4 *)u;                    // [TILE][C4] swizzled
    float4 *Vs = Ks + MG_ATT_TILE * C4;          // [TILE][C4]
    float *qs = (float *)(Vs + MG_ATT_TILE * C4); // [group][HS] rotated queries
    float *ps = qs + MG_MAX_GROUP * HS;          // [TILE][MAX_GROUP]
    float *hm = ps + MG_ATT_TILE * MG_MAX_GROUP; // running max / sum / correction per head
    float *hl = hm + MG_MAX_GROUP;
    float *hc = hl + MG_MAX_GROUP;
    float *comb = hc + MG_MAX_GROUP + 8;         // [PARTS][MAX_GROUP][HS] P.V combine buffer

    const int session = P.sessions[m], pos = P.positions[m];
    const int n = pos + 1;
    const int S = P.splits;
    const int per = (((n + S - 1) / S) + MG_ATT_TILE - 1) / MG_ATT_TILE * MG_ATT_TILE;
    const int t0 = split * per, t1 = min(n, t0 + per);
    const int hp = HS / 2;
    const int h0 = kvh * group, xoff = kvh * HS, dt = P.kv.kv_dtype;
    const size_t poffset = (size_t)pos * hp;

    // RoPE on this group's queries (CausalSelfAttention.java:260-268: table index poffset + kvh_global*hs + j)
    for (int idx = tid; idx < group * hp; idx += MG_CONSUMERS) {
        const int h = idx / hp, j = idx % hp;
        const float2 f = ((const float2 *)P.rope)[poffset + (size_t)(P.kv_head0_global + kvh) * HS + j];
        const float *qr = P.q + (size_t)m * P.attn_seg + (h0 + h) * HS;
        const float q0 = __ldcg(qr + j), q1 = __ldcg(qr + j + hp);
        qs[h * HS + j] = __fsub_rn(__fmul_rn(q0, f.x), __fmul_rn(q1, f.y));
        qs[h * HS + j + hp] = __fadd_rn(__fmul_rn(q0, f.y), __fmul_rn(q1, f.x));
    }
    // the split that contains `pos` appends the rotated key and the value to the page (:230-243,279-285)
    if (pos >= t0 && pos < t1) {
        char *krow = (char *)mg_kv_row(P.kv, session, layer, pos, 0);
        char *vrow = (char *)mg_kv_row(P.kv, session, layer, pos, 1);
        for (int j = tid; j < hp; j += MG_CONSUMERS) {
            const float2 f = ((const float2 *)P.rope)[poffset + (size_t)(P.kv_head0_global + kvh) * HS + j];
            const float *kr = P.k + (size_t)m * P.kv_seg + xoff, *vr = P.v + (size_t)m * P.kv_seg + xoff;
            const float k0 = __ldcg(kr + j), k1 = __ldcg(kr + j + hp);
            const float r0 = __fsub_rn(__fmul_rn(k0, f.x), __fmul_rn(k1, f.y));
            const float r1 = __fadd_rn(__fmul_rn(k0, f.y), __fmul_rn(k1, f.x));
            const float v0 = __ldcg(vr + j), v1 = __ldcg(vr + j + hp);
            if (dt == JL_F32) {
                ((float *)krow)[xoff + j] = r0, ((float *)krow)[xoff + j + hp] = r1;
                ((float *)vrow)[xoff + j] = v0, ((float *)vrow)[xoff + j + hp] = v1;
            } else {
                ((uint16_t *)krow)[xoff + j] = mg_bf16(r0), ((uint16_t *)krow)[xoff + j + hp] = mg_bf16(r1);
                ((uint16_t *)vrow)[xoff + j] = mg_bf16(v0), ((uint16_t *)vrow)[xoff + j + hp] = mg_bf16(v1);
            }
        }
    }
    if (tid < group) hm[tid] = -INFINITY, hl[tid] = 0.0f;
    float acc[MG_MAX_GROUP];
#pragma unroll
    for (int h = 0; h < MG_MAX_GROUP; h++) acc[h] = 0.0f;
    const int part = tid / HS, d = tid % HS;
    consumer_bar();

    for (int tb = t0; tb < t1; tb += MG_ATT_TILE) {
        const int cnt = min(MG_ATT_TILE, t1 - tb);
        // stage K and V rows of the tile (coalesced 128-bit loads through L2)
        for (int f = tid; f < cnt * C4; f += MG_CONSUMERS) {
            const int r = f / C4, c4 = f % C4;
            const char *kr = mg_kv_row(P.kv, session, layer, tb + r, 0);
            const char *vr = mg_kv_row(P.kv, session, layer, tb + r, 1);
            float4 k4, v4;
            if (dt == JL_F32) {
                k4 = __ldcg((const float4 *)((const float *)kr + xoff + c4 * 4));
                v4 = __ldcg((const float4 *)((const float *)vr + xoff + c4 * 4));
            } else {
                const uint2 uk = __ldcg((const uint2 *)((const uint16_t *)kr + xoff + c4 * 4));
                const uint2 uv = __ldcg((const uint2 *)((const uint16_t *)vr + xoff + c4 * 4));
                k4 = make_float4(__uint_as_float(uk.x << 16), __uint_as_float(uk.x & 0xffff0000u), __uint_as_float(uk.y << 16),
                                 __uint_as_float(uk.y & 0xffff0000u));
                v4 = make_float4(__uint_as_float(uv.x << 16), __uint_as_float(uv.x & 0xffff0000u), __uint_as_float(uv.y << 16),
                                 __uint_as_float(uv.y & 0xffff0000u));
            }
            Ks[r * C4 + (c4 ^ (r & 7))] = k4;
            Vs[r * C4 + c4] = v4;
        }
        consumer_bar();
        // scores (batchDotProduct :324-330, scale :332)
        for (int idx = tid; idx < group * MG_ATT_TILE; idx += MG_CONSUMERS) {
            const int h = idx / MG_ATT_TILE, t = idx % MG_ATT_TILE;
            float s = -INFINITY;
            if (t < cnt) {
                float a = 0.0f;
                const float4 *q4 = (const float4 *)(qs + h * HS);
#pragma unroll 8
                for (int c4 = 0; c4 < C4; c4++) {
                    const float4 k4 = Ks[t * C4 + (c4 ^ (t & 7))];
                    const float4 qq = q4[c4];
                    a = fmaf(qq.x, k4.x, a);
                    a = fmaf(qq.y, k4.y, a);
                    a = fmaf(qq.z, k4.z, a);
                    a = fmaf(qq.w, k4.w, a);
                }
                s = __fmul_rn(a, P.attn_scale);
            }
            ps[t * MG_MAX_GROUP + h] = s;
        }
        consumer_bar();
        // online softmax: warp h owns head h (TILE == 32: one score per lane)
        if (warp < group) {
            const int h = warp;
            const float s0 = ps[lane * MG_MAX_GROUP + h];
            const float m_old = hm[h];
            const float m_new = fmaxf(m_old, warp_max(s0));
            const float e0 = s0 == -INFINITY ? 0.0f : (float)exp((double)__fsub_rn(s0, m_new));
            const float ts = warp_sum(e0);
            ps[lane * MG_MAX_GROUP + h] = e0;
            if (lane == 0) {
                const float corr = m_old == -INFINITY ? 0.0f : (float)exp((double)__fsub_rn(m_old, m_new));
                hc[h] = corr, hm[h] = m_new, hl[h] = fmaf(hl[h], corr, ts);
            }
        }
        consumer_bar();
        // P.V: thread (part, d) accumulates positions t = part, part+PARTS, ...
#pragma unroll
        for (int h = 0; h < MG_MAX_GROUP; h++)
            if (h < group) acc[h] *= hc[h];
        for (int t = part; t < cnt; t += PARTS) {
            const float v = ((const float *)Vs)[t * HS + d];
#pragma unroll
            for (int h = 0; h < MG_MAX_GROUP; h++)
                if (h < group) acc[h] = fmaf(v, ps[t * MG_MAX_GROUP + h], acc[h]);
        }
        consumer_bar(); // tile fully consumed before the next one overwrites Ks/Vs/ps
    }
    // combine the PARTS partial sums
    for (int h = 0; h < group; h++) comb[((size_t)part * MG_MAX_GROUP + h) * HS + d] = acc[h];
    consumer_bar();
    if (part == 0) {
        for (int h = 0; h < group; h++) {
            float a = 0.0f;
            for (int pp = 0; pp < PARTS; pp++) a += comb[((size_t)pp * MG_MAX_GROUP + h) * HS + d];
            if (S == 1) {
                P.att[(size_t)m * P.attn_seg + (h0 + h) * HS + d] = t0 < t1 ? __fdiv_rn(a, hl[h]) : 0.0f;
            } else {
                float *w = P.attn_ws + (((size_t)m * P.heads + h0 + h) * S + split) * (HS + 2);
                w[d] = a;
                if (d == 0) w[HS] = hm[h], w[HS + 1] = hl[h];
            }
        }
    }
}

// merge the split partials of one (row, kv head): out = sum_s acc_s*exp(m_s-M) / sum_s l_s*exp(m_s-M)
template <int HS>
__device__ void attention_merge(const MegaParams &P, int m, int kvh) {
    const int group = P.heads / P.kv_heads, S = P.splits;
    for (int idx = threadIdx.x; idx < group * HS; idx += MG_CONSUMERS) {
        const int h = kvh * group + idx / HS, d = idx % HS;
        const float *w = P.attn_ws + ((size_t)m * P.heads + h) * S * (HS + 2);
        float M = -INFINITY;
        for (int s = 0; s < S; s++) M = fmaxf(M, __ldcg(w + s * (HS + 2) + HS));
        float num = 0.0f, den = 0.0f;
        for (int s = 0; s < S; s++) {
            const float ms = __ldcg(w + s * (HS + 2) + HS);
            if (ms == -INFINITY) continue;
            const float f = (float)exp((double)__fsub_rn(ms, M));
            num = fmaf(__ldcg(w + s * (HS + 2) + d), f, num);
            den = fmaf(__ldcg(w + s * (HS + 2) + HS + 1), f, den);
        }
        P.att[(size_t)m * P.attn_seg + h * HS + d] = __fdiv_rn(num, den);
    }
}

// ---- the kernel ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long pack_arg(float v, int idx) {
    if (!(v == v)) return 0ull; // NaN never wins (AbstractModel.java:465 'v > maxv' is false)
    uint32_t b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}

template <int MM, int NSTAGE>
__global__ void __launch_bounds__(MG_THREADS, 1) mega_decode_kernel(const MegaParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *ring = smem;
    unsigned char *uarea = smem + (size_t)NSTAGE * MG_STAGE_BYTES;
    __shared__ uint64_t full[NSTAGE], empty[NSTAGE];
    __shared__ double red[MM * MG_CWARPS];
    __shared__ float rs[MM];
    __shared__ unsigned long long wbest[MM][MG_CWARPS];
    __shared__ int s_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], MG_CWARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == MG_CWARPS) { // ===== producer warp =====
        producer_loop<NSTAGE>(P, ring, full, empty, lane);
        return;
    }

    // ===== consumers =====
    unsigned *sync = P.sync;
    const int n_ops = P.layers * 4 + 1;
    // embedding lookup (LlamaModel.java:68-100): columns split over the grid
    {
        const int c_a = (int)(((long long)P.E * cta) / G), c_b = (int)(((long long)P.E * (cta + 1)) / G);
        for (int m = 0; m < P.M; m++) {
            const size_t tok = (size_t)P.tokens[m];
            for (int c = c_a + tid; c < c_b; c += MG_CONSUMERS) {
                float v;
                if (P.embed_dt == JL_F32) v = ((const float *)P.embed_w)[tok * P.E + c];
                else if (P.embed_dt == JL_BF16) v = bf16_bits_to_f32(((const uint16_t *)P.embed_w)[tok * P.E + c]);
                else if (P.embed_dt == JL_Q4) {
                    const int blk = c / 32, in = c % 32;
                    const uint8_t byte = ((const uint8_t *)P.embed_w)[(tok * P.E + blk * 32) / 2 + (in & 15)];
                    const int nib = in < 16 ? (byte & 0x0F) : (byte >> 4);
                    v = __fmul_rn((float)(nib - 8), P.embed_s[tok * (P.E / 32) + blk]);
                } else {
                    v = __fmul_rn((float)((const int8_t *)P.embed_w)[tok * P.E + c], P.embed_s[tok * (P.E / 32) + c / 32]);
                }
                P.x[(size_t)m * P.E + c] = v;
            }
        }
        op_signal(&sync[0]);
    }

    Sched s = {0, 0, 0, 0};
    Stage d;
    unsigned cit = 0;
    float best_v[MM];
    int best_i[MM];
#pragma unroll
    for (int m = 0; m < MM; m++) best_v[m] = -INFINITY, best_i[m] = 0x7fffffff;

    for (int op = 0; op < n_ops; op++) {
        const OpInfo oi = op_info(P, op);
        const int L = oi.layer;
        unsigned *cnt = &sync[1 + (op < P.layers * 4 ? L * 5 + (oi.type == OP_QKV ? 0 : oi.type == OP_O ? 2 : oi.type == OP_GATEUP ? 3 : 4)
                                                      : P.layers * 5)];
        // ---- attention phase sits between QKV and O ----
        if (oi.type == OP_O) {
            const int ntasks = P.M * P.kv_heads * P.splits;
            if (cta < ntasks) {
                const int split = cta % P.splits, kvh = (cta / P.splits) % P.kv_heads, m = cta / (P.splits * P.kv_heads);
                op_wait(&sync[1 + L * 5 + 0], (unsigned)G);
                switch (P.head_size) {
                    case 32: attention_task<32>(P, L, m, kvh, split, uarea); break;
                    case 64: attention_task<64>(P, L, m, kvh, split, uarea); break;
                    default: attention_task<128>(P, L, m, kvh, split, uarea); break;
                }
                if (P.splits > 1) {
                    consumer_bar();
                    if (tid == 0) {
                        __threadfence();
                        const unsigned old = atomicAdd(&P.att_done[(size_t)L * P.M * P.kv_heads + m * P.kv_heads + kvh], 1u);
                        s_last = (old == (unsigned)P.splits - 1);
                        __threadfence();
                    }
                    consumer_bar();
                    if (s_last) {
                        switch (P.head_size) {
                            case 32: attention_merge<32>(P, m, kvh); break;
                            case 64: attention_merge<64>(P, m, kvh); break;
                            default: attention_merge<128>(P, m, kvh); break;
                        }
                        op_signal(&sync[1 + L * 5 + 1]);
                    }
                } else {
                    op_signal(&sync[1 + L * 5 + 1]);
                }
            }
        }
        // ---- does this CTA own rows of the op? ----
        Sched probe = s;
        const bool has = sched_get(P, probe, d, op + 1) && d.op == op;
        if (has) {
            // dependency + prologue
            const int K = oi.K, nblk = K / 32;
            ActView av;
            if (oi.type == OP_QKV) {
                op_wait(L == 0 ? &sync[0] : &sync[1 + (L - 1) * 5 + 4], (unsigned)G);
                stage_acts<MM, true>(P, uarea, P.x, P.E, K, P.lw[L].attn_norm, P.lw[L].attn_norm_dt, red, rs);
            } else if (oi.type == OP_O) {
                op_wait(&sync[1 + L * 5 + 1], (unsigned)(P.M * P.kv_heads));
                stage_acts<MM, true>(P, uarea, P.att, P.attn_seg, K, nullptr, 0, red, rs);
            } else if (oi.type == OP_GATEUP) {
                op_wait(&sync[1 + L * 5 + 2], (unsigned)G);
                stage_acts<MM, true>(P, uarea, P.xb, P.E, K, P.lw[L].ffn_norm, P.lw[L].ffn_norm_dt, red, rs);
            } else if (oi.type == OP_DOWN) {
                op_wait(&sync[1 + L * 5 + 3], (unsigned)G);
                stage_acts<MM, true>(P, uarea, P.h, P.H, K, nullptr, 0, red, rs);
            } else {
                op_wait(&sync[1 + (P.layers - 1) * 5 + 4], (unsigned)G);
                stage_acts<MM, false>(P, uarea, P.x, P.E, K, P.out_norm, P.out_norm_dt, red, rs);
            }
            av = act_view<MM>(uarea, nblk);

            float acc[2][MM];
#pragma unroll
            for (int m = 0; m < MM; m++) acc[0][m] = 0.0f, acc[1][m] = 0.0f;
            while (sched_get(P, s, d, op + 1) && d.op == op) {
                const unsigned slot = cit % NSTAGE, use = cit / NSTAGE;
                mbar_wait(&full[slot], use & 1);
                const unsigned char *sp = ring + (size_t)slot * MG_STAGE_BYTES;
                if (oi.type == OP_LMHEAD) consume_f32<MM>(d, sp, av, nblk, acc, warp, lane);
                else consume_q8<MM>(d, sp, av, nblk, acc, warp, lane);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
                cit++;
                if (d.last_slice) {
                    if (warp < d.nrows) {
                        const int row = d.row0 + warp;
#pragma unroll
                        for (int m = 0; m < MM; m++) {
                            acc[0][m] = warp_sum(acc[0][m]);
                            if (oi.pair) acc[1][m] = warp_sum(acc[1][m]);
                        }
                        if (lane == 0) {
#pragma unroll
                            for (int m = 0; m < MM; m++) {
                                if (m >= P.M) continue;
                                const float v = acc[0][m];
                                switch (oi.type) {
                                    case OP_QKV: {
                                        float *out = d.seg == 0 ? P.q + (size_t)m * P.attn_seg
                                                                : (d.seg == 1 ? P.k : P.v) + (size_t)m * P.kv_seg;
                                        out[row] = v;
                                    } break;
                                    case OP_O: // TransformerBlock.java:185
                                        P.xb[(size_t)m * P.E + row] = __fadd_rn(v, __ldcg(P.x + (size_t)m * P.E + row));
                                        break;
                                    case OP_GATEUP: // MLPBlock.java:132-141
                                        P.h[(size_t)m * P.H + row] = __fmul_rn(silu_ref(v), acc[1][m]);
                                        break;
                                    case OP_DOWN: // TransformerBlock.java:203
                                        P.x[(size_t)m * P.E + row] = __fadd_rn(v, __ldcg(P.xb + (size_t)m * P.E + row));
                                        break;
                                    default: // logits + running arg-max (strict '>', lowest index wins)
                                        P.logits[(size_t)m * P.vocab + row] = v;
                                        if (v > best_v[m] || (v == best_v[m] && row < best_i[m])) best_v[m] = v, best_i[m] = row;
                                        break;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int m = 0; m < MM; m++) acc[0][m] = 0.0f, acc[1][m] = 0.0f;
                }
                sched_advance(s, d);
            }
        } else {
            s = probe; // nothing of this op here: cursor already skipped past it
        }
        if (oi.type == OP_LMHEAD) {
            // CTA-level arg-max, published as one packed 64-bit candidate per row
            if (lane == 0)
                for (int m = 0; m < MM; m++) wbest[m][warp] = best_i[m] == 0x7fffffff ? 0ull : pack_arg(best_v[m], best_i[m]);
            consumer_bar();
            if (tid < P.M) {
                unsigned long long b = 0ull;
                for (int w = 0; w < MG_CWARPS; w++) b = wbest[tid][w] > b ? wbest[tid][w] : b;
                P.argmax_slots[(size_t)tid * G + cta] = b;
            }
        }
        op_signal(cnt);
    }

    // ---- final arg-max across CTAs + device-side feedback for the resident loop ----
    if (cta == 0) {
        op_wait(&sync[1 + P.layers * 5], (unsigned)G);
        if (warp < P.M) {
            const int m = warp;
            unsigned long long b = 0ull;
            for (int i = lane; i < G; i += 32) {
                const unsigned long long c = __ldcg(&P.argmax_slots[(size_t)m * G + i]);
                b = c > b ? c : b;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long c = __shfl_xor_sync(0xffffffffu, b, o);
                b = c > b ? c : b;
            }
            if (lane == 0) {
                const int tok = b == 0ull ? 0 : (int)(0xFFFFFFFFu - (uint32_t)(b & 0xFFFFFFFFull));
                P.next[m] = tok;
                if (P.resident) {
                    const int cnt = *P.counter;
                    P.tokens[m] = tok;
                    P.positions[m] += 1;
                    if (cnt * P.M + m < P.hist_cap) P.hist[cnt * P.M + m] = tok;
                }
            }
        }
        consumer_bar();
        if (tid == 0 && P.resident) *P.counter = *P.counter + 1;
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------------
size_t jl_mega_sync_words(int layers) { return (size_t)layers * 5 + 3; }

static size_t uarea_bytes(const MegaParams &p, int MM) {
    int Kmax = p.E;
    if (p.H > Kmax) Kmax = p.H;
    if (p.attn_seg > Kmax) Kmax = p.attn_seg;
    size_t acts = (size_t)MM * (Kmax / 32) * 40;
    size_t lm = (size_t)MM * p.E * 4;
    if (lm > acts) acts = lm;
    const int hs = p.head_size;
    size_t att = (size_t)2 * MG_ATT_TILE * hs * 4 + (size_t)MG_MAX_GROUP * hs * 4 + (size_t)MG_ATT_TILE * MG_MAX_GROUP * 4 + 256;
    att += (size_t)(MG_CONSUMERS / hs) * MG_MAX_GROUP * hs * 4; // P.V combine buffer
    return (acts > att ? acts : att) + 256;
}

bool jl_mega_supported(const MegaParams &p) {
    if (p.M < 1 || p.M > MEGA_MAX_M) return false;
    if (p.head_size != 32 && p.head_size != 64 && p.head_size != 128) return false;
    if (p.heads % p.kv_heads || p.heads / p.kv_heads > MG_MAX_GROUP) return false;
    if ((p.E % 128) || (p.H % 128) || (p.attn_seg % 128)) return false; // TMA: 16-byte aligned scale slices
    const int MM = p.M <= 1 ? 1 : (p.M <= 2 ? 2 : 4);
    const int nstage = MM == 1 ? 4 : 3;
    return (size_t)nstage * MG_STAGE_BYTES + uarea_bytes(p, MM) <= 225 * 1024;
}

template <int MM, int NSTAGE>
static int launch_mega(jl_ctx *ctx, cudaStream_t stream, const MegaParams &p) {
    auto kern = mega_decode_kernel<MM, NSTAGE>;
    const size_t smem = (size_t)NSTAGE * MG_STAGE_BYTES + uarea_bytes(p, MM);
    static size_t configured = 0;
    if (smem > configured) {
        JL_CUDA_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctx->sm_count);
    cfg.blockDim = dim3(MG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; // all CTAs co-resident: the spin waits cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    JL_CUDA_CHECK(ctx, cudaLaunchKernelEx(&cfg, kern, p));
    ctx->launches++;
    return JL_OK;
}

int jl_launch_mega(jl_ctx *ctx, cudaStream_t stream, const MegaParams &p) {
    if (!jl_mega_supported(p)) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "megakernel: unsupported shape");
    const size_t words = jl_mega_sync_words(p.layers);
    JL_CUDA_CHECK(ctx, cudaMemsetAsync(p.sync, 0, words * sizeof(unsigned), stream));
    if (p.splits > 1)
        JL_CUDA_CHECK(ctx, cudaMemsetAsync(p.att_done, 0, (size_t)p.layers * p.M * p.kv_heads * sizeof(unsigned), stream));
    if (p.M <= 1) return launch_mega<1, 4>(ctx, stream, p);
    if (p.M <= 2) return launch_mega<2, 3>(ctx, stream, p);
    return launch_mega<4, 3>(ctx, stream, p);
}
