// Persistent cooperative decode megakernel: one launch = one decoded token for M <= 4 sessions.
//
// It executes the whole of AbstractModel.forward + sample (core/model/AbstractModel.java:314-329,443-473;
// TransformerBlock.java:158-215; CausalSelfAttention.java:145-385; MLPBlock.java:106-166) in a single grid of
// one CTA per SM.  Why: at batch 1 every Llama-3-8B GEMV is 2-11 us of HBM streaming, so per-kernel launch,
// prologue and drain latencies (measured ~10 us per launch) dominate a kernel-per-op design.
//
// Data path ("L2-staged"):
//   * every CTA owns a fixed, contiguous row range of every weight matrix.  20 CONSUMER warps run the same
//     inner loop as jl_gemv.cu: 128-bit ld.global.nc loads straight into registers, four row-slices in flight
//     per warp, dp4a against Q8 activations staged in shared memory, warp-shuffle reduction.  The first
//     row-slices of an op are requested BEFORE the op's dependency is waited for.
//   * one HELPER warp walks the same schedule two ops ahead and pulls the CTA's rows of those ops into L2 with
//     cp.async.bulk.prefetch.L2, so HBM streams continuously (through attention, barriers and prologues) and
//     the consumers mostly hit L2; the 126 MB L2 is the staging buffer instead of a shared-memory ring.
//   * ops are ordered by per-op completion counters in global memory (red.release / ld.acquire), not by
//     kernel boundaries; RMSNorm + Q8 quantisation are recomputed per CTA in the op prologue (one L2 round
//     trip); K-partitions of a row and the gate/up halves meet in a shared-memory buffer and the op-end pass
//     applies the fused epilogues (residual add, SiLU*up) for all rows of the CTA at once.
//   * attention (RoPE, KV append, scores, softmax, P.V over the paged KV cache) runs as (row, kv-head, split)
//     tasks on the first CTAs.
//
// Arithmetic is the same as the kernel-per-op path (jl_gemv.cu / jl_attention.cu) up to summation order of the
// K-partitions; the launch is cooperative so the spin waits cannot deadlock.
// (An earlier TMA + shared-memory-ring variant streamed at 5.8 TB/s raw but was consumer-latency bound with
// 8-16 consumer warps; see git history and tools/micro/stream_bench.cu.)
#include "jl_mega.cuh"
#include "jl_attn_task.cuh"

#define MG_CWARPS 20
#define MG_CONSUMERS (MG_CWARPS * 32)
#define MG_THREADS (MG_CONSUMERS + 32)
#define MG_SLICE 2048   // columns per row-slice: 64 blocks = two 16-byte blocks per lane
#define MG_MAXPARTS 8   // K-partitions of a row handled by different warps
#define MG_DEPTH 4      // row-slices in flight per warp
#define MG_EROWS 112    // max rows (or pairs) of one op owned by one CTA (lm_head excluded)
#define MG_LOOKAHEAD 2  // ops the helper warp prefetches ahead of the consumers
#define MG_MAX_LAYERS 64

__device__ __forceinline__ void l2_prefetch(const void *src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(MG_CONSUMERS) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

enum { OP_QKV = 0, OP_O, OP_GATEUP, OP_DOWN, OP_LMHEAD };

// ---- per-op description (layer pointers come from the shared-memory copy of the MegaLayer table) ----------------------
struct OpDesc {
    int type, L, K, pair, nseg, kparts, ipu; // ipu = row-slices per unit
    int seg_rows[3];
    const uint8_t *w[3];
    const float *s[3];
    int a, b; // this CTA's range in the concatenated row (pair) space
};
__device__ __forceinline__ OpDesc op_desc(const MegaParams &P, const MegaLayer *slw, int op, int cta, int G) {
    OpDesc o;
    o.pair = 0, o.nseg = 1, o.seg_rows[1] = o.seg_rows[2] = 0;
    if (op >= P.layers * 4) {
        o.type = OP_LMHEAD, o.L = P.layers, o.K = P.E, o.seg_rows[0] = P.vocab, o.w[0] = P.lm_w, o.s[0] = P.lm_s;
    } else {
        o.L = op >> 2, o.type = op & 3;
        const MegaLayer &L = slw[o.L];
        if (o.type == OP_QKV) {
            o.nseg = 3, o.K = P.E;
            o.seg_rows[0] = P.attn_seg, o.seg_rows[1] = P.kv_seg, o.seg_rows[2] = P.kv_seg;
            o.w[0] = L.w[MW_Q], o.w[1] = L.w[MW_K], o.w[2] = L.w[MW_V];
            o.s[0] = L.s[MW_Q], o.s[1] = L.s[MW_K], o.s[2] = L.s[MW_V];
        } else if (o.type == OP_O) {
            o.K = P.attn_seg, o.seg_rows[0] = P.E, o.w[0] = L.w[MW_O], o.s[0] = L.s[MW_O];
        } else if (o.type == OP_GATEUP) { // w[1]/s[1] = up rows of the same pair
            o.K = P.E, o.pair = 1, o.seg_rows[0] = P.H;
            o.w[0] = L.w[MW_GATE], o.s[0] = L.s[MW_GATE], o.w[1] = L.w[MW_UP], o.s[1] = L.s[MW_UP];
        } else {
            o.K = P.H, o.seg_rows[0] = P.E, o.w[0] = L.w[MW_DOWN], o.s[0] = L.s[MW_DOWN];
        }
    }
    const unsigned T = (unsigned)(o.seg_rows[0] + o.seg_rows[1] + o.seg_rows[2]);
    o.a = (int)((T * (unsigned)cta) / (unsigned)G), o.b = (int)((T * (unsigned)(cta + 1)) / (unsigned)G);
    const int nslices = (o.K + MG_SLICE - 1) / MG_SLICE;
    o.kparts = o.type == OP_LMHEAD ? 1 : (nslices <= MG_MAXPARTS ? nslices : MG_MAXPARTS);
    o.ipu = (nslices + o.kparts - 1) / o.kparts;
    return o;
}
// concatenated row -> (segment, row inside the segment)
__device__ __forceinline__ void seg_of(const OpDesc &o, int r, int &seg, int &local) {
    seg = 0, local = r;
    if (o.nseg > 1) {
        if (local >= o.seg_rows[0]) {
            local -= o.seg_rows[0], seg = 1;
            if (local >= o.seg_rows[1]) local -= o.seg_rows[1], seg = 2;
        }
    }
}

// ---- helper warp: L2 prefetch of this CTA's rows, MG_LOOKAHEAD ops ahead --------------------------------------------------
__device__ void prefetch_range(const void *base, size_t bytes, int lane) {
    const size_t CH = 16384;
    const unsigned char *p = (const unsigned char *)base;
    for (size_t off = (size_t)lane * CH; off < bytes; off += 32 * CH) {
        const size_t n = bytes - off < CH ? bytes - off : CH;
        l2_prefetch(p + off, (uint32_t)((n + 15) & ~(size_t)15));
    }
}
__device__ void helper_loop(const MegaParams &P, const MegaLayer *slw, volatile int *s_cur_op, int lane) {
    const int G = gridDim.x, cta = blockIdx.x, n_ops = P.layers * 4 + 1;
    // + MG_LOOKAHEAD: the first ops of the NEXT token's launch are pulled in while lm_head is being computed
    for (int pf = 0; pf < n_ops + MG_LOOKAHEAD; pf++) {
        while (*s_cur_op + MG_LOOKAHEAD < pf) __nanosleep(200);
        const OpDesc o = op_desc(P, slw, pf < n_ops ? pf : pf - n_ops, cta, G);
        int seg_start = 0;
        for (int seg = 0; seg < o.nseg; seg++) {
            const int seg_end = seg_start + o.seg_rows[seg];
            const int p0 = max(o.a, seg_start) - seg_start, p1 = min(o.b, seg_end) - seg_start;
            if (p1 > p0) {
                const size_t rb = o.K / 2, sb = o.K / 8;
                prefetch_range(o.w[seg] + (size_t)p0 * rb, (size_t)(p1 - p0) * rb, lane);
                prefetch_range((const unsigned char *)o.s[seg] + (size_t)p0 * sb, (size_t)(p1 - p0) * sb, lane);
                if (o.pair) {
                    prefetch_range(o.w[1] + (size_t)p0 * rb, (size_t)(p1 - p0) * rb, lane);
                    prefetch_range((const unsigned char *)o.s[1] + (size_t)p0 * sb, (size_t)(p1 - p0) * sb, lane);
                }
            }
            seg_start = seg_end;
        }
    }
}

// ---- cross-CTA ordering ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void op_signal(unsigned *cnt) { // all consumer threads call
    consumer_bar();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(cnt) : "memory");
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
// Four threads own one 32-element block (8 consecutive elements each): global loads are fully coalesced, the
// block max / sum are two xor-shuffles, and for the RMSNorm ops the values stay in registers between the
// sum-of-squares pass and the quantisation pass, so a prologue costs ONE L2 round trip and one barrier.
#define MG_MAXP 4
template <int MM, bool ACTQ8>
__device__ void stage_acts(const MegaParams &P, unsigned char *acts, const float *src, int ld, int K, const void *norm_w,
                           int norm_dt, double *red /*[MM][16]*/) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = K / 32;
    ActView av = act_view<MM>(acts, nblk);
    const int per_row = K / 8;            // 8-element tasks per row
    const int total = MM * per_row;
    const int passes = (total + MG_CONSUMERS - 1) / MG_CONSUMERS;
    const bool keep = passes <= MG_MAXP;   // values fit in registers across the norm barrier
    float xv[MG_MAXP][8];
    float rsv[MM];
#pragma unroll
    for (int m = 0; m < MM; m++) rsv[m] = 1.0f;

    auto load8 = [&](int task, float (&v)[8]) {
        const int m = task / per_row, j = task - m * per_row;
        if (task < total && m < P.M) {
            const float4 a = ldcg4(src + (size_t)m * ld + j * 8), b = ldcg4(src + (size_t)m * ld + j * 8 + 4);
            v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = 0.0f;
        }
    };

    if (norm_w) { // RMSNorm.java:41-52: float products summed in double, /E, +eps, 1/sqrt, cast to float
        double ss[MM];
#pragma unroll
        for (int m = 0; m < MM; m++) ss[m] = 0.0;
        if (keep) {
#pragma unroll
            for (int p = 0; p < MG_MAXP; p++)
                if (p < passes) load8(p * MG_CONSUMERS + tid, xv[p]);
#pragma unroll
            for (int p = 0; p < MG_MAXP; p++)
                if (p < passes) {
                    const int m = (p * MG_CONSUMERS + tid) / per_row;
                    double t = 0.0;
#pragma unroll
                    for (int i = 0; i < 8; i++) t += (double)__fmul_rn(xv[p][i], xv[p][i]);
#pragma unroll
                    for (int mm = 0; mm < MM; mm++)
                        if (mm == m) ss[mm] += t;
                }
        } else {
            for (int task = tid; task < total; task += MG_CONSUMERS) {
                float v[8];
                load8(task, v);
                const int m = task / per_row;
                double t = 0.0;
#pragma unroll
                for (int i = 0; i < 8; i++) t += (double)__fmul_rn(v[i], v[i]);
#pragma unroll
                for (int mm = 0; mm < MM; mm++)
                    if (mm == m) ss[mm] += t;
            }
        }
#pragma unroll
        for (int m = 0; m < MM; m++) {
            ss[m] = warp_sum_d(ss[m]);
            if (lane == 0) red[m * MG_CWARPS + warp] = ss[m];
        }
        consumer_bar();
#pragma unroll
        for (int m = 0; m < MM; m++) { // every thread finishes the reduction itself: no second barrier
            double t = 0;
#pragma unroll
            for (int w = 0; w < MG_CWARPS; w++) t += red[m * MG_CWARPS + w];
            t /= (double)P.E;
            t += (double)P.eps;
            rsv[m] = (float)(1.0 / sqrt(t));
        }
    }

    auto process = [&](int task, float (&v)[8]) {
        const int m = task / per_row, j = task - m * per_row;
        const int b = j >> 2, sub = j & 3; // block, 8-element part of the block
        if (norm_w && task < total) {
            float rsf = 1.0f;
#pragma unroll
            for (int mm = 0; mm < MM; mm++)
                if (mm == m) rsf = rsv[mm];
            float w[8];
            if (norm_dt == JL_BF16) {
                const uint4 u = __ldg((const uint4 *)((const uint16_t *)norm_w + j * 8));
                w[0] = __uint_as_float(u.x << 16), w[1] = __uint_as_float(u.x & 0xffff0000u);
                w[2] = __uint_as_float(u.y << 16), w[3] = __uint_as_float(u.y & 0xffff0000u);
                w[4] = __uint_as_float(u.z << 16), w[5] = __uint_as_float(u.z & 0xffff0000u);
                w[6] = __uint_as_float(u.w << 16), w[7] = __uint_as_float(u.w & 0xffff0000u);
            } else {
                const float4 a = __ldg((const float4 *)((const float *)norm_w + j * 8));
                const float4 c = __ldg((const float4 *)((const float *)norm_w + j * 8 + 4));
                w[0] = a.x, w[1] = a.y, w[2] = a.z, w[3] = a.w, w[4] = c.x, w[5] = c.y, w[6] = c.z, w[7] = c.w;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = __fmul_rn(__fadd_rn(0.0f, w[i]), __fmul_rn(rsf, v[i]));
        }
        if (ACTQ8) { // PanamaTensorOperations.java:1696-1710
            float mx = 0.0f;
#pragma unroll
            for (int i = 0; i < 8; i++) mx = fmaxf(mx, fabsf(v[i]));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float d = __fdiv_rn(mx, 127.0f);
            const float id = mx != 0.0f ? __fdiv_rn(127.0f, mx) : 0.0f;
            int q[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                q[i] = (int)__fadd_rn(__fmul_rn(v[i], id), 0.5f);
                sum += q[i];
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            if (task < total) {
                uint2 w2;
                w2.x = (uint32_t)(q[0] & 0xFF) | ((uint32_t)(q[1] & 0xFF) << 8) | ((uint32_t)(q[2] & 0xFF) << 16) | ((uint32_t)(q[3] & 0xFF) << 24);
                w2.y = (uint32_t)(q[4] & 0xFF) | ((uint32_t)(q[5] & 0xFF) << 8) | ((uint32_t)(q[6] & 0xFF) << 16) | ((uint32_t)(q[7] & 0xFF) << 24);
                *(uint2 *)(av.aq + (((size_t)m * 2 + (sub >> 1)) * nblk + b) * 16 + (sub & 1) * 8) = w2;
                if (sub == 0) {
                    av.asc[m * nblk + b] = d;
                    av.asum[m * nblk + b] = sum;
                }
            }
        } else if (task < total) {
            av.af4[((size_t)m * 8 + sub * 2) * nblk + b] = make_float4(v[0], v[1], v[2], v[3]);
            av.af4[((size_t)m * 8 + sub * 2 + 1) * nblk + b] = make_float4(v[4], v[5], v[6], v[7]);
        }
    };
    if (norm_w && keep) {
#pragma unroll
        for (int p = 0; p < MG_MAXP; p++)
            if (p < passes) process(p * MG_CONSUMERS + tid, xv[p]);
    } else {
        for (int p = 0; p < passes; p += 2) { // two independent tasks in flight per thread
            float v0[8], v1[8];
            load8(p * MG_CONSUMERS + tid, v0);
            if (p + 1 < passes) load8((p + 1) * MG_CONSUMERS + tid, v1);
            process(p * MG_CONSUMERS + tid, v0);
            if (p + 1 < passes) process((p + 1) * MG_CONSUMERS + tid, v1);
        }
    }
    consumer_bar();
}


// ---- the GEMV phase of one op: pipelined row-slices ----------------------------------------------------------------------
// A warp's work list: units u = warp, warp + 20, ... of the CTA's (rows x K-partitions); a unit is `ipu` row-slices
// (x2 for gate/up pairs: gate slice then up slice).  Slot = one row-slice in flight: two 16-byte blocks per lane.
struct Slot {
    uint4 q[2];
    float s[2];
    int meta; // gb0 (first block, 10 bits) | nbl (7 bits) << 10 | part << 17 | up << 20 | last << 21 | erow << 22
};
struct ItemCursor {
    int u, it, side; // unit, slice within the unit, 0 = row/gate, 1 = up
};
__device__ __forceinline__ bool cursor_valid(const ItemCursor &c, int nunits) { return c.u < nunits; }
__device__ __forceinline__ void cursor_next(ItemCursor &c, const OpDesc &o) {
    if (o.pair && c.side == 0) {
        c.side = 1;
        return;
    }
    c.side = 0;
    if (++c.it == o.ipu) c.it = 0, c.u += MG_CWARPS;
}
__device__ __forceinline__ void slot_load(Slot &sl, const ItemCursor &c, const OpDesc &o, int lane) {
    const int erow = c.u / o.kparts, part = c.u - erow * o.kparts;
    int seg, local;
    seg_of(o, o.a + erow, seg, local);
    const int slice = part * o.ipu + c.it;
    const int c0 = slice * MG_SLICE;
    int ncols = o.K - c0;
    ncols = ncols > MG_SLICE ? MG_SLICE : (ncols < 0 ? 0 : ncols);
    const int nbl = ncols >> 5, gb0 = c0 >> 5;
    const uint8_t *w = o.pair ? o.w[c.side] : o.w[seg];
    const float *s = o.pair ? o.s[c.side] : o.s[seg];
    const uint8_t *wrow = w + (size_t)local * (size_t)(o.K >> 1);
    const float *srow = s + (size_t)local * (size_t)(o.K >> 5);
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int bl = lane + 32 * j;
        if (bl < nbl) {
            sl.q[j] = ldg_nc_u4(wrow + (size_t)(gb0 + bl) * 16);
            sl.s[j] = ldg_nc_f32(srow + gb0 + bl);
        } else {
            sl.q[j] = make_uint4(0, 0, 0, 0), sl.s[j] = 0.0f;
        }
    }
    const int last = (c.it == o.ipu - 1);
    sl.meta = gb0 | (nbl << 10) | (part << 17) | (c.side << 20) | (last << 21) | (erow << 22);
}

template <int MM>
__device__ __forceinline__ void slot_math_q8(const Slot &sl, const ActView &av, int nblk_total, float (&acc)[MM], int lane) {
    const int gb0 = sl.meta & 1023, nbl = (sl.meta >> 10) & 127;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int bl = lane + 32 * j;
        if (bl >= nbl) continue;
        const int b = gb0 + bl;
        const uint4 q = sl.q[j];
#pragma unroll
        for (int m = 0; m < MM; m++) {
            const uint4 alo = *(const uint4 *)(av.aq + (((size_t)m * 2 + 0) * nblk_total + b) * 16);
            const uint4 ahi = *(const uint4 *)(av.aq + (((size_t)m * 2 + 1) * nblk_total + b) * 16);
            int s0 = 0, s1 = 0;
            s0 = __dp4a((int)(q.x & 0x0F0F0F0Fu), (int)alo.x, s0);
            s1 = __dp4a((int)((q.x >> 4) & 0x0F0F0F0Fu), (int)ahi.x, s1);
            s0 = __dp4a((int)(q.y & 0x0F0F0F0Fu), (int)alo.y, s0);
            s1 = __dp4a((int)((q.y >> 4) & 0x0F0F0F0Fu), (int)ahi.y, s1);
            s0 = __dp4a((int)(q.z & 0x0F0F0F0Fu), (int)alo.z, s0);
            s1 = __dp4a((int)((q.z >> 4) & 0x0F0F0F0Fu), (int)ahi.z, s1);
            s0 = __dp4a((int)(q.w & 0x0F0F0F0Fu), (int)alo.w, s0);
            s1 = __dp4a((int)((q.w >> 4) & 0x0F0F0F0Fu), (int)ahi.w, s1);
            const int sm = s0 + s1 - 8 * av.asum[m * nblk_total + b];
            acc[m] = fmaf(__fmul_rn(av.asc[m * nblk_total + b], sl.s[j]), (float)sm, acc[m]);
        }
    }
}
// F32 activations x Q4 weights (lm_head; AbstractModel.java:444-449)
template <int MM>
__device__ __forceinline__ void slot_math_f32(const Slot &sl, const ActView &av, int nblk_total, float (&acc)[MM], int lane) {
    const int gb0 = sl.meta & 1023, nbl = (sl.meta >> 10) & 127;
#pragma unroll
    for (int j = 0; j < 2; j++) {
        const int bl = lane + 32 * j;
        if (bl >= nbl) continue;
        const int b = gb0 + bl;
        const uint4 q = sl.q[j];
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
                const float4 a4 = av.af4[((size_t)m * 8 + c4) * nblk_total + b];
                part = fmaf(a4.x, wf[c4 * 4 + 0], part);
                part = fmaf(a4.y, wf[c4 * 4 + 1], part);
                part = fmaf(a4.z, wf[c4 * 4 + 2], part);
                part = fmaf(a4.w, wf[c4 * 4 + 3], part);
            }
            acc[m] = fmaf(sl.s[j], part, acc[m]);
        }
    }
}

// ---- attention task: shared with the stand-alone fused decode attention kernel (jl_attn_task.cuh) ---------------------
__device__ __forceinline__ AttnTask mega_attn_task(const MegaParams &P) {
    AttnTask t;
    t.heads = P.heads, t.kv_heads = P.kv_heads, t.head_size = P.head_size, t.attn_seg = P.attn_seg, t.kv_seg = P.kv_seg;
    t.kv_head0_global = P.kv_head0_global, t.splits = P.splits, t.attn_scale = P.attn_scale;
    t.q = P.q, t.k = P.k, t.v = P.v, t.att = P.att, t.attn_ws = P.attn_ws, t.rope = P.rope, t.kv = P.kv;
    t.sessions = P.sessions, t.positions = P.positions;
    return t;
}

__device__ __forceinline__ unsigned long long pack_arg(float v, int idx) {
    if (!(v == v)) return 0ull; // NaN never wins (AbstractModel.java:465 'v > maxv' is false)
    uint32_t b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)idx);
}


#define MG_TRACE(slot_)                                                       \
    do {                                                                      \
        if (tr && tid == 0) tr[(size_t)op * 8 + (slot_)] = clock64();         \
    } while (0)

template <int MM>
__global__ void __launch_bounds__(MG_THREADS, 1) mega_decode_kernel(const MegaParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *uarea = smem;                                     // activations / attention scratch
    MegaLayer *slw = (MegaLayer *)(smem + P.uarea_bytes);            // copy of the layer pointer table
    __shared__ double red[MM * MG_CWARPS];
    __shared__ unsigned long long wbest[MM][MG_CWARPS];
    __shared__ int s_last;
    __shared__ float ebuf[MG_EROWS * 2 * MG_MAXPARTS * MM];
    __shared__ int emeta[MG_EROWS];
    __shared__ int s_cur_op;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    const int n_ops = P.layers * 4 + 1;
    {
        const int words = P.layers * (int)(sizeof(MegaLayer) / 4);
        for (int i = tid; i < words; i += MG_THREADS) ((uint32_t *)slw)[i] = __ldg((const uint32_t *)P.lw + i);
        if (tid == 0) s_cur_op = 0;
    }
    __syncthreads();

    if (warp == MG_CWARPS) { // ===== helper warp: L2 prefetch ahead of the consumers =====
        if (P.l2_ahead) helper_loop(P, slw, &s_cur_op, lane);
        return;
    }

    // ===== consumers =====
    unsigned *sync = P.sync;
    long long *tr = nullptr;
    if (P.trace) {
        const int which = cta == 0 ? 0 : (cta == G / 2 ? 1 : (cta == G - 1 ? 2 : -1));
        if (which >= 0) tr = P.trace + (size_t)which * n_ops * 8;
    }
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

    float best_v[MM];
    int best_i[MM];
#pragma unroll
    for (int m = 0; m < MM; m++) best_v[m] = -INFINITY, best_i[m] = 0x7fffffff;

    for (int op = 0; op < n_ops; op++) {
        const OpDesc oi = op_desc(P, slw, op, cta, G);
        const int L = oi.L;
        unsigned *cnt = &sync[1 + (op < P.layers * 4 ? L * 5 + (oi.type == OP_QKV ? 0 : oi.type == OP_O ? 2 : oi.type == OP_GATEUP ? 3 : 4)
                                                      : P.layers * 5)];
        if (tid == 0) s_cur_op = op;
        MG_TRACE(0);
        // ---- request this warp's first row-slices now: weights do not depend on the previous op ----
        const int nrows = oi.b - oi.a;            // rows (pairs) of this op owned by the CTA
        const int nunits = nrows * oi.kparts;
        ItemCursor lc = {warp, 0, 0}, cc = {warp, 0, 0};
        Slot sl[MG_DEPTH];
#pragma unroll
        for (int k = 0; k < MG_DEPTH; k++) {
            if (cursor_valid(lc, nunits)) {
                slot_load(sl[k], lc, oi, lane);
                cursor_next(lc, oi);
            } else {
                sl[k].meta = 0;
            }
        }
        // ---- attention phase sits between QKV and O ----
        if (oi.type == OP_O && !(P.dbg & 2)) {
            const int ntasks = P.M * P.kv_heads * P.splits;
            if (cta < ntasks) {
                const int split = cta % P.splits, kvh = (cta / P.splits) % P.kv_heads, m = cta / (P.splits * P.kv_heads);
                op_wait(&sync[1 + L * 5 + 0], (unsigned)G);
                const AttnTask at = mega_attn_task(P);
                switch (P.head_size) {
                    case 32: attention_task<32, MG_CONSUMERS>(at, L, m, kvh, split, uarea); break;
                    case 64: attention_task<64, MG_CONSUMERS>(at, L, m, kvh, split, uarea); break;
                    default: attention_task<128, MG_CONSUMERS>(at, L, m, kvh, split, uarea); break;
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
                            case 32: attention_merge<32, MG_CONSUMERS>(at, m, kvh, uarea); break;
                            case 64: attention_merge<64, MG_CONSUMERS>(at, m, kvh, uarea); break;
                            default: attention_merge<128, MG_CONSUMERS>(at, m, kvh, uarea); break;
                        }
                        op_signal(&sync[1 + L * 5 + 1]);
                    }
                } else {
                    op_signal(&sync[1 + L * 5 + 1]);
                }
            }
        }
        MG_TRACE(1);
        if (nrows > 0) {
            // ---- dependency + prologue ----
            const int K = oi.K, nblk = K / 32;
            if (P.dbg & 2) {
            } else if (oi.type == OP_QKV) {
                op_wait(L == 0 ? &sync[0] : &sync[1 + (L - 1) * 5 + 4], (unsigned)G);
                MG_TRACE(2);
                stage_acts<MM, true>(P, uarea, P.x, P.E, K, slw[L].attn_norm, slw[L].attn_norm_dt, red);
            } else if (oi.type == OP_O) {
                op_wait(&sync[1 + L * 5 + 1], (unsigned)(P.M * P.kv_heads));
                MG_TRACE(2);
                stage_acts<MM, true>(P, uarea, P.att, P.attn_seg, K, nullptr, 0, red);
            } else if (oi.type == OP_GATEUP) {
                op_wait(&sync[1 + L * 5 + 2], (unsigned)G);
                MG_TRACE(2);
                stage_acts<MM, true>(P, uarea, P.xb, P.E, K, slw[L].ffn_norm, slw[L].ffn_norm_dt, red);
            } else if (oi.type == OP_DOWN) {
                op_wait(&sync[1 + L * 5 + 3], (unsigned)G);
                MG_TRACE(2);
                stage_acts<MM, true>(P, uarea, P.h, P.H, K, nullptr, 0, red);
            } else {
                op_wait(&sync[1 + (P.layers - 1) * 5 + 4], (unsigned)G);
                MG_TRACE(2);
                stage_acts<MM, false>(P, uarea, P.x, P.E, K, P.out_norm, P.out_norm_dt, red);
            }
            const ActView av = act_view<MM>(uarea, nblk);
            MG_TRACE(3);

            // ---- pipelined row-slices ----
            float acc[MM];
#pragma unroll
            for (int m = 0; m < MM; m++) acc[m] = 0.0f;
            while (cursor_valid(cc, nunits)) {
#pragma unroll
                for (int k = 0; k < MG_DEPTH; k++) {
                    if (!cursor_valid(cc, nunits)) break;
                    if (!(P.dbg & 1)) {
                        if (oi.type == OP_LMHEAD) slot_math_f32<MM>(sl[k], av, nblk, acc, lane);
                        else slot_math_q8<MM>(sl[k], av, nblk, acc, lane);
                    }
                    const int meta = sl[k].meta;
                    // refill the slot with the item MG_DEPTH ahead
                    if (cursor_valid(lc, nunits)) {
                        slot_load(sl[k], lc, oi, lane);
                        cursor_next(lc, oi);
                    }
                    if ((meta >> 21) & 1) { // last slice of this unit(-side): reduce and park / store
#pragma unroll
                        for (int m = 0; m < MM; m++) acc[m] = warp_sum(acc[m]);
                        const int erow = (int)((unsigned)meta >> 22), part = (meta >> 17) & 7, up = (meta >> 20) & 1;
                        if (lane == 0) {
                            if (oi.type == OP_LMHEAD) { // logits + running arg-max (strict '>', lowest index wins)
                                const int row = oi.a + erow;
#pragma unroll
                                for (int m = 0; m < MM; m++) {
                                    if (m >= P.M) continue;
                                    const float v = acc[m];
                                    P.logits[(size_t)m * P.vocab + row] = v;
                                    if (v > best_v[m] || (v == best_v[m] && row < best_i[m])) best_v[m] = v, best_i[m] = row;
                                }
                            } else {
#pragma unroll
                                for (int m = 0; m < MM; m++) ebuf[(((size_t)erow * 2 + up) * MG_MAXPARTS + part) * MM + m] = acc[m];
                            }
                        }
#pragma unroll
                        for (int m = 0; m < MM; m++) acc[m] = 0.0f;
                    }
                    cursor_next(cc, oi);
                }
            }
            MG_TRACE(4);
            if (oi.type != OP_LMHEAD) {
                // ---- op-end pass: combine K-partitions (fixed order) and gate/up, apply the fused epilogue ----
                consumer_bar();
                for (int t = tid; t < nrows * MM; t += MG_CONSUMERS) {
                    const int e = t / MM, m = t - e * MM;
                    if (m >= P.M) continue;
                    int seg, row;
                    seg_of(oi, oi.a + e, seg, row);
                    float v = ebuf[(((size_t)e * 2 + 0) * MG_MAXPARTS + 0) * MM + m];
                    for (int pt = 1; pt < oi.kparts; pt++) v += ebuf[(((size_t)e * 2 + 0) * MG_MAXPARTS + pt) * MM + m];
                    switch (oi.type) {
                        case OP_QKV: {
                            float *out = seg == 0 ? P.q + (size_t)m * P.attn_seg : (seg == 1 ? P.k : P.v) + (size_t)m * P.kv_seg;
                            out[row] = v;
                        } break;
                        case OP_O: // TransformerBlock.java:185
                            P.xb[(size_t)m * P.E + row] = __fadd_rn(v, __ldcg(P.x + (size_t)m * P.E + row));
                            break;
                        case OP_GATEUP: { // MLPBlock.java:132-141
                            float u = ebuf[(((size_t)e * 2 + 1) * MG_MAXPARTS + 0) * MM + m];
                            for (int pt = 1; pt < oi.kparts; pt++) u += ebuf[(((size_t)e * 2 + 1) * MG_MAXPARTS + pt) * MM + m];
                            P.h[(size_t)m * P.H + row] = __fmul_rn(silu_ref(v), u);
                        } break;
                        default: // OP_DOWN, TransformerBlock.java:203
                            P.x[(size_t)m * P.E + row] = __fadd_rn(v, __ldcg(P.xb + (size_t)m * P.E + row));
                            break;
                    }
                }
            }
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
        MG_TRACE(5);
    }
    if (tid == 0) s_cur_op = n_ops + MG_LOOKAHEAD; // release the helper warp

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
                    const int cntv = *P.counter;
                    P.tokens[m] = tok;
                    P.positions[m] += 1;
                    if (cntv * P.M + m < P.hist_cap) P.hist[cntv * P.M + m] = tok;
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
    att += (size_t)2 * hs * 4;                                  // current position's key / value row
    return (((acts > att ? acts : att) + 256) + 1023) & ~(size_t)1023;
}

bool jl_mega_supported(const MegaParams &p) {
    if (p.M < 1 || p.M > MEGA_MAX_M) return false;
    if (p.head_size != 32 && p.head_size != 64 && p.head_size != 128) return false;
    if (p.heads % p.kv_heads || p.heads / p.kv_heads > MG_MAX_GROUP) return false;
    if ((p.E % 32) || (p.H % 32) || (p.attn_seg % 32)) return false;
    if (p.layers > MG_MAX_LAYERS) return false;
    if (p.E > MG_SLICE * MG_MAXPARTS) return false; // gate/up pairs: one row-slice per K-partition
    for (int K : {p.E, p.H, p.attn_seg})
        if (K > MG_SLICE * MG_MAXPARTS * 4 || (K >> 5) > 1023) return false; // slot meta: 10-bit block index
    if (p.grid > 0) {
        const int G = p.grid;
        if ((p.attn_seg + 2 * p.kv_seg + G - 1) / G + 1 > MG_EROWS || (p.H + G - 1) / G + 1 > MG_EROWS || (p.E + G - 1) / G + 1 > MG_EROWS)
            return false;
        if ((p.vocab + G - 1) / G + 1 >= (1 << 10)) return false; // erow field of the slot meta
    }
    const int MM = p.M <= 1 ? 1 : (p.M <= 2 ? 2 : 4);
    return uarea_bytes(p, MM) + (size_t)p.layers * sizeof(MegaLayer) <= 200 * 1024;
}

template <int MM>
static int launch_mega(jl_ctx *ctx, cudaStream_t stream, MegaParams p) {
    auto kern = mega_decode_kernel<MM>;
    p.uarea_bytes = (int)uarea_bytes(p, MM);
    const size_t smem = (size_t)p.uarea_bytes + (size_t)p.layers * sizeof(MegaLayer);
    static size_t configured[JL_MAX_DEVICES] = {};
    JL_CUDA_CHECK(ctx, jl_ensure_dyn_smem(kern, ctx->device, smem, configured));
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
    if (p.M <= 1) return launch_mega<1>(ctx, stream, p);
    if (p.M <= 2) return launch_mega<2>(ctx, stream, p);
    return launch_mega<4>(ctx, stream, p);
}
