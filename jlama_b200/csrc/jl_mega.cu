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
//   * 8 CONSUMER warps (two row-units each) wait on the ring's mbarriers, run the dp4a block dot products against the Q8
//     activations staged in shared memory, and apply the fused epilogues (residual add, SiLU*up, arg-max);
//   * ops are ordered by per-op completion counters in global memory (release/acquire), not kernel boundaries;
//     RMSNorm + Q8 quantisation are recomputed per CTA in the op prologue (one L2 round trip);
//   * attention (RoPE, KV append, scores, softmax, P.V over the paged KV cache) runs as (row, kv-head, split)
//     tasks on the first CTAs while every producer keeps prefetching the following matrices.
//
// Arithmetic is the same as the kernel-per-op path (jl_gemv.cu / jl_attention.cu): identical per-lane block
// order, so GEMV results are bit-identical; the launch is cooperative so the spin waits cannot deadlock.
#include "jl_mega.cuh"

#define MG_CWARPS 8   // consumer warps; each handles two of the 16 row-units of a stage
#define MG_UNITS 16
#define MG_CONSUMERS (MG_CWARPS * 32)
#define MG_THREADS (MG_CONSUMERS + 32)
#define MG_STAGE_NIB 32768
#define MG_STAGE_SC 8192
#define MG_STAGE_BYTES (MG_STAGE_NIB + MG_STAGE_SC)
#define MG_REC_BYTES 96
#define MG_EROWS 112   // max rows (or pairs) of one op owned by one CTA (lm_head excluded)
#define MG_MAX_OPS 520 // 4 * layers + 1
#define MG_DB 4 // stage records per descriptor batch
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

// ---- static schedule (host-built table) --------------------------------------------------------------------------------
// ops: for each layer QKV, O, GATEUP, DOWN; then LMHEAD.
enum { OP_QKV = 0, OP_O, OP_GATEUP, OP_DOWN, OP_LMHEAD };
// A stage = R FULL weight rows (contiguous in HBM -> one bulk copy for the nibbles, one for the scales):
//   plain op: R = min(16, 32 KB / row bytes) rows, 16/R units per row (each a K-range of the row);
//   pair op (gate/up): R pairs = R gate rows + R up rows (two + two copies), 16/(2R) units per row.
// A stage always has 16 units; consumer warp w owns units w and w+8, which share their K-range (so the
// activations of that range live in the warp's registers) and, for pair ops, are the gate and up row of one pair.
// slot layout: nibbles of row i at i*(K/2) (pair: gate rows first, up rows at R*(K/2)); scales at
// MG_STAGE_NIB + i*(K/8) (pair: up scales at MG_STAGE_NIB + R*(K/8)).
struct Stage {
    int seg, row0, nrows, R, wpr, K, pair, type, wpr_shift, r_shift;
};
__device__ __forceinline__ int stage_rows_dev(int K, int pair) {
    const int rb = K / 2;
    int R = pair ? 8 : 16;
    const int budget = pair ? MG_STAGE_NIB / 2 : MG_STAGE_NIB;
    while (R > 1 && R * rb > budget) R >>= 1;
    return R;
}
static int stage_rows(int K, int pair) {
    const int rb = K / 2;
    int R = pair ? 8 : 16;
    const int budget = pair ? MG_STAGE_NIB / 2 : MG_STAGE_NIB;
    while (R > 1 && R * rb > budget) R >>= 1;
    return R;
}

void jl_mega_build_table(const MegaParams &P, const MegaLayer *layers, int G, std::vector<unsigned char> &records,
                         std::vector<int> &cta_first, std::vector<int> &op_first) {
    const int n_ops = P.layers * 4 + 1;
    std::vector<MegaCopy> copies;
    std::vector<MegaMeta> metas;
    cta_first.assign(G + 1, 0);
    op_first.assign((size_t)G * (n_ops + 1), 0);
    for (int cta = 0; cta < G; cta++) {
        cta_first[cta] = (int)metas.size();
        for (int op = 0; op < n_ops; op++) {
            op_first[(size_t)cta * (n_ops + 1) + op] = (int)metas.size();
            int type, K, pair = 0, nseg = 1, seg_rows[3] = {0, 0, 0};
            const uint8_t *w[3] = {nullptr, nullptr, nullptr};
            const float *sc[3] = {nullptr, nullptr, nullptr};
            if (op == n_ops - 1) {
                type = OP_LMHEAD, K = P.E, seg_rows[0] = P.vocab, w[0] = P.lm_w, sc[0] = P.lm_s;
            } else {
                const MegaLayer &L = layers[op >> 2];
                type = op & 3;
                if (type == OP_QKV) {
                    nseg = 3, K = P.E;
                    seg_rows[0] = P.attn_seg, seg_rows[1] = P.kv_seg, seg_rows[2] = P.kv_seg;
                    w[0] = L.w[MW_Q], w[1] = L.w[MW_K], w[2] = L.w[MW_V];
                    sc[0] = L.s[MW_Q], sc[1] = L.s[MW_K], sc[2] = L.s[MW_V];
                } else if (type == OP_O) {
                    K = P.attn_seg, seg_rows[0] = P.E, w[0] = L.w[MW_O], sc[0] = L.s[MW_O];
                } else if (type == OP_GATEUP) {
                    K = P.E, pair = 1, seg_rows[0] = P.H;
                    w[0] = L.w[MW_GATE], sc[0] = L.s[MW_GATE], w[1] = L.w[MW_UP], sc[1] = L.s[MW_UP];
                } else {
                    K = P.H, seg_rows[0] = P.E, w[0] = L.w[MW_DOWN], sc[0] = L.s[MW_DOWN];
                }
            }
            long long T = 0;
            for (int i = 0; i < nseg; i++) T += seg_rows[i];
            const int a = (int)((T * cta) / G), b = (int)((T * (cta + 1)) / G);
            const int R = stage_rows(K, pair);
            const uint32_t rb = K / 2, sbp = K / 8;
            int seg_start = 0;
            for (int seg = 0; seg < nseg; seg++) {
                const int seg_end = seg_start + seg_rows[seg];
                const int p0 = std::max(a, seg_start) - seg_start, p1 = std::min(b, seg_end) - seg_start;
                for (int rg = p0; rg < p1; rg += R) {
                    const int nrows = std::min(R, p1 - rg);
                    MegaMeta md = {};
                    md.total_bytes = (uint32_t)nrows * (rb + sbp) * (pair ? 2u : 1u);
                    md.row0 = rg, md.nrows = (int16_t)nrows, md.R = (int16_t)R, md.wpr = (int16_t)(MG_UNITS / (pair ? 2 * R : R));
                    md.seg = (int16_t)seg, md.pair = (uint8_t)pair, md.type = (uint8_t)type, md.op = op, md.K = K;
                    for (int t = md.wpr; t > 1; t >>= 1) md.wpr_shift++;
                    for (int t = R; t > 1; t >>= 1) md.r_shift++;
                    metas.push_back(md);
                    for (int c = 0; c < 4; c++) {
                        MegaCopy cp = {0, 0, 0};
                        const int up = c >> 1, is_sc = c & 1;
                        if (c < (pair ? 4 : 2)) {
                            const uint8_t *wp = pair ? (up ? w[1] : w[0]) : w[seg];
                            const float *sp = pair ? (up ? sc[1] : sc[0]) : sc[seg];
                            cp.src = is_sc ? (unsigned long long)(sp + (size_t)rg * (K / 32)) : (unsigned long long)(wp + (size_t)rg * rb);
                            cp.bytes = (uint32_t)nrows * (is_sc ? sbp : rb);
                            cp.dst = is_sc ? (uint32_t)(MG_STAGE_NIB + (size_t)up * R * sbp) : (uint32_t)((size_t)up * R * rb);
                        }
                        copies.push_back(cp);
                    }
                }
                seg_start = seg_end;
            }
        }
        op_first[(size_t)cta * (n_ops + 1) + n_ops] = (int)metas.size();
    }
    cta_first[G] = (int)metas.size();
    records.resize(metas.size() * MG_REC_BYTES);
    for (size_t i = 0; i < metas.size(); i++) {
        memcpy(&records[i * MG_REC_BYTES], &copies[i * 4], 4 * sizeof(MegaCopy));
        memcpy(&records[i * MG_REC_BYTES + 64], &metas[i], sizeof(MegaMeta));
    }
}

// ---- producer ----------------------------------------------------------------------------------------------------------
// The warp never touches global memory on its critical path: stage records arrive in shared memory in
// batches of MG_DB through their own bulk copies (double buffered), two batches ahead of use.
template <int NSTAGE>
__device__ void producer_loop(const MegaParams &P, unsigned char *ring, uint64_t *full, uint64_t *empty, MegaMeta *sdesc,
                              unsigned char *dbuf /*[2][MG_DB*96]*/, uint64_t *dfull /*[2]*/, int lane) {
    const int s0 = __ldg(&P.cta_first[blockIdx.x]), s1 = __ldg(&P.cta_first[blockIdx.x + 1]);
    const int nst = s1 - s0, nbatch = (nst + MG_DB - 1) / MG_DB;
    auto fetch_batch = [&](int b) {
        if (b < nbatch && lane == 0) {
            const int cnt = min(MG_DB, nst - b * MG_DB);
            mbar_expect_tx(&dfull[b & 1], (uint32_t)cnt * MG_REC_BYTES);
            tma_load_1d(dbuf + (size_t)(b & 1) * MG_DB * MG_REC_BYTES, P.records + (size_t)(s0 + b * MG_DB) * MG_REC_BYTES,
                        (uint32_t)cnt * MG_REC_BYTES, &dfull[b & 1]);
        }
    };
    auto prefetch_batch = [&](int b) { // L2 prefetch of every stage of batch b (its records must have arrived)
        if (!P.l2_ahead || b >= nbatch) return;
        const int cnt = min(MG_DB, nst - b * MG_DB);
        const unsigned char *rec = dbuf + (size_t)(b & 1) * MG_DB * MG_REC_BYTES;
        for (int i = lane; i < cnt * 4; i += 32) {
            const uint4 v = *(const uint4 *)(rec + (size_t)(i >> 2) * MG_REC_BYTES + (i & 3) * 16);
            if (v.z) l2_prefetch((const void *)(((unsigned long long)v.y << 32) | v.x), v.z);
        }
    };
    fetch_batch(0);
    fetch_batch(1);
    if (nbatch > 0) {
        while (!mbar_try_wait(&dfull[0], 0)) {
        }
        prefetch_batch(0);
    }
    for (int b = 0; b < nbatch; b++) {
        if (b + 1 < nbatch) {
            while (!mbar_try_wait(&dfull[(b + 1) & 1], ((b + 1) >> 1) & 1)) {
            }
            prefetch_batch(b + 1);
        }
        const int cnt = min(MG_DB, nst - b * MG_DB);
        const unsigned char *rec = dbuf + (size_t)(b & 1) * MG_DB * MG_REC_BYTES;
        for (int j = 0; j < cnt; j++) {
            const unsigned k = (unsigned)(b * MG_DB + j), slot = k % NSTAGE, use = k / NSTAGE;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (lane < 6) v = *(const uint4 *)(rec + (size_t)j * MG_REC_BYTES + lane * 16);
            const uint32_t total_bytes = __shfl_sync(0xffffffffu, v.x, 4); // first word of the MegaMeta
            while (!mbar_try_wait(&empty[slot], (use & 1) ^ 1)) {
                __nanosleep(32);
            }
            if (lane == 4) ((uint4 *)&sdesc[slot])[0] = v;
            if (lane == 5) ((uint4 *)&sdesc[slot])[1] = v;
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&full[slot], total_bytes);
            __syncwarp();
            if (lane < 4 && v.z)
                tma_load_1d(ring + (size_t)slot * MG_STAGE_BYTES + v.w, (const void *)(((unsigned long long)v.y << 32) | v.x), v.z,
                            &full[slot]);
        }
        __syncwarp();
        fetch_batch(b + 2); // this buffer is free again
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

// ---- consumer math on one ring stage -----------------------------------------------------------------------------------
// unit u -> row slot i = u / wpr, K-range qd = u % wpr.
struct WarpRow {
    const unsigned char *nib;
    const float *sc;
    int b0, nb; // first block and number of blocks of this unit's K-range
    int row;    // row (or pair) index inside the stage
    int up;     // pair stages: 1 = up row
    bool valid;
};
__device__ __forceinline__ WarpRow warp_row(const Stage &d, const unsigned char *slot, int unit) {
    WarpRow r;
    const int i = unit >> d.wpr_shift, qd = unit & (d.wpr - 1);
    r.up = d.pair ? (i >= d.R) : 0;
    r.row = d.pair ? (i & (d.R - 1)) : i;
    r.valid = r.row < d.nrows && i < (d.pair ? 2 * d.R : d.R);
    r.nb = (d.K >> 5) >> d.wpr_shift;
    r.b0 = qd * r.nb;
    r.nib = slot + (size_t)i * (d.K >> 1);
    r.sc = (const float *)(slot + MG_STAGE_NIB + (size_t)i * (d.K >> 3));
    return r;
}

// Q8 activations of the (up to four) blocks a lane owns in its warp's K-range, kept in registers for a whole op
struct ActRegs {
    uint4 lo[4], hi[4];
    float sc[4];
    int sum[4];
};
__device__ __forceinline__ void load_act_regs(ActRegs &ar, const ActView &av, int nblk_total, int b0, int nb, int lane) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int b = b0 + lane + 32 * j;
        if (lane + 32 * j < nb) {
            ar.lo[j] = *(const uint4 *)(av.aq + ((size_t)0 * nblk_total + b) * 16);
            ar.hi[j] = *(const uint4 *)(av.aq + ((size_t)1 * nblk_total + b) * 16);
            ar.sc[j] = av.asc[b];
            ar.sum[j] = av.asum[b];
        } else {
            ar.lo[j] = make_uint4(0, 0, 0, 0), ar.hi[j] = ar.lo[j], ar.sc[j] = 0.0f, ar.sum[j] = 0;
        }
    }
}

// Q8 activations x Q4 weights for the warp's two units (same K-range).  All weight loads are issued before the
// math, the two halves of a block run on independent dp4a chains, and the per-row accumulation order (blocks
// lane, lane+32, ...) is the same as in jl_gemv.cu.  REGS: activations come from ActRegs (MM == 1).
template <int MM, bool REGS>
__device__ __forceinline__ void consume2_q8(const WarpRow (&r)[2], const ActView &av, const ActRegs &ar, int nblk_total,
                                            float (&acc)[2][MM], int lane) {
    const int b0 = r[0].b0, nb = r[0].nb;
    for (int off = lane; off < nb; off += 128) {
        uint4 q[2][4];
        float sb[2][4];
#pragma unroll
        for (int u = 0; u < 2; u++)
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int b = b0 + off + 32 * j;
                if (r[u].valid && off + 32 * j < nb) {
                    q[u][j] = *(const uint4 *)(r[u].nib + (size_t)b * 16);
                    sb[u][j] = r[u].sc[b];
                } else {
                    q[u][j] = make_uint4(0, 0, 0, 0), sb[u][j] = 0.0f;
                }
            }
#pragma unroll
        for (int m = 0; m < MM; m++) {
            int s[2][4];
            float sc[2][4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int b = b0 + off + 32 * j;
                const bool ok = off + 32 * j < nb;
                uint4 alo, ahi;
                float asc;
                int asum;
                if (REGS) {
                    alo = ar.lo[j], ahi = ar.hi[j], asc = ar.sc[j], asum = ar.sum[j];
                } else if (ok) {
                    alo = *(const uint4 *)(av.aq + (((size_t)m * 2 + 0) * nblk_total + b) * 16);
                    ahi = *(const uint4 *)(av.aq + (((size_t)m * 2 + 1) * nblk_total + b) * 16);
                    asc = av.asc[m * nblk_total + b];
                    asum = av.asum[m * nblk_total + b];
                } else {
                    alo = make_uint4(0, 0, 0, 0), ahi = alo, asc = 0.0f, asum = 0;
                }
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    int s0 = 0, s1 = 0;
                    s0 = __dp4a((int)(q[u][j].x & 0x0F0F0F0Fu), (int)alo.x, s0);
                    s1 = __dp4a((int)((q[u][j].x >> 4) & 0x0F0F0F0Fu), (int)ahi.x, s1);
                    s0 = __dp4a((int)(q[u][j].y & 0x0F0F0F0Fu), (int)alo.y, s0);
                    s1 = __dp4a((int)((q[u][j].y >> 4) & 0x0F0F0F0Fu), (int)ahi.y, s1);
                    s0 = __dp4a((int)(q[u][j].z & 0x0F0F0F0Fu), (int)alo.z, s0);
                    s1 = __dp4a((int)((q[u][j].z >> 4) & 0x0F0F0F0Fu), (int)ahi.z, s1);
                    s0 = __dp4a((int)(q[u][j].w & 0x0F0F0F0Fu), (int)alo.w, s0);
                    s1 = __dp4a((int)((q[u][j].w >> 4) & 0x0F0F0F0Fu), (int)ahi.w, s1);
                    s[u][j] = s0 + s1 - 8 * asum;
                    sc[u][j] = __fmul_rn(asc, sb[u][j]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (off + 32 * j < nb) acc[u][m] = fmaf(sc[u][j], (float)s[u][j], acc[u][m]);
        }
    }
}

// F32 activations x Q4 weights (lm_head; AbstractModel.java:444-449): the two rows share every activation load
template <int MM>
__device__ __forceinline__ void consume2_f32(const WarpRow (&r)[2], const ActView &av, int nblk_total, float (&acc)[2][MM],
                                             int lane) {
    const int b0 = r[0].b0, nb = r[0].nb;
    for (int off = lane; off < nb; off += 32) {
        const int b = b0 + off;
        float wf[2][32], sb[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            uint4 q = make_uint4(0x88888888u, 0x88888888u, 0x88888888u, 0x88888888u);
            sb[u] = 0.0f;
            if (r[u].valid) {
                q = *(const uint4 *)(r[u].nib + (size_t)b * 16);
                sb[u] = r[u].sc[b];
            }
            const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t lo = qw[i] & 0x0F0F0F0Fu, hi = (qw[i] >> 4) & 0x0F0F0F0Fu;
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    wf[u][i * 4 + t] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7540 | t)) - 8388616.0f;
                    wf[u][16 + i * 4 + t] = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7540 | t)) - 8388616.0f;
                }
            }
        }
#pragma unroll
        for (int m = 0; m < MM; m++) {
            float part[2] = {0.0f, 0.0f};
#pragma unroll
            for (int c4 = 0; c4 < 8; c4++) {
                const float4 a4 = av.af4[((size_t)m * 8 + c4) * nblk_total + b];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    part[u] = fmaf(a4.x, wf[u][c4 * 4 + 0], part[u]);
                    part[u] = fmaf(a4.y, wf[u][c4 * 4 + 1], part[u]);
                    part[u] = fmaf(a4.z, wf[u][c4 * 4 + 2], part[u]);
                    part[u] = fmaf(a4.w, wf[u][c4 * 4 + 3], part[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) acc[u][m] = fmaf(sb[u], part[u], acc[u][m]);
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
// Latency plan: the RoPE inputs, the KV append and the first K/V tile are all requested before the first
// barrier; the next tile is prefetched into registers while the current one is processed; the row of the
// current position is taken from shared memory (never re-read from the page it was just written to).
template <int HS>
__device__ void attention_task(const MegaParams &P, int layer, int m, int kvh, int split, unsigned char *u) {
    constexpr int C4 = HS / 4;
    constexpr int PARTS = MG_CONSUMERS / HS; // P.V position groups
    constexpr int NF = (MG_ATT_TILE * C4 + MG_CONSUMERS - 1) / MG_CONSUMERS; // float4 per thread per tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int group = P.heads / P.kv_heads;
    float4 *Ks = (float4 *)u;                    // [TILE][C4] swizzled
    float4 *Vs = Ks + MG_ATT_TILE * C4;          // [TILE][C4]
    float *qs = (float *)(Vs + MG_ATT_TILE * C4); // [group][HS] rotated queries
    float *ps = qs + MG_MAX_GROUP * HS;          // [TILE][MAX_GROUP]
    float *hm = ps + MG_ATT_TILE * MG_MAX_GROUP; // running max / sum / correction per head
    float *hl = hm + MG_MAX_GROUP;
    float *hc = hl + MG_MAX_GROUP;
    float *knew = hc + MG_MAX_GROUP + 8;         // [HS] rotated key of the current position
    float *vnew = knew + HS;                     // [HS]
    float *comb = vnew + HS;                     // [PARTS][MAX_GROUP][HS] P.V combine buffer

    const int session = P.sessions[m], pos = P.positions[m];
    const int n = pos + 1;
    const int S = P.splits;
    const int per = (((n + S - 1) / S) + MG_ATT_TILE - 1) / MG_ATT_TILE * MG_ATT_TILE;
    const int t0 = split * per, t1 = min(n, t0 + per);
    const int hp = HS / 2;
    const int h0 = kvh * group, xoff = kvh * HS, dt = P.kv.kv_dtype;
    const size_t poffset = (size_t)pos * hp;
    const bool owner = pos >= t0 && pos < t1;

    float4 kreg[NF], vreg[NF];
    auto fetch_tile = [&](int tb) {
#pragma unroll
        for (int i = 0; i < NF; i++) {
            const int f = tid + i * MG_CONSUMERS;
            const int r = f / C4, c4 = f % C4;
            kreg[i] = make_float4(0.f, 0.f, 0.f, 0.f), vreg[i] = kreg[i];
            if (f < MG_ATT_TILE * C4 && tb + r < t1 && tb + r != pos) {
                const char *kr = mg_kv_row(P.kv, session, layer, tb + r, 0);
                const char *vr = mg_kv_row(P.kv, session, layer, tb + r, 1);
                if (dt == JL_F32) {
                    kreg[i] = __ldcg((const float4 *)((const float *)kr + xoff + c4 * 4));
                    vreg[i] = __ldcg((const float4 *)((const float *)vr + xoff + c4 * 4));
                } else {
                    const uint2 uk = __ldcg((const uint2 *)((const uint16_t *)kr + xoff + c4 * 4));
                    const uint2 uv = __ldcg((const uint2 *)((const uint16_t *)vr + xoff + c4 * 4));
                    kreg[i] = make_float4(__uint_as_float(uk.x << 16), __uint_as_float(uk.x & 0xffff0000u),
                                          __uint_as_float(uk.y << 16), __uint_as_float(uk.y & 0xffff0000u));
                    vreg[i] = make_float4(__uint_as_float(uv.x << 16), __uint_as_float(uv.x & 0xffff0000u),
                                          __uint_as_float(uv.y << 16), __uint_as_float(uv.y & 0xffff0000u));
                }
            }
        }
    };
    if (t0 < t1) fetch_tile(t0);

    // RoPE on this group's queries (CausalSelfAttention.java:260-268: table index poffset + kvh_global*hs + j)
    for (int idx = tid; idx < group * hp; idx += MG_CONSUMERS) {
        const int h = idx / hp, j = idx % hp;
        const float2 f = __ldg((const float2 *)P.rope + poffset + (size_t)(P.kv_head0_global + kvh) * HS + j);
        const float *qr = P.q + (size_t)m * P.attn_seg + (h0 + h) * HS;
        const float q0 = __ldcg(qr + j), q1 = __ldcg(qr + j + hp);
        qs[h * HS + j] = __fsub_rn(__fmul_rn(q0, f.x), __fmul_rn(q1, f.y));
        qs[h * HS + j + hp] = __fadd_rn(__fmul_rn(q0, f.y), __fmul_rn(q1, f.x));
    }
    // the split that contains `pos` rotates the key, appends key and value to the page (:230-243,279-285)
    // and keeps both in shared memory for its own scores
    if (owner) {
        for (int j = MG_CONSUMERS - 1 - tid; j < hp; j += MG_CONSUMERS) { // use the warps the q loop leaves idle
            const float2 f = __ldg((const float2 *)P.rope + poffset + (size_t)(P.kv_head0_global + kvh) * HS + j);
            const float *kr = P.k + (size_t)m * P.kv_seg + xoff, *vr = P.v + (size_t)m * P.kv_seg + xoff;
            const float k0 = __ldcg(kr + j), k1 = __ldcg(kr + j + hp);
            const float v0 = __ldcg(vr + j), v1 = __ldcg(vr + j + hp);
            float r0 = __fsub_rn(__fmul_rn(k0, f.x), __fmul_rn(k1, f.y));
            float r1 = __fadd_rn(__fmul_rn(k0, f.y), __fmul_rn(k1, f.x));
            char *krow = (char *)mg_kv_row(P.kv, session, layer, pos, 0);
            char *vrow = (char *)mg_kv_row(P.kv, session, layer, pos, 1);
            if (dt == JL_F32) {
                ((float *)krow)[xoff + j] = r0, ((float *)krow)[xoff + j + hp] = r1;
                ((float *)vrow)[xoff + j] = v0, ((float *)vrow)[xoff + j + hp] = v1;
                knew[j] = r0, knew[j + hp] = r1, vnew[j] = v0, vnew[j + hp] = v1;
            } else { // the scores see the values as stored (bf16-rounded), like a later read of the page would
                const uint16_t b0 = mg_bf16(r0), b1 = mg_bf16(r1), c0 = mg_bf16(v0), c1 = mg_bf16(v1);
                ((uint16_t *)krow)[xoff + j] = b0, ((uint16_t *)krow)[xoff + j + hp] = b1;
                ((uint16_t *)vrow)[xoff + j] = c0, ((uint16_t *)vrow)[xoff + j + hp] = c1;
                knew[j] = bf16_bits_to_f32(b0), knew[j + hp] = bf16_bits_to_f32(b1);
                vnew[j] = bf16_bits_to_f32(c0), vnew[j + hp] = bf16_bits_to_f32(c1);
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
        // registers -> shared (the current position's row comes from knew/vnew)
#pragma unroll
        for (int i = 0; i < NF; i++) {
            const int f = tid + i * MG_CONSUMERS;
            const int r = f / C4, c4 = f % C4;
            if (f < MG_ATT_TILE * C4 && r < cnt) {
                float4 k4 = kreg[i], v4 = vreg[i];
                if (tb + r == pos) {
                    k4 = *(const float4 *)(knew + c4 * 4);
                    v4 = *(const float4 *)(vnew + c4 * 4);
                }
                Ks[r * C4 + (c4 ^ (r & 7))] = k4;
                Vs[r * C4 + c4] = v4;
            }
        }
        if (tb + MG_ATT_TILE < t1) fetch_tile(tb + MG_ATT_TILE); // in flight while this tile is processed
        consumer_bar();
        // scores (batchDotProduct :324-330, scale :332): 4 threads per (head, position), 1/4 of the head each
        for (int idx = tid; idx < group * MG_ATT_TILE * 4; idx += MG_CONSUMERS) {
            const int qd = idx & 3, t = (idx >> 2) % MG_ATT_TILE, h = idx / (4 * MG_ATT_TILE);
            float a = 0.0f;
            if (t < cnt) {
                const float4 *q4 = (const float4 *)(qs + h * HS);
#pragma unroll
                for (int c = 0; c < C4 / 4; c++) {
                    const int c4 = qd * (C4 / 4) + c;
                    const float4 k4 = Ks[t * C4 + (c4 ^ (t & 7))];
                    const float4 qq = q4[c4];
                    a = fmaf(qq.x, k4.x, a);
                    a = fmaf(qq.y, k4.y, a);
                    a = fmaf(qq.z, k4.z, a);
                    a = fmaf(qq.w, k4.w, a);
                }
            }
            a += __shfl_xor_sync(0xffffffffu, a, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2);
            if (qd == 0) ps[t * MG_MAX_GROUP + h] = t < cnt ? __fmul_rn(a, P.attn_scale) : -INFINITY;
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

#define MG_TRACE(slot_)                                                                     \
    do {                                                                                    \
        if (tr && tid == 0) tr[(size_t)op * 8 + (slot_)] = clock64();                      \
    } while (0)

template <int MM, int NSTAGE>
__global__ void __launch_bounds__(MG_THREADS, 1) mega_decode_kernel(const MegaParams P) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *ring = smem;
    unsigned char *uarea = smem + (size_t)NSTAGE * MG_STAGE_BYTES;
    __shared__ uint64_t full[NSTAGE], empty[NSTAGE];
    __shared__ double red[MM * MG_CWARPS];
    __shared__ unsigned long long wbest[MM][MG_CWARPS];
    __shared__ int s_last;
    __shared__ float ebuf[MG_EROWS * 2 * 8 * MM];
    __shared__ int emeta[MG_EROWS];
    __shared__ int s_op_first[MG_MAX_OPS + 1];
    __shared__ __align__(16) MegaMeta sdesc[NSTAGE];
    __shared__ __align__(16) unsigned char dbuf[2 * MG_DB * MG_REC_BYTES];
    __shared__ uint64_t dfull[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], MG_CWARPS);
        }
        mbar_init(&dfull[0], 1);
        mbar_init(&dfull[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == MG_CWARPS) { // ===== producer warp =====
        producer_loop<NSTAGE>(P, ring, full, empty, sdesc, dbuf, dfull, lane);
        return;
    }

    // ===== consumers =====
    unsigned *sync = P.sync;
    const int n_ops = P.layers * 4 + 1;
    for (int i = tid; i <= n_ops; i += MG_CONSUMERS) s_op_first[i] = __ldg(&P.op_first[(size_t)cta * (n_ops + 1) + i]);
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

    Stage d;
    unsigned cit = 0;
    long long *tr = nullptr;
    if (P.trace) {
        const int which = cta == 0 ? 0 : (cta == G / 2 ? 1 : (cta == G - 1 ? 2 : -1));
        if (which >= 0) tr = P.trace + (size_t)which * (P.layers * 4 + 1) * 8;
    }
    float best_v[MM];
    int best_i[MM];
#pragma unroll
    for (int m = 0; m < MM; m++) best_v[m] = -INFINITY, best_i[m] = 0x7fffffff;

    for (int op = 0; op < n_ops; op++) {
        struct {
            int type, K, pair;
        } oi;
        oi.type = op < P.layers * 4 ? (op & 3) : OP_LMHEAD;
        oi.pair = oi.type == OP_GATEUP;
        oi.K = oi.type == OP_O ? P.attn_seg : (oi.type == OP_DOWN ? P.H : P.E);
        const int L = op < P.layers * 4 ? (op >> 2) : P.layers;
        unsigned *cnt = &sync[1 + (op < P.layers * 4 ? L * 5 + (oi.type == OP_QKV ? 0 : oi.type == OP_O ? 2 : oi.type == OP_GATEUP ? 3 : 4)
                                                      : P.layers * 5)];
        MG_TRACE(0);
        // ---- attention phase sits between QKV and O ----
        if (oi.type == OP_O && !(P.dbg & 2)) {
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
        MG_TRACE(1);
        // ---- does this CTA own rows of the op? ----
        const void *nw_attn = nullptr, *nw_ffn = nullptr;
        int nw_attn_dt = 0, nw_ffn_dt = 0;
        if (op < P.layers * 4) {
            nw_attn = P.lw[L].attn_norm, nw_attn_dt = P.lw[L].attn_norm_dt;
            nw_ffn = P.lw[L].ffn_norm, nw_ffn_dt = P.lw[L].ffn_norm_dt;
        }
        const int nst = s_op_first[op + 1] - s_op_first[op];
        if (nst > 0) {
            // dependency + prologue
            const int K = oi.K, nblk = K / 32;
            ActView av;
            if (P.dbg & 2) {
            } else if (oi.type == OP_QKV) {
                op_wait(L == 0 ? &sync[0] : &sync[1 + (L - 1) * 5 + 4], (unsigned)G);
                MG_TRACE(2);
                stage_acts<MM, true>(P, uarea, P.x, P.E, K, nw_attn, nw_attn_dt, red);
            } else if (oi.type == OP_O) {
                op_wait(&sync[1 + L * 5 + 1], (unsigned)(P.M * P.kv_heads));
                MG_TRACE(2);
                stage_acts<MM, true>(P, uarea, P.att, P.attn_seg, K, nullptr, 0, red);
            } else if (oi.type == OP_GATEUP) {
                op_wait(&sync[1 + L * 5 + 2], (unsigned)G);
                MG_TRACE(2);
                stage_acts<MM, true>(P, uarea, P.xb, P.E, K, nw_ffn, nw_ffn_dt, red);
            } else if (oi.type == OP_DOWN) {
                op_wait(&sync[1 + L * 5 + 3], (unsigned)G);
                MG_TRACE(2);
                stage_acts<MM, true>(P, uarea, P.h, P.H, K, nullptr, 0, red);
            } else {
                op_wait(&sync[1 + (P.layers - 1) * 5 + 4], (unsigned)G);
                MG_TRACE(2);
                stage_acts<MM, false>(P, uarea, P.x, P.E, K, P.out_norm, P.out_norm_dt, red);
            }
            av = act_view<MM>(uarea, nblk);
            // the warp's fixed K-range for this op, and (single session) its activations in registers
            const int op_wpr = MG_UNITS / (oi.pair ? 2 * stage_rows_dev(oi.K, 1) : stage_rows_dev(oi.K, 0));
            ActRegs ar;
            if (MM == 1 && oi.type != OP_LMHEAD) {
                const int nbw = nblk / op_wpr;
                load_act_regs(ar, av, nblk, (warp & (op_wpr - 1)) * nbw, nbw, lane);
            }
            MG_TRACE(3);

            int erow = 0;
            for (int st = 0; st < nst; st++) {
                const unsigned slot = cit % NSTAGE, use = cit / NSTAGE;
#define MG_ST(k_) do { if (tr && (P.dbg & 4) && cta == 0 && tid == 0 && cit < 128) P.trace[(size_t)3 * n_ops * 8 + cit * 8 + (k_)] = clock64(); } while (0)
                MG_ST(0);
                mbar_wait(&full[slot], use & 1);
                MG_ST(1);
                {
                    const MegaMeta md = sdesc[slot];
                    d.seg = md.seg, d.row0 = md.row0, d.nrows = md.nrows, d.R = md.R, d.wpr = md.wpr, d.K = md.K, d.pair = md.pair, d.type = md.type, d.wpr_shift = md.wpr_shift, d.r_shift = md.r_shift;
                }
                const unsigned char *sp = ring + (size_t)slot * MG_STAGE_BYTES;
                const WarpRow wr[2] = {warp_row(d, sp, warp), warp_row(d, sp, warp + MG_CWARPS)};
                MG_ST(2);
                float acc[2][MM];
#pragma unroll
                for (int m = 0; m < MM; m++) acc[0][m] = 0.0f, acc[1][m] = 0.0f;
                if ((wr[0].valid || wr[1].valid) && !(P.dbg & 1)) {
                    if (oi.type == OP_LMHEAD) consume2_f32<MM>(wr, av, nblk, acc, lane);
                    else consume2_q8<MM, MM == 1>(wr, av, ar, nblk, acc, lane);
                }
                MG_ST(3);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);
                MG_ST(4);
#pragma unroll
                for (int m = 0; m < MM; m++) acc[0][m] = warp_sum(acc[0][m]), acc[1][m] = warp_sum(acc[1][m]);
                MG_ST(5);
                cit++;
                if (oi.type == OP_LMHEAD) {
                    // logits + running arg-max (strict '>', lowest index wins); K = E is one unit per row
                    if (lane == 0) {
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            if (!wr[u].valid) continue;
                            const int row = d.row0 + wr[u].row;
#pragma unroll
                            for (int m = 0; m < MM; m++) {
                                if (m >= P.M) continue;
                                const float v = acc[u][m];
                                P.logits[(size_t)m * P.vocab + row] = v;
                                if (v > best_v[m] || (v == best_v[m] && row < best_i[m])) best_v[m] = v, best_i[m] = row;
                            }
                        }
                    }
                } else {
                    // park the (partial) sums in shared memory; the op-end pass combines K-ranges / gate+up and
                    // applies the epilogue for all rows at once (no barrier, load or double math per stage)
                    if (lane == 0) {
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            if (!wr[u].valid) continue;
                            const int e = erow + wr[u].row;
                            const int qd = (warp + u * MG_CWARPS) & (d.wpr - 1);
#pragma unroll
                            for (int m = 0; m < MM; m++) ebuf[(((size_t)e * 2 + wr[u].up) * 8 + qd) * MM + m] = acc[u][m];
                            if (qd == 0 && !wr[u].up) emeta[e] = (d.seg << 28) | (d.row0 + wr[u].row);
                        }
                    }
                    erow += d.nrows;
                }
            }
            if (oi.type != OP_LMHEAD) {
                consumer_bar();
                const int wpr = MG_UNITS / (oi.pair ? 2 * stage_rows_dev(oi.K, 1) : stage_rows_dev(oi.K, 0));
                for (int t = tid; t < erow * MM; t += MG_CONSUMERS) {
                    const int e = t / MM, m = t - e * MM;
                    if (m >= P.M) continue;
                    const int seg = emeta[e] >> 28, row = emeta[e] & 0x0FFFFFFF;
                    float v = ebuf[(((size_t)e * 2 + 0) * 8 + 0) * MM + m];
                    for (int qd = 1; qd < wpr; qd++) v += ebuf[(((size_t)e * 2 + 0) * 8 + qd) * MM + m];
                    switch (oi.type) {
                        case OP_QKV: {
                            float *out = seg == 0 ? P.q + (size_t)m * P.attn_seg : (seg == 1 ? P.k : P.v) + (size_t)m * P.kv_seg;
                            out[row] = v;
                        } break;
                        case OP_O: // TransformerBlock.java:185
                            P.xb[(size_t)m * P.E + row] = __fadd_rn(v, __ldcg(P.x + (size_t)m * P.E + row));
                            break;
                        case OP_GATEUP: { // MLPBlock.java:132-141
                            float u = ebuf[(((size_t)e * 2 + 1) * 8 + 0) * MM + m];
                            for (int qd = 1; qd < wpr; qd++) u += ebuf[(((size_t)e * 2 + 1) * 8 + qd) * MM + m];
                            P.h[(size_t)m * P.H + row] = __fmul_rn(silu_ref(v), u);
                        } break;
                        default: // OP_DOWN, TransformerBlock.java:203
                            P.x[(size_t)m * P.E + row] = __fadd_rn(v, __ldcg(P.xb + (size_t)m * P.E + row));
                            break;
                    }
                }
            }
        }
        MG_TRACE(4);
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
    att += (size_t)2 * hs * 4;                                  // current position's key / value row
    return (acts > att ? acts : att) + 256;
}

bool jl_mega_supported(const MegaParams &p) {
    if (p.M < 1 || p.M > MEGA_MAX_M) return false;
    if (p.head_size != 32 && p.head_size != 64 && p.head_size != 128) return false;
    if (p.heads % p.kv_heads || p.heads / p.kv_heads > MG_MAX_GROUP) return false;
    if ((p.E % 128) || (p.H % 128) || (p.attn_seg % 128)) return false; // TMA: 16-byte aligned rows of scales
    for (int K : {p.E, p.H, p.attn_seg}) {
        if (K / 2 > MG_STAGE_NIB) return false;
        int R = 16;
        while (R > 1 && R * (K / 2) > MG_STAGE_NIB) R >>= 1;
        if ((K / 32) % (MG_UNITS / R) || MG_UNITS / R > 8) return false; // K-ranges of the units sharing a row
    }
    if (p.layers * 4 + 1 > MG_MAX_OPS || p.E > 4096) return false; // lm_head rows are one warp each (K = E <= 4096)
    if (p.grid > 0) {
        const int G = p.grid;
        if ((p.attn_seg + 2 * p.kv_seg + G - 1) / G + 3 > MG_EROWS || (p.H + G - 1) / G + 1 > MG_EROWS || (p.E + G - 1) / G + 1 > MG_EROWS)
            return false;
    }
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
