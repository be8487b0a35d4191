// Block-quantised GEMV / small-M GEMM for decode:  out[m, n] = sum_k A[m,k] * W[n,k]
//
// Replaces, for M <= 8, the reference's batchDotProduct back-ends
//   I8 x Q4  : PanamaTensorOperations.java:768-1044 (GemmerI8Q4_512), vector_simd.c:261-437
//   F32 x Q4 : PanamaTensorOperations.java:289-548, vector_simd.c:770-964
// and the WebGPU GEMV native/res/gemm_i8q4_v5.wgsl:57-137, and fuses the scalar Java loops that
// surround them in the decode step: RMSNorm (RMSNorm.java:34-56), the Q8 activation quantiser
// (PanamaTensorOperations.java:1684-1723), the residual add (TransformerBlock.java:185,203) and
// SiLU*up (MLPBlock.java:132-141).
//
// HBM-bound by design: every weight byte is read exactly once with 128-bit coalesced
// ld.global.nc loads straight into registers (one 16-byte Q4 block per lane, 512 B per warp
// request), activations are staged once per CTA in shared memory in a bank-conflict-free
// [half][block][16B] layout, integer dot products use dp4a, the reduction is a warp shuffle.
//
// Decode shape of the kernel (M = 1, Q4 weights, Q8 activations; DESIGN.md "GEMV v3"): ONE 512-thread CTA per SM with
// a three-deep register ring (two 2.5 KB chunks per warp = 80 KB per SM in flight; a B200 SM sustains about 64 KB).
// More CTAs per SM only repeat the prologue: with three 256-thread CTAs the RMSNorm + Q8 quantisation ran three times
// per SM and its loads queued behind the weight requests (5 us of a 8.6 us QKV launch, tools/ktrace.py).  Here the
// first warps ("stagers") request the hidden row BEFORE any weight load, every warp then puts its first chunks in
// flight, and the stagers normalise / quantise into shared memory while the weights stream.  The gate+up launch uses
// 640 threads with a two-deep ring instead (more issue slots for the long stream); finished rows are parked one per lane
// and their epilogues run together.  DESIGN.md section 5 has the measurements behind each of these choices.
#include "jl_common.cuh"
#include <stdlib.h>

#define GEMV_THREADS 256 // generic kernels
#define GEMV_WARPS 8
// CH = 32-element blocks per lane per chunk (a chunk is CH*32 blocks = CH*1024 weights of one row), NBUF = register
// chunk buffers per warp (NBUF-1 chunks are in flight while one is being consumed)

template <int WDT, int CH>
struct WBuf {
    uint4 q[CH * (WDT == JL_I8 ? 2 : 1)];
    float s[CH];
};

// Which weight row does item-row `row` (index within the launch) map to?
__device__ __forceinline__ void seg_lookup(const GemvParams &p, int row, int &seg, int &local) {
    seg = 0;
    local = row;
#pragma unroll
    for (int i = 0; i < 2; i++) {
        if (seg == i && i + 1 < p.nseg && local >= p.seg[i].rows) {
            local -= p.seg[i].rows;
            seg = i + 1;
        }
    }
}

template <int WDT, int CH>
__device__ __forceinline__ void load_chunk(WBuf<WDT, CH> &b, const uint8_t *wrow, const float *srow, int blk0, int nblk,
                                           int lane, unsigned long long pol) {
#pragma unroll
    for (int j = 0; j < CH; j++) {
        int bi = blk0 + j * 32 + lane;
        if (bi < nblk) {
            if (WDT == JL_Q4) {
                b.q[j] = ldg_stream_u4(wrow + (size_t)bi * 16, pol);
            } else {
                b.q[2 * j] = ldg_stream_u4(wrow + (size_t)bi * 32, pol);
                b.q[2 * j + 1] = ldg_stream_u4(wrow + (size_t)bi * 32 + 16, pol);
            }
            b.s[j] = ldg_stream_f32(srow + bi, pol);
        }
    }
}

// ---- activation staging -------------------------------------------------------------------------
// Q8 layout in smem:  q[m][half][blk][16] int8, then sc[m][blk] f32, then sum[m][blk] int32
// F32 layout in smem: f[m][c4(8)][blk][4] floats
__device__ __forceinline__ size_t q8_bytes_per_row(int nblk) { return (size_t)nblk * 32 + (size_t)nblk * 8; }

// One thread owns one 32-element block: 8 independent 128-bit loads, the block max, the Q8 rounding
// and the packing all stay in registers (no shuffles), so the prologue costs about one L2 round trip.
template <bool ACTQ8, int MM>
__device__ void stage_activations(const GemvParams &p, int prologue, unsigned char *smem, int nblk) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double red[MM][GEMV_WARPS];
    __shared__ float rs_sh[MM];
    int8_t *aq = (int8_t *)smem;
    float *asc = (float *)(smem + (size_t)MM * nblk * 32);
    int *asum = (int *)(smem + (size_t)MM * nblk * 32 + (size_t)MM * nblk * 4);
    float4 *af4 = (float4 *)smem;
    const int K = p.K;
    const bool norm = (prologue == PRO_RMSNORM_QUANT || prologue == PRO_RMSNORM_F32);

    if (norm) {
        // RMSNorm.java:41-52: float products summed in double over [0, E), all rows in one pass
        double ss[MM];
#pragma unroll
        for (int m = 0; m < MM; m++) ss[m] = 0.0;
        for (int i4 = tid; i4 < K / 4; i4 += GEMV_THREADS) {
            float4 v[MM];
#pragma unroll
            for (int m = 0; m < MM; m++)
                v[m] = m < p.M ? *(const float4 *)((const float *)p.a + (size_t)m * p.lda + p.a_col_off + i4 * 4)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
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
            if (lane == 0) red[m][warp] = ss[m];
        }
        __syncthreads();
        if (tid < MM) {
            double t = 0;
            for (int w = 0; w < GEMV_WARPS; w++) t += red[tid][w];
            t /= (double)p.norm_E;
            t += (double)p.norm_eps;
            rs_sh[tid] = (float)(1.0 / sqrt(t));
        }
        __syncthreads();
    }

    for (int idx = tid; idx < MM * nblk; idx += GEMV_THREADS) {
        const int m = idx / nblk, b = idx - m * nblk;
        const bool live = m < p.M;
        if (prologue == PRO_Q8_GLOBAL) {
            uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
            float sc = 0.0f;
            if (live) {
                const int8_t *src = (const int8_t *)p.a + (size_t)m * p.lda + p.a_col_off + b * 32;
                lo = *(const uint4 *)src;
                hi = *(const uint4 *)(src + 16);
                sc = p.a_scales[(size_t)m * (p.lda / 32) + p.a_col_off / 32 + b];
            }
            int sum = 0;
            sum = __dp4a((int)lo.x, 0x01010101, sum);
            sum = __dp4a((int)lo.y, 0x01010101, sum);
            sum = __dp4a((int)lo.z, 0x01010101, sum);
            sum = __dp4a((int)lo.w, 0x01010101, sum);
            sum = __dp4a((int)hi.x, 0x01010101, sum);
            sum = __dp4a((int)hi.y, 0x01010101, sum);
            sum = __dp4a((int)hi.z, 0x01010101, sum);
            sum = __dp4a((int)hi.w, 0x01010101, sum);
            *(uint4 *)(aq + (((size_t)m * 2 + 0) * nblk + b) * 16) = lo;
            *(uint4 *)(aq + (((size_t)m * 2 + 1) * nblk + b) * 16) = hi;
            asc[m * nblk + b] = sc;
            asum[m * nblk + b] = sum;
            continue;
        }
        float v[32];
        if (!live) {
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = 0.0f;
        } else if (prologue == PRO_BF16_GLOBAL) {
            const uint4 *src = (const uint4 *)((const uint16_t *)p.a + (size_t)m * p.lda + p.a_col_off + b * 32);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint4 u = src[i];
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int t = 0; t < 4; t++) {
                    v[i * 8 + t * 2] = __uint_as_float(w[t] << 16);
                    v[i * 8 + t * 2 + 1] = __uint_as_float(w[t] & 0xffff0000u);
                }
            }
        } else {
            const float4 *src = (const float4 *)((const float *)p.a + (size_t)m * p.lda + p.a_col_off + b * 32);
            float4 x4[8];
#pragma unroll
            for (int i = 0; i < 8; i++) x4[i] = src[i];
#pragma unroll
            for (int i = 0; i < 8; i++) v[i * 4] = x4[i].x, v[i * 4 + 1] = x4[i].y, v[i * 4 + 2] = x4[i].z, v[i * 4 + 3] = x4[i].w;
            if (norm) {
                const float rsf = rs_sh[m];
                if (p.norm_w_dtype == JL_BF16) {
                    const uint16_t *wp = (const uint16_t *)p.norm_w + b * 32;
#pragma unroll
                    for (int i = 0; i < 32; i++)
                        v[i] = __fmul_rn(__fadd_rn(p.norm_adj, bf16_bits_to_f32(wp[i])), __fmul_rn(rsf, v[i]));
                } else {
                    const float4 *wp = (const float4 *)((const float *)p.norm_w + b * 32);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float4 w4 = wp[i];
                        v[i * 4] = __fmul_rn(__fadd_rn(p.norm_adj, w4.x), __fmul_rn(rsf, v[i * 4]));
                        v[i * 4 + 1] = __fmul_rn(__fadd_rn(p.norm_adj, w4.y), __fmul_rn(rsf, v[i * 4 + 1]));
                        v[i * 4 + 2] = __fmul_rn(__fadd_rn(p.norm_adj, w4.z), __fmul_rn(rsf, v[i * 4 + 2]));
                        v[i * 4 + 3] = __fmul_rn(__fadd_rn(p.norm_adj, w4.w), __fmul_rn(rsf, v[i * 4 + 3]));
                    }
                }
            }
        }
        if (ACTQ8) {
            // PanamaTensorOperations.java:1696-1710: d = max/127, q = (byte)(x*(127/max) + 0.5f), F2B truncates
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
            *(uint4 *)(aq + (((size_t)m * 2 + 0) * nblk + b) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            *(uint4 *)(aq + (((size_t)m * 2 + 1) * nblk + b) * 16) = make_uint4(w[4], w[5], w[6], w[7]);
            asc[m * nblk + b] = d;
            asum[m * nblk + b] = sum;
        } else {
#pragma unroll
            for (int c4 = 0; c4 < 8; c4++)
                af4[((size_t)m * 8 + c4) * nblk + b] = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
        }
    }
    __syncthreads();
}


// ---- fast F32 -> (RMSNorm) -> Q8 staging ------------------------------------------------------------------------
// Thread pair (t, t^1) owns one 32-element Q8 block: thread t holds the 16 contiguous floats that become one 16-byte
// half of the smem layout, so a 4096-element tile is quantised by all 256 threads at once with a single shuffle for the
// block max and one for the block sum.  For the decode case (one row, K <= 4096) the hidden row and the norm weights
// are requested together and stay in registers: the whole prologue costs one L2 round trip instead of three.
__device__ __forceinline__ void quant_half_block(const float (&v)[16], int8_t *aq, float *asc, int *asum, int m, int nblk,
                                                 int blk, int half, bool in) {
    // PanamaTensorOperations.java:1696-1710: d = max/127, q = (byte)(x*(127/max) + 0.5f), F2B truncates
    float mx = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; i++) mx = fmaxf(mx, fabsf(v[i]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    const float d = __fdiv_rn(mx, 127.0f);
    const float id = mx != 0.0f ? __fdiv_rn(127.0f, mx) : 0.0f;
    uint32_t w[4];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int q0 = (int)__fadd_rn(__fmul_rn(v[i * 4], id), 0.5f);
        const int q1 = (int)__fadd_rn(__fmul_rn(v[i * 4 + 1], id), 0.5f);
        const int q2 = (int)__fadd_rn(__fmul_rn(v[i * 4 + 2], id), 0.5f);
        const int q3 = (int)__fadd_rn(__fmul_rn(v[i * 4 + 3], id), 0.5f);
        sum += q0 + q1 + q2 + q3;
        w[i] = (uint32_t)(q0 & 0xFF) | ((uint32_t)(q1 & 0xFF) << 8) | ((uint32_t)(q2 & 0xFF) << 16) | ((uint32_t)(q3 & 0xFF) << 24);
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    if (in) {
        *(uint4 *)(aq + (((size_t)m * 2 + half) * nblk + blk) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        if (half == 0) {
            asc[m * nblk + blk] = d;
            asum[m * nblk + blk] = sum;
        }
    }
}

__device__ __forceinline__ void load16(float (&v)[16], const float *src, bool in) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float4 t = in ? *(const float4 *)(src + i * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[i * 4] = t.x, v[i * 4 + 1] = t.y, v[i * 4 + 2] = t.z, v[i * 4 + 3] = t.w;
    }
}
__device__ __forceinline__ void load16_normw(float (&v)[16], const GemvParams &p, int e0, bool in) {
    const unsigned long long pol = l2_evict_last_policy();
    if (p.norm_w_dtype == JL_BF16) {
        const uint4 *src = (const uint4 *)((const uint16_t *)p.norm_w + e0);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const uint4 u = in ? ldg_keep_u4(src + i, pol) : make_uint4(0, 0, 0, 0);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int t = 0; t < 4; t++) {
                v[i * 8 + t * 2] = __uint_as_float(w[t] << 16);
                v[i * 8 + t * 2 + 1] = __uint_as_float(w[t] & 0xffff0000u);
            }
        }
    } else {
        const uint4 *src = (const uint4 *)((const float *)p.norm_w + e0);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 u = in ? ldg_keep_u4(src + i, pol) : make_uint4(0, 0, 0, 0);
            v[i * 4] = __uint_as_float(u.x), v[i * 4 + 1] = __uint_as_float(u.y);
            v[i * 4 + 2] = __uint_as_float(u.z), v[i * 4 + 3] = __uint_as_float(u.w);
        }
    }
}

// RMSNorm.java:41-52 scale factor from the double sum of float squares.  inv_E = 1.0 / E (exact for the power-of-two
// embedding lengths of every Llama-family model; otherwise within 1 ulp(double) of the division, far below float
// resolution).  rsqrt() is within 1 ulp(double) of 1.0 / sqrt(t); after the cast to float the two agree.
__device__ __forceinline__ float rms_scale(double sumsq, double inv_E, float eps) {
    double t = sumsq * inv_E;
    t += (double)eps;
    return (float)rsqrt(t);
}

// Stagers = the first ceil(K/16/32) warps (at most all of them).  The registers were loaded by stage_q8_issue() before
// the weight loads were queued.  NT = threads per CTA.  LONG (K > 16*NT, no norm): a second register tile and a tail loop.
template <bool NORM, bool LONG>
struct StageRegs {
    float x0[16], x1[LONG ? 16 : 1], w[NORM ? 16 : 1];
};

template <bool NORM, bool LONG, int NT>
__device__ __forceinline__ void stage_q8_issue(const GemvParams &p, StageRegs<NORM, LONG> &r) {
    const int tid = threadIdx.x;
    const float *x0 = (const float *)p.a + p.a_col_off;
    const int e0 = tid * 16;
    load16(r.x0, x0 + e0, e0 < p.K);
    if constexpr (NORM) load16_normw(r.w, p, e0, e0 < p.K);
    if constexpr (LONG) load16(r.x1, x0 + (NT + tid) * 16, (NT + tid) * 16 < p.K);
}

template <bool NORM, bool LONG, int NT>
__device__ __forceinline__ void stage_q8_finish(const GemvParams &p, StageRegs<NORM, LONG> &r, unsigned char *smem, const int nblk) {
    constexpr int NWARP = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double red[NORM ? NWARP : 1];
    int8_t *aq = (int8_t *)smem;
    float *asc = (float *)(smem + (size_t)nblk * 32);
    int *asum = (int *)(smem + (size_t)nblk * 32 + (size_t)nblk * 4);
    const int K = p.K;
    const int half = tid & 1;
    const int e0 = tid * 16;
    // warps that hold part of the row (whole warps, so the pair shuffles and the named barrier see full warps)
    const int nsw = min(NWARP, (K / 16 + 31) / 32);
    if (warp < nsw) {
        if constexpr (NORM) {
            double ss = 0.0;
#pragma unroll
            for (int i = 0; i < 16; i++) ss += (double)__fmul_rn(r.x0[i], r.x0[i]); // float products, double sum
            ss = warp_sum_d(ss);
            if (lane == 0) red[warp] = ss;
            asm volatile("bar.sync 1, %0;" ::"r"(nsw * 32) : "memory");
            double t = 0.0;
            for (int i = 0; i < nsw; i++) t += red[i];
            const float rsf = rms_scale(t, p.norm_inv_E, p.norm_eps);
#pragma unroll
            for (int i = 0; i < 16; i++) r.x0[i] = __fmul_rn(__fadd_rn(p.norm_adj, r.w[i]), __fmul_rn(rsf, r.x0[i])); // RMSNorm.java:50-52
        }
        quant_half_block(r.x0, aq, asc, asum, 0, nblk, e0 >> 5, half, e0 < K);
    }
    if constexpr (LONG) {
        // second register tile, then the rest of the row tile by tile (uniform trip count: all warps)
        const int e1 = (NT + tid) * 16;
        quant_half_block(r.x1, aq, asc, asum, 0, nblk, e1 >> 5, half, e1 < K);
        const float *x0 = (const float *)p.a + p.a_col_off;
        for (int e = (2 * NT + tid) * 16; e - tid * 16 < K; e += NT * 16) {
            load16(r.x1, x0 + e, e < K);
            quant_half_block(r.x1, aq, asc, asum, 0, nblk, e >> 5, half, e < K);
        }
    }
    __syncthreads();
}

template <int MM>
__device__ void stage_q8_pairs(const GemvParams &p, const bool norm, unsigned char *smem, const int nblk) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double red[MM][GEMV_WARPS];
    int8_t *aq = (int8_t *)smem;
    float *asc = (float *)(smem + (size_t)MM * nblk * 32);
    int *asum = (int *)(smem + (size_t)MM * nblk * 32 + (size_t)MM * nblk * 4);
    const int K = p.K;
    const int half = tid & 1;
    const float *x0 = (const float *)p.a + p.a_col_off;

    float rsf[MM];
#pragma unroll
    for (int m = 0; m < MM; m++) rsf[m] = 1.0f;
    if (norm) {
        double ss[MM];
#pragma unroll
        for (int m = 0; m < MM; m++) ss[m] = 0.0;
        for (int e0 = tid * 16; e0 < K; e0 += 16 * GEMV_THREADS) {
#pragma unroll
            for (int m = 0; m < MM; m++) {
                float x[16];
                load16(x, x0 + (size_t)m * p.lda + e0, m < p.M);
#pragma unroll
                for (int i = 0; i < 16; i++) ss[m] += (double)__fmul_rn(x[i], x[i]);
            }
        }
#pragma unroll
        for (int m = 0; m < MM; m++) {
            ss[m] = warp_sum_d(ss[m]);
            if (lane == 0) red[m][warp] = ss[m];
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < MM; m++) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < GEMV_WARPS; i++) t += red[m][i];
            t /= (double)p.norm_E;
            t += (double)p.norm_eps;
            rsf[m] = (float)(1.0 / sqrt(t));
        }
    }
    // tiles of 16*GEMV_THREADS elements; the trip count is uniform so the pair shuffles see full warps
    const int ntiles = (K + 16 * GEMV_THREADS - 1) / (16 * GEMV_THREADS);
    for (int j = 0; j < ntiles; j++) {
        const int e0 = j * 16 * GEMV_THREADS + tid * 16;
        const bool in = e0 < K;
        float w[16];
        if (norm) load16_normw(w, p, e0, in);
#pragma unroll
        for (int m = 0; m < MM; m++) {
            float x[16];
            load16(x, x0 + (size_t)m * p.lda + e0, in && m < p.M);
            if (norm) {
#pragma unroll
                for (int i = 0; i < 16; i++) x[i] = __fmul_rn(__fadd_rn(p.norm_adj, w[i]), __fmul_rn(rsf[m], x[i]));
            }
            quant_half_block(x, aq, asc, asum, m, nblk, e0 >> 5, half, in);
        }
    }
    __syncthreads();
}

// ---- per-chunk math -------------------------------------------------------------------------------
template <int WDT, bool ACTQ8, int MM, int CH>
__device__ __forceinline__ void compute_chunk(const WBuf<WDT, CH> &w, float (&acc)[MM], const unsigned char *smem, int blk0,
                                              int nblk, int lane) {
    const int8_t *aq = (const int8_t *)smem;
    const float *asc = (const float *)(smem + (size_t)MM * nblk * 32);
    const int *asum = (const int *)(smem + (size_t)MM * nblk * 32 + (size_t)MM * nblk * 4);
    const float *af = (const float *)smem;
#pragma unroll
    for (int j = 0; j < CH; j++) {
        int bi = blk0 + j * 32 + lane;
        if (bi >= nblk) continue;
        const float sb = w.s[j];
        if (ACTQ8) {
#pragma unroll
            for (int m = 0; m < MM; m++) {
                const uint4 alo = *(const uint4 *)(aq + (((size_t)m * 2 + 0) * nblk + bi) * 16);
                const uint4 ahi = *(const uint4 *)(aq + (((size_t)m * 2 + 1) * nblk + bi) * 16);
                int s = 0;
                if (WDT == JL_Q4) {
                    const uint4 q = w.q[j];
                    s = __dp4a((int)(q.x & 0x0F0F0F0Fu), (int)alo.x, s);
                    s = __dp4a((int)((q.x >> 4) & 0x0F0F0F0Fu), (int)ahi.x, s);
                    s = __dp4a((int)(q.y & 0x0F0F0F0Fu), (int)alo.y, s);
                    s = __dp4a((int)((q.y >> 4) & 0x0F0F0F0Fu), (int)ahi.y, s);
                    s = __dp4a((int)(q.z & 0x0F0F0F0Fu), (int)alo.z, s);
                    s = __dp4a((int)((q.z >> 4) & 0x0F0F0F0Fu), (int)ahi.z, s);
                    s = __dp4a((int)(q.w & 0x0F0F0F0Fu), (int)alo.w, s);
                    s = __dp4a((int)((q.w >> 4) & 0x0F0F0F0Fu), (int)ahi.w, s);
                    s -= 8 * asum[m * nblk + bi]; // sum a*(nib-8) = sum a*nib - 8*sum a   (exact)
                } else {
                    const uint4 q0 = w.q[2 * j], q1 = w.q[2 * j + 1];
                    s = __dp4a((int)q0.x, (int)alo.x, s);
                    s = __dp4a((int)q0.y, (int)alo.y, s);
                    s = __dp4a((int)q0.z, (int)alo.z, s);
                    s = __dp4a((int)q0.w, (int)alo.w, s);
                    s = __dp4a((int)q1.x, (int)ahi.x, s);
                    s = __dp4a((int)q1.y, (int)ahi.y, s);
                    s = __dp4a((int)q1.z, (int)ahi.z, s);
                    s = __dp4a((int)q1.w, (int)ahi.w, s);
                }
                // acc += (sa*sb) * isum   (vector_simd.c:384-420)
                acc[m] = fmaf(__fmul_rn(asc[m * nblk + bi], sb), (float)s, acc[m]);
            }
        } else {
            // F32 activations: acc += sb * sum_j a_j * (w_j)   with w_j = nib-8 (Q4) or int8 (I8)
            float wf[32];
            if (WDT == JL_Q4) {
                const uint4 q = w.q[j];
                const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t lo = qw[i] & 0x0F0F0F0Fu, hi = (qw[i] >> 4) & 0x0F0F0F0Fu;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        // byte t -> float(2^23 + nib) via PRMT, minus (2^23 + 8)
                        wf[i * 4 + t] = __uint_as_float(__byte_perm(lo, 0x4B000000u, 0x7540 | t)) - 8388616.0f;
                        wf[16 + i * 4 + t] = __uint_as_float(__byte_perm(hi, 0x4B000000u, 0x7540 | t)) - 8388616.0f;
                    }
                }
            } else {
                const uint4 q0 = w.q[2 * j], q1 = w.q[2 * j + 1];
                const uint32_t qw[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                for (int i = 0; i < 8; i++)
#pragma unroll
                    for (int t = 0; t < 4; t++) wf[i * 4 + t] = (float)(int)(int8_t)((qw[i] >> (8 * t)) & 0xFF);
            }
#pragma unroll
            for (int m = 0; m < MM; m++) {
                float part = 0.0f;
#pragma unroll
                for (int c4 = 0; c4 < 8; c4++) {
                    const float4 a4 = *(const float4 *)(af + (((size_t)m * 8 + c4) * nblk + bi) * 4);
                    part = fmaf(a4.x, wf[c4 * 4 + 0], part);
                    part = fmaf(a4.y, wf[c4 * 4 + 1], part);
                    part = fmaf(a4.z, wf[c4 * 4 + 2], part);
                    part = fmaf(a4.w, wf[c4 * 4 + 3], part);
                }
                acc[m] = fmaf(sb, part, acc[m]);
            }
        }
    }
}

// End of an output row.
template <int EPI, int MM>
__device__ __forceinline__ void finish_row_fn(const GemvParams &p, int rr, int wrr, float (&acc)[MM], float (&gate)[MM]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int m = 0; m < MM; m++) acc[m] = warp_sum(acc[m]);
    if (EPI == EPI_SILU_MUL && wrr == 0) {
#pragma unroll
        for (int m = 0; m < MM; m++) gate[m] = acc[m], acc[m] = 0.0f;
        return;
    }
    if (lane == 0) {
        int seg = 0, local = rr;
        if (EPI != EPI_SILU_MUL) seg_lookup(p, rr, seg, local);
        const GemvSeg &sg = p.seg[seg];
        const int col = p.row0 + local + sg.out_off;
#pragma unroll
        for (int m = 0; m < MM; m++) {
            if (m < p.M) {
                float v = acc[m];
                if (EPI == EPI_ADD_RESIDUAL) v = __fadd_rn(v, p.residual[(size_t)m * p.res_ld + (p.row0 + local)]);
                if (EPI == EPI_SILU_MUL) v = __fmul_rn(silu_ref(gate[m]), v);
                sg.out[(size_t)m * sg.out_ld + col] = v;
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MM; m++) acc[m] = 0.0f;
}

// ---- generic kernel: any M <= 8, any prologue (run time), 256 threads -------------------------------------------
template <int WDT, bool ACTQ8, int EPI, int MM, int CH, int NBUF, int MINB>
__global__ void __launch_bounds__(GEMV_THREADS, MINB) gemv_kernel(const GemvParams p, const int prologue) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = p.K / 32;
    const int nchunks = (nblk + 32 * CH - 1) / (32 * CH);
    constexpr int NW = (EPI == EPI_SILU_MUL) ? 2 : 1; // weight rows per output row
    const int wbytes_per_blk = (WDT == JL_Q4) ? 16 : 32;
    ktrace_begin(p.trace, 0x100u | (unsigned)EPI | ((unsigned long long)prologue << 4) | ((unsigned long long)p.total_rows << 16) |
                              ((unsigned long long)p.K << 40));

    // balanced static partition of output rows over all warps of the grid
    const long long gw = (long long)blockIdx.x * GEMV_WARPS + warp;
    const long long tw = (long long)gridDim.x * GEMV_WARPS;
    const int r0 = (int)(((long long)p.total_rows * gw) / tw);
    const int r1 = (int)(((long long)p.total_rows * (gw + 1)) / tw);

    // item iterator: (row r, weight-row wr, chunk c)
    struct It {
        int r, wr, c;
    };
    auto row_ptrs = [&](const It &it, const uint8_t *&wrow, const float *&srow) {
        int seg, local;
        if (EPI == EPI_SILU_MUL) {
            seg = it.wr;
            local = it.r;
        } else {
            seg_lookup(p, it.r, seg, local);
        }
        const GemvSeg &sg = p.seg[seg];
        const size_t grow = (size_t)(p.row0 + local);
        wrow = (const uint8_t *)sg.w + (grow * (size_t)(p.ldw / 32) + (size_t)(p.w_col_off / 32)) * wbytes_per_blk;
        srow = sg.ws + grow * (size_t)(p.ldw / 32) + (size_t)(p.w_col_off / 32);
    };
    auto advance = [&](It &it) {
        if (++it.c == nchunks) {
            it.c = 0;
            if (++it.wr == NW) {
                it.wr = 0;
                ++it.r;
            }
        }
    };

    WBuf<WDT, CH> buf[NBUF];
    const uint8_t *wrow;
    const float *srow;
    const unsigned long long pol = l2_evict_first_policy();
    It cur = {r0, 0, 0}, ld = cur;
    // NBUF-1 chunks per warp go in flight before anything else: the weight stream does not depend on the previous kernel
#pragma unroll
    for (int b = 0; b < NBUF - 1; b++) {
        if (ld.r < r1) {
            row_ptrs(ld, wrow, srow);
            load_chunk<WDT, CH>(buf[b], wrow, srow, ld.c * 32 * CH, nblk, lane, pol);
            advance(ld);
        }
    }
    pdl_launch_dependents();
    pdl_wait(); // activations are produced by the previous kernel
    if (ACTQ8 && (prologue == PRO_F32_QUANT || prologue == PRO_RMSNORM_QUANT))
        stage_q8_pairs<MM>(p, prologue == PRO_RMSNORM_QUANT, smem, nblk);
    else
        stage_activations<ACTQ8, MM>(p, prologue, smem, nblk);
    ktrace_mid(p.trace);

    float acc[MM];
    float gate[MM];
#pragma unroll
    for (int m = 0; m < MM; m++) acc[m] = 0.0f, gate[m] = 0.0f;

    while (cur.r < r1) {
#pragma unroll
        for (int b = 0; b < NBUF; b++) {
            if (cur.r < r1) {
                if (ld.r < r1) {
                    row_ptrs(ld, wrow, srow);
                    load_chunk<WDT, CH>(buf[(b + NBUF - 1) % NBUF], wrow, srow, ld.c * 32 * CH, nblk, lane, pol);
                    advance(ld);
                }
                compute_chunk<WDT, ACTQ8, MM, CH>(buf[b], acc, smem, cur.c * 32 * CH, nblk, lane);
                if (cur.c == nchunks - 1) finish_row_fn<EPI, MM>(p, cur.r, cur.wr, acc, gate);
                advance(cur);
            }
        }
    }
    if (p.trace) {
        __syncthreads();
        ktrace_end(p.trace);
    }
}

// ---- decode kernel: M = 1, Q8 activations, prologue fixed at compile time, one big CTA per SM -----------------------
// Work split: output rows are dealt to the CTAs; inside a CTA the (row, weight-row, chunk) items are dealt to the warps.
// LONG = false (a weight row is one chunk, K <= 1024*CH): whole rows go to warps, a warp finishes its rows alone.
// LONG = true (down_proj: K = 14336 = 3.5 chunks): rows are split at chunk granularity, so every warp streams the same
// number of bytes (+-1 chunk) however few rows the CTA has; a warp leaves one partial sum per (row, weight-row) it
// touched in shared memory and, after a barrier, one thread per output row adds the partials in chunk order
// (deterministic).  The kernel is kept small on purpose: five different kernels alternate every ~50 us, each starts
// with a cold instruction cache, and every kilobyte on the prologue path showed up in the timeline (tools/ktrace.py).
__device__ __forceinline__ int item_owner(int i, int items, int nwarp) { return (int)((((long long)(i + 1)) * nwarp - 1) / items); }

template <int WDT, int EPI, int PRO, int CH, int NBUF, int NT, bool PDL, bool LONG>
__global__ void __launch_bounds__(NT, 1) gemv_decode_kernel(const GemvParams p, const int) {
    extern __shared__ __align__(16) unsigned char smem[];
    constexpr int NWARP = NT / 32;
    constexpr bool NORM = PRO == PRO_RMSNORM_QUANT;
    constexpr int NW = (EPI == EPI_SILU_MUL) ? 2 : 1; // weight rows per output row
    constexpr int WB = (WDT == JL_Q4) ? 16 : 32;      // weight bytes per 32-element block
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nblk = p.K / 32;
    const int nchunks = LONG ? (nblk + 32 * CH - 1) / (32 * CH) : 1;
    ktrace_begin(p.trace, 0x100u | (unsigned)EPI | ((unsigned long long)PRO << 4) | ((unsigned long long)p.total_rows << 16) |
                              ((unsigned long long)p.K << 40));

    // rows of this CTA, items of this warp
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
    // load cursor: (row relative to R0, weight-row, chunk) plus the byte/element offset of that weight row
    struct It {
        int r, wr, c;
    };
    const size_t row_blocks = (size_t)(p.ldw / 32), col_blocks = (size_t)(p.w_col_off / 32);
    auto row_ptrs = [&](const It &it, const uint8_t *&wrow, const float *&srow) {
        int seg, local;
        if (EPI == EPI_SILU_MUL) {
            seg = it.wr;
            local = R0 + it.r;
        } else {
            seg_lookup(p, R0 + it.r, seg, local);
        }
        const size_t blk = (size_t)(p.row0 + local) * row_blocks + col_blocks;
        wrow = (const uint8_t *)p.seg[seg].w + blk * WB;
        srow = p.seg[seg].ws + blk;
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

    StageRegs<NORM, LONG && !NORM> sr;
    WBuf<WDT, CH> buf[NBUF];
    const uint8_t *wrow;
    const float *srow;
    const unsigned long long pol = l2_evict_first_policy();
    int ci = i0, li = i0;
    if (PDL) {
        // programmatic dependent launch: this CTA may start while the producer of the hidden row is still running, so the
        // weight stream goes first and the row is only touched after the wait
#pragma unroll
        for (int b = 0; b < NBUF - 1; b++) {
            if (li < i1) {
                row_ptrs(ld, wrow, srow);
                load_chunk<WDT, CH>(buf[b], wrow, srow, ld.c * 32 * CH, nblk, lane, pol);
                advance(ld);
                li++;
            }
        }
        pdl_launch_dependents();
        pdl_wait();
        stage_q8_issue<NORM, LONG && !NORM, NT>(p, sr);
    } else {
        // the hidden row (and the norm weights) are requested before this SM's ~80 KB of weight requests
        stage_q8_issue<NORM, LONG && !NORM, NT>(p, sr);
#pragma unroll
        for (int b = 0; b < NBUF - 1; b++) {
            if (li < i1) {
                row_ptrs(ld, wrow, srow);
                load_chunk<WDT, CH>(buf[b], wrow, srow, ld.c * 32 * CH, nblk, lane, pol);
                advance(ld);
                li++;
            }
        }
    }
    stage_q8_finish<NORM, LONG && !NORM, NT>(p, sr, smem, nblk);
    ktrace_mid(p.trace);

    float *parts = (float *)(smem + (((size_t)nblk * 40 + 15) & ~(size_t)15));
    const int maxsplit = nchunks < NWARP ? nchunks : NWARP;
    float acc[1] = {0.0f};
    // !LONG: finished rows are parked one per lane (lane i holds the i-th finished row of this warp) and the epilogue --
    // residual add or the double-precision SiLU -- runs for all parked rows at once with coalesced stores, instead of
    // lane 0 doing it serially after every row (gate+up: 18.7 vs 19.9 us per launch)
    float park0 = 0.0f, park1 = 0.0f; // value (or gate), up
    int nparked = 0, park_row0 = R0 + cur.r;
    auto flush = [&]() {
        if (lane < nparked) {
            int seg = 0, local = park_row0 + lane;
            if (EPI != EPI_SILU_MUL) seg_lookup(p, park_row0 + lane, seg, local);
            const GemvSeg &sg = p.seg[seg];
            float v = NW == 2 ? park1 : park0;
            if (EPI == EPI_ADD_RESIDUAL) v = __fadd_rn(v, p.residual[p.row0 + local]);
            if (EPI == EPI_SILU_MUL) v = __fmul_rn(silu_ref(park0), v);
            sg.out[p.row0 + local + sg.out_off] = v;
        }
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
                    const float v = warp_sum(acc[0]); // every lane has the total
                    acc[0] = 0.0f;
                    if (NW == 2 && cur.wr == 0) {
                        if (lane == nparked) park0 = v;
                    } else {
                        if (lane == nparked) (NW == 2 ? park1 : park0) = v;
                        ++nparked;
                    }
                } else if (cur.c == nchunks - 1 || ci == i1 - 1) {
                    // partial of (row, weight-row) from this warp
                    const float v = warp_sum(acc[0]);
                    const int rw = cur.r * NW + cur.wr;
                    if (lane == 0) parts[rw * maxsplit + (warp - item_owner(rw * nchunks, items, NWARP))] = v;
                    acc[0] = 0.0f;
                }
                advance(cur);
                ci++;
            }
        }
        if (!LONG && nparked > 32 - NBUF) flush(); // checked once per ring revolution: at most NBUF rows arrive in between
    }
    if (!LONG) flush();
    if (LONG) {
        __syncthreads();
        for (int o = tid; o < nrows; o += NT) {
            float sums[NW];
#pragma unroll
            for (int wr = 0; wr < NW; wr++) {
                const int rw = o * NW + wr;
                const int wa = item_owner(rw * nchunks, items, NWARP), wb = item_owner(rw * nchunks + nchunks - 1, items, NWARP);
                float t = 0.0f;
                for (int w = wa; w <= wb; w++) t = __fadd_rn(t, parts[rw * maxsplit + (w - wa)]);
                sums[wr] = t;
            }
            int seg = 0, local = R0 + o;
            if (EPI != EPI_SILU_MUL) seg_lookup(p, R0 + o, seg, local);
            const GemvSeg &sg = p.seg[seg];
            const int col = p.row0 + local + sg.out_off;
            float v = sums[NW - 1];
            if (EPI == EPI_ADD_RESIDUAL) v = __fadd_rn(v, p.residual[p.row0 + local]);
            if (EPI == EPI_SILU_MUL) v = __fmul_rn(silu_ref(sums[0]), v);
            sg.out[col] = v;
        }
    }
    if (p.trace) {
        __syncthreads();
        ktrace_end(p.trace);
    }
}

// ---- dense (F32 / BF16 weights) GEMV: GemmerF32 / GemmerF32BF16 (PanamaTensorOperations.java:1046-1231,1466-1539)
template <int WDT, int EPI, int MM>
__global__ void __launch_bounds__(GEMV_THREADS) gemv_dense_kernel(const GemvParams p, const int prologue) {
    extern __shared__ __align__(16) unsigned char smem[];
    float *af = (float *)smem; // [MM][K]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ double red[GEMV_WARPS];
    __shared__ float rs_sh;
    pdl_launch_dependents();
    pdl_wait();
    const int K = p.K;
    for (int m = 0; m < MM; m++) {
        const bool live = m < p.M;
        float rsf = 1.0f;
        const bool norm = (prologue == PRO_RMSNORM_F32);
        const float *x = (const float *)p.a + (size_t)m * p.lda + p.a_col_off;
        if (norm && live) {
            double ss = 0.0;
            for (int i = tid; i < K; i += GEMV_THREADS) {
                float v = x[i];
                ss += (double)__fmul_rn(v, v);
            }
            ss = warp_sum_d(ss);
            if (lane == 0) red[warp] = ss;
            __syncthreads();
            if (tid == 0) {
                double t = 0;
                for (int w = 0; w < GEMV_WARPS; w++) t += red[w];
                t /= (double)p.norm_E;
                t += (double)p.norm_eps;
                rs_sh = (float)(1.0 / sqrt(t));
            }
            __syncthreads();
            rsf = rs_sh;
        }
        for (int i = tid; i < K; i += GEMV_THREADS) {
            float v = 0.0f;
            if (live) {
                if (prologue == PRO_BF16_GLOBAL)
                    v = bf16_bits_to_f32(((const uint16_t *)p.a)[(size_t)m * p.lda + p.a_col_off + i]);
                else
                    v = x[i];
                if (norm) {
                    float w = p.norm_w_dtype == JL_BF16 ? bf16_bits_to_f32(((const uint16_t *)p.norm_w)[i])
                                                        : ((const float *)p.norm_w)[i];
                    v = __fmul_rn(__fadd_rn(p.norm_adj, w), __fmul_rn(rsf, v));
                }
            }
            af[(size_t)m * K + i] = v;
        }
    }
    __syncthreads();
    const long long gw = (long long)blockIdx.x * GEMV_WARPS + warp;
    const long long tw = (long long)gridDim.x * GEMV_WARPS;
    const int r0 = (int)(((long long)p.total_rows * gw) / tw);
    const int r1 = (int)(((long long)p.total_rows * (gw + 1)) / tw);
    for (int r = r0; r < r1; r++) {
        int seg, local;
        seg_lookup(p, r, seg, local);
        const GemvSeg &sg = p.seg[seg];
        const size_t grow = (size_t)(p.row0 + local);
        float acc[MM];
#pragma unroll
        for (int m = 0; m < MM; m++) acc[m] = 0.0f;
        if (WDT == JL_F32) {
            const float *wrow = (const float *)sg.w + grow * (size_t)p.ldw + p.w_col_off;
            for (int i = lane; i < K; i += 32) {
                float w = ldg_nc_f32(wrow + i);
#pragma unroll
                for (int m = 0; m < MM; m++) acc[m] = fmaf(af[(size_t)m * K + i], w, acc[m]);
            }
        } else {
            const uint16_t *wrow = (const uint16_t *)sg.w + grow * (size_t)p.ldw + p.w_col_off;
            for (int i = lane; i < K; i += 32) {
                float w = bf16_bits_to_f32(wrow[i]);
#pragma unroll
                for (int m = 0; m < MM; m++) acc[m] = fmaf(af[(size_t)m * K + i], w, acc[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < MM; m++) acc[m] = warp_sum(acc[m]);
        if (lane == 0) {
            const int col = p.row0 + local + sg.out_off;
#pragma unroll
            for (int m = 0; m < MM; m++)
                if (m < p.M) {
                    float v = acc[m];
                    if (EPI == EPI_ADD_RESIDUAL) v = __fadd_rn(v, p.residual[(size_t)m * p.res_ld + (p.row0 + local)]);
                    sg.out[(size_t)m * sg.out_ld + col] = v;
                }
        }
    }
}

// ---- host launcher ---------------------------------------------------------------------------------
template <typename KERN>
static int launch_kern(jl_ctx *ctx, cudaStream_t stream, KERN kern, const GemvParams &p, int prologue, bool pdl, int grid,
                       int threads, size_t smem, size_t (&configured)[JL_MAX_DEVICES]) {
    JL_CUDA_CHECK(ctx, jl_ensure_dyn_smem(kern, ctx->device, smem, configured));
    JL_CUDA_CHECK(ctx, jl_launch_kernel(kern, dim3(grid), dim3(threads), smem, stream, pdl, p, prologue));
    ctx->launches++;
    return JL_OK;
}

// grid for the generic 256-thread kernels: a multiple of the SM count; every warp gets >= 1 row when possible
static int generic_grid(jl_ctx *ctx, int rows, int max_per_sm) {
    int per_sm = rows >= ctx->sm_count * GEMV_WARPS * 4 ? 3 : (rows >= ctx->sm_count * GEMV_WARPS * 2 ? 2 : 1);
    if (per_sm > max_per_sm) per_sm = max_per_sm;
    int grid = ctx->sm_count * per_sm;
    const int max_grid = (rows + GEMV_WARPS - 1) / GEMV_WARPS;
    if (grid > max_grid) grid = max_grid;
    return grid < 1 ? 1 : grid;
}

// decode hot path: one big CTA per SM, prologue fixed at compile time
template <int WDT, int EPI, int PRO, int NT, int NBUF, int CH, bool PDL, bool LONG>
static int launch_decode_k(jl_ctx *ctx, cudaStream_t stream, const GemvParams &p, int grid, size_t smem) {
    static size_t configured[JL_MAX_DEVICES] = {};
    return launch_kern(ctx, stream, gemv_decode_kernel<WDT, EPI, PRO, CH, NBUF, NT, PDL, LONG>, p, 0, PDL, grid, NT, smem, configured);
}

template <int WDT, int EPI, int PRO, int NT, int NBUF, int CH>
static int launch_decode(jl_ctx *ctx, cudaStream_t stream, const GemvParams &p, bool pdl, size_t smem) {
    if (PRO == PRO_RMSNORM_QUANT && p.K > 16 * NT)
        return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "gemv: RMSNorm prologue supports rows up to %d", 16 * NT);
    int grid = ctx->sm_count;
    const int max_grid = (p.total_rows + NT / 32 - 1) / (NT / 32);
    if (grid > max_grid) grid = max_grid;
    const int nchunks = (p.K / 32 + 32 * CH - 1) / (32 * CH);
    const bool lng = nchunks > 1;
    size_t smem_d = ((smem + 15) & ~(size_t)15);
    if (lng) { // partial sums of split rows live behind the staged activations
        const int rows_per_cta = (p.total_rows + grid - 1) / grid;
        const int maxsplit = nchunks < NT / 32 ? nchunks : NT / 32;
        smem_d += (size_t)rows_per_cta * (EPI == EPI_SILU_MUL ? 2 : 1) * maxsplit * 4;
    }
    if (smem_d > 200 * 1024) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "gemv: shape exceeds shared memory");
    if (pdl) {
        if (lng) return launch_decode_k<WDT, EPI, PRO, NT, NBUF, CH, true, true>(ctx, stream, p, grid, smem_d);
        return launch_decode_k<WDT, EPI, PRO, NT, NBUF, CH, true, false>(ctx, stream, p, grid, smem_d);
    }
    if (lng) return launch_decode_k<WDT, EPI, PRO, NT, NBUF, CH, false, true>(ctx, stream, p, grid, smem_d);
    return launch_decode_k<WDT, EPI, PRO, NT, NBUF, CH, false, false>(ctx, stream, p, grid, smem_d);
}

template <int WDT, bool ACTQ8, int EPI, int MM>
static int launch_q(jl_ctx *ctx, cudaStream_t stream, const GemvParams &p, int prologue, bool pdl, size_t smem) {
    if constexpr (WDT == JL_Q4 && ACTQ8 && MM == 1) {
#ifdef JL_DECODE_NT
        constexpr int NT = JL_DECODE_NT, NB = JL_DECODE_NBUF; // diagnostic builds (tools/build_variant.sh)
#else
        // CTA shape per launch kind, from A/B timelines on the 8B step (tools/ktrace.py; us per launch QKV / o / gate+up / down):
        //   512 threads x 3-deep ring 5.92 / 4.43 / 16.58 / 10.41      640 x 2: 6.28 / 4.53 / 15.38 / 10.53
        //   704 x 2: 5.98 / 4.62 / 15.93 / 11.10    768 x 2: 6.17 / 4.77 / 16.07 / 11.12    384 x 4: 6.32 / 4.75 / 17.01 / 11.42
        // The long gate+up stream wants warps (issue slots), the short launches want the deeper ring.
        constexpr int NT = EPI == EPI_SILU_MUL ? 640 : 512, NB = EPI == EPI_SILU_MUL ? 2 : 3;
#endif
        if (prologue == PRO_F32_QUANT) return launch_decode<WDT, EPI, PRO_F32_QUANT, NT, NB, 4>(ctx, stream, p, pdl, smem);
        if (prologue == PRO_RMSNORM_QUANT && p.K <= 16 * NT)
            return launch_decode<WDT, EPI, PRO_RMSNORM_QUANT, NT, NB, 4>(ctx, stream, p, pdl, smem);
    }
    constexpr int MINB = (MM <= 2) ? 2 : 1;
    static size_t cfg[JL_MAX_DEVICES] = {};
    return launch_kern(ctx, stream, gemv_kernel<WDT, ACTQ8, EPI, MM, 4, 2, MINB>, p, prologue, pdl,
                       generic_grid(ctx, p.total_rows, MINB), GEMV_THREADS, smem, cfg);
}

template <int WDT, bool ACTQ8, int EPI>
static int launch_m(jl_ctx *ctx, cudaStream_t stream, const GemvParams &p, int prologue, bool pdl, size_t smem1) {
    if (p.M <= 1) return launch_q<WDT, ACTQ8, EPI, 1>(ctx, stream, p, prologue, pdl, smem1);
    if (p.M <= 2) return launch_q<WDT, ACTQ8, EPI, 2>(ctx, stream, p, prologue, pdl, smem1 * 2);
    if (p.M <= 4) return launch_q<WDT, ACTQ8, EPI, 4>(ctx, stream, p, prologue, pdl, smem1 * 4);
    return launch_q<WDT, ACTQ8, EPI, 8>(ctx, stream, p, prologue, pdl, smem1 * 8);
}

template <int WDT, bool ACTQ8>
static int launch_e(jl_ctx *ctx, cudaStream_t stream, const GemvParams &p, int prologue, int epi, bool pdl, size_t smem1) {
    switch (epi) {
        case EPI_STORE: return launch_m<WDT, ACTQ8, EPI_STORE>(ctx, stream, p, prologue, pdl, smem1);
        case EPI_ADD_RESIDUAL: return launch_m<WDT, ACTQ8, EPI_ADD_RESIDUAL>(ctx, stream, p, prologue, pdl, smem1);
        case EPI_SILU_MUL: return launch_m<WDT, ACTQ8, EPI_SILU_MUL>(ctx, stream, p, prologue, pdl, smem1);
    }
    return jl_set_error(ctx, JL_ERR_INVALID, "bad epilogue %d", epi);
}

template <int WDT, int EPI, int MM>
static int launch_dense(jl_ctx *ctx, cudaStream_t stream, const GemvParams &p, int prologue, bool pdl, int grid) {
    auto kern = gemv_dense_kernel<WDT, EPI, MM>;
    size_t smem = (size_t)MM * p.K * 4;
    if (smem > 200 * 1024) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "dense gemv: M*K too large for shared memory");
    static size_t cfg[JL_MAX_DEVICES] = {};
    JL_CUDA_CHECK(ctx, jl_ensure_dyn_smem(kern, ctx->device, smem, cfg));
    JL_CUDA_CHECK(ctx, jl_launch_kernel(kern, dim3(grid), dim3(GEMV_THREADS), smem, stream, pdl, p, prologue));
    ctx->launches++;
    return JL_OK;
}

template <int WDT, int EPI>
static int launch_dense_m(jl_ctx *ctx, cudaStream_t stream, const GemvParams &p, int prologue, bool pdl, int grid) {
    if (p.M <= 1) return launch_dense<WDT, EPI, 1>(ctx, stream, p, prologue, pdl, grid);
    if (p.M <= 2) return launch_dense<WDT, EPI, 2>(ctx, stream, p, prologue, pdl, grid);
    if (p.M <= 4) return launch_dense<WDT, EPI, 4>(ctx, stream, p, prologue, pdl, grid);
    return launch_dense<WDT, EPI, 8>(ctx, stream, p, prologue, pdl, grid);
}

// Largest activation-row chunk (1, 2, 4 or 8) whose staged activations fit the kernels' shared-memory budget; 0 = even one
// row does not fit.  Callers split M into chunks of this size (the reference's F32 x Q4 GEMM has no such limit,
// PanamaTensorOperations.java:289-548).
int jl_gemv_max_m(int w_dtype, int prologue, int K) {
    const bool quant_w = (w_dtype == JL_Q4 || w_dtype == JL_I8);
    const bool actq8 = quant_w && (prologue == PRO_Q8_GLOBAL || prologue == PRO_F32_QUANT || prologue == PRO_RMSNORM_QUANT);
    const size_t smem1 = actq8 ? (size_t)(K / 32) * 40 : (size_t)K * 4;
    for (int mm = GEMV_MAX_M; mm >= 1; mm >>= 1)
        if (smem1 * mm <= 200 * 1024) return mm;
    return 0;
}

int jl_launch_gemv(jl_ctx *ctx, cudaStream_t stream, const GemvParams &p_in, int prologue, int epilogue, bool use_pdl) {
    GemvParams p = p_in;
    p.trace = jl_ktrace_slot(ctx);
    p.norm_inv_E = p.norm_E > 0 ? 1.0 / (double)p.norm_E : 0.0;

    if (p.M < 1 || p.M > GEMV_MAX_M) return jl_set_error(ctx, JL_ERR_INVALID, "gemv: M=%d out of range", p.M);
    if (p.K <= 0 || p.total_rows <= 0) return jl_set_error(ctx, JL_ERR_INVALID, "gemv: empty problem");
    const bool quant_w = (p.w_dtype == JL_Q4 || p.w_dtype == JL_I8);
    const int rows = p.total_rows;
    const int grid = generic_grid(ctx, rows, 3);

    if (!quant_w) {
        if (prologue != PRO_F32 && prologue != PRO_RMSNORM_F32 && prologue != PRO_BF16_GLOBAL)
            return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "dense weights need f32/bf16 activations");
        if (epilogue == EPI_SILU_MUL) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "dense silu epilogue");
        if (p.w_dtype == JL_F32)
            return epilogue == EPI_STORE ? launch_dense_m<JL_F32, EPI_STORE>(ctx, stream, p, prologue, use_pdl, grid)
                                         : launch_dense_m<JL_F32, EPI_ADD_RESIDUAL>(ctx, stream, p, prologue, use_pdl, grid);
        return epilogue == EPI_STORE ? launch_dense_m<JL_BF16, EPI_STORE>(ctx, stream, p, prologue, use_pdl, grid)
                                     : launch_dense_m<JL_BF16, EPI_ADD_RESIDUAL>(ctx, stream, p, prologue, use_pdl, grid);
    }
    if ((p.K % 32) || (p.w_col_off % 32) || (p.ldw % 32))
        return jl_set_error(ctx, JL_ERR_INVALID, "quantised gemv needs K, offsets and ld multiples of 32");
    const bool actq8 = (prologue == PRO_Q8_GLOBAL || prologue == PRO_F32_QUANT || prologue == PRO_RMSNORM_QUANT);
    const int nblk = p.K / 32;
    const size_t smem1 = actq8 ? (size_t)nblk * 40 : (size_t)p.K * 4;
    const int mm = p.M <= 1 ? 1 : (p.M <= 2 ? 2 : (p.M <= 4 ? 4 : 8));
    if (smem1 * mm > 200 * 1024)
        return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "gemv: activations (M=%d,K=%d) exceed shared memory", p.M, p.K);
    if (p.w_dtype == JL_Q4)
        return actq8 ? launch_e<JL_Q4, true>(ctx, stream, p, prologue, epilogue, use_pdl, smem1)
                     : launch_e<JL_Q4, false>(ctx, stream, p, prologue, epilogue, use_pdl, smem1);
    return actq8 ? launch_e<JL_I8, true>(ctx, stream, p, prologue, epilogue, use_pdl, smem1)
                 : launch_e<JL_I8, false>(ctx, stream, p, prologue, epilogue, use_pdl, smem1);
}
