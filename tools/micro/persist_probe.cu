// Micro-benchmark for the round-2 persistent decode kernel: how much of the per-launch cost of
// the kernel-per-op decode step disappears when the five ops of a layer become phases of ONE cooperative kernel?
//
// One "token" = 32 layers x {QKV 6144x4096, attention stub, O 4096x4096, gate+up 28672x4096, down 4096x14336} Q4 blocks
// (16 B nibbles + 4 B scale per 32 weights, the Llama-3-8B decode stream: 136 MB per layer).  The consumer is the
// register-ring GEMV body of jl_gemv.cu reduced to its memory pattern and instruction mix (128-bit evict-first loads,
// dp4a against Q8 activations in shared memory, warp-shuffle reduction); results are meaningless numbers.
// Variants:
//   mode 0: phase boundary = grid barrier, weights requested after the barrier     (what a kernel boundary does)
//   mode 1: the first ring chunks of the next phase are requested BEFORE waiting    (priming)
//   mode 2: mode 1 + a helper warp that pulls this CTA's future chunks into L2 with a bounded lead
// plus the cost of the bare grid barrier.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o persist_probe
// persist_probe.cu ; run on the GPU box (tools/micro/run_probe.sh).
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define NT 512
#define NWARP (NT / 32)
#define CH 4
#define NBUF 3
#define CHUNK_BLK (32 * CH)            // Q4 blocks per chunk
#define CHUNK_WB (CHUNK_BLK * 16)      // weight bytes per chunk

struct Phase {
    const uint8_t *w;   // nibbles
    const float *s;     // scales
    int chunks;         // total chunks of this phase (rows * K/4096)
    int kblk;           // blocks per row (activation row length / 32)
};
struct Params {
    Phase ph[4];
    int layers;
    size_t layer_stride_w, layer_stride_s; // bytes / floats between layers
    unsigned long long *cnt;      // barrier counters [8]
    float *act;                   // [14336] activations in global
    float *out;
    unsigned long long *stamps;   // [layers][8] CTA0 phase stamps
    int mode;
    int attn_ns;                  // attention stub duration
    int lead_kb;                  // helper lead per CTA
    int prologue;                 // 1: re-stage activations after every barrier
};

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long pol_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint4 ldg_stream(const void *p, unsigned long long pol) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float ldg_stream_f(const float *p, unsigned long long pol) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void arrive(unsigned long long *c) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(c) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long *c) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(c) : "memory");
    return v;
}
// bounded spin: a protocol bug must not hang the box (returns after ~2^27 polls)
#define SPIN_UNTIL(cond)                                   \
    do {                                                   \
        unsigned _n = 0;                                   \
        while (!(cond) && ++_n < (1u << 22)) {}            \
    } while (0)
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

struct WB {
    uint4 q[CH];
    float s[CH];
};

// stage activations: K floats -> Q8 in smem (thread pairs, as jl_gemv.cu)
__device__ void stage(const float *act, int kblk, unsigned char *smem) {
    int8_t *aq = (int8_t *)smem;
    float *asc = (float *)(smem + (size_t)kblk * 32);
    int *asum = (int *)(smem + (size_t)kblk * 36);
    const int tid = threadIdx.x;
    for (int e0 = tid * 16; e0 < kblk * 32; e0 += NT * 16) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            float4 t = __ldcg((const float4 *)(act + e0 + i * 4));
            v[i * 4] = t.x, v[i * 4 + 1] = t.y, v[i * 4 + 2] = t.z, v[i * 4 + 3] = t.w;
        }
        float mx = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) mx = fmaxf(mx, fabsf(v[i]));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        const float id = mx != 0.f ? 127.f / mx : 0.f;
        uint32_t w[4];
        int sum = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int q0 = (int)(v[i * 4] * id + .5f), q1 = (int)(v[i * 4 + 1] * id + .5f), q2 = (int)(v[i * 4 + 2] * id + .5f),
                q3 = (int)(v[i * 4 + 3] * id + .5f);
            sum += q0 + q1 + q2 + q3;
            w[i] = (q0 & 255) | ((q1 & 255) << 8) | ((q2 & 255) << 16) | ((q3 & 255) << 24);
        }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        const int blk = e0 >> 5, half = tid & 1;
        *(uint4 *)(aq + ((size_t)half * kblk + blk) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        if (!half) asc[blk] = mx / 127.f, asum[blk] = sum;
    }
}

__device__ __forceinline__ void load_chunk(WB &b, const Phase &ph, long long chunk, int lane, unsigned long long pol) {
    const uint8_t *w = ph.w + (size_t)chunk * CHUNK_WB;
    const float *s = ph.s + (size_t)chunk * CHUNK_BLK;
#pragma unroll
    for (int j = 0; j < CH; j++) {
        b.q[j] = ldg_stream(w + (size_t)(j * 32 + lane) * 16, pol);
        b.s[j] = ldg_stream_f(s + j * 32 + lane, pol);
    }
}
__device__ __forceinline__ float compute_chunk(const WB &b, const unsigned char *smem, int kblk, int blk0, int lane) {
    const int8_t *aq = (const int8_t *)smem;
    const float *asc = (const float *)(smem + (size_t)kblk * 32);
    const int *asum = (const int *)(smem + (size_t)kblk * 36);
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < CH; j++) {
        const int bi = blk0 + j * 32 + lane;
        const uint4 alo = *(const uint4 *)(aq + ((size_t)bi) * 16);
        const uint4 ahi = *(const uint4 *)(aq + ((size_t)kblk + bi) * 16);
        const uint4 q = b.q[j];
        int s = 0;
        s = __dp4a((int)(q.x & 0x0F0F0F0Fu), (int)alo.x, s);
        s = __dp4a((int)((q.x >> 4) & 0x0F0F0F0Fu), (int)ahi.x, s);
        s = __dp4a((int)(q.y & 0x0F0F0F0Fu), (int)alo.y, s);
        s = __dp4a((int)((q.y >> 4) & 0x0F0F0F0Fu), (int)ahi.y, s);
        s = __dp4a((int)(q.z & 0x0F0F0F0Fu), (int)alo.z, s);
        s = __dp4a((int)((q.z >> 4) & 0x0F0F0F0Fu), (int)ahi.z, s);
        s = __dp4a((int)(q.w & 0x0F0F0F0Fu), (int)alo.w, s);
        s = __dp4a((int)((q.w >> 4) & 0x0F0F0F0Fu), (int)ahi.w, s);
        s -= 8 * asum[bi];
        acc = fmaf(asc[bi] * b.s[j], (float)s, acc);
    }
    return acc;
}

__global__ void __launch_bounds__(NT + 32, 1) probe_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ volatile long long s_consumed; // bytes of weights consumed by this CTA so far (for the helper)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, cta = blockIdx.x;
    if (tid == 0) s_consumed = 0;
    __syncthreads();
    const unsigned long long base = P.cnt[7]; // epoch base written by the host: counters are cumulative
    if (warp == NWARP) {
        // helper: walk this CTA's chunk ranges phase by phase, keep `lead` bytes ahead of the consumers
        if (P.mode < 2) return;
        long long issued = 0;
        const long long lead = (long long)P.lead_kb * 1024;
        for (int L = 0; L < P.layers; L++)
            for (int p = 0; p < 4; p++) {
                const Phase ph = P.ph[p];
                const long long c0 = (long long)ph.chunks * cta / G, c1 = (long long)ph.chunks * (cta + 1) / G;
                const uint8_t *w = ph.w + (size_t)L * P.layer_stride_w;
                const float *s = ph.s + (size_t)L * P.layer_stride_s;
                for (long long c = c0 + lane; c - lane < c1; c += 32) {
                    { unsigned n_ = 0; while (issued - s_consumed > lead && ++n_ < (1u << 16)) __nanosleep(100); }
                    if (c < c1) {
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(w + (size_t)c * CHUNK_WB), "r"(CHUNK_WB) : "memory");
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(s + (size_t)c * CHUNK_BLK), "r"(CHUNK_BLK * 4) : "memory");
                    }
                    issued += 32LL * (CHUNK_WB + CHUNK_BLK * 4);
                }
            }
        return;
    }
    const unsigned long long pol = pol_first();
    unsigned long long nbar = 0; // barriers passed so far
    float total = 0.f;
    // initial staging
    stage(P.act, P.ph[0].kblk, smem);
    bar_consumers();
    for (int L = 0; L < P.layers; L++) {
        for (int p = 0; p < 4; p++) {
            Phase ph = P.ph[p];
            ph.w += (size_t)L * P.layer_stride_w;
            ph.s += (size_t)L * P.layer_stride_s;
            const long long c0 = (long long)ph.chunks * cta / G, c1 = (long long)ph.chunks * (cta + 1) / G;
            const int cpr = ph.kblk / CHUNK_BLK > 0 ? (ph.kblk + CHUNK_BLK - 1) / CHUNK_BLK : 1; // chunks per row (activation wrap)
            WB buf[NBUF];
            long long lc = c0 + warp, cc = c0 + warp;
            auto prime = [&]() {
#pragma unroll
                for (int b = 0; b < NBUF - 1; b++)
                    if (lc < c1) {
                        load_chunk(buf[b], ph, lc, lane, pol);
                        lc += NWARP;
                    }
            };
            if (P.mode >= 1) prime();
            // ---- dependency: grid barrier on the previous phase (+ attention stub before O) ----
            if (!(L == 0 && p == 0)) {
                if (p == 1) { // attention between QKV and O: CTAs 0..7 "compute" after the QKV barrier, everyone waits for them
                    if (cta < 8) {
                        if (tid == 0) {
                            SPIN_UNTIL(ld_acq(&P.cnt[0]) >= base + (nbar + 1) * G);
                            const unsigned long long t0 = gtimer();
                            while (gtimer() - t0 < (unsigned long long)P.attn_ns) {}
                        }
                        bar_consumers();
                        if (tid == 0) arrive(&P.cnt[1]);
                    }
                    if (tid == 0) SPIN_UNTIL(ld_acq(&P.cnt[1]) >= base / G * 8 + (unsigned long long)(L + 1) * 8);
                    nbar++;
                } else {
                    if (tid == 0) SPIN_UNTIL(ld_acq(&P.cnt[0]) >= base + (nbar + 1) * G);
                    nbar++;
                }
                bar_consumers();
                if (P.prologue) {
                    stage(P.act, ph.kblk, smem);
                    bar_consumers();
                }
            }
            if (cta == 0 && tid == 0 && P.stamps) P.stamps[(size_t)L * 8 + p] = gtimer();
            if (P.mode == 0) prime();
            float acc = 0.f;
            while (cc < c1) {
#pragma unroll
                for (int b = 0; b < NBUF; b++) {
                    if (cc < c1) {
                        if (lc < c1) {
                            load_chunk(buf[(b + NBUF - 1) % NBUF], ph, lc, lane, pol);
                            lc += NWARP;
                        }
                        {
                            int blk0 = (int)(cc % cpr) * CHUNK_BLK;
                            if (blk0 > ph.kblk - CHUNK_BLK) blk0 = ph.kblk - CHUNK_BLK;
                            acc += compute_chunk(buf[b], smem, ph.kblk, blk0, lane);
                        }
                        cc += NWARP;
                        if (P.mode >= 2 && lane == 0) atomicAdd((unsigned long long *)&s_consumed, (unsigned long long)(CHUNK_WB + CHUNK_BLK * 4));
                    }
                }
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 16);
            acc += __shfl_xor_sync(0xffffffffu, acc, 8);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            total += acc;
            if (lane == 0) P.out[(size_t)cta * NWARP + warp] = total;
            // ---- phase end: arrive ----
            bar_consumers();
            if (tid == 0) arrive(&P.cnt[0]);
        }
    }
    if (cta == 0 && tid == 0 && P.stamps) P.stamps[(size_t)P.layers * 8] = gtimer();
}

__global__ void __launch_bounds__(NT, 1) barrier_kernel(unsigned long long *cnt, int n, unsigned long long base, unsigned long long *out) {
    const unsigned long long t0 = gtimer();
    for (int i = 0; i < n; i++) {
        __syncthreads();
        if (threadIdx.x == 0) {
            arrive(cnt);
            SPIN_UNTIL(ld_acq(cnt) >= base + (unsigned long long)(i + 1) * gridDim.x);
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = gtimer() - t0;
}

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int layers = 32;
    // per-layer blocks: QKV 6144*128, O 4096*128, GU 28672*128, DOWN 4096*448
    const long long blk[4] = {6144LL * 128, 4096LL * 128, 28672LL * 128, 4096LL * 448};
    const int kblk[4] = {128, 128, 128, 448};
    long long per_layer_blk = 0;
    for (int i = 0; i < 4; i++) per_layer_blk += blk[i];
    uint8_t *w;
    float *s, *act, *out;
    unsigned long long *cnt, *stamps;
    CK(cudaMalloc(&w, (size_t)per_layer_blk * 16 * layers));
    CK(cudaMalloc(&s, (size_t)per_layer_blk * 4 * layers));
    CK(cudaMemset(w, 0x5a, (size_t)per_layer_blk * 16 * layers));
    CK(cudaMemset(s, 0, (size_t)per_layer_blk * 4 * layers));
    CK(cudaMalloc(&act, 14336 * 4));
    CK(cudaMemset(act, 0, 14336 * 4));
    CK(cudaMalloc(&out, 148 * 64 * 4));
    CK(cudaMalloc(&cnt, 64));
    CK(cudaMalloc(&stamps, (size_t)(layers + 1) * 8 * 8));
    Params P = {};
    long long off = 0;
    for (int i = 0; i < 4; i++) {
        P.ph[i].w = w + (size_t)off * 16;
        P.ph[i].s = s + (size_t)off;
        P.ph[i].chunks = (int)(blk[i] / CHUNK_BLK);
        P.ph[i].kblk = kblk[i];
        off += blk[i];
    }
    P.layers = layers;
    P.layer_stride_w = (size_t)per_layer_blk * 16;
    P.layer_stride_s = (size_t)per_layer_blk;
    P.cnt = cnt, P.act = act, P.out = out, P.stamps = stamps;
    const size_t smem = 448 * 40 + 64;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const double bytes = (double)per_layer_blk * 20 * layers;
    printf("SMs %d, %.3f GB per token (32 layers)\n", sms, bytes / 1e9);

    // bare barrier
    {
        unsigned long long *o;
        CK(cudaMalloc(&o, 8));
        CK(cudaMemset(cnt, 0, 64));
        const int n = 2000;
        void *args[] = {&cnt, (void *)&n, nullptr, &o};
        unsigned long long base = 0;
        args[2] = &base;
        for (int rep = 0; rep < 2; rep++) {
            base = (unsigned long long)rep * n * sms;
            CK(cudaLaunchCooperativeKernel((void *)barrier_kernel, dim3(sms), dim3(NT), args, 0, 0));
            CK(cudaDeviceSynchronize());
        }
        unsigned long long ns;
        CK(cudaMemcpy(&ns, o, 8, cudaMemcpyDeviceToHost));
        printf("grid barrier (148 x 512, red.release + ld.acquire poll): %.3f us each\n", ns / 1e3 / n);
    }
    struct V {
        int mode, attn_ns, lead_kb, prologue;
    };
    const V vs[] = {{0, 0, 0, 0},   {0, 0, 0, 1},    {1, 0, 0, 0},    {1, 0, 0, 1},   {1, 4000, 0, 1}, {0, 4000, 0, 1},
                    {2, 0, 128, 1}, {2, 4000, 64, 1}, {2, 4000, 128, 1}, {2, 4000, 256, 1}, {2, 4000, 512, 1}, {2, 4000, 1024, 1},
                    {2, 8000, 256, 1}, {1, 8000, 0, 1}};
    unsigned long long epoch = 0;
    for (const V &v : vs) {
        P.mode = v.mode, P.attn_ns = v.attn_ns, P.lead_kb = v.lead_kb, P.prologue = v.prologue;
        float best = 1e9f;
        for (int rep = 0; rep < 4; rep++) {
            // cumulative counters: cnt[0] advances by 4*layers*G per token, cnt[1] by 8*layers; cnt[7] = base of cnt[0]
            unsigned long long h[8] = {0};
            CK(cudaMemcpy(cnt, h, 64, cudaMemcpyHostToDevice));
            void *args[] = {(void *)&P};
            CK(cudaEventRecord(e0));
            CK(cudaLaunchCooperativeKernel((void *)probe_kernel, dim3(sms), dim3(NT + 32), args, smem, 0));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
            epoch++;
        }
        static unsigned long long hs[33 * 8];
        CK(cudaMemcpy(hs, stamps, sizeof hs, cudaMemcpyDeviceToHost));
        double ph[4] = {0, 0, 0, 0};
        for (int L = 1; L < layers; L++) {
            for (int p = 0; p < 3; p++) ph[p] += (double)(hs[L * 8 + p + 1] - hs[L * 8 + p]);
            ph[3] += (double)(hs[(L + 1) * 8] - hs[L * 8 + 3]);
        }
        printf("mode %d attn %4d ns lead %4d KB prologue %d : %.3f ms/token  %.1f GB/s (%.1f us/layer; CTA0 phases qkv %.2f o(+attn) %.2f gu %.2f down %.2f us)\n",
               v.mode, v.attn_ns, v.lead_kb, v.prologue, best, bytes / 1e9 / (best * 1e-3) , best * 1e3 / layers, ph[0] / 31e3, ph[1] / 31e3,
               ph[2] / 31e3, ph[3] / 31e3);
    }
    return 0;
}
