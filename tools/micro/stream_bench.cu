// Micro-benchmark: how fast can one CTA per SM stream HBM through a shared-memory ring with cp.async.bulk?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream_bench stream_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

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
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// each CTA streams `per_cta` bytes starting at base + cta*per_cta, in stages of `stage` bytes split in `ncopy` copies
__global__ void __launch_bounds__(160) ring_kernel(const unsigned char *base, size_t per_cta, int stage, int nstage, int ncopy,
                                                   unsigned long long *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t full[16], empty[16];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < nstage; i++) mbar_init(&full[i], 1), mbar_init(&empty[i], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned char *src = base + (size_t)blockIdx.x * per_cta;
    const int n = (int)(per_cta / stage);
    if (warp == 4) {
        for (int it = 0; it < n; it++) {
            const int slot = it % nstage, use = it / nstage;
            while (!mbar_try_wait(&empty[slot], (use & 1) ^ 1)) {}
            if (lane == 0) mbar_expect_tx(&full[slot], stage);
            __syncwarp();
            if (lane < ncopy) tma_load_1d(smem + (size_t)slot * stage + (size_t)lane * (stage / ncopy),
                                          src + (size_t)it * stage + (size_t)lane * (stage / ncopy), stage / ncopy, &full[slot]);
        }
    } else {
        unsigned long long acc = 0;
        for (int it = 0; it < n; it++) {
            const int slot = it % nstage, use = it / nstage;
            while (!mbar_try_wait(&full[slot], use & 1)) {}
            acc += *(const unsigned long long *)(smem + (size_t)slot * stage + tid * 8);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
        if (acc == 0x1234567) sink[0] = acc;
    }
}

// plain LDG streaming: grid-stride uint4 loads, `unroll` independent loads per thread in flight
template <int U>
__global__ void ldg_kernel(const uint4 *p, size_t n, unsigned long long *sink) {
    unsigned long long acc = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (; i + (size_t)(U - 1) * blockDim.x < n; i += stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            const uint4 *q = p + i + (size_t)u * blockDim.x;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(q));
        }
#pragma unroll
        for (int u = 0; u < U; u++) acc += v[u].x ^ v[u].w;
    }
    if (acc == 0x1234567) sink[0] = acc;
}

int main() {
    const size_t total = (size_t)148 * 32 * 1024 * 1024; // 4.6 GB, far larger than L2
    unsigned char *buf;
    unsigned long long *sink;
    cudaMalloc(&buf, total);
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, total);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    const int stages[] = {16384, 32768, 40960, 65536};
    const int nst[] = {2, 3, 4, 5, 6, 8, 12};
    const int ncp[] = {1, 2, 4};
    for (int stage : stages)
        for (int ns : nst)
            for (int nc : ncp) {
                size_t smem = (size_t)stage * ns;
                if (smem > 220 * 1024) continue;
                if (nc != 2 && !(stage == 32768 && ns == 4)) continue;
                cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                size_t per_cta = (total / sms) / stage * stage;
                for (int rep = 0; rep < 2; rep++) {
                    cudaEventRecord(e0);
                    ring_kernel<<<sms, 160, smem>>>(buf, per_cta, stage, ns, nc, sink);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                }
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                printf("ring stage=%6d nstage=%2d ncopy=%d inflight/SM=%4zu KB : %7.1f GB/s (%s)\n", stage, ns, nc, smem / 1024,
                       per_cta * sms / 1e9 / (ms * 1e-3), cudaGetErrorString(cudaGetLastError()));
            }
    // 2 CTAs per SM variant: smaller rings
    for (int ns : {2, 3}) {
        const int stage = 32768;
        size_t smem = (size_t)stage * ns;
        cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        size_t per_cta = (total / (2 * sms)) / stage * stage;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            ring_kernel<<<2 * sms, 160, smem>>>(buf, per_cta, stage, ns, 2, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("ring 2CTA/SM stage=%d nstage=%d : %7.1f GB/s\n", stage, ns, per_cta * 2 * sms / 1e9 / (ms * 1e-3));
    }
    const size_t n16 = total / 16;
    for (int blocks_per_sm : {1, 2, 4, 8}) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            ldg_kernel<8><<<sms * blocks_per_sm, 256>>>((const uint4 *)buf, n16, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("ldg U=8 blocks/SM=%d : %7.1f GB/s\n", blocks_per_sm, total / 1e9 / (ms * 1e-3));
    }
    return 0;
}
