// Prefill GEMM on the 5th-generation tensor cores:  C[T, N] (+residual) = A[T, K] * W[N, K]^T
//
// The prefill counterpart of jl_gemv.cu for M = T tokens >= 16 (AbstractModel.batchForward chunks of up to 256
// tokens, core/model/AbstractModel.java:295-312).  Replaces the reference's GemmerF32Q4 / GemmerI8Q4 tiles
// (PanamaTensorOperations.java:148-1044, vector_simd.c:261-964) and the WebGPU gemm_q4.wgsl / gemm_i8q4.wgsl.
//
//   * tcgen05.mma kind::f16 (BF16 x BF16 -> F32) issued by one elected thread, accumulator in TMEM;
//   * the WEIGHT tile is the UMMA "A" operand (M = 128 weight rows), the token tile the "B" operand (N = 128
//     tokens), so the instruction's M is always full and D comes out as [weight row][token];
//   * per-block dequantisation (nibble - 8) * f32 scale -> BF16 is fused into the shared-memory fill of the
//     weight tile: eight producer warps read the packed Q4 blocks with 128-bit loads straight from HBM/L2
//     and write the 128-byte-swizzled K-major tile the tensor core consumes; their packed blocks are requested three
//     pipeline stages ahead (registers), so the dequantisation never waits for HBM/L2;
//   * the activation tile (already BF16 in HBM) arrives by TMA: one thread of a tenth warp issues
//     cp.async.bulk.tensor.2d with a SWIZZLE_128B tensor map ([T, lda] BF16, box 64 x BN) that completes on the stage's
//     full barrier (expect_tx); rows past T are zero-filled by the copy engine;
//   * 4-stage mbarrier pipeline: producers + TMA -> full[s] -> MMA warp -> tcgen05.commit -> empty[s];
//   * epilogue: tcgen05.ld (32 lanes x 16 columns per warp) -> optional residual add -> coalesced f32 stores.
//
// Numerics: operands are rounded to BF16 (8-bit mantissa), accumulation is F32.  This is the "F32/BF16 x Q4"
// flavour of the reference (tolerance class 1e-2 rel on logits); the exact-integer Q8 x Q4 arithmetic stays
// available through the GEMV path (jl_model: prefill_tensor_core = 0).
#include "jl_common.cuh"
#include <cuda.h> // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint, no libcuda link)

#define TC_PWARPS 8                      // producer / epilogue warps
#define TC_THREADS (TC_PWARPS * 32 + 64) // + warp 8: TMEM allocator and MMA issuer, warp 9: TMA issuer
#define TC_PF 3                          // weight blocks requested this many stages ahead
#define TC_BM 128                        // weight rows per CTA (UMMA M)
#define TC_BK 64                         // K per pipeline stage: 64 bf16 = one 128-byte swizzle atom
#define TC_STAGES 4
#define TC_WTILE_BYTES (TC_BM * TC_BK * 2) // 16 KB weight tile

__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(tc_smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (tcgen05): rows are 128 bytes, 8-row groups are 1024 bytes
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, 16-byte units
    d |= (uint64_t)0 << 16;                       // leading byte offset: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                       // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = BN
template <int BN>
__device__ __forceinline__ uint32_t tc_instr_desc() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *(uint32_t *)&v;
}

struct TcParams {
    const uint16_t *a; // activations bf16 [T, lda]
    int lda, T;
    const uint8_t *w;  // Q4 nibbles (row pitch K_total/2) or int8 (row pitch K_total)
    const float *ws;   // scales, row pitch K_total/32
    int ldw;           // K_total (elements)
    int w_col_off;     // first weight column (multiple of 64)
    int K;             // reduction length (multiple of 64)
    int N;             // weight rows handled (multiple of 128)
    float *out;        // [T, ldc]
    int ldc, out_col_off;
    const float *residual; // nullable [T, res_ld]
    int res_ld;
};

// BN = tokens per CTA (UMMA N): 128, or 256 to amortise the weight dequantisation over twice the tokens
// WDT = JL_Q4 or JL_I8 (Q8_0: int8 + one f32 scale per 32 elements; the block is 32 bytes in natural element order)
template <int BN, int WDT>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_q4_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap amap) {
    constexpr int ATILE = BN * TC_BK * 2;
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char *wtile = tc_smem;                                // [STAGES][16 KB] weight tiles (UMMA A)
    unsigned char *atile = tc_smem + TC_STAGES * TC_WTILE_BYTES;   // [STAGES][BN x 128 B] token tiles (UMMA B)
    __shared__ uint64_t full[TC_STAGES], empty[TC_STAGES], tmem_full;
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.x * TC_BM; // first weight row of this CTA
    const int t0 = blockIdx.y * BN;    // first token of this CTA
    const int nk = p.K / TC_BK;

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; s++) {
            tc_mbar_init(&full[s], TC_PWARPS); // one arrival per producer warp
            tc_mbar_init(&empty[s], 1);        // tcgen05.commit
        }
        tc_mbar_init(&tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_PWARPS) { // TMEM: BN columns x 128 lanes of f32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_sh)), "n"(BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_sh;

    if (warp < TC_PWARPS) {
        // ===== producers.  Weight tile: thread (r = tid & 127, half = tid >> 7) dequantises one 32-element block of
        // row r (four 16-byte chunks), XOR-swizzled by (row & 7) -- the SWIZZLE_128B K-major layout the tensor core
        // reads (and the layout TMA writes for the token tile). =====
        const int r = tid & 127, half = tid >> 7;
        const size_t wrow = (size_t)(n0 + r);
        constexpr int WB = WDT == JL_Q4 ? 16 : 32; // bytes of a 32-element block
        constexpr int NQ = WB / 16;                 // 128-bit loads per block
        const uint8_t *wq = p.w + (wrow * (size_t)(p.ldw / 32) + p.w_col_off / 32 + half) * WB;
        const float *wsc = p.ws + wrow * (size_t)(p.ldw / 32) + p.w_col_off / 32 + half;
        uint4 qn[TC_PF][NQ];
        float scn[TC_PF];
#pragma unroll
        for (int i = 0; i < TC_PF; i++) {
            const int kk = i < nk ? i : nk - 1;
#pragma unroll
            for (int j = 0; j < NQ; j++) qn[i][j] = ldg_nc_u4(wq + (size_t)kk * 2 * WB + j * 16);
            scn[i] = ldg_nc_f32(wsc + kk * 2);
        }
        for (int kc = 0; kc < nk; kc++) {
            const int s = kc % TC_STAGES, use = kc / TC_STAGES;
            uint4 q[NQ];
#pragma unroll
            for (int j = 0; j < NQ; j++) q[j] = qn[0][j];
            const float sc = scn[0];
#pragma unroll
            for (int i = 0; i + 1 < TC_PF; i++) {
#pragma unroll
                for (int j = 0; j < NQ; j++) qn[i][j] = qn[i + 1][j];
                scn[i] = scn[i + 1];
            }
            {
                const int kk = kc + TC_PF < nk ? kc + TC_PF : nk - 1;
#pragma unroll
                for (int j = 0; j < NQ; j++) qn[TC_PF - 1][j] = ldg_nc_u4(wq + (size_t)kk * 2 * WB + j * 16);
                scn[TC_PF - 1] = ldg_nc_f32(wsc + kk * 2);
            }
            tc_mbar_wait(&empty[s], (use & 1) ^ 1);
            unsigned char *wdst = wtile + (size_t)s * TC_WTILE_BYTES + (size_t)r * 128;
            float lo[16], hi[16]; // elements 0..15 and 16..31 of the block, scaled
            if (WDT == JL_Q4) {
                // block = 16 bytes: element j = low nibble of byte j, element j+16 = high nibble of byte j.
                // nibble -> float through the 2^23 magic number (PRMT + FADD), * scale, round once to BF16.
                const uint32_t qw[4] = {q[0].x, q[0].y, q[0].z, q[0].w};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t l4 = qw[i] & 0x0F0F0F0Fu, h4 = (qw[i] >> 4) & 0x0F0F0F0Fu;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        lo[i * 4 + t] = __fmul_rn(__uint_as_float(__byte_perm(l4, 0x4B000000u, 0x7540 | t)) - 8388616.0f, sc);
                        hi[i * 4 + t] = __fmul_rn(__uint_as_float(__byte_perm(h4, 0x4B000000u, 0x7540 | t)) - 8388616.0f, sc);
                    }
                }
            } else {
                // int8 -> float: byte ^ 0x80 is the value + 128 as an unsigned byte; 2^23 + that, minus (2^23 + 128)
                const uint32_t qw[8] = {q[0].x, q[0].y, q[0].z, q[0].w, q[NQ - 1].x, q[NQ - 1].y, q[NQ - 1].z, q[NQ - 1].w};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const uint32_t u = qw[i] ^ 0x80808080u;
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const float f = __fmul_rn(__uint_as_float(__byte_perm(u, 0x4B000000u, 0x7540 | t)) - 8388736.0f, sc);
                        if (i < 4) lo[i * 4 + t] = f;
                        else hi[(i - 4) * 4 + t] = f;
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 4; c++) { // chunks half*4 + {0,1}: elements 0..15; {2,3}: elements 16..31
                const float *src = c < 2 ? lo + c * 8 : hi + (c - 2) * 8;
                uint4 o;
                o.x = pack_bf16x2(src[0], src[1]);
                o.y = pack_bf16x2(src[2], src[3]);
                o.z = pack_bf16x2(src[4], src[5]);
                o.w = pack_bf16x2(src[6], src[7]);
                *(uint4 *)(wdst + (((half * 4 + c) ^ (r & 7)) * 16)) = o;
            }
            // make the generic-proxy writes visible to the tensor core (async proxy), then signal
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&full[s]);
        }
        // ===== epilogue: warp w reads TMEM lanes 32*(w%4) .. +31 (weight rows), column half w/4 =====
        tc_mbar_wait(&tmem_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lq = warp & 3, chalf = warp >> 2;
        const int nrow = n0 + lq * 32 + lane;
        const int col_out = p.out_col_off + nrow;
        for (int c0 = chalf * (BN / 2); c0 < (chalf + 1) * (BN / 2); c0 += 16) {
            if (t0 + c0 >= p.T) break;
            uint32_t v[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int tok = t0 + c0 + i;
                if (tok < p.T) {
                    float x = __uint_as_float(v[i]);
                    if (p.residual) x = __fadd_rn(x, p.residual[(size_t)tok * p.res_ld + nrow]);
                    p.out[(size_t)tok * p.ldc + col_out] = x;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (warp == TC_PWARPS + 1) {
        // ===== TMA issuer: the BN x 64 BF16 token tile of every K stage, as far ahead as the ring allows =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&amap) : "memory");
            for (int kc = 0; kc < nk; kc++) {
                const int s = kc % TC_STAGES, use = kc / TC_STAGES;
                tc_mbar_wait(&empty[s], (use & 1) ^ 1);
                asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&full[s])), "r"((uint32_t)ATILE) : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(tc_smem_u32(atile + (size_t)s * ATILE)), "l"((uint64_t)&amap), "r"(kc * TC_BK), "r"(t0), "r"(tc_smem_u32(&full[s]))
                             : "memory");
            }
        }
    } else {
        // ===== MMA issuer (one elected lane of warp 8) =====
        const uint32_t idesc = tc_instr_desc<BN>();
        for (int kc = 0; kc < nk; kc++) {
            const int s = kc % TC_STAGES, use = kc / TC_STAGES;
            tc_mbar_wait(&full[s], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t wa = tc_smem_u32(wtile + (size_t)s * TC_WTILE_BYTES);
                const uint32_t aa = tc_smem_u32(atile + (size_t)s * ATILE);
#pragma unroll
                for (int k = 0; k < TC_BK / 16; k++) {
                    const uint64_t adesc = tc_smem_desc(wa + k * 32); // +16 bf16 along K inside the swizzle atom
                    const uint64_t bdesc = tc_smem_desc(aa + k * 32);
                    const uint32_t accum = (kc | k) ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem_base), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
                        : "memory");
                }
                // frees the smem slot when these MMAs have consumed it (implies fence::before_thread_sync)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(&empty[s]))
                             : "memory");
                if (kc == nk - 1)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(&tmem_full))
                                 : "memory");
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == TC_PWARPS) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
    }
}

typedef CUresult (*tc_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                 const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tc_encode_fn tc_encoder() {
    static tc_encode_fn fn = nullptr;
    if (!fn) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
            fn = (tc_encode_fn)sym;
    }
    return fn;
}

template <int BN, int WDT>
static int launch_tc(jl_ctx *ctx, cudaStream_t stream, const TcParams &p) {
    const size_t smem = (size_t)TC_STAGES * (TC_WTILE_BYTES + (size_t)BN * TC_BK * 2) + 1024;
    static size_t configured[JL_MAX_DEVICES] = {};
    JL_CUDA_CHECK(ctx, (jl_ensure_dyn_smem(gemm_q4_tc_kernel<BN, WDT>, ctx->device, smem, configured)));
    // tensor map of the BF16 activations [T rows, K cols] (row pitch lda): box = 64 columns (one 128-byte swizzle atom) x BN rows
    tc_encode_fn enc = tc_encoder();
    if (!enc) return jl_set_error(ctx, JL_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled is not available");
    CUtensorMap amap;
    const cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.T};
    const cuuint64_t strides[1] = {(cuuint64_t)p.lda * 2};
    const cuuint32_t box[2] = {TC_BK, (cuuint32_t)BN};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult cr = enc(&amap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void *)p.a, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return jl_set_error(ctx, JL_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)cr);
    dim3 grid(p.N / TC_BM, (p.T + BN - 1) / BN);
    gemm_q4_tc_kernel<BN, WDT><<<grid, TC_THREADS, smem, stream>>>(p, amap);
    ctx->launches++;
    JL_CUDA_CHECK(ctx, cudaGetLastError());
    return JL_OK;
}

// C[T, out_col_off + n] (+= residual) = sum_k A_bf16[T, a_col_off + k] * dequant(W[n, w_col_off + k]),  n in [0, N)
int jl_launch_gemm_tc(jl_ctx *ctx, cudaStream_t stream, const uint16_t *a_bf16, int lda, int T, const DevTensor &W, int n_rows,
                      int w_col_off, int K, float *out, int ldc, int out_col_off, const float *residual, int res_ld) {
    if (W.dtype != JL_Q4 && W.dtype != JL_I8) return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "gemm_tc: weights must be Q4 or Q8_0 (I8)");
    if ((K % TC_BK) || (n_rows % TC_BM) || (w_col_off % TC_BK) || T <= 0 || (lda % 8) || ((uintptr_t)a_bf16 % 16))
        return jl_set_error(ctx, JL_ERR_UNSUPPORTED, "gemm_tc: needs K %% 64 == 0 and N %% 128 == 0 (K=%d N=%d)", K, n_rows);
    TcParams p;
    p.a = a_bf16, p.lda = lda, p.T = T;
    p.w = (const uint8_t *)W.data, p.ws = W.scales, p.ldw = (int)W.cols, p.w_col_off = w_col_off, p.K = K, p.N = n_rows;
    p.out = out, p.ldc = ldc, p.out_col_off = out_col_off, p.residual = residual, p.res_ld = res_ld;
    // 256-token tiles amortise the dequantisation better; use them when they do not starve the grid
    const bool wide = T > 128 && (size_t)(n_rows / TC_BM) * ((T + 255) / 256) >= (size_t)ctx->sm_count / 2;
    if (W.dtype == JL_Q4) return wide ? launch_tc<256, JL_Q4>(ctx, stream, p) : launch_tc<128, JL_Q4>(ctx, stream, p);
    return wide ? launch_tc<256, JL_I8>(ctx, stream, p) : launch_tc<128, JL_I8>(ctx, stream, p);
}
