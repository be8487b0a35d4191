/*
 * jlama_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the arithmetic of tjake/Jlama's quantized
 * forward pass (the path BASELINE.json's north_star names).  It exists so the
 * CUDA path in jlama_b200/ can be checked against the reference's semantics.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product never does.
 *
 * Parity pinning: the GEMM restatements are checked against the reference's
 * own C kernels (oracle/_ref/libjlama.so, compiled from
 * /root/reference/jlama-native/src/main/c/simd/vector_simd.c by oracle/Makefile)
 * and the RoPE table against the golden vectors of
 * jlama-tests/.../model/TestCorrectness.java:93-115 (tests/test_oracle.py).
 * Everything that depends on net.jafama FastMath (exp/tanh/sqrt/pow/cos/sin,
 * a ~1-ulp double libm replacement not vendored in /root/reference) is
 * restated with C libm in double and is pinned only by the TestRope vectors:
 * "parity unpinned" for exp/sqrt/pow beyond that.
 *
 * Every function cites the reference file:line it follows.  Paths are relative
 * to /root/reference/jlama-core/src/main/java/com/github/tjake/jlama/ unless
 * they start with jlama-native/ or jlama-tests/.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#include <float.h>
#include <dlfcn.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define JO_F32 0
#define JO_BF16 1
#define JO_Q4 2
#define JO_I8 3 /* Q8 block-quantised int8 + f32 scale per 32 */

#define QBLOCK 32
#define HALF_BLOCK 16

/* ------------------------------------------------------------------------ */
/* Float conversions: math/FloatConversions.java:31-90                       */
/* ------------------------------------------------------------------------ */
static inline float bf16_to_f32(uint16_t raw) { /* :31-33 */
    uint32_t u = ((uint32_t)raw) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static inline uint16_t f32_to_bf16(float n) { /* :35-61, round() :73-90 */
    uint32_t nbits;
    memcpy(&nbits, &n, 4);
    uint32_t s = (nbits >> 16) & 0x8000u;
    uint32_t e = (nbits >> 16) & 0x7f80u;
    uint32_t m = nbits & 0x7fffffu;
    if (e != 0x7f80u) {
        int mid = 1 << 15, mask = (1 << 16) - 1;
        int mshift = (int)(m >> 16);
        int masked = (int)(m & mask);
        int cmp = masked - mid;
        int m1;
        if (cmp > 0) m1 = mshift + 1;
        else if (cmp < 0) m1 = mshift;
        else m1 = (mshift & 1) ? mshift + 1 : mshift;
        return (uint16_t)(s | (e + (uint32_t)m1));
    }
    return m != 0 ? (uint16_t)0x7fc0 : (uint16_t)(nbits >> 16);
}

void jo_f32_to_bf16(const float *x, uint16_t *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = f32_to_bf16(x[i]);
}
void jo_bf16_to_f32(const uint16_t *x, float *out, int64_t n) {
    for (int64_t i = 0; i < n; i++) out[i] = bf16_to_f32(x[i]);
}

/* ------------------------------------------------------------------------ */
/* Block quantisers                                                          */
/* ------------------------------------------------------------------------ */

/* Q4 weight quantiser: tensor/Q4ByteBufferTensor.java:66-120.
 * scale = signed max-abs / -8 (:83), q = min(15, (byte)(x*iscale + 8.5f)) (:103-104),
 * byte j of a block = q[j] | q[j+16] << 4 (:96-106). */
void jo_quantize_q4(const float *x, int64_t rows, int64_t cols, uint8_t *q, float *scales) {
    int64_t nblk = cols / QBLOCK;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        for (int64_t b = 0; b < nblk; b++) {
            const float *p = x + r * cols + b * QBLOCK;
            float max = FLT_TRUE_MIN, amax = FLT_TRUE_MIN; /* Float.MIN_VALUE */
            for (int i = 0; i < QBLOCK; i++) {
                float v = p[i];
                float absv = v < 0 ? -v : v;
                if (absv > amax) { max = v; amax = absv; }
            }
            volatile float scale = max / -8.0f;
            float iscale = scale != 0.0f ? 1.0f / scale : 0.0f;
            scales[r * nblk + b] = scale;
            uint8_t *qb = q + (r * cols + b * QBLOCK) / 2;
            for (int j = 0; j < HALF_BLOCK; j++) {
                volatile float f0 = p[j] * iscale;
                volatile float f1 = p[j + HALF_BLOCK] * iscale;
                volatile float g0 = f0 + 8.5f, g1 = f1 + 8.5f;
                int8_t b0 = (int8_t)(int)g0; /* Java (byte)(float): f2i then i2b */
                int8_t b1 = (int8_t)(int)g1;
                if (b0 > 15) b0 = 15;
                if (b1 > 15) b1 = 15;
                qb[j] = (uint8_t)((b0) | (b1 << 4));
            }
        }
    }
}

/* Q4 get(): tensor/Q4ByteBufferTensor.java:179-197 */
static inline float q4_get(const uint8_t *q, const float *scales, int64_t cols, int64_t r, int64_t c) {
    int64_t blk = c / QBLOCK;
    int in = (int)(c % QBLOCK);
    float scale = scales[r * (cols / QBLOCK) + blk];
    const uint8_t *qb = q + (r * cols + blk * QBLOCK) / 2;
    int x = in < HALF_BLOCK ? (qb[in] & 0x0F) - 8 : ((qb[in - HALF_BLOCK] >> 4) & 0x0F) - 8;
    return x * scale;
}

void jo_dequantize_q4(const uint8_t *q, const float *scales, int64_t rows, int64_t cols, float *out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++) out[r * cols + c] = q4_get(q, scales, cols, r, c);
}

/* Q8 *weight* quantiser: tensor/Q8ByteBufferTensor.java:68-90
 * (iscale = 127f/max, scale = 1/iscale, q = (byte)Math.round(x*iscale)). */
void jo_quantize_q8_weights(const float *x, int64_t rows, int64_t cols, int8_t *q, float *scales) {
    int64_t nblk = cols / QBLOCK;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; r++) {
        for (int64_t b = 0; b < nblk; b++) {
            const float *p = x + r * cols + b * QBLOCK;
            float max = FLT_TRUE_MIN;
            for (int i = 0; i < QBLOCK; i++) {
                float absv = p[i] < 0 ? -p[i] : p[i];
                if (absv > max) max = absv;
            }
            volatile float iscale = 127.0f / max;
            float scale = iscale != 0.0f ? 1.0f / iscale : 0.0f;
            scales[r * nblk + b] = scale;
            for (int j = 0; j < QBLOCK; j++) {
                volatile float f0 = p[j] * iscale;
                int v;
                if (f0 != f0) v = 0; /* Math.round(NaN) == 0 */
                else v = (int)floor((double)f0 + 0.5); /* Math.round(float) */
                q[r * cols + b * QBLOCK + j] = (int8_t)v;
            }
        }
    }
}

/* Q8 *activation* quantiser (the default working path):
 * tensor/operations/PanamaTensorOperations.java:1684-1723 (quantizeQ8_512).
 * d = max/127, id = max != 0 ? 127/max : 0, q = (byte)(x*id + 0.5f) with F2B
 * truncating toward zero (mul and add are separate roundings). */
void jo_quantize_q8_act(const float *x, int64_t rows, int64_t cols, int64_t ld, int8_t *q, float *scales) {
    int64_t nblk = cols / QBLOCK;
    for (int64_t r = 0; r < rows; r++) {
        for (int64_t b = 0; b < nblk; b++) {
            const float *p = x + r * ld + b * QBLOCK;
            float max = 0.0f;
            for (int i = 0; i < QBLOCK; i++) {
                float a = fabsf(p[i]);
                if (a > max) max = a;
            }
            float d = max / 127.0f;
            volatile float id = (max != 0.0f) ? 127.0f / max : 0.0f;
            for (int j = 0; j < QBLOCK; j++) {
                volatile float m = p[j] * id;
                volatile float a = m + 0.5f;
                q[r * cols + b * QBLOCK + j] = (int8_t)(int)a;
            }
            scales[r * nblk + b] = d;
        }
    }
}

void jo_dequantize_q8(const int8_t *q, const float *scales, int64_t rows, int64_t cols, float *out) {
    /* tensor/Q8ByteBufferTensor.java:139-145: value = q * scale */
    int64_t nblk = cols / QBLOCK;
    for (int64_t r = 0; r < rows; r++)
        for (int64_t c = 0; c < cols; c++) out[r * cols + c] = q[r * cols + c] * scales[r * nblk + c / QBLOCK];
}

/* ------------------------------------------------------------------------ */
/* Tensor view used by the op restatements                                   */
/* ------------------------------------------------------------------------ */
typedef struct {
    int dtype;
    int64_t rows, cols; /* logical shape; row stride == cols (dense) */
    const void *data;   /* f32 / bf16(u16) / q4 nibbles / int8 */
    const float *scales; /* [rows, cols/32] for Q4 / I8 */
} jo_tensor;

static inline float t_get(const jo_tensor *t, int64_t r, int64_t c) {
    switch (t->dtype) {
        case JO_F32: return ((const float *)t->data)[r * t->cols + c];
        case JO_BF16: return bf16_to_f32(((const uint16_t *)t->data)[r * t->cols + c]);
        case JO_Q4: return q4_get((const uint8_t *)t->data, t->scales, t->cols, r, c);
        case JO_I8:
            return ((const int8_t *)t->data)[r * t->cols + c] * t->scales[r * (t->cols / QBLOCK) + c / QBLOCK];
    }
    return 0.0f;
}

/* ------------------------------------------------------------------------ */
/* Optional: the reference's own C kernels (oracle/_ref/libjlama.so)         */
/* jlama-native/src/main/c/simd/vector_simd.h:22-39                          */
/* ------------------------------------------------------------------------ */
typedef void (*ref_gemm_q8_q4_t)(int, const float *, const char *, int, const float *, const char *, int, float *, int,
                                 int, int, int, int, int, int, int, int, int);
typedef void (*ref_gemm_f32_q4_t)(int, const float *, int, const float *, const char *, int, float *, int, int, int, int,
                                  int, int, int, int, int);
typedef void (*ref_gemm_f32_t)(int, const float *, int, const float *, int, float *, int, int, int, int, int, int, int,
                               int);
/* gemm_f32_bf16 (vector_simd.h:38): `cr` is the optional BF16 result, NULL here (F32 result in r) */
typedef void (*ref_gemm_f32_bf16_t)(int, const float *, int, const short *, int, short *, float *, int, int, int, int, int,
                                    int, int, int);
static ref_gemm_f32_bf16_t ref_f32_bf16 = NULL;
static ref_gemm_q8_q4_t ref_q8_q4 = NULL;
static ref_gemm_f32_q4_t ref_f32_q4 = NULL;
static ref_gemm_f32_t ref_f32 = NULL;
static int ref_flags = 0;

/* returns 0 when all three symbols were bound */
int jo_load_reference_kernels(const char *path, int flags) {
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) return -1;
    ref_q8_q4 = (ref_gemm_q8_q4_t)dlsym(h, "gemm_q8_q4");
    ref_f32_q4 = (ref_gemm_f32_q4_t)dlsym(h, "gemm_f32_q4");
    ref_f32 = (ref_gemm_f32_t)dlsym(h, "gemm_f32");
    ref_f32_bf16 = (ref_gemm_f32_bf16_t)dlsym(h, "gemm_f32_bf16");
    ref_flags = flags;
    return (ref_q8_q4 && ref_f32_q4 && ref_f32) ? 0 : -2;
}
int jo_reference_kernels_loaded(void) { return ref_q8_q4 != NULL; }
static int use_ref_kernels = 0;
void jo_use_reference_kernels(int on) { use_ref_kernels = on && ref_q8_q4; }

/* ------------------------------------------------------------------------ */
/* batchDotProduct restatements                                              */
/* result[i, j + rRowOff] for j in [bRowOff, bRowOff+N): the production        */
/* semantics of TensorOperations.java:62-72 as implemented by Panama            */
/* (c.set(.., i, j + rOffset), PanamaTensorOperations.java:848, contract         */
/* rOffset==0 or >= bRowOff :108) and by the natives (r[ldc*i + j - roffset]     */
/* with roffset = -rRowOffset, NativeSimdTensorOperations.java:100-107,          */
/* vector_simd.c:344).  NaiveTensorOperations.java:94-100 instead writes         */
/* result[i, rRowOff + (j - bRowOff)]; the two agree whenever bRowOff == 0,      */
/* which is the only case the reference's tests exercise with Naive.             */
/* ------------------------------------------------------------------------ */

/* Naive semantic: float sequential sum of get()*get()  (NaiveTensorOperations.java:64-79) */
static float dot_naive(const jo_tensor *a, int64_t ar, const jo_tensor *b, int64_t br, int aoff, int boff, int K) {
    float s = 0;
    for (int t = 0; t < K; t++) s += t_get(a, ar, aoff + t) * t_get(b, br, boff + t);
    return s;
}

/* I8 x Q4: per 32-block integer dot, then acc += (sa*sb) * (float)isum
 * (jlama-native/src/main/c/simd/vector_simd.c:384-420;
 *  PanamaTensorOperations.java:836-848).  Integer part is exact; the float
 *  accumulation here is sequential over blocks. */
static float dot_q8_q4(const jo_tensor *a, int64_t ar, const jo_tensor *b, int64_t br, int aoff, int boff, int K) {
    const int8_t *aq = (const int8_t *)a->data + ar * a->cols;
    const float *as = a->scales + ar * (a->cols / QBLOCK);
    const uint8_t *bq = (const uint8_t *)b->data + (br * b->cols) / 2;
    const float *bs = b->scales + br * (b->cols / QBLOCK);
    float acc = 0.0f;
    for (int t = 0; t < K; t += QBLOCK) {
        const int8_t *ap = aq + aoff + t;
        const uint8_t *bp = bq + (boff + t) / 2;
        int isum = 0;
        for (int j = 0; j < HALF_BLOCK; j++) {
            isum += ap[j] * ((bp[j] & 0x0F) - 8);
            isum += ap[j + HALF_BLOCK] * (((bp[j] >> 4) & 0x0F) - 8);
        }
        float scale = as[(aoff + t) / QBLOCK] * bs[(boff + t) / QBLOCK];
        acc = fmaf(scale, (float)isum, acc);
    }
    return acc;
}

/* F32 x Q4: acc += a[j] * ((nib-8)*s)  (vector_simd.c:770-878;
 * PanamaTensorOperations.java:347-370) */
static float dot_f32_q4(const jo_tensor *a, int64_t ar, const jo_tensor *b, int64_t br, int aoff, int boff, int K) {
    const float *ap0 = (const float *)a->data + ar * a->cols + aoff;
    const uint8_t *bq = (const uint8_t *)b->data + (br * b->cols) / 2;
    const float *bs = b->scales + br * (b->cols / QBLOCK);
    float acc = 0.0f;
    for (int t = 0; t < K; t += QBLOCK) {
        const float *ap = ap0 + t;
        const uint8_t *bp = bq + (boff + t) / 2;
        float s = bs[(boff + t) / QBLOCK];
        for (int j = 0; j < HALF_BLOCK; j++) {
            float lo = (float)((bp[j] & 0x0F) - 8) * s;
            float hi = (float)(((bp[j] >> 4) & 0x0F) - 8) * s;
            acc = fmaf(ap[j], lo, acc);
            acc = fmaf(ap[j + HALF_BLOCK], hi, acc);
        }
    }
    return acc;
}

static float dot_f32_f32(const float *a, const float *b, int K) {
    float acc = 0.0f;
    for (int t = 0; t < K; t++) acc = fmaf(a[t], b[t], acc);
    return acc;
}

void jo_batch_dot(float *result, int64_t ldc, const jo_tensor *a, const jo_tensor *b, int aColOff, int bColOff, int K,
                  int rRowOff, int bRowOff, int N) {
    int64_t M = a->rows;
    /* Fast path through the reference's own kernels when requested.
     * NOTE (reference quirk, not reproduced): vector_simd.c:180-183 recurses on
     * (rows mp..m, cols n0..np) and (rows m0..mp, cols np..n) and therefore never
     * computes the corner tile (rows mp..m, cols np..n) when both the row and the
     * column counts leave a remainder modulo the tile size.  We only hand it column
     * ranges whose length is a multiple of 5 or smaller than 5, for which the
     * recursion is complete. */
    if (use_ref_kernels && b->dtype == JO_Q4 &&
        ((a->dtype == JO_I8 && (K % 256) == 0) || a->dtype == JO_F32)) {
        /* NativeSimdTensorOperations.java:176-193 argument marshalling */
#pragma omp parallel
        {
#ifdef _OPENMP
            int nt = omp_get_num_threads(), tid = omp_get_thread_num();
#else
            int nt = 1, tid = 0;
#endif
            int chunk = (N + nt - 1) / nt;
            chunk = (chunk + 4) / 5 * 5;
            int c0 = bRowOff + tid * chunk;
            int cn = chunk;
            if (c0 + cn > bRowOff + N) cn = bRowOff + N - c0;
            for (int part = 0; part < 2 && cn > 0; part++) {
                int n0 = part == 0 ? c0 : c0 + cn / 5 * 5;
                int n = part == 0 ? cn / 5 * 5 : cn % 5;
                if (M == 1) { n0 = c0; n = part == 0 ? cn : 0; }
                if (n <= 0) continue;
                if (a->dtype == JO_I8)
                    ref_q8_q4(ref_flags, a->scales, (const char *)a->data, aColOff, b->scales, (const char *)b->data,
                              bColOff / 2, result, -rRowOff, (int)M, n0, n, K, (int)a->cols,
                              (int)(a->cols / QBLOCK), (int)(b->cols / 2), (int)(b->cols / QBLOCK), (int)ldc);
                else
                    ref_f32_q4(ref_flags, (const float *)a->data, aColOff, b->scales, (const char *)b->data,
                               bColOff / 2, result, -rRowOff, (int)M, n0, n, K, (int)a->cols, (int)(b->cols / 2),
                               (int)(b->cols / QBLOCK), (int)ldc);
            }
        }
        return;
    }
#pragma omp parallel for schedule(static) if (N >= 64)
    for (int j = 0; j < N; j++) {
        int64_t br = bRowOff + j;
        for (int64_t i = 0; i < M; i++) {
            float d;
            if (a->dtype == JO_I8 && b->dtype == JO_Q4) d = dot_q8_q4(a, i, b, br, aColOff, bColOff, K);
            else if (a->dtype == JO_F32 && b->dtype == JO_Q4) d = dot_f32_q4(a, i, b, br, aColOff, bColOff, K);
            else if (a->dtype == JO_F32 && b->dtype == JO_F32)
                d = dot_f32_f32((const float *)a->data + i * a->cols + aColOff,
                                (const float *)b->data + br * b->cols + bColOff, K);
            else d = dot_naive(a, i, b, br, aColOff, bColOff, K);
            result[i * ldc + rRowOff + bRowOff + j] = d;
        }
    }
}

/* Always-naive variant (NaiveTensorOperations control implementation) */
void jo_batch_dot_naive(float *result, int64_t ldc, const jo_tensor *a, const jo_tensor *b, int aColOff, int bColOff,
                        int K, int rRowOff, int bRowOff, int N) {
    for (int64_t i = 0; i < a->rows; i++)
        for (int j = 0; j < N; j++)
            result[i * ldc + rRowOff + bRowOff + j] = dot_naive(a, i, b, bRowOff + j, aColOff, bColOff, K);
}

/* Direct wrappers over the reference kernels, for pinning the restatement */
int jo_ref_gemm_q8_q4(const float *af, const int8_t *a, const float *bf, const uint8_t *b, float *r, int m, int n0,
                      int n, int k, int lda, int ldb_elems, int ldc) {
    if (!ref_q8_q4) return -1;
    ref_q8_q4(ref_flags, af, (const char *)a, 0, bf, (const char *)b, 0, r, 0, m, n0, n, k, lda, lda / QBLOCK,
              ldb_elems / 2, ldb_elems / QBLOCK, ldc);
    return 0;
}
int jo_ref_gemm_f32_q4(const float *a, const float *bf, const uint8_t *b, float *r, int m, int n0, int n, int k,
                       int lda, int ldb_elems, int ldc) {
    if (!ref_f32_q4) return -1;
    ref_f32_q4(ref_flags, a, 0, bf, (const char *)b, 0, r, 0, m, n0, n, k, lda, ldb_elems / 2, ldb_elems / QBLOCK, ldc);
    return 0;
}
int jo_ref_gemm_f32(const float *a, const float *b, float *r, int m, int n0, int n, int k, int lda, int ldb, int ldc) {
    if (!ref_f32) return -1;
    ref_f32(ref_flags, a, 0, b, 0, r, 0, m, n0, n, k, lda, ldb, ldc);
    return 0;
}

int jo_ref_gemm_f32_bf16(const float *a, const uint16_t *b, float *r, int m, int n0, int n, int k, int lda, int ldb, int ldc) {
    if (!ref_f32_bf16) return -1;
    ref_f32_bf16(ref_flags, a, 0, (const short *)b, 0, NULL, r, 0, m, n0, n, k, lda, ldb, ldc);
    return 0;
}
/* ------------------------------------------------------------------------ */
/* Element-wise ops (NaiveTensorOperations.java:34-125 semantics)            */
/* ------------------------------------------------------------------------ */
/* a[r, off:off+len] += b[(r or 0), off:off+len]  (:34-46; Panama :2151-2218) */
void jo_accumulate(float *a, int64_t arows, int64_t lda, const jo_tensor *b, int off, int len) {
    int batch = b->rows > 1;
    for (int64_t r = 0; r < arows; r++)
        for (int i = off; i < off + len; i++) a[r * lda + i] = a[r * lda + i] + t_get(b, batch ? r : 0, i);
}
/* a *= b (:49-61) */
void jo_maccumulate(float *a, int64_t arows, int64_t lda, const float *b, int64_t brows, int64_t ldb, int off, int len) {
    int batch = brows > 1;
    for (int64_t r = 0; r < arows; r++)
        for (int i = off; i < off + len; i++) a[r * lda + i] = a[r * lda + i] * b[(batch ? r : 0) * ldb + i];
}
/* x[b, off:off+len] *= f (:113-119) */
void jo_scale(float f, float *x, int64_t rows, int64_t ld, int off, int len) {
    for (int64_t r = 0; r < rows; r++)
        for (int i = off; i < off + len; i++) x[r * ld + i] = x[r * ld + i] * f;
}
/* y[yoff+i] = alpha*x[xoff+i] + y[yoff+i]  (:106-111) */
void jo_saxpy(float alpha, const float *x, float *y, int xoff, int yoff, int limit) {
    for (int i = 0; i < limit; i++) y[yoff + i] = (alpha * x[xoff + i]) + y[yoff + i];
}
/* batched saxpy: TensorOperations.java:122-137 (order of rows preserved; Panama
 * fuses 4 rows with FMA, PanamaTensorOperations.java:2648-2698) */
void jo_saxpy_batch(const float *alpha, const float *x, int64_t ldx, float *y, int xoff, int yoff, int limit, int aOff,
                    int xRowOff, int batch) {
    for (int r = 0; r < batch; r++) {
        float al = alpha[aOff + r];
        const float *xr = x + (int64_t)(xRowOff + r) * ldx;
        for (int i = 0; i < limit; i++) y[yoff + i] = fmaf(xr[xoff + i], al, y[yoff + i]);
    }
}

/* math/VectorMath.java:69-90 */
void jo_softmax(float *x, int offset, int length) {
    long size = offset + length;
    float max_val = x[offset];
    for (long i = offset + 1; i < size; i++)
        if (x[i] > max_val) max_val = x[i];
    float sum = 0.0f;
    for (long i = offset; i < size; i++) {
        x[i] = (float)exp((double)(x[i] - max_val));
        sum += x[i];
    }
    for (long i = 0; i < size; i++) x[i] = x[i] / sum; /* starts at 0, :87 */
}

/* math/ActivationFunction.java:29-37 */
float jo_silu(float x) { return (float)(x * (1.0f / (1.0f + exp(-(double)x)))); }
float jo_gelu(float x) {
    return (float)(0.5 * x * (1 + tanh(sqrt(2 / M_PI) * (x + 0.044715 * pow((double)x, 3)))));
}

/* model/RMSNorm.java:34-56: float products summed in double, /embeddingLength,
 * +eps, 1/sqrt in double, (float) cast before the multiply. */
void jo_rmsnorm(const float *x, int64_t rows, int64_t ld, const jo_tensor *w, float adj, float eps, int E, int off,
                int len, float *out) {
    for (int64_t b = 0; b < rows; b++) {
        double ss = 0.0;
        for (int j = off; j < off + len; j++) {
            float v = x[b * ld + j];
            volatile float vv = v * v;
            ss += vv;
        }
        ss /= E;
        ss += eps;
        ss = 1.0 / sqrt(ss);
        for (int j = off; j < off + len; j++) {
            volatile float n = (float)ss * x[b * ld + j];
            out[b * ld + j] = (adj + t_get(w, 0, j)) * n;
        }
    }
}

/* model/LayerNorm.java:41-67 (GPT-2) */
void jo_layernorm(const float *x, int64_t rows, int64_t ld, const jo_tensor *w, const jo_tensor *bias, float eps, int E,
                  int off, int len, float *out) {
    for (int64_t b = 0; b < rows; b++) {
        float sum = 0, sumSq = 0;
        for (int i = off; i < off + len; i++) {
            float v = x[b * ld + i];
            sum += v;
            sumSq += v * v;
        }
        float mean = sum / E;
        float variance = sumSq / E - mean * mean;
        float invStddev = 1.0f / (float)sqrt((double)(variance + eps));
        for (int i = off; i < off + len; i++) {
            float v = (x[b * ld + i] - mean) * invStddev * t_get(w, 0, i) + t_get(bias, 0, i);
            out[b * ld + i] = v;
        }
    }
}

/* math/VectorMath.java:148-165: table[(pos*(dim/2) + i)] = (cos, sin) */
void jo_precompute_freqs_cis(int dim, int end, double theta, double scaling, float *out /* [end*dim/2][2] */) {
    int half = dim / 2;
    float *freqs = (float *)malloc(sizeof(float) * half);
    float step = 0.0f;
    for (int i = 0; i < half; i++, step += 2.0f) freqs[i] = (float)((1.0 / pow(theta, (double)(step / dim))) / scaling);
    for (int64_t p = 0; p < end; p++) {
        float t = (float)p;
        for (int i = 0; i < half; i++) {
            volatile float ang = t * freqs[i];
            out[(p * half + i) * 2 + 0] = (float)cos((double)ang);
            out[(p * half + i) * 2 + 1] = (float)sin((double)ang);
        }
    }
    free(freqs);
}

/* model/DistributedContext.java:60-98 */
typedef struct {
    int embeddingSegmentStart, embeddingSegmentLength;
    int attentionSegmentStart, attentionSegmentLength;
    int hiddenSegmentStart, hiddenSegmentLength;
    int kvSegmentStart, kvSegmentLength;
    int headStart, headEnd, groupHeadStart, groupHeadEnd;
    int numberOfLayers, layerStart, layerEnd;
} jo_dctx;

void jo_dctx_build(int E, int attentionLength, int H, int headSize, int headGroupSize, int numLayers, int modelShard,
                   int numModelShards, int layerShard, int numLayerShards, jo_dctx *d) {
    d->numberOfLayers = numLayers / numLayerShards;
    d->layerStart = d->numberOfLayers * layerShard;
    d->layerEnd = d->layerStart + d->numberOfLayers;
    d->embeddingSegmentLength = E / numModelShards;
    d->embeddingSegmentStart = d->embeddingSegmentLength * modelShard;
    d->attentionSegmentLength = attentionLength / numModelShards;
    d->attentionSegmentStart = d->attentionSegmentLength * modelShard;
    d->hiddenSegmentLength = H / numModelShards;
    d->hiddenSegmentStart = d->hiddenSegmentLength * modelShard;
    d->kvSegmentStart = d->attentionSegmentStart / headGroupSize;
    d->kvSegmentLength = d->attentionSegmentLength / headGroupSize;
    int embEnd = d->embeddingSegmentStart + d->embeddingSegmentLength;
    d->headStart = d->embeddingSegmentStart / headSize;
    d->headEnd = embEnd / headSize;
    d->groupHeadStart = d->kvSegmentStart / headSize;
    d->groupHeadEnd = (d->kvSegmentStart + d->kvSegmentLength) / headSize;
}

/* tensor/KvBufferCache.java:224-280 page-size solver; returns layersPerPage, ctxPerPage */
void jo_kv_page_solver(int numLayers, int contextLength, int kvSegmentLength, int dtypeSize, int64_t maxPageBytes,
                       int *layersPerPage, int *ctxPerPage) {
    int64_t s = 2LL * dtypeSize * kvSegmentLength;
    int optL = 1, optC = 1;
    int64_t maxProduct = 0;
    for (int x = numLayers; x >= 1; x--) {
        int64_t y = maxPageBytes / (x * s);
        if (y >= 1 && y <= contextLength) {
            int64_t product = x * y;
            if (product > maxProduct) { optL = x; optC = (int)y; maxProduct = product; }
            if (product < maxProduct) break;
        }
    }
    *layersPerPage = optL;
    *ctxPerPage = optC;
}

/* ------------------------------------------------------------------------ */
/* Llama model (model/llama/LlamaModel.java, model/AbstractModel.java,        */
/* model/TransformerBlock.java, model/CausalSelfAttention.java,               */
/* model/MLPBlock.java)                                                        */
/* ------------------------------------------------------------------------ */
enum { T_EMBED = 0, T_OUT_NORM, T_LM_HEAD, T_GLOBAL_COUNT };
enum { L_ATTN_NORM = 0, L_Q, L_K, L_V, L_O, L_FFN_NORM, L_GATE, L_DOWN, L_UP, L_COUNT };

typedef struct {
    /* config (safetensors/Config.java:230-287) */
    int ctx, E, H, heads, kv_heads, layers, vocab, head_size;
    float eps;
    double rope_theta, rope_scale;
    int act_q8; /* workingQType == I8 (AbstractModel.java:119-176) */
    /* derived */
    int attn_len, kv_len, group;
    jo_tensor g[T_GLOBAL_COUNT];
    jo_tensor *l; /* [layers][L_COUNT] */
    float *rope;  /* [(ctx + pad) * hs/2][2] */
    int rope_positions;
    /* KV pages (tensor/KvBufferCache.java:99-112,307-352) */
    int layers_per_page, ctx_per_page, n_layer_pages, n_ctx_pages;
    float **pages; /* [n_layer_pages * n_ctx_pages] of [lpp,2,cpp,kv_len] f32, lazily calloc'd */
    /* tensor-parallel simulation: numModelShards partial sums reduced in rank order */
    int tp;
    /* mixture of experts (model/MoEBlock.java, model/mixtral/MixtralModel.java:63-120); n_experts == 0: dense MLP */
    int n_experts, n_experts_per_tok;
    jo_tensor *moe_gate; /* [layers] router weight [n_experts, E] */
    jo_tensor *moe_w;    /* [layers][n_experts][3]: w1 (gate_proj), w2 (down_proj), w3 (up_proj) */
} jo_model;

jo_model *jo_model_create(int ctx, int E, int H, int heads, int kv_heads, int layers, int vocab, float eps,
                          double rope_theta, double rope_scale, int act_q8) {
    jo_model *m = (jo_model *)calloc(1, sizeof(jo_model));
    m->ctx = ctx; m->E = E; m->H = H; m->heads = heads; m->kv_heads = kv_heads; m->layers = layers; m->vocab = vocab;
    m->head_size = E / heads; /* Config.java:254 (headSize defaults to E/heads) */
    m->eps = eps; m->rope_theta = rope_theta; m->rope_scale = rope_scale; m->act_q8 = act_q8;
    m->attn_len = heads * m->head_size;
    m->kv_len = kv_heads * m->head_size;
    m->group = heads / kv_heads;
    m->l = (jo_tensor *)calloc((size_t)layers * L_COUNT, sizeof(jo_tensor));
    /* The reference table has ctx positions; head h looks up position pos+2*kvh
     * (CausalSelfAttention.java:260-268) and would throw past the end; we pad. */
    m->rope_positions = ctx + 2 * kv_heads;
    m->rope = (float *)malloc(sizeof(float) * 2 * (size_t)m->rope_positions * (m->head_size / 2));
    jo_precompute_freqs_cis(m->head_size, m->rope_positions, rope_theta, rope_scale, m->rope);
    jo_kv_page_solver(layers, ctx, m->kv_len, 4, 1 << 23, &m->layers_per_page, &m->ctx_per_page);
    m->n_layer_pages = (layers + m->layers_per_page - 1) / m->layers_per_page;
    m->n_ctx_pages = (ctx + m->ctx_per_page - 1) / m->ctx_per_page;
    m->pages = (float **)calloc((size_t)m->n_layer_pages * m->n_ctx_pages, sizeof(float *));
    m->tp = 1;
    return m;
}

void jo_model_set_tp(jo_model *m, int tp) { m->tp = tp; }
void jo_model_set_moe(jo_model *m, int n_experts, int n_experts_per_tok) {
    m->n_experts = n_experts;
    m->n_experts_per_tok = n_experts_per_tok;
    free(m->moe_gate);
    free(m->moe_w);
    m->moe_gate = (jo_tensor *)calloc((size_t)m->layers, sizeof(jo_tensor));
    m->moe_w = (jo_tensor *)calloc((size_t)m->layers * n_experts * 3, sizeof(jo_tensor));
}
/* expert < 0: the router ("block_sparse_moe.gate.weight"); which: 0 = w1, 1 = w2, 2 = w3 */
void jo_model_set_expert(jo_model *m, int layer, int expert, int which, int dtype, int64_t rows, int64_t cols, const void *data,
                         const float *scales) {
    jo_tensor *t = expert < 0 ? &m->moe_gate[layer] : &m->moe_w[((size_t)layer * m->n_experts + expert) * 3 + which];
    t->dtype = dtype; t->rows = rows; t->cols = cols; t->data = data; t->scales = scales;
}
void jo_model_kv_geometry(jo_model *m, int *lpp, int *cpp) { *lpp = m->layers_per_page; *cpp = m->ctx_per_page; }

void jo_model_reset_kv(jo_model *m) {
    for (int i = 0; i < m->n_layer_pages * m->n_ctx_pages; i++) {
        free(m->pages[i]);
        m->pages[i] = NULL;
    }
}

void jo_model_free(jo_model *m) {
    jo_model_reset_kv(m);
    free(m->pages); free(m->rope); free(m->l); free(m->moe_gate); free(m->moe_w); free(m);
}

/* layer < 0 -> global tensor `kind`; data is NOT copied */
void jo_model_set_tensor(jo_model *m, int layer, int kind, int dtype, int64_t rows, int64_t cols, const void *data,
                         const float *scales) {
    jo_tensor *t = layer < 0 ? &m->g[kind] : &m->l[(size_t)layer * L_COUNT + kind];
    t->dtype = dtype; t->rows = rows; t->cols = cols; t->data = data; t->scales = scales;
}

static float *kv_row(jo_model *m, int layer, int pos, int which) {
    int lp = layer / m->layers_per_page, cp = pos / m->ctx_per_page;
    int rl = layer % m->layers_per_page, rc = pos % m->ctx_per_page;
    float **pg = &m->pages[(size_t)lp * m->n_ctx_pages + cp];
    if (!*pg) *pg = (float *)calloc((size_t)m->layers_per_page * 2 * m->ctx_per_page * m->kv_len, sizeof(float));
    return *pg + (((size_t)rl * 2 + which) * m->ctx_per_page + rc) * m->kv_len;
}

/* debug/test access to a stored K or V row */
const float *jo_model_kv_row(jo_model *m, int layer, int pos, int which) { return kv_row(m, layer, pos, which); }

/* maybeQuantize (LlamaModel.java:176-184) then GEMM; scratch q/s sized by caller */
static void gemm_act(jo_model *m, float *result, int64_t ldc, const float *x, int64_t M, int64_t xcols,
                     const jo_tensor *w, int colOff, int K, int rowOff, int N, int8_t *qbuf, float *sbuf) {
    jo_tensor a;
    a.rows = M; a.cols = xcols;
    if (m->act_q8 && w->dtype == JO_Q4) {
        jo_quantize_q8_act(x, M, xcols, xcols, qbuf, sbuf);
        a.dtype = JO_I8; a.data = qbuf; a.scales = sbuf;
    } else {
        a.dtype = JO_F32; a.data = x; a.scales = NULL;
    }
    jo_batch_dot(result, ldc, &a, w, colOff, colOff, K, 0, rowOff, N);
}

/* Sum the per-shard partial GEMMs in rank order (JlamaService.combine :300-359 adds in
 * arrival order; rank order is one admissible order).  For tp==1 this is the plain GEMM. */
static void gemm_colsharded(jo_model *m, float *result, const float *x, int64_t M, int64_t xcols, const jo_tensor *w,
                            int N, int8_t *qbuf, float *sbuf, float *tmp) {
    if (m->tp == 1) {
        gemm_act(m, result, N, x, M, xcols, w, 0, (int)xcols, 0, N, qbuf, sbuf);
        return;
    }
    int seg = (int)(xcols / m->tp);
    for (int s = 0; s < m->tp; s++) {
        float *dst = s == 0 ? result : tmp;
        gemm_act(m, dst, N, x, M, xcols, w, s * seg, seg, 0, N, qbuf, sbuf);
        if (s > 0)
            for (int64_t i = 0; i < M * N; i++) result[i] += tmp[i];
    }
}

/* One AbstractModel.forward (AbstractModel.java:314-329) over M rows starting at startPos.
 * x: [M,E] f32 in/out (hidden state). */
static void forward_rows(jo_model *m, float *x, int M, int startPos) {
    int E = m->E, H = m->H, hs = m->head_size, hp = hs / 2;
    int64_t maxc = H > m->attn_len ? H : m->attn_len;
    if (E > maxc) maxc = E;
    float *ln = (float *)malloc(sizeof(float) * M * E);
    float *q = (float *)malloc(sizeof(float) * M * m->attn_len);
    float *k = (float *)malloc(sizeof(float) * M * m->kv_len);
    float *v = (float *)malloc(sizeof(float) * M * m->kv_len);
    float *val = (float *)malloc(sizeof(float) * M * m->attn_len);
    float *res = (float *)malloc(sizeof(float) * M * E);
    float *tmp = (float *)malloc(sizeof(float) * M * E);
    float *buf = (float *)malloc(sizeof(float) * M * H);
    float *buf2 = (float *)malloc(sizeof(float) * M * H);
    int8_t *qb = (int8_t *)malloc((size_t)M * maxc);
    float *sb = (float *)malloc(sizeof(float) * M * (maxc / QBLOCK + 1));
    float attnScale = (float)(1.0 / sqrt((double)hs)); /* CausalSelfAttention.java:134 */

    for (int L = 0; L < m->layers; L++) {
        jo_tensor *lw = &m->l[(size_t)L * L_COUNT];
        /* TransformerBlock.java:158-215 */
        jo_rmsnorm(x, M, E, &lw[L_ATTN_NORM], 0.0f, m->eps, E, 0, E, ln);
        /* CausalSelfAttention.java:161-171 (row-sharded q/k/v == plain full GEMM) */
        gemm_act(m, q, m->attn_len, ln, M, E, &lw[L_Q], 0, E, 0, m->attn_len, qb, sb);
        gemm_act(m, k, m->kv_len, ln, M, E, &lw[L_K], 0, E, 0, m->kv_len, qb, sb);
        gemm_act(m, v, m->kv_len, ln, M, E, &lw[L_V], 0, E, 0, m->kv_len, qb, sb);
        memset(val, 0, sizeof(float) * M * m->attn_len); /* TensorCache zero-init, :158 */

        for (int bi = 0; bi < M; bi++) { /* :199 one position at a time */
            int pos = startPos + bi;
            float *key = kv_row(m, L, pos, 0), *vrow = kv_row(m, L, pos, 1);
            memcpy(key, k + (size_t)bi * m->kv_len, sizeof(float) * m->kv_len); /* :230-243 */
            memcpy(vrow, v + (size_t)bi * m->kv_len, sizeof(float) * m->kv_len);
            float *query = q + (size_t)bi * m->attn_len;
            /* RoPE :247-311 with the table-index quirk (poffset + kvh*hs + j) */
            int64_t poffset = (int64_t)pos * hp;
            for (int h = 0; h < m->heads; h++) {
                int offset = h * hs;
                int goffset = (h / m->group) * hs;
                for (int i = offset, g = goffset; i < offset + hp; i++, g++) {
                    float q0 = query[i], q1 = query[i + hp];
                    const float *f = m->rope + (poffset + g) * 2;
                    float fcr = f[0], fci = f[1];
                    volatile float a0 = q0 * fcr, a1 = q1 * fci, b0 = q0 * fci, b1 = q1 * fcr;
                    query[i] = a0 - a1;
                    query[i + hp] = b0 + b1;
                }
            }
            for (int h = 0; h < m->kv_heads; h++) {
                int offset = h * hs;
                for (int i = offset; i < offset + hp; i++) {
                    float k0 = key[i], k1 = key[i + hp];
                    const float *f = m->rope + (poffset + i) * 2;
                    float fcr = f[0], fci = f[1];
                    volatile float a0 = k0 * fcr, a1 = k1 * fci, b0 = k0 * fci, b1 = k1 * fcr;
                    key[i] = a0 - a1;
                    key[i + hp] = b0 + b1;
                }
            }
            /* attention :314-356 */
#pragma omp parallel for schedule(static)
            for (int h = 0; h < m->heads; h++) {
                int xoffset = (h / m->group) * hs, yoffset = h * hs;
                float *attn = (float *)malloc(sizeof(float) * (pos + 1));
                for (int t = 0; t <= pos; t++)
                    attn[t] = dot_f32_f32(query + yoffset, kv_row(m, L, t, 0) + xoffset, hs);
                jo_scale(attnScale, attn, 1, pos + 1, 0, pos + 1);
                jo_softmax(attn, 0, pos + 1);
                float *out = val + (size_t)bi * m->attn_len;
                for (int t = 0; t <= pos; t++) {
                    const float *vr = kv_row(m, L, t, 1);
                    float al = attn[t];
                    for (int i = 0; i < hs; i++) out[yoffset + i] = fmaf(vr[xoffset + i], al, out[yoffset + i]);
                }
                free(attn);
            }
        }
        /* o_proj :363-378 (column-sharded; reducer sums shards) */
        gemm_colsharded(m, res, val, M, m->attn_len, &lw[L_O], E, qb, sb, tmp);
        /* residual TransformerBlock.java:185: lnattn = attn_out + embedding */
        for (int64_t i = 0; i < (int64_t)M * E; i++) res[i] = res[i] + x[i];
        /* pre-FF norm, MLP (MLPBlock.java:106-166) */
        jo_rmsnorm(res, M, E, &lw[L_FFN_NORM], 0.0f, m->eps, E, 0, E, ln);
        if (m->n_experts > 0) {
            /* MoEBlock.forward (MoEBlock.java:73-149), one row at a time.  The input is the maybeQuantize'd pre-FF norm
             * (TransformerBlock.java:192-196), so the router dot products and the expert GEMVs are Q8 x Q4 when the model is.
             * Quirk NOT reproduced: for expert 0 the reference copies the expert result into row 0 of `result` whatever the
             * batch row (MoEBlock.java:139-143 copyFrom(moeResult, 0, 0, E)); with one row per call (decode, or
             * jlama.max_batch_size = 1) that is row b, which is what is computed here for every row. */
            const int NE = m->n_experts, KE = m->n_experts_per_tok;
            float er[64];
            int sel[64];
            float *moe = (float *)malloc(sizeof(float) * E);
            for (int b = 0; b < M; b++) {
                const float *lnb = ln + (size_t)b * E;
                gemm_act(m, er, NE, lnb, 1, E, &m->moe_gate[L], 0, E, 0, NE, qb, sb); /* :82-88 dotProduct per expert */
                jo_softmax(er, 0, NE);                                                  /* :91 */
                for (int i = 0; i < KE; i++) sel[i] = i;                                /* topk :151-168: replace the minimum */
                for (int i = KE; i < NE; i++) {
                    int mn = 0;
                    for (int j = 1; j < KE; j++)
                        if (er[sel[j]] < er[sel[mn]]) mn = j;
                    if (er[i] > er[sel[mn]]) sel[mn] = i;
                }
                float *xrow = x + (size_t)b * E;
                for (int i = 0; i < KE; i++) {
                    const jo_tensor *ew = &m->moe_w[((size_t)L * NE + sel[i]) * 3];
                    gemm_act(m, buf, H, lnb, 1, E, &ew[0], 0, E, 0, H, qb, sb);  /* w1 */
                    gemm_act(m, buf2, H, lnb, 1, E, &ew[2], 0, E, 0, H, qb, sb); /* w3 */
                    for (int j = 0; j < H; j++) buf[j] = jo_silu(buf[j]) * buf2[j]; /* :118-125 */
                    /* w1/w3 are row-sharded and the [1,H] buffer is reduced before the un-sharded w2 (:127-134): disjoint
                     * segments, so sharding does not change the arithmetic */
                    gemm_act(m, moe, E, buf, 1, H, &ew[1], 0, H, 0, E, qb, sb);
                    if (i == 0) memcpy(xrow, moe, sizeof(float) * E);
                    else
                        for (int j = 0; j < E; j++) xrow[j] = xrow[j] + moe[j]; /* unweighted sum :139-143 */
                }
            }
            free(moe);
            for (int64_t i = 0; i < (int64_t)M * E; i++) x[i] = x[i] + res[i]; /* TransformerBlock.java:203 */
            continue;
        }
        gemm_act(m, buf, H, ln, M, E, &lw[L_GATE], 0, E, 0, H, qb, sb);
        gemm_act(m, buf2, H, ln, M, E, &lw[L_UP], 0, E, 0, H, qb, sb);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < (int64_t)M * H; i++) {
            float a = jo_silu(buf[i]);
            buf[i] = a * buf2[i]; /* maccumulate :141 */
        }
        gemm_colsharded(m, x, buf, M, H, &lw[L_DOWN], E, qb, sb, tmp);
        for (int64_t i = 0; i < (int64_t)M * E; i++) x[i] = x[i] + res[i]; /* :203 */
    }
    free(ln); free(q); free(k); free(v); free(val); free(res); free(tmp); free(buf); free(buf2); free(qb); free(sb);
}

/* LlamaModel.java:68-100: embedding row in its stored dtype, consumed through get() */
static void embed(jo_model *m, const int *tokens, int n, float *x) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < m->E; j++) x[(size_t)i * m->E + j] = t_get(&m->g[T_EMBED], tokens[i], j);
}

/* AbstractModel.batchForward :295-312 (chunks of max_batch) -> last chunk's hidden rows.
 * hidden_out: [min(n,max_batch) rows of last chunk, E]; returns rows in last chunk. */
int jo_model_batch_forward(jo_model *m, const int *tokens, int n, int startPos, int max_batch, float *hidden_last_row) {
    int rows_last = 0;
    float *x = (float *)malloc(sizeof(float) * (size_t)max_batch * m->E);
    for (int i = 0; i < n; i += max_batch) {
        int cnt = n - i < max_batch ? n - i : max_batch;
        embed(m, tokens + i, cnt, x);
        forward_rows(m, x, cnt, startPos + i);
        rows_last = cnt;
    }
    memcpy(hidden_last_row, x + (size_t)(rows_last - 1) * m->E, sizeof(float) * m->E);
    free(x);
    return rows_last;
}

/* AbstractModel.sample :443-473 at temperature 0: final RMSNorm, lm_head GEMV with
 * un-quantised F32 activations, argmax with strict '>' (lowest index wins). */
int jo_model_sample(jo_model *m, const float *hidden_row, float *logits) {
    float *e = (float *)malloc(sizeof(float) * m->E);
    jo_rmsnorm(hidden_row, 1, m->E, &m->g[T_OUT_NORM], 0.0f, m->eps, m->E, 0, m->E, e);
    jo_tensor a = {JO_F32, 1, m->E, e, NULL};
    const jo_tensor *w = m->g[T_LM_HEAD].data ? &m->g[T_LM_HEAD] : &m->g[T_EMBED];
    jo_batch_dot(logits, m->vocab, &a, w, 0, 0, m->E, 0, 0, m->vocab);
    int maxi = -1;
    double maxv = -INFINITY;
    for (int i = 0; i < m->vocab; i++)
        if (logits[i] > maxv) { maxi = i; maxv = logits[i]; }
    free(e);
    return maxi;
}

/* AbstractModel.sample :475-489, the part after the arg-max when temperature != 0: v = (float) exp((logit - maxv) / temperature)
 * with maxv a double and FastMath.exp restated by libm (parity unpinned, see the header), `sum` a float accumulated in index order,
 * then the first index whose float running sum of v / sum reaches uniformSample; vocab - 1 if none does.  `logits` is overwritten with
 * the exponentials exactly like the reference's logits tensor. */
int jo_sample_temperature(float *logits, int vocab, float temperature, float uniform) {
    double maxv = -INFINITY;
    for (int i = 0; i < vocab; i++)
        if (logits[i] > maxv) maxv = logits[i];
    float sum = 0;
    for (int i = 0; i < vocab; i++) {
        const float v = (float)exp(((double)logits[i] - maxv) / (double)temperature);
        sum += v;
        logits[i] = v;
    }
    float acc = 0;
    for (int i = 0; i < vocab; i++) {
        const float v = logits[i] / sum;
        acc += v;
        if (acc >= uniform) return i;
    }
    return vocab - 1;
}

/* AbstractModel.generate :516-646 at temperature 0 over token ids:
 * prefill prompt, then decode until n_total positions. out_tokens gets the
 * sampled tokens (first = sample after the prompt).  If logits_out != NULL it
 * receives the logits of every sampling step ([n_generated, vocab]). */
int jo_model_generate(jo_model *m, const int *prompt, int n_prompt, int n_new, int max_batch, int *out_tokens,
                      float *logits_out) {
    float *hidden = (float *)malloc(sizeof(float) * m->E);
    float *logits = (float *)malloc(sizeof(float) * m->vocab);
    jo_model_batch_forward(m, prompt, n_prompt, 0, max_batch, hidden);
    int next = jo_model_sample(m, hidden, logits);
    int produced = 0;
    out_tokens[produced] = next;
    if (logits_out) memcpy(logits_out, logits, sizeof(float) * m->vocab);
    produced++;
    for (int pos = n_prompt; produced < n_new; pos++) {
        jo_model_batch_forward(m, &next, 1, pos, max_batch, hidden);
        next = jo_model_sample(m, hidden, logits);
        out_tokens[produced] = next;
        if (logits_out) memcpy(logits_out + (size_t)produced * m->vocab, logits, sizeof(float) * m->vocab);
        produced++;
    }
    free(hidden); free(logits);
    return produced;
}

int jo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void jo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}
