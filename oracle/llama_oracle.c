/* TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT.
 *
 * CPU restatement of the reference's Q4_0/Q4_1 LLaMA decode path: what llama_eval()
 * (Sources/llamaObjCxx/bridge/LlamaPredictOperation.mm:510-735, "PO.mm") makes ggml
 * (Sources/cpp/ggml.c) compute, written as plain scalar C with the SIMD lane structure of the
 * reference's x86 AVX2 build spelled out (the flags tools/Makefile:34-35,78-98 selects:
 * -O3 -DNDEBUG -std=c11 -mavx -mavx2 -mfma -mf16c -msse3, i.e. FMA in the vector kernels, no
 * contraction in scalar code).  Every function cites the reference lines it follows.
 *
 * PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path
 * (SURVEY.md section 4), so this restatement is pinned against the reference itself: oracle/Makefile
 * compiles the unmodified reference sources where they lie into oracle/_ref/libllama_ref.so and
 * tests/test_oracle_vs_ref.py requires bit-identical results op by op and for whole llama_eval
 * calls (logits and KV cache), at several thread counts.  Small outputs of the reference are also
 * committed as fixtures under tests/golden/ (made by tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the
 * product (llama.swift_b200/) never links, imports or executes anything under oracle/.
 */
#define _GNU_SOURCE
#include "llama_oracle.h"

#include <immintrin.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define QK 32 /* ggml.c:360 */

/* ---- fp16 <-> fp32: GGML_COMPUTE_FP32_TO_FP16 = _cvtss_sh(x, 0) (round-to-nearest-even), ggml.c:159-162 */
uint16_t ora_fp32_to_fp16(float f) { return _cvtss_sh(f, 0); }
float ora_fp16_to_fp32(uint16_t h) { return _cvtsh_ss(h); }

/* ---- the two fp16 tables llama_eval uses, ggml.c:2377-2389 (silu: ggml.c:1944-1946) ------------------ */
static uint16_t g_table_silu[1 << 16];
static uint16_t g_table_exp[1 << 16];
static int g_tables_ready = 0;

static float silu_f32(float x) { return x / (1.0 + exp(-x)); } /* double math, result rounded to f32 */

static void init_tables(void) {
    if (g_tables_ready) return;
    for (int i = 0; i < (1 << 16); ++i) {
        const float f = ora_fp16_to_fp32((uint16_t) i);
        g_table_silu[i] = ora_fp32_to_fp16(silu_f32(f));
        g_table_exp[i] = ora_fp32_to_fp16((float) exp(f)); /* exp(double) -> f32 -> f16 */
    }
    g_tables_ready = 1;
}
const uint16_t *ora_table_silu_f16(void) { init_tables(); return g_table_silu; }
const uint16_t *ora_table_exp_f16(void) { init_tables(); return g_table_exp; }

/* ---- activation quantizers --------------------------------------------------------------------------- */

/* quantize_row_q4_0, AVX2 branch, ggml.c:456-523: d = amax/7.0f; id = amax ? 7.0f/amax : 0;
 * q = round-to-nearest-EVEN(x*id) + 8; byte j = q[2j] | q[2j+1] << 4; block = [f32 d][16 B]. */
void ora_quantize_row_q4_0(const float *x, void *y, int k) {
    const int nb = k / QK;
    uint8_t *out = (uint8_t *) y;
    for (int i = 0; i < nb; i++) {
        const float *xb = x + i * QK;
        float amax = 0.0f;
        for (int l = 0; l < QK; l++) {
            const float a = fabsf(xb[l]);
            if (a > amax) amax = a;
        }
        const float d = amax / 7.0f;
        const float id = (amax != 0.0f) ? 7.0f / amax : 0.0f;
        memcpy(out + i * 20, &d, 4);
        for (int l = 0; l < QK; l += 2) {
            const float v0 = xb[l + 0] * id;
            const float v1 = xb[l + 1] * id;
            const int q0 = (int) rintf(v0) + 8; /* default rounding mode = nearest-even, as _MM_ROUND_NEAREST */
            const int q1 = (int) rintf(v1) + 8;
            out[i * 20 + 4 + l / 2] = (uint8_t) (q0 | (q1 << 4));
        }
    }
}

/* quantize_row_q4_1, scalar only, ggml.c:606-648: row layout [nb f32 min][nb f32 d][nb*16 B];
 * d = (max-min)/15; id = d ? 1.0f/d : 0; q = round-half-away((x-min)*id). */
void ora_quantize_row_q4_1(const float *x, void *y, int k) {
    const int nb = k / QK;
    float *pm = (float *) y;
    float *pd = pm + nb;
    uint8_t *pb = (uint8_t *) (pd + nb);
    for (int i = 0; i < nb; i++) {
        float mn = 3.402823466e+38F, mx = -3.402823466e+38F;
        for (int l = 0; l < QK; l++) {
            const float v = x[i * QK + l];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        const float d = (mx - mn) / 15;
        const float id = d ? 1.0f / d : 0.0f;
        pm[i] = mn;
        pd[i] = d;
        for (int l = 0; l < QK; l += 2) {
            const float v0 = (x[i * QK + l + 0] - mn) * id;
            const float v1 = (x[i * QK + l + 1] - mn) * id;
            const uint8_t vi0 = (uint8_t) round(v0); /* C round(): half away from zero, in double */
            const uint8_t vi1 = (uint8_t) round(v1);
            pb[i * QK / 2 + l / 2] = vi0 | (vi1 << 4);
        }
    }
}

/* dequantize_row_q4_0, ggml.c:651-684: y = (q - 8) * d */
void ora_dequantize_row_q4_0(const void *x, float *y, int k) {
    const int nb = k / QK;
    const uint8_t *in = (const uint8_t *) x;
    for (int i = 0; i < nb; i++) {
        float d;
        memcpy(&d, in + i * 20, 4);
        for (int l = 0; l < QK; l += 2) {
            const uint8_t vi = in[i * 20 + 4 + l / 2];
            y[i * QK + l + 0] = (float) ((int) (vi & 0xf) - 8) * d;
            y[i * QK + l + 1] = (float) ((int) (vi >> 4) - 8) * d;
        }
    }
}

/* dequantize_row_q4_1, ggml.c:686-717: y = q*d + m (mul then add, no contraction) */
void ora_dequantize_row_q4_1(const void *x, float *y, int k) {
    const int nb = k / QK;
    const float *pm = (const float *) x;
    const float *pd = pm + nb;
    const uint8_t *pb = (const uint8_t *) (pd + nb);
    for (int i = 0; i < nb; i++) {
        for (int l = 0; l < QK; l += 2) {
            const uint8_t vi = pb[i * QK / 2 + l / 2];
            const float t0 = (float) (vi & 0xf) * pd[i];
            const float t1 = (float) (vi >> 4) * pd[i];
            y[i * QK + l + 0] = t0 + pm[i];
            y[i * QK + l + 1] = t1 + pm[i];
        }
    }
}

/* ---- dot products -------------------------------------------------------------------------------------- */

/* ggml_vec_dot_q4_0, AVX2 branch, ggml.c:1415-1466.  Eight f32 accumulator lanes; lane l of a block
 * holds the exact integer sum over elements {2l, 2l+1, 16+2l, 17+2l} of (qw-8)(qx-8) (madd_epi16 pairs of
 * the low 16 bytes plus pairs of the high 16 bytes), accumulated as acc[l] = fma(dw*dx, (float)isum, acc[l]);
 * horizontal sum: (acc[k]+acc[k+4]) for k<4, then (r0+r2)+(r1+r3). */
float ora_vec_dot_q4_0(int n, const void *vx, const void *vy) {
    const int nb = n / QK;
    const uint8_t *x = (const uint8_t *) vx;
    const uint8_t *y = (const uint8_t *) vy;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < nb; i++) {
        float d0, d1;
        memcpy(&d0, x + i * 20, 4);
        memcpy(&d1, y + i * 20, 4);
        const float scale = d0 * d1;
        const uint8_t *p0 = x + i * 20 + 4;
        const uint8_t *p1 = y + i * 20 + 4;
        int8_t a[32], b[32];
        for (int j = 0; j < 16; j++) { /* bytesFromNibbles, ggml.c:367-382: element 2j = low nibble of byte j */
            a[2 * j] = (int8_t) ((p0[j] & 0xf) - 8);
            a[2 * j + 1] = (int8_t) ((p0[j] >> 4) - 8);
            b[2 * j] = (int8_t) ((p1[j] & 0xf) - 8);
            b[2 * j + 1] = (int8_t) ((p1[j] >> 4) - 8);
        }
        for (int l = 0; l < 8; l++) {
            const int isum = a[2 * l] * b[2 * l] + a[2 * l + 1] * b[2 * l + 1] +
                             a[16 + 2 * l] * b[16 + 2 * l] + a[17 + 2 * l] * b[17 + 2 * l];
            acc[l] = fmaf(scale, (float) isum, acc[l]);
        }
    }
    const float r0 = acc[4] + acc[0], r1 = acc[5] + acc[1], r2 = acc[6] + acc[2], r3 = acc[7] + acc[3];
    const float s0 = r0 + r2, s1 = r1 + r3;
    return s0 + s1;
}

/* ggml_vec_dot_q4_1, scalar only, ggml.c:1584-1626: one sequential f32 chain
 * sumf += (d0*q0+m0)*(d1*q2+m1) + (d0*q1+m0)*(d1*q3+m1), no FMA (ISO C mode => -ffp-contract=off). */
float ora_vec_dot_q4_1(int n, const void *vx, const void *vy) {
    const int nb = n / QK;
    const float *pm0 = (const float *) vx, *pm1 = (const float *) vy;
    const float *pd0 = pm0 + nb, *pd1 = pm1 + nb;
    const uint8_t *pb0 = (const uint8_t *) (pd0 + nb), *pb1 = (const uint8_t *) (pd1 + nb);
    float sumf = 0.0f;
    for (int i = 0; i < nb; i++) {
        const float m0 = pm0[i], m1 = pm1[i], d0 = pd0[i], d1 = pd1[i];
        const uint8_t *p0 = pb0 + i * QK / 2, *p1 = pb1 + i * QK / 2;
        for (int j = 0; j < QK / 2; j++) {
            const uint8_t v0 = p0[j], v1 = p1[j];
            const float t0 = d0 * (float) (v0 & 0xf), t1 = d0 * (float) (v0 >> 4);
            const float t2 = d1 * (float) (v1 & 0xf), t3 = d1 * (float) (v1 >> 4);
            const float f0 = t0 + m0, f1 = t1 + m0, f2 = t2 + m1, f3 = t3 + m1;
            const float pa = f0 * f2, pb = f1 * f3;
            const float ps = pa + pb;
            sumf = sumf + ps;
        }
    }
    return sumf;
}

/* ggml_vec_dot_f32 with the AVX mapping (GGML_F32_STEP 32, EPR 8, ARR 4), ggml.c:1223-1258, reduce at
 * ggml.c:872-887: 4 vectors x 8 lanes of FMA chains, (s0+s1)+(s2+s3) lane-wise, then lanes k and k+4,
 * then hadd: (t0+t1)+(t2+t3).  Leftovers (n % 32) accumulate in double. */
float ora_vec_dot_f32(int n, const float *x, const float *y) {
    const int np = n & ~31;
    float sum[4][8];
    memset(sum, 0, sizeof(sum));
    for (int i = 0; i < np; i += 32)
        for (int j = 0; j < 4; j++)
            for (int l = 0; l < 8; l++)
                sum[j][l] = fmaf(x[i + 8 * j + l], y[i + 8 * j + l], sum[j][l]);
    float s[8];
    for (int l = 0; l < 8; l++) {
        const float a = sum[0][l] + sum[1][l];
        const float b = sum[2][l] + sum[3][l];
        s[l] = a + b;
    }
    float t[4];
    for (int k = 0; k < 4; k++) t[k] = s[k] + s[k + 4];
    const float h0 = t[0] + t[1], h1 = t[2] + t[3];
    double sumf = (float) (h0 + h1);
    for (int i = np; i < n; ++i) {
        const float p = x[i] * y[i];
        sumf += p;
    }
    return (float) sumf;
}

/* ggml_vec_mad_f32, ggml.c:1683-1712: y[i] = fma(x[i], v, y[i]) (leftovers: mul then add) */
void ora_vec_mad_f32(int n, float *y, const float *x, float v) {
    const int np = n & ~31;
    for (int i = 0; i < np; i++) y[i] = fmaf(x[i], v, y[i]);
    for (int i = np; i < n; i++) {
        const float p = x[i] * v;
        y[i] = y[i] + p;
    }
}

/* ggml_compute_forward_mul_mat_q4_0_f32 / _q4_1_f32, row-parallel branch, ggml.c:6134-6151 + 6182-6222
 * (6434-6451 + 6482-6522): quantize every src1 column, then out[ic*M + r] = vec_dot(W row r, xq col ic).
 * Rows are independent, so the result does not depend on the reference's thread count. */
void ora_mul_mat_q4(int type, const void *W, int M, int K, const float *x, int N, float *out) {
    const size_t row_bytes = (size_t) K / QK * (type == 2 ? 20 : 24);
    uint8_t *xq = (uint8_t *) malloc(row_bytes * (size_t) N);
    for (int ic = 0; ic < N; ic++) {
        if (type == 2) ora_quantize_row_q4_0(x + (size_t) ic * K, xq + ic * row_bytes, K);
        else ora_quantize_row_q4_1(x + (size_t) ic * K, xq + ic * row_bytes, K);
    }
#pragma omp parallel for schedule(static)
    for (int r = 0; r < M; r++) {
        const uint8_t *wr = (const uint8_t *) W + (size_t) r * row_bytes;
        for (int ic = 0; ic < N; ic++) {
            out[(size_t) ic * M + r] = type == 2 ? ora_vec_dot_q4_0(K, wr, xq + ic * row_bytes)
                                                 : ora_vec_dot_q4_1(K, wr, xq + ic * row_bytes);
        }
    }
    free(xq);
}

/* ---- small ops ------------------------------------------------------------------------------------------- */

/* ggml_compute_forward_norm_f32, ggml.c:5355-5381: LayerNorm (mean-subtracting!), double accumulators in
 * index order, eps = (double)1e-5f, scale rounded to f32, then ggml_vec_scale_f32 (f32 mul). */
void ora_norm(const float *x, float *y, int n) {
    double mean = 0.0;
    for (int i = 0; i < n; i++) mean += x[i];
    mean /= n;
    double sum2 = 0.0;
    for (int i = 0; i < n; i++) {
        const double v = x[i] - mean;
        y[i] = (float) v;
        sum2 += v * v;
    }
    const double eps = 1e-5f;
    const float scale = (float) (1.0 / sqrt(sum2 / n + eps));
    for (int i = 0; i < n; i++) y[i] = y[i] * scale;
}

/* ggml_compute_forward_rope_f32, ggml.c:7110-7127, one row of [head_dim, n_head] at position pos:
 * theta = pow(10000, -i0/n_dims), cos/sin(pos*theta) in double (gcc merges them into one sincos call in
 * the reference object code), rotate the interleaved pair in double, round to f32. */
void ora_rope(float *x, int n_head, int head_dim, int pos) {
    for (int h = 0; h < n_head; h++) {
        for (int i0 = 0; i0 < head_dim; i0 += 2) {
            const double theta = pow(10000.0, ((double) -i0) / head_dim);
            double s, c;
            sincos(pos * theta, &s, &c);
            float *p = x + (size_t) h * head_dim + i0;
            const double x0 = p[0], x1 = p[1];
            p[0] = (float) (x0 * c - x1 * s);
            p[1] = (float) (x0 * s + x1 * c);
        }
    }
}

/* ggml_compute_forward_soft_max_f32, ggml.c:7019-7041: max; e = f16->f32(table_exp_f16[f32->f16(x-max)]);
 * -inf -> 0; sum in double; scale by (float)(1.0/sum). */
void ora_soft_max(float *p, int n) {
    init_tables();
    double mx = -INFINITY;
    for (int i = 0; i < n; i++) mx = mx > p[i] ? mx : p[i];
    const float max = (float) mx;
    double sum = 0.0;
    for (int i = 0; i < n; i++) {
        if (p[i] == -INFINITY) {
            p[i] = 0.0f;
        } else {
            const uint16_t s = ora_fp32_to_fp16(p[i] - max);
            const float val = ora_fp16_to_fp32(g_table_exp[s]);
            sum += val;
            p[i] = val;
        }
    }
    const float inv = (float) (1.0 / sum);
    for (int i = 0; i < n; i++) p[i] = p[i] * inv;
}

/* ggml_vec_silu_f32 with GGML_SILU_FP16, ggml.c:1955-1963 */
void ora_silu(const float *x, float *y, int n) {
    init_tables();
    for (int i = 0; i < n; i++) y[i] = ora_fp16_to_fp32(g_table_silu[ora_fp32_to_fp16(x[i])]);
}

/* ---- model ------------------------------------------------------------------------------------------------ */

typedef struct {
    float *attention_norm, *ffn_norm;
    uint8_t *wq, *wk, *wv, *wo, *w1, *w2, *w3;
} ora_layer;

struct ora_model {
    int n_vocab, n_ctx, n_embd, n_mult, n_head, n_layer, n_rot, f16, n_ff;
    uint8_t *tok_embeddings, *output;
    float *norm;
    ora_layer *layers;
    float *memory_k, *memory_v; /* [n_layer][n_ctx][n_embd] f32, PO.mm:290-304 */
};

static void set_err(char *err, size_t errlen, const char *fmt, ...) {
    if (!err || !errlen) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, errlen, fmt, ap);
    va_end(ap);
}

static size_t q_row_bytes(int f16, int k) { return (size_t) k / QK * (f16 == 2 ? 20 : 24); }

/* name -> destination + shape; returns 0 if unknown (PO.mm:252-285) */
static int find_tensor(ora_model *m, const char *name, void **data, int *ne0, int *ne1, int *is2d) {
    const int e = m->n_embd, f = m->n_ff, v = m->n_vocab;
    *is2d = 1;
    if (!strcmp(name, "tok_embeddings.weight")) { *data = m->tok_embeddings; *ne0 = e; *ne1 = v; return 1; }
    if (!strcmp(name, "norm.weight")) { *data = m->norm; *ne0 = e; *ne1 = 1; *is2d = 0; return 1; }
    if (!strcmp(name, "output.weight")) { *data = m->output; *ne0 = e; *ne1 = v; return 1; }
    int il = -1, off = 0;
    if (sscanf(name, "layers.%d.%n", &il, &off) < 1 || il < 0 || il >= m->n_layer || off == 0) return 0;
    const char *s = name + off;
    ora_layer *L = &m->layers[il];
    if (!strcmp(s, "attention_norm.weight")) { *data = L->attention_norm; *ne0 = e; *ne1 = 1; *is2d = 0; return 1; }
    if (!strcmp(s, "ffn_norm.weight")) { *data = L->ffn_norm; *ne0 = e; *ne1 = 1; *is2d = 0; return 1; }
    if (!strcmp(s, "attention.wq.weight")) { *data = L->wq; *ne0 = e; *ne1 = e; return 1; }
    if (!strcmp(s, "attention.wk.weight")) { *data = L->wk; *ne0 = e; *ne1 = e; return 1; }
    if (!strcmp(s, "attention.wv.weight")) { *data = L->wv; *ne0 = e; *ne1 = e; return 1; }
    if (!strcmp(s, "attention.wo.weight")) { *data = L->wo; *ne0 = e; *ne1 = e; return 1; }
    if (!strcmp(s, "feed_forward.w1.weight")) { *data = L->w1; *ne0 = e; *ne1 = f; return 1; }
    if (!strcmp(s, "feed_forward.w2.weight")) { *data = L->w2; *ne0 = f; *ne1 = e; return 1; }
    if (!strcmp(s, "feed_forward.w3.weight")) { *data = L->w3; *ne0 = e; *ne1 = f; return 1; }
    return 0;
}

static int n_parts_for(int n_embd) { /* LLAMA_N_PARTS, PO.mm:33-38 */
    switch (n_embd) {
        case 4096: return 1;
        case 5120: return 2;
        case 6656: return 4;
        case 8192: return 8;
        default: return 0;
    }
}

void ora_free(ora_model *m) {
    if (!m) return;
    if (m->layers) {
        for (int i = 0; i < m->n_layer; i++) {
            ora_layer *L = &m->layers[i];
            free(L->attention_norm); free(L->ffn_norm);
            free(L->wq); free(L->wk); free(L->wv); free(L->wo); free(L->w1); free(L->w2); free(L->w3);
        }
        free(m->layers);
    }
    free(m->tok_embeddings); free(m->output); free(m->norm); free(m->memory_k); free(m->memory_v);
    free(m);
}

/* llama_model_load, PO.mm:98-498 */
ora_model *ora_load(const char *path, int n_ctx, char *err, size_t errlen) {
    init_tables();
    FILE *fin = fopen(path, "rb");
    if (!fin) { set_err(err, errlen, "failed to open '%s'", path); return NULL; }
    uint32_t magic = 0;
    if (fread(&magic, 4, 1, fin) != 1 || magic != 0x67676d6c) {
        set_err(err, errlen, "invalid model file '%s' (bad magic)", path);
        fclose(fin);
        return NULL;
    }
    ora_model *m = (ora_model *) calloc(1, sizeof(ora_model));
    int32_t hp[7];
    if (fread(hp, 4, 7, fin) != 7) { set_err(err, errlen, "truncated header"); fclose(fin); free(m); return NULL; }
    m->n_vocab = hp[0]; m->n_embd = hp[1]; m->n_mult = hp[2]; m->n_head = hp[3];
    m->n_layer = hp[4]; m->n_rot = hp[5]; m->f16 = hp[6];
    m->n_ctx = n_ctx;
    m->n_ff = ((2 * (4 * m->n_embd) / 3 + m->n_mult - 1) / m->n_mult) * m->n_mult; /* PO.mm:135 */
    const int n_parts = n_parts_for(m->n_embd);
    if (n_parts == 0) { set_err(err, errlen, "unsupported n_embd %d", m->n_embd); fclose(fin); free(m); return NULL; }
    for (int i = 0; i < m->n_vocab; i++) { /* vocab strings are skipped here (host-side concern) */
        uint32_t len;
        if (fread(&len, 4, 1, fin) != 1) { set_err(err, errlen, "truncated vocab"); fclose(fin); free(m); return NULL; }
        fseek(fin, len, SEEK_CUR);
    }
    if (m->f16 != 2 && m->f16 != 3) { /* the oracle restates the Q4_0 / Q4_1 paths only */
        set_err(err, errlen, "invalid model file '%s' (bad f16 value %d)", path, m->f16);
        fclose(fin); free(m); return NULL;
    }
    const long file_offset = ftell(fin);
    fclose(fin);

    const int e = m->n_embd, f = m->n_ff, v = m->n_vocab;
    m->tok_embeddings = (uint8_t *) malloc(q_row_bytes(m->f16, e) * v);
    m->output = (uint8_t *) malloc(q_row_bytes(m->f16, e) * v);
    m->norm = (float *) malloc(sizeof(float) * e);
    m->layers = (ora_layer *) calloc(m->n_layer, sizeof(ora_layer));
    for (int i = 0; i < m->n_layer; i++) {
        ora_layer *L = &m->layers[i];
        L->attention_norm = (float *) malloc(sizeof(float) * e);
        L->ffn_norm = (float *) malloc(sizeof(float) * e);
        L->wq = (uint8_t *) malloc(q_row_bytes(m->f16, e) * e);
        L->wk = (uint8_t *) malloc(q_row_bytes(m->f16, e) * e);
        L->wv = (uint8_t *) malloc(q_row_bytes(m->f16, e) * e);
        L->wo = (uint8_t *) malloc(q_row_bytes(m->f16, e) * e);
        L->w1 = (uint8_t *) malloc(q_row_bytes(m->f16, e) * f);
        L->w2 = (uint8_t *) malloc(q_row_bytes(m->f16, f) * e);
        L->w3 = (uint8_t *) malloc(q_row_bytes(m->f16, e) * f);
    }
    m->memory_k = (float *) calloc((size_t) m->n_layer * n_ctx * e, sizeof(float));
    m->memory_v = (float *) calloc((size_t) m->n_layer * n_ctx * e, sizeof(float));

    for (int part = 0; part < n_parts; part++) { /* PO.mm:312-495 */
        char fname[4096];
        if (part == 0) snprintf(fname, sizeof(fname), "%s", path);
        else snprintf(fname, sizeof(fname), "%s.%d", path, part);
        fin = fopen(fname, "rb");
        if (!fin) { set_err(err, errlen, "failed to open '%s'", fname); ora_free(m); return NULL; }
        fseek(fin, file_offset, SEEK_SET);
        for (;;) {
            int32_t hdr[3];
            if (fread(hdr, 4, 3, fin) != 3) break;
            const int n_dims = hdr[0], length = hdr[1], ftype = hdr[2];
            int32_t ne[2] = {1, 1};
            for (int i = 0; i < n_dims; i++) if (fread(&ne[i], 4, 1, fin) != 1) break;
            char name[256] = {0};
            if (length >= (int) sizeof(name) || fread(name, 1, length, fin) != (size_t) length) {
                set_err(err, errlen, "bad tensor name"); fclose(fin); ora_free(m); return NULL;
            }
            void *data; int t0, t1, is2d;
            if (!find_tensor(m, name, &data, &t0, &t1, &is2d)) {
                set_err(err, errlen, "unknown tensor '%s' in model file", name);
                fclose(fin); ora_free(m); return NULL;
            }
            int split_type = 0; /* PO.mm:358-388: 0 = by columns, 1 = by rows */
            if (strstr(name, "tok_embeddings")) split_type = 0;
            else if (strstr(name, "layers")) {
                if (strstr(name, "attention.wo.weight")) split_type = 0;
                else if (strstr(name, "feed_forward.w2.weight")) split_type = 0;
                else split_type = 1;
            } else if (strstr(name, "output")) split_type = 1;

            if (n_dims == 1) {
                if (t0 != ne[0] || ftype != 0) {
                    set_err(err, errlen, "tensor '%s' has wrong shape in model file", name);
                    fclose(fin); ora_free(m); return NULL;
                }
                if (part == 0) { if (fread(data, 4, t0, fin) != (size_t) t0) break; }
                else fseek(fin, 4L * t0, SEEK_CUR);
                continue;
            }
            if (ftype != m->f16) {
                set_err(err, errlen, "tensor '%s': ftype %d does not match model type %d", name, ftype, m->f16);
                fclose(fin); ora_free(m); return NULL;
            }
            const int ok = split_type == 0 ? (t0 / n_parts == ne[0] && t1 == ne[1])
                                           : (t0 == ne[0] && t1 / n_parts == ne[1]);
            if (!ok) {
                set_err(err, errlen, "tensor '%s' has wrong shape in model file", name);
                fclose(fin); ora_free(m); return NULL;
            }
            const size_t row_size = q_row_bytes(m->f16, t0);
            if (n_parts == 1) {
                if (fread(data, 1, row_size * t1, fin) != row_size * t1) break;
            } else if (split_type == 0) {
                const size_t bpb = m->f16 == 2 ? 20 : 24;
                for (int i1 = 0; i1 < ne[1]; i1++) {
                    const size_t offset = (size_t) i1 * row_size + ((size_t) part * ne[0] / QK) * bpb;
                    if (fread((uint8_t *) data + offset, 1, row_size / n_parts, fin) != row_size / n_parts) break;
                }
            } else {
                for (int i1 = 0; i1 < ne[1]; i1++) {
                    const size_t offset_row = ((size_t) i1 + (size_t) part * ne[1]) * row_size;
                    if (fread((uint8_t *) data + offset_row, 1, row_size, fin) != row_size) break;
                }
            }
        }
        fclose(fin);
    }
    return m;
}

int ora_n_vocab(const ora_model *m) { return m->n_vocab; }
int ora_n_ctx(const ora_model *m) { return m->n_ctx; }
int ora_n_embd(const ora_model *m) { return m->n_embd; }
int ora_n_layer(const ora_model *m) { return m->n_layer; }
int ora_n_head(const ora_model *m) { return m->n_head; }
int ora_ftype(const ora_model *m) { return m->f16; }

void ora_kv_export(const ora_model *m, int layer, int which, int n_rows, float *out) {
    const float *base = (which == 0 ? m->memory_k : m->memory_v) + (size_t) layer * m->n_ctx * m->n_embd;
    memcpy(out, base, sizeof(float) * (size_t) n_rows * m->n_embd);
}
void ora_kv_import(ora_model *m, int layer, int which, int n_rows, const float *in) {
    float *base = (which == 0 ? m->memory_k : m->memory_v) + (size_t) layer * m->n_ctx * m->n_embd;
    memcpy(base, in, sizeof(float) * (size_t) n_rows * m->n_embd);
}

/* norm (ggml.c:5327) followed by ggml_mul with the repeated weight row (PO.mm:570-575; ggml.c:4555-4580) */
static void norm_mul(const float *x, const float *w, float *y, int n, int N) {
    for (int i = 0; i < N; i++) {
        ora_norm(x + (size_t) i * n, y + (size_t) i * n, n);
        for (int j = 0; j < n; j++) y[(size_t) i * n + j] = w[j] * y[(size_t) i * n + j];
    }
}

/* llama_eval, PO.mm:510-735 (graph of PO.mm:561-706 executed in node order) */
int ora_eval(ora_model *m, int n_threads, int n_past, const int32_t *tokens, int N, float *logits,
             char *err, size_t errlen) {
    const int e = m->n_embd, f = m->n_ff, v = m->n_vocab, nh = m->n_head, hd = e / nh;
    const int P = n_past + N;
    if (N < 1 || n_past < 0 || P > m->n_ctx) { set_err(err, errlen, "bad n_past/n_tokens"); return -1001; }
    if (n_threads < 1) n_threads = 1;
    const int type = m->f16;

    float *inpL = (float *) malloc(sizeof(float) * (size_t) N * e);
    float *cur = (float *) malloc(sizeof(float) * (size_t) N * e);
    float *q = (float *) malloc(sizeof(float) * (size_t) N * e);
    float *kv = (float *) malloc(sizeof(float) * (size_t) N * e);
    float *att = (float *) malloc(sizeof(float) * (size_t) N * e);
    float *inpFF = (float *) malloc(sizeof(float) * (size_t) N * e);
    float *h1 = (float *) malloc(sizeof(float) * (size_t) N * f);
    float *h3 = (float *) malloc(sizeof(float) * (size_t) N * f);
    float *kq = (float *) malloc(sizeof(float) * (size_t) P);
    float *part = (float *) malloc(sizeof(float) * (size_t) hd * n_threads);
    float *out = (float *) malloc(sizeof(float) * (size_t) N * v);

    /* get_rows: dequantize embedding rows, ggml.c:6760-6812 */
    for (int i = 0; i < N; i++) {
        const uint8_t *row = m->tok_embeddings + (size_t) tokens[i] * q_row_bytes(type, e);
        if (type == 2) ora_dequantize_row_q4_0(row, inpL + (size_t) i * e, e);
        else ora_dequantize_row_q4_1(row, inpL + (size_t) i * e, e);
    }

    const float kq_scale = 1.0f / sqrtf((float) e / nh); /* PO.mm:620 (float sqrt overload) */

    for (int il = 0; il < m->n_layer; il++) {
        const ora_layer *L = &m->layers[il];
        float *Kc = m->memory_k + (size_t) il * m->n_ctx * e;
        float *Vc = m->memory_v + (size_t) il * m->n_ctx * e;

        norm_mul(inpL, L->attention_norm, cur, e, N);                              /* PO.mm:570-575 */
        ora_mul_mat_q4(type, L->wq, e, e, cur, N, q);                              /* PO.mm:580 */
        ora_mul_mat_q4(type, L->wk, e, e, cur, N, kv);                             /* PO.mm:581 */
        memcpy(Kc + (size_t) n_past * e, kv, sizeof(float) * (size_t) N * e);      /* PO.mm:586,589 */
        ora_mul_mat_q4(type, L->wv, e, e, cur, N, kv);                             /* PO.mm:582 */
        memcpy(Vc + (size_t) n_past * e, kv, sizeof(float) * (size_t) N * e);      /* PO.mm:587,590 */
        for (int i = 0; i < N; i++) {
            ora_rope(q + (size_t) i * e, nh, hd, n_past + i);                      /* PO.mm:594-601, mode 0 */
            ora_rope(Kc + (size_t) (n_past + i) * e, nh, hd, n_past + i);          /* PO.mm:604-611, mode 1, in place */
        }
        /* attention per (head, token): KQ (PO.mm:614), scale (617-621), mask (624), soft_max (627), V*P (638) */
        const int dc = (P + n_threads - 1) / n_threads; /* columns per thread, ggml.c:5628 */
        for (int h = 0; h < nh; h++) {
            for (int i = 0; i < N; i++) {
                const float *qv = q + (size_t) i * e + (size_t) h * hd;
                for (int j = 0; j < P; j++) {
                    float s = ora_vec_dot_f32(hd, Kc + (size_t) j * e + (size_t) h * hd, qv);
                    s = s * kq_scale;                                               /* ggml_vec_scale_f32 */
                    if (j > n_past + i) s = -INFINITY;                              /* ggml.c:6946-6953 */
                    kq[j] = s;
                }
                ora_soft_max(kq, P);
                /* transposed mul_mat branch: thread t accumulates columns [t*dc, min((t+1)*dc, P)) into its own
                 * zeroed buffer with vec_mad; FINALIZE copies buffer 0 and adds buffers 1.. in order (ggml.c:5570-5574) */
                memset(part, 0, sizeof(float) * (size_t) hd * n_threads);
                for (int t = 0; t < n_threads; t++) {
                    const int ic0 = dc * t, ic1 = (ic0 + dc < P) ? ic0 + dc : P;
                    for (int j = ic0; j < ic1; j++)
                        ora_vec_mad_f32(hd, part + (size_t) t * hd, Vc + (size_t) j * e + (size_t) h * hd, kq[j]);
                }
                float *dst = att + (size_t) i * e + (size_t) h * hd;               /* KQV_merged, PO.mm:641-646 */
                for (int d = 0; d < hd; d++) dst[d] = part[d];
                for (int t = 1; t < n_threads; t++)
                    for (int d = 0; d < hd; d++) dst[d] = dst[d] + part[(size_t) t * hd + d];
            }
        }
        ora_mul_mat_q4(type, L->wo, e, e, att, N, cur);                            /* PO.mm:649-651 */
        for (size_t i = 0; i < (size_t) N * e; i++) inpFF[i] = cur[i] + inpL[i];   /* PO.mm:654 */
        norm_mul(inpFF, L->ffn_norm, cur, e, N);                                   /* PO.mm:660-665 */
        ora_mul_mat_q4(type, L->w3, f, e, cur, N, h3);                             /* PO.mm:668-670 */
        ora_mul_mat_q4(type, L->w1, f, e, cur, N, h1);                             /* PO.mm:673-675 */
        ora_silu(h1, h1, N * f);                                                   /* PO.mm:678 */
        for (size_t i = 0; i < (size_t) N * f; i++) h1[i] = h1[i] * h3[i];         /* PO.mm:680 */
        ora_mul_mat_q4(type, L->w2, e, f, h1, N, cur);                             /* PO.mm:682-684 */
        for (size_t i = 0; i < (size_t) N * e; i++) inpL[i] = cur[i] + inpFF[i];   /* PO.mm:687 */
    }
    norm_mul(inpL, m->norm, cur, e, N);                                            /* PO.mm:694-701 */
    ora_mul_mat_q4(type, m->output, v, e, cur, N, out);                            /* PO.mm:705 (all N rows) */
    memcpy(logits, out + (size_t) (N - 1) * v, sizeof(float) * v);                 /* PO.mm:724-725 */

    free(inpL); free(cur); free(q); free(kv); free(att); free(inpFF); free(h1); free(h3);
    free(kq); free(part); free(out);
    return 0;
}
